mkdir -p gpurun_out
rm -f gpurun_out/s4p_stream.log
for i in 1 2; do
( timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_nway_gpu.py tests/test_oracle_golden.py tests/test_stream_gpu.py -q -m gpu --timeout 300 -k "not 5way" 2>&1 | grep -E "^E  |passed|failed|^FAILED" | cut -c1-1200 ) >> gpurun_out/s4p_stream.log 2>&1
echo "---- $i" >> gpurun_out/s4p_stream.log
done
( timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 ) > gpurun_out/s4p_suite.log 2>&1
echo done
