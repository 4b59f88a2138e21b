mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_nway_gpu.py -m gpu -q -s --timeout 600 2>&1 | grep -E "^\[|passed|failed|Error|error|assert" | cut -c1-400 ) > gpurun_out/s4f_nway.log 2>&1
( timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_stream_gpu.py -m gpu -q -k "torch_module or stream_matches" --timeout 600 2>&1 | tail -30 ) > gpurun_out/s4f_encmod.log 2>&1
for i in 1 2; do ( timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -8 ) > gpurun_out/s4f_tests$i.log 2>&1; done
echo done
