mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_stream_gpu.py -q --timeout 300 2>&1 | tail -25 > gpurun_out/r2_stream2_test.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench2_ref.json 2> gpurun_out/r2_bench2_ref.err
echo done
