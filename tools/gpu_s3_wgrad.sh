mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_wgrad_gpu.py -q --timeout 200 2>&1 | tail -40 > gpurun_out/s3_wgrad_test.log
echo done
