mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:blur_stream -c 2 -o gpurun_out/s3_blur_ncu -f python tests/prof_elem.py 8 256 128 > gpurun_out/s3_blur_ncu.log 2>&1
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_configs_gpu.py tests/test_stream_gpu.py -q --timeout 600 2>&1 | tail -15 > gpurun_out/s3_blur_tests2.log
echo done
