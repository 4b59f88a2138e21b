mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 300 -k "tconv_blur" 2>&1 | tail -5 > gpurun_out/s3_blur_test.log
(for a in "8 256 128" "8 512 64" "8 128 256" "8 64 512" "8 1024 32"; do timeout 120 python tests/prof_elem.py $a 2>&1 | grep -i "blur\|error\|Trace"; done) > gpurun_out/s3_blur_prof.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:blur_stream -c 6 -o gpurun_out/s3_blur_ncu -f python tests/prof_elem.py 8 256 128 > gpurun_out/s3_blur_ncu.log 2>&1
echo done
