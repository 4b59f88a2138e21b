mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 300 -k "tconv_blur" 2>&1 | tail -5 > gpurun_out/s3_blur_test.log
(for a in "8 256 128" "8 512 64" "8 128 256" "8 64 512" "8 1024 32"; do timeout 120 python tests/prof_elem.py $a 2>&1 | grep -i "blur\|error\|Trace"; done) > gpurun_out/s3_blur_prof.log 2>&1
for r in 256 512 1024 4096; do
SFK_FUSED_UP_RES=$r timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-calls 1 --dump-launches gpurun_out/s3_bl${r}_launches.json > gpurun_out/s3_bl${r}.json 2> gpurun_out/s3_bl${r}.err
done
echo done
