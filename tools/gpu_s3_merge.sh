mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_tf32_gpu.py tests/test_engine_gpu.py -q -x --timeout 600 2>&1 | tail -8 > gpurun_out/s3_merge_test.log
for r in 1024 4096; do
SFK_FUSED_UP_RES=$r timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-calls 1 --dump-launches gpurun_out/s3_mg${r}_launches.json > gpurun_out/s3_mg${r}.json 2> gpurun_out/s3_mg${r}.err
done
SFK_M2_TCONV=0 SFK_FUSED_UP_RES=1024 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-calls 1 > gpurun_out/s3_mg_off.json 2> gpurun_out/s3_mg_off.err
echo done
