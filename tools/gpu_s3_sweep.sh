mkdir -p gpurun_out
for r in 512 1024 4096; do
SFK_FUSED_UP_RES=$r timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-calls 1 --dump-launches gpurun_out/s3_fu${r}_launches.json > gpurun_out/s3_fu${r}.json 2> gpurun_out/s3_fu${r}.err
done
echo done
