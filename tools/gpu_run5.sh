mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -rA -s > gpurun_out/r2_t5.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
echo done
