mkdir -p gpurun_out
python -c "import torch; print(torch.cuda.get_device_name(0))" > gpurun_out/r2_env.log 2>&1
timeout 900 python -m pytest tests/test_tf32_gpu.py -q --timeout 300 -rA 2>&1 | tail -80 > gpurun_out/r2_t_tf32.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -rA --deselect tests/test_tf32_gpu.py -s 2>&1 | grep -v "^$" | tail -250 > gpurun_out/r2_t_all.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-launches gpurun_out/r2_launches_a.json > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
SFK_SPECIALIZE=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_nospec.json 2> gpurun_out/r2_bench_nospec.err
SFK_LOWRES_BN=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_nolowres.json 2> gpurun_out/r2_bench_nolowres.err
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1
echo done
