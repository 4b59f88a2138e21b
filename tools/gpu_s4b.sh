# session 4, call b: first conv on the tensor cores (packed hi/lo operand) -- kernel test, full GPU suite, bench A/B
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "c3" --timeout 300 2>&1 | tail -15 ) > gpurun_out/s4b_c3.log 2>&1
( timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -15 ) > gpurun_out/s4b_tests.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c3-line --dump-launches gpurun_out/s4b_launches.json > gpurun_out/s4b_bench.json 2> gpurun_out/s4b_bench.err
SFK_C3_TC=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c3-line > gpurun_out/s4b_bench_off.json 2> gpurun_out/s4b_bench_off.err
echo done
