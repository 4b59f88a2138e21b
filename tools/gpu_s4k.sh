# early release of the TMEM accumulator stage (epilogue): A/B per shape and on the step
mkdir -p gpurun_out
{
for er in 0 1; do
echo "== SFK_EARLY_RELEASE=$er"
for spec in "8 1024 1024 32 32 fwd" "8 512 512 64 64 fwd" "8 256 256 128 128 fwd" "8 64 64 512 512 fwd" "8 512 512 64 32 fup"; do
  SFK_EARLY_RELEASE=$er SFK_FLAGS=518 python tests/prof_igemm.py $spec 10
done
SFK_EARLY_RELEASE=$er SFK_FLAGS=12 python tests/prof_igemm.py 8 256 256 64 64 fwd 10
for spec in "8 1024 1024 32 32 dgrad" "8 512 512 64 64 dgrad" "8 256 256 128 128 dgrad" "8 256 256 128 64 tconv"; do
  SFK_EARLY_RELEASE=$er SFK_FLAGS=0 python tests/prof_igemm.py $spec 10
done
done
} > gpurun_out/s4k_prof.log 2>&1
( timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_tf32_gpu.py -m gpu -q --timeout 600 2>&1 | tail -4 ) > gpurun_out/s4k_tests.log 2>&1
for v in 1 0 1 0; do
SFK_EARLY_RELEASE=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c3-line --e2e-calls 2 > gpurun_out/s4k_bench_$v.json 2> gpurun_out/s4k_bench_$v.err
python -c "
import json;d=json.load(open('gpurun_out/s4k_bench_$v.json'));print('early=$v', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'])" >> gpurun_out/s4k_ab.log
done
echo done
