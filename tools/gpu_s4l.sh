# four accumulator stages for narrow tiles: A/B per shape (with role cycles) and on the step
mkdir -p gpurun_out
{
for a4 in 0 1; do
echo "== SFK_ACC4=$a4"
for spec in "8 1024 1024 32 32 fwd" "8 512 512 64 64 fwd" "8 256 256 64 64 fwd" "8 512 512 64 32 fup"; do
  SFK_ACC4=$a4 SFK_FLAGS=518 python tests/prof_igemm.py $spec 10
done
SFK_ACC4=$a4 SFK_FLAGS=12 python tests/prof_igemm.py 8 256 256 64 64 fwd 10
SFK_ACC4=$a4 SFK_FLAGS=12 python tests/prof_igemm.py 16 256 256 16 32 fwd 10
for spec in "8 1024 1024 32 32 dgrad" "8 512 512 64 64 dgrad" "8 256 256 64 16 dgrad"; do
  SFK_ACC4=$a4 SFK_FLAGS=0 python tests/prof_igemm.py $spec 10
done
SFK_ACC4=$a4 SFK_FLAGS=518 SFK_ROLES=1 python tests/prof_igemm.py 8 1024 1024 32 32 fwd 5
SFK_ACC4=$a4 SFK_FLAGS=0 SFK_ROLES=1 python tests/prof_igemm.py 8 512 512 64 64 dgrad 5
done
} > gpurun_out/s4l_prof.log 2>&1
( timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_tf32_gpu.py tests/test_engine_gpu.py -m gpu -q --timeout 600 2>&1 | tail -4 ) > gpurun_out/s4l_tests.log 2>&1
for v in 1 0 1 0; do
SFK_ACC4=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c3-line --e2e-calls 2 > gpurun_out/s4l_bench_$v.json 2> gpurun_out/s4l_bench_$v.err
python -c "
import json;d=json.load(open('gpurun_out/s4l_bench_$v.json'));print('acc4=$v', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'])" >> gpurun_out/s4l_ab.log
done
echo done
