# session 4, call a: A/B of the epilogue micro-optimisations (packed fp32, incremental ring counters) and the blur barrier probe
mkdir -p gpurun_out
{
for spec in "8 1024 1024 32 32 fwd" "8 512 512 64 64 fwd" "8 256 256 64 64 fwd" "8 256 256 128 128 fwd" "8 64 64 512 512 fwd" "8 512 512 64 32 fup"; do
  SFK_FLAGS=518 python tests/prof_igemm.py $spec 10
done
SFK_FLAGS=12 python tests/prof_igemm.py 8 256 256 64 64 fwd 10
SFK_FLAGS=12 python tests/prof_igemm.py 8 128 128 128 128 fwd 10
for spec in "8 1024 1024 32 32 dgrad" "8 512 512 64 64 dgrad" "8 256 256 128 128 dgrad" "8 128 128 128 64 tconv" "8 256 256 128 64 tconv"; do
  SFK_FLAGS=0 python tests/prof_igemm.py $spec 10
done
SFK_FLAGS=518 SFK_ROLES=1 python tests/prof_igemm.py 8 1024 1024 32 32 fwd 5
python tests/prof_elem.py 8 1024 32 | grep blur
python tests/prof_elem.py 8 512 64 | grep blur
python tests/prof_elem.py 8 256 128 | grep blur
python tests/prof_elem.py 8 128 256 | grep blur
} > gpurun_out/s4a_prof.log 2>&1
( timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -8 ) > gpurun_out/s4a_tests.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c3-line > gpurun_out/s4a_bench.json 2> gpurun_out/s4a_bench.err
echo done
