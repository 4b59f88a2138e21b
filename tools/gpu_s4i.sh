mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_stream_gpu.py tests/test_configs_gpu.py -m gpu -q --timeout 600 2>&1 | tail -6 ) > gpurun_out/s4i_tests.log 2>&1
for v in 1 0 1 0; do
SFK_MOD_OVERLAP=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c3-line --e2e-calls 2 > gpurun_out/s4i_bench_$v.json 2> gpurun_out/s4i_bench_$v.err
python -c "
import json;d=json.load(open('gpurun_out/s4i_bench_$v.json'));print('overlap=$v', d['ms_per_step'], d['value'], d['e2e']['value'])" >> gpurun_out/s4i_ab.log
done
echo done
