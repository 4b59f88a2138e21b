mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -25 ) > gpurun_out/s3_tests.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --dump-launches gpurun_out/s3_launches.json > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s3_ncu_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --e2e-calls 1 > gpurun_out/s3_ncu.log 2>&1
echo done
