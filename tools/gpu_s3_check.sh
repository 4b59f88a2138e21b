mkdir -p gpurun_out
for i in 1 2; do ( timeout 1200 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6 ) > gpurun_out/s3c_tests$i.log 2>&1; done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s3c_bench.json 2> gpurun_out/s3c_bench.err
timeout 300 python __graft_entry__.py smoke > gpurun_out/s3c_smoke.log 2>&1
echo done
