# what the driver runs at round end, on the committed tree
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 ) > gpurun_out/s4v_tests.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/s4v_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/s4v_smoke.log
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/s4v_bench_ref.json 2> gpurun_out/s4v_bench_ref.err; echo "ref rc=$?" >> gpurun_out/s4v_smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/s4v_bench.json 2> gpurun_out/s4v_bench.err; echo "bench rc=$?" >> gpurun_out/s4v_smoke.log
echo done
