mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_stream_gpu.py -q --timeout 300 2>&1 | tail -25 > gpurun_out/r2_stream2_test.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_bench_n2_ref.json 2> gpurun_out/r2_bench_n2_ref.err
echo done
