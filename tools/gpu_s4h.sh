# session 4: 1/2/4/8-GPU scaling of the final kernels on ONE 8-GPU box (device-timed value and end-to-end, strong = configs[2])
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s4h_n1.json 2> gpurun_out/s4h_n1.err
for n in 2 4 8; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/s4h_n$n.json 2> gpurun_out/s4h_n$n.err
done
echo done
