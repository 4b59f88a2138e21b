mkdir -p gpurun_out
timeout 600 python tools/time_configs.py > gpurun_out/s4r_configs.jsonl 2> gpurun_out/s4r_configs.err
echo done
