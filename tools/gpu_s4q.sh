mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_nway_gpu.py -q -s -m gpu --timeout 300 -k "torch_module or (nway_gradient and bf16)" 2>&1 | grep -E "^\[|passed|failed" | cut -c1-330 ) > gpurun_out/s4q_margins.log 2>&1
echo done
