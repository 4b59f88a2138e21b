mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_surface_gpu.py -q -m gpu --timeout 300 -k "torch_module_encoder or loss_menus" 2>&1 | grep -E "^E  |passed|failed|^FAILED" | cut -c1-600 ) > gpurun_out/s4q_recon.log 2>&1
( timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 ) > gpurun_out/s4q_suite.log 2>&1
echo done
