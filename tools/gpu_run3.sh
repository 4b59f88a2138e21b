mkdir -p gpurun_out
timeout 300 python tools/diag_tf32_accuracy.py > gpurun_out/r2_diag_tf32.log 2>&1
timeout 300 python tools/diag_patch512.py 512 2e3 > gpurun_out/r2_diag_patch.log 2>&1
timeout 300 python tools/diag_patch512.py 512 20 >> gpurun_out/r2_diag_patch.log 2>&1
echo done
