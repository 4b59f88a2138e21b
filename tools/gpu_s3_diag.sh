mkdir -p gpurun_out
python tools/diag_wgrad_stack.py > gpurun_out/s3_diag3.log 2>&1
echo done
