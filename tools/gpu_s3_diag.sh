mkdir -p gpurun_out
python tools/diag_blur_repeat.py > gpurun_out/s3_diag2.log 2>&1
echo done
