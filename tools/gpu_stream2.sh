mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "act_bwd_and_torgb" --timeout 300 2>&1 | tail -15 > gpurun_out/r2_stream_test.log
{
for a in "8 1024 32" "8 512 64" "8 256 128" "8 128 256"; do
  python tests/prof_elem.py $a | grep "torgb_fwd"
done
} > gpurun_out/r2_stream_elem.log 2>&1
echo done
