mkdir -p gpurun_out
SFK_FLAGS=518 ncu --set full --clock-control none --import-source on -k regex:igemm_tc2 -s 2 -c 1 -o gpurun_out/r2_src_fwd1024 -f python tests/prof_igemm.py 8 1024 1024 32 32 fwd 2 > gpurun_out/r2_src_fwd1024.log 2>&1
SFK_FLAGS=0 ncu --set full --clock-control none --import-source on -k regex:igemm_tc2 -s 2 -c 1 -o gpurun_out/r2_src_dg1024 -f python tests/prof_igemm.py 8 1024 1024 32 32 dgrad 2 > gpurun_out/r2_src_dg1024.log 2>&1
echo done
