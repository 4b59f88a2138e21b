# session 4 final: full GPU suite, bench (both arms), the other configs, ncu launch list + full captures of the final kernels
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -12 ) > gpurun_out/s4z_tests.log 2>&1

timeout 600 python bench.py --steps 10 --warmup 3 --dump-launches gpurun_out/s4z_launches.json > gpurun_out/s4z_bench.json 2> gpurun_out/s4z_bench.err


SFK_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s4z_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-calls 1 > gpurun_out/s4z_ncu.log 2>&1
SFK_NCU_RANGE=1 timeout 1500 ncu --profile-from-start off --set full --clock-control none -k regex:igemm_tc2 -c 62 --csv --page raw --log-file gpurun_out/s4z_igemm_full.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-calls 1 > gpurun_out/s4z_ncu2.log 2>&1
SFK_NCU_RANGE=1 timeout 1500 ncu --profile-from-start off --set full --clock-control none -k regex:"stream_kernel|update_linf|maxpool2|c3_pack|c3_unpack|torgb_fwd|mse_tap|image_loss|modulate_weights|avgpool|rgb_down" -c 90 --csv --page raw --log-file gpurun_out/s4z_elem_full.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-calls 1 > gpurun_out/s4z_ncu3.log 2>&1
echo done
