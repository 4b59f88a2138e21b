# knob re-check on the final kernels (each: one bench line, no CPU leg)
mkdir -p gpurun_out
rm -f gpurun_out/s4n_knobs.log
run() { env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c3-line --e2e-calls 1 > gpurun_out/s4n_tmp.json 2> gpurun_out/s4n_tmp.err; python -c "
import json,sys;d=json.load(open('gpurun_out/s4n_tmp.json'));print(' '.join(sys.argv[1:]), d['ms_per_step'], d['value'])" "$@" >> gpurun_out/s4n_knobs.log; }
run SFK_NONE=1
run SFK_FUSED_UP_RES=512
run SFK_FUSED_UP_RES=4096
run SFK_XSTORE=2
run SFK_NONE=1
echo done
