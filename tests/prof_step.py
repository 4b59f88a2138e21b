"""Per-launch CUDA-event timing of every tensor-core conv of one attack step (no profiler): python tests/prof_step.py [pairs] [size]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sfattack import lib
from sfattack.engine import AttackEngine, LossCfg

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    spec, GP, es, EP, vsd = bench.build_models(size)
    eng = AttackEngine(spec, GP, es, EP, vsd, None, batch=B, device="cuda:0", loss=LossCfg(1.0, 1.0))
    xa, xb = bench.synthetic_pairs(B, size)
    eng.set_inputs(xa.cuda(), xb.cuda()); eng.compute_reference()
    def step():
        _, g = eng.forward_backward()
        lib.attack_update_linf(eng.x, eng.x0, g, bench.ALPHA, bench.EPS, 1.0, 0.0, 1.0, eng.stats, eng.k_in)
    for _ in range(3): step()
    if os.environ.get("SFK_PROF_ALL"):
        # CUDA-event timing of EVERY wrapper call of one step, summed per entry point
        import collections, types
        rec = []
        names = [k for k, v in vars(lib).items() if isinstance(v, types.FunctionType) and not k.startswith("_") and k not in
                 ("load", "make_igemm_desc", "igemm_flops", "profile_igemm", "role_cycles", "pick_block_n", "conv3x3_taps",
                  "conv3x3_dgrad_taps", "tconv_taps", "tconv_dgrad_taps", "set_activation_dtype", "activation_dtype", "version")]
        orig = {k: getattr(lib, k) for k in names}
        def wrap(k, f):
            def g(*a, **kw):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); r = f(*a, **kw); e1.record(); rec.append((k, e0, e1)); return r
            return g
        for k in names: setattr(lib, k, wrap(k, orig[k]))
        step(); torch.cuda.synchronize()
        for k in names: setattr(lib, k, orig[k])
        agg = collections.defaultdict(lambda: [0, 0.0])
        for k, e0, e1 in rec:
            agg[k][0] += 1; agg[k][1] += e0.elapsed_time(e1) * 1e3
        tot = sum(v for _, v in agg.values())
        print(f"{len(rec)} calls, {tot:.0f} us")
        only = os.environ.get("SFK_PROF_NAME")
        if only:
            for k, e0, e1 in rec:
                if k in only.split(","): print(f"    {k}: {e0.elapsed_time(e1) * 1e3:.1f} us")
        for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
            print(f"  {k:24s} {c:3d} {v:8.1f} us {100 * v / tot:5.1f}%")
        return
    prof = lib.profile_igemm(step)
    tot = prof["ms"]
    print(f"{prof['launches']} conv launches, {tot:.3f} ms, {prof['flops']/tot/1e9:.1f} TFLOP/s")
    print("  n   out_h out_w  a_c out_c taps acc |    us   TFLOP/s")
    for ms, fl, shp in prof["per_launch"]:
        print("  %d %6d %5d %4d %5d %4d %3d | %7.1f %8.1f" % (*shp, ms * 1e3, fl / ms / 1e9))

if __name__ == "__main__":
    main()
