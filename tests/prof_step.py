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
    prof = lib.profile_igemm(step)
    tot = prof["ms"]
    print(f"{prof['launches']} conv launches, {tot:.3f} ms, {prof['flops']/tot/1e9:.1f} TFLOP/s")
    print("  n   out_h out_w  a_c out_c taps acc |    us   TFLOP/s")
    for ms, fl, shp in prof["per_launch"]:
        print("  %d %6d %5d %4d %5d %4d %3d | %7.1f %8.1f" % (*shp, ms * 1e3, fl / ms / 1e9))

if __name__ == "__main__":
    main()
