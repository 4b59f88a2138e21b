"""Tiny end-to-end workload (a few seconds): every attack rule, both fusion modes, fused and unfused up-layers, odd batch, on a
64x64 model.  Meant for memory checkers / debuggers where one is available: python tests/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from sfattack.attack_loop import AttackCfg, run_attack
from sfattack.engine import AttackEngine, LossCfg
from test_engine_gpu import _small_setup

DEV = "cuda:0"
for fused_res in ("8", "1000"):
    os.environ["SFK_FUSED_UP_RES"] = fused_res
    for fusion in ("arithmetic", "spatial"):
        spec, GP, es, EP, vsd, FP, xa, xb = _small_setup(size=64, fusion=fusion, B=3)
        eng = AttackEngine(spec, GP, es, EP, vsd, FP, fusion=fusion, batch=3, device=DEV, loss=LossCfg(1.0, 1.0, 0.1), vgg_res=64,
                           vgg_width_div=4)
        g = torch.Generator().manual_seed(3)
        noise = torch.rand(2, 3, 3, 64, 64, generator=g) * 2 - 1
        mask = torch.zeros(1, 3, 64, 64); mask[:, :, 20:40, 20:40] = 1
        for cfg, kw in ((AttackCfg(kind="linf", steps=2), dict(start_noise=noise)),
                        (AttackCfg(kind="l2", steps=2, eps=1.0, alpha=0.2), dict(start_noise=noise * 0.01)),
                        (AttackCfg(kind="patch", steps=2, alpha=1.0), dict(mask=mask, patch0=torch.rand(1, 3, 64, 64, generator=g))),
                        (AttackCfg(kind="adam", steps=2, lr=0.01, targeted=True), dict(target=(xb.to(DEV), xa.to(DEV))))):
            out = run_attack(eng, xa.to(DEV), xb.to(DEV), cfg, **kw)
            assert torch.isfinite(out["x_adv"]).all(), cfg.kind
        eng.check()
        print("ok", fused_res, fusion, flush=True)
print("sanitize_small: done")
