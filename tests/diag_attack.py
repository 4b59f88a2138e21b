"""Per-iteration comparison of the CUDA attack loop with the oracle (prints, never asserts)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_engine_gpu import _small_setup, _cos, _relerr, DEV  # noqa
from oracle.pipeline import AttackCfg as OCfg, LossCfg as OLoss, OraclePipeline, run_attack as oracle_run
from sfattack.attack_loop import AttackCfg, run_attack
from sfattack.engine import AttackEngine, LossCfg


def main():
    for kind in sys.argv[1:] or ["linf", "l2", "patch", "adam"]:
        spec, GP, es, EP, vsd, FP, xa, xb = _small_setup(size=32)
        B = xa.shape[0]
        creg = 0.5 if kind == "l2" else 0.0
        pipe = OraclePipeline(spec, GP, es, EP, vsd, FP, vgg_res=32)
        eng = AttackEngine(spec, GP, es, EP, vsd, FP, batch=B, device=DEV, loss=LossCfg(1.0, 1.0, creg), vgg_res=32, vgg_width_div=4)
        g = torch.Generator().manual_seed(12)
        noise = torch.rand(2, B, 3, 32, 32, generator=g) * 2 - 1
        kw = {}
        if kind == "linf":
            c = dict(kind="linf", steps=5)
        elif kind == "l2":
            c = dict(kind="l2", steps=3, eps=1.0, alpha=0.3)
        elif kind == "patch":
            mask = torch.zeros(1, 3, 32, 32); mask[..., 10:20, 10:20] = 1
            patch0 = torch.rand(2 * B, 3, 32, 32, generator=g)
            c = dict(kind="patch", steps=3, lr=50.0)
            kw = dict(mask=mask, patch0=patch0)
        else:
            c = dict(kind="adam", steps=3, lr=5e-3, random_start=False, targeted=True)
            kw = dict(target=(xb.flip(0), xa.flip(0)))
        r1, r2 = [], []
        kw2 = {k: (tuple(t.to(DEV) for t in v) if isinstance(v, tuple) else v) for k, v in kw.items()}
        o_ref = oracle_run(pipe, xa, xb, OCfg(loss=OLoss(1.0, 1.0, creg), **c), start_noise=noise, record=r1, **kw)
        o = run_attack(eng, xa.to(DEV), xb.to(DEV), AttackCfg(**c), start_noise=noise, record=r2, **kw2)
        print(f"==== {kind}")
        for it, (a, b) in enumerate(zip(r1, r2)):
            print(f" it{it}: loss ref {a['loss'].tolist()} gpu {b['loss'].tolist()}  x rel {_relerr(b['x'], a['x']):.3e} "
                  f"grad cos {_cos(b['grad'], a['grad']):.4f} rel {_relerr(b['grad'], a['grad']):.3e} |g| {a['grad'].abs().max():.3e} img rel {_relerr(b['img'], a['img']):.3e}")
        X0 = torch.cat([xa, xb])
        d_ref, d = o_ref["x_adv"] - X0, o["x_adv"].cpu() - X0
        print(f" final: delta rel {_relerr(d, d_ref):.3e} max|delta_ref| {d_ref.abs().max():.4f} max err {(d - d_ref).abs().max():.4f} "
              f"frac<1e-3 {((d - d_ref).abs() < 1e-3).float().mean():.4f}")


if __name__ == "__main__":
    main()
