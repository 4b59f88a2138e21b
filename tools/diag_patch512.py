"""diagnostic: per-iteration gradient agreement of the 512 patch loop, oracle (GPU fp32) vs CUDA path"""
import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_fullsize_gpu import _models, _pairs, _to, _set_mode, _cos, _rel, DEV
from oracle.pipeline import AttackCfg as OCfg, LossCfg as OLoss, OraclePipeline, run_attack as oracle_run
from sfattack.attack_loop import AttackCfg, run_attack
from sfattack.engine import AttackEngine, LossCfg
S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
lr = float(sys.argv[2]) if len(sys.argv) > 2 else 2e3
spec, GP, es, EP, vsd, FP = _models(S)
xa, xb, g = _pairs(1, S, 71)
side = int(math.sqrt(0.1) * S); o = (S - side) // 2
mask = torch.zeros(1, 3, S, S); mask[:, :, o:o + side, o:o + side] = 1.0
patch0 = torch.rand(1, 3, S, S, generator=g)
cfg = dict(kind="patch", steps=3, lr=lr)
pipe = OraclePipeline(spec, _to(GP, DEV), es, _to(EP, DEV), _to(vsd, DEV), None)
rec = []
want = oracle_run(pipe, xa.to(DEV), xb.to(DEV), OCfg(loss=OLoss(1.0, 1.0), **cfg), mask=mask.to(DEV), patch0=patch0.expand(2, -1, -1, -1).contiguous().to(DEV), record=rec)
for mode in ("fp32", "bf16"):
    _set_mode(mode)
    eng = AttackEngine(spec, GP, es, EP, vsd, None, batch=1, device=DEV, loss=LossCfg(1.0, 1.0))
    rg = []
    got = run_attack(eng, xa.to(DEV), xb.to(DEV), AttackCfg(**cfg), mask=mask, patch0=patch0, record=rg)
    _set_mode("bf16")
    for it in range(3):
        a, b = rg[it], rec[it]
        print(mode, "it", it, "loss", a["loss"].tolist(), b["loss"].tolist(), "x rel", _rel(a["x"], b["x"]), "img rel", _rel(a["img"], b["img"]),
              "grad cos a/b", _cos(a["grad"][0], b["grad"][0]), _cos(a["grad"][1], b["grad"][1]), "gnorm", a["grad"][0].norm().item(), b["grad"][0].norm().item(),
              a["grad"][1].norm().item(), b["grad"][1].norm().item())
    del eng
