"""eager vs eager vs graph agreement of the Linf attack on the small engine (diagnostic)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_engine_gpu import _small_setup, DEV
from sfattack.attack_loop import AttackCfg, run_attack
from sfattack.engine import AttackEngine, LossCfg
spec, GP, es, EP, vsd, FP, xa, xb = _small_setup(size=64, B=2)
eng = AttackEngine(spec, GP, es, EP, vsd, FP, fusion="arithmetic", batch=2, device=DEV, loss=LossCfg(1.0, 1.0), vgg_res=64, vgg_width_div=4)
g = torch.Generator().manual_seed(5)
noise = torch.rand(2, 2, 3, 64, 64, generator=g) * 2 - 1
outs = []
for graph in (False, False, True, True):
    outs.append(run_attack(eng, xa.to(DEV), xb.to(DEV), AttackCfg(kind="linf", graph=graph, steps=6), start_noise=noise))
names = ["eager1", "eager2", "graph1", "graph2"]
for i in range(4):
    for j in range(i + 1, 4):
        same = ((outs[i]["x_adv"] - outs[j]["x_adv"]).abs() < 1e-6).float().mean().item()
        print(names[i], names[j], f"{same:.4f}", (outs[i]["losses"] - outs[j]["losses"]).abs().max().item())
print(outs[0]["losses"][:, 0].tolist())
