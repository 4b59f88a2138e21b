"""Per-iteration time of BASELINE.json configs[2..4] on one GPU (resident batch), as (T(K2) - T(K1)) / (K2 - K1) of run_attack so the
reference fusion and the final re-synthesis cancel.  Prints one JSON line per config with the roofline fraction against SURVEY 8(d)'s
algorithmic work (C3 360.3, C4 301.9, C5 423.6 GFLOP per iteration*pair) and the measured sustained bf16 peak."""
import json, math, os, sys, time
import torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sfattack.attack_loop import AttackCfg, run_attack
from sfattack.engine import AttackEngine, LossCfg
from sfattack.params import EncSpec, gen_spec, make_encoder_params, make_fusion_params, make_generator_params, make_vgg_state_dict
DEV = "cuda:0"

def models(size, seed=0):
    spec = gen_spec(size)
    es = EncSpec(n_latent=spec.n_latent)
    return spec, make_generator_params(spec, seed), es, make_encoder_params(es, seed + 1), make_vgg_state_dict(seed + 2), make_fusion_params(spec.s_dim, seed + 3)

def pairs(B, size, seed):
    g = torch.Generator().manual_seed(seed)
    mk = lambda: F.avg_pool2d(torch.rand(B, 3, size + 4, size + 4, generator=g), 5, 1)
    return mk().to(DEV), mk().to(DEV), g

def timed(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)

peak = 1393.1
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
except Exception:
    pass

def report(name, S, B, gflop, eng, mk_cfg, **kw):
    xa, xb, g = pairs(B, S, 5)
    k1, k2 = 4, 12
    t1 = timed(lambda: run_attack(eng, xa, xb, mk_cfg(k1), compute_final=False, **kw))
    t2 = timed(lambda: run_attack(eng, xa, xb, mk_cfg(k2), compute_final=False, **kw))
    ms = (t2 - t1) / (k2 - k1)
    rate = B / ms * 1e3
    print(json.dumps({"config": name, "size": S, "pairs": B, "ms_per_iteration": round(ms, 3), "iter_img_per_s": round(rate, 1),
                      "gflop_per_iter_pair": gflop, "tflops": round(rate * gflop / 1e3, 1), "frac_of_bf16_peak": round(rate * gflop / 1e3 / peak, 3)}), flush=True)

which = sys.argv[1:] or ["c3", "c4", "c5"]
if "c3" in which:
    S, B = 1024, 8
    spec, GP, es, EP, vsd, FP = models(S)
    eng = AttackEngine(spec, GP, es, EP, vsd, FP, fusion="spatial", batch=B, device=DEV, loss=LossCfg(1.0, 1.0))
    report("C3 PGD Linf, spatial fusion, 1024 (8 pairs resident of the 64 / N per GPU)", S, B, 360.3, eng,
           lambda k: AttackCfg(kind="linf", steps=k, graph=True), seed=1)
    del eng; torch.cuda.empty_cache()
if "c4" in which:
    S, B = 512, 4
    spec, GP, es, EP, vsd, FP = models(S)
    side = int(math.sqrt(0.1) * S); o = (S - side) // 2
    mask = torch.zeros(1, 3, S, S); mask[:, :, o:o + side, o:o + side] = 1.0
    eng = AttackEngine(spec, GP, es, EP, vsd, FP, fusion="spatial", batch=B, device=DEV, loss=LossCfg(1.0, 1.0))
    patch0 = torch.rand(1, 3, S, S)
    report("C4 adversarial patch 161x161 (raw-gradient step), spatial fusion, 512 (4 pairs per GPU = 32 / 8)", S, B, 301.9, eng,
           lambda k: AttackCfg(kind="patch", steps=k, alpha=2e3, graph=True), mask=mask, patch0=patch0)
    del eng; torch.cuda.empty_cache()
if "c5" in which:
    S, B = 1024, 8
    spec, GP, es, EP, vsd, FP = models(S)
    eps2 = 0.5 * math.sqrt(3 * S * S) / 255
    eng = AttackEngine(spec, GP, es, EP, vsd, FP, fusion="arithmetic", batch=B, device=DEV, loss=LossCfg(1.0, 1.0, 0.05))
    g = torch.Generator().manual_seed(3)
    start = (torch.rand(2, B, 3, S, S, generator=g) * 2 - 1) * ((1 / 255) / eps2)
    report("C5 L2 ball + VGG regulariser on the inputs, arithmetic fusion, 1024 (8 pairs resident of the 32 per GPU)", S, B, 423.6, eng,
           lambda k: AttackCfg(kind="l2", steps=k, eps=eps2, alpha=eps2 / 10, graph=True), start_noise=start)
