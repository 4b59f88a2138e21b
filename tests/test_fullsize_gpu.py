"""Oracle comparisons AT the sizes BASELINE.json names and the attack-success criterion.

  * 1024x1024 (configs 2, 3, 5) and 512x512 (config 4), one pair: the first loss / fused image / gradient and a two-step PGD (a
    three-step patch loop at 512) against the oracle.  The oracle is plain torch, so for these sizes it is run on the SAME GPU in
    fp32 with TF32 switched off (it is the checker, never the thing measured); both storage modes of the product path are held to
    their stated tolerance: fp32 storage + split-tf32 tensor-core conv to north_star's 1e-3, bf16 storage to the bf16 numbers.
    The test asserts that the launch variants that only exist at these sizes were exercised: two M tiles per stage (m2), halo
    loads with resident weights, depth-to-space / space-to-depth fused upsampling.
  * success criterion: attack succeeds  <=>  pixel MSE between the benign and the adversarial fusion (cal_result's `or_f_ad_f`,
    code/attack/interpolation.py:1082) exceeds tau.  16 fixed-seed pairs are attacked with budgets eps in {1,2,4,8}/255 (so the
    outcomes are spread); tau sits in the widest gap of the ORACLE's outcome distribution around its median (so the criterion is
    not decided by rounding noise); the CUDA path must reproduce the oracle's 16 success bits in both storage modes.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _to(d, dev):
    return {k: v.to(dev) for k, v in d.items()} if d is not None else None


def _models(size, seed=0):
    from sfattack.params import (EncSpec, gen_spec, make_encoder_params, make_fusion_params, make_generator_params,
                                 make_vgg_state_dict)
    spec = gen_spec(size)
    GP = make_generator_params(spec, seed)
    es = EncSpec(n_latent=spec.n_latent)
    return spec, GP, es, make_encoder_params(es, seed + 1), make_vgg_state_dict(seed + 2), make_fusion_params(spec.s_dim, seed + 3)


def _pairs(B, size, seed):
    g = torch.Generator().manual_seed(seed)
    mk = lambda: F.avg_pool2d(torch.rand(B, 3, size + 4, size + 4, generator=g), 5, 1)
    return mk(), mk(), g


def _cos(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _set_mode(mode):
    from sfattack import lib
    lib.set_activation_dtype(torch.bfloat16 if mode == "bf16" else torch.float32)
    lib.set_conv_math("auto")


def _variants(eng):
    """planner decisions of every tensor-core launch of the engine (lib.plan_info)"""
    from sfattack import lib
    descs = []
    for st in (eng.enc, eng.vgg):
        descs += list(st._fwd_desc.values()) + list(st._bwd_desc.values())
    for e in eng.syn.L:
        descs += [e[k] for k in ("fwd", "bwd") if k in e]
    return [lib.plan_info(d) for d in descs]


@pytest.mark.parametrize("fusion", ["arithmetic", "spatial"])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_1024_gradient_and_pgd2_vs_oracle(mode, fusion):
    from oracle.pipeline import AttackCfg as OCfg, LossCfg as OLoss, OraclePipeline, run_attack as oracle_run
    from sfattack.attack_loop import AttackCfg, run_attack
    from sfattack.engine import AttackEngine, LossCfg
    S = 1024
    spec, GP, es, EP, vsd, FP = _models(S)
    import os
    xa, xb, g = _pairs(1, S, int(os.environ.get("SFK_TEST_SEED", "61")))      # (the knob is for tools/gpu_s4e.sh: other pairs)
    noise = torch.rand(2, 1, 3, S, S, generator=g) * 2 - 1
    eps, alpha = 8 / 255, 2 / 255
    pipe = OraclePipeline(spec, _to(GP, DEV), es, _to(EP, DEV), _to(vsd, DEV), _to(FP, DEV), fusion=fusion)
    rec = []
    want = oracle_run(pipe, xa.to(DEV), xb.to(DEV), OCfg(kind="linf", steps=2, eps=eps, alpha=alpha, loss=OLoss(1.0, 1.0)),
                      start_noise=noise.to(DEV), record=rec)
    torch.cuda.empty_cache()
    _set_mode(mode)
    try:
        eng = AttackEngine(spec, GP, es, EP, vsd, FP, fusion=fusion, batch=1, device=DEV, loss=LossCfg(1.0, 1.0))
        rec_g = []
        got = run_attack(eng, xa.to(DEV), xb.to(DEV), AttackCfg(kind="linf", steps=2, eps=eps, alpha=alpha), start_noise=noise, record=rec_g)
        info = _variants(eng)
    finally:
        _set_mode("bf16")
    # --- the launch variants that only trigger at these sizes were on the path
    assert any(i["m2"] for i in info), "no two-M-tile launch"
    assert any(i["halo"] and i["resident"] for i in info), "no halo/resident-weight launch"
    assert any(i["d2s"] for i in info) and any(i["s2d"] for i in info), "no fused-upsample launch"
    assert all(i["passes"] == (3 if mode == "fp32" else 1) and not i["cuda_cores"] for i in info)
    scale = want["fused_ref"].abs().max().item()
    ref_err = (got["fused_ref"] - want["fused_ref"]).abs().max().item()
    loss_rel = _rel(got["losses"][0], want["losses"][0])
    g_ref, g_got = rec[0]["grad"], rec_g[0]["grad"]
    cos = _cos(g_got, g_ref)
    band = g_ref.abs() > (1e-3 if mode == "fp32" else 5e-2) * g_ref.abs().mean()
    agree = (torch.sign(g_got)[band] == torch.sign(g_ref)[band]).float().mean().item()
    within = ((got["x_adv"] - want["x_adv"]).abs() < 1e-3).float().mean().item()
    fused_err = (got["fused_adv"] - want["fused_adv"]).abs().max().item()
    d_ref = ((want["fused_adv"] - want["fused_ref"]) ** 2).mean().item()
    d_got = ((got["fused_adv"] - got["fused_ref"]) ** 2).mean().item()
    # first step alone (x before the second update), outside the sign tie band of the oracle's first gradient
    step1 = ((rec_g[1]["x"] - rec[1]["x"]).abs() < 1e-3)[band].float().mean().item()
    loss_abs = (got["losses"][0] - want["losses"][0]).abs().max().item()
    print(f"[1024 {fusion} {mode}] image range +-{scale:.2f}: clean fusion max-abs err {ref_err:.2e}, loss rel {loss_rel:.2e} (abs {loss_abs:.2e}), "
          f"grad cos {cos:.6f}, sign agreement {agree:.4f} (band excludes {(~band).float().mean():.4f}), x after step 1 within 1e-3 (outside band): "
          f"{step1:.5f}, x_adv after 2 steps within 1e-3 (all pixels): {within:.4f}, adversarial fusion max-abs err {fused_err:.2e}, "
          f"outcome MSE {d_got:.4e} vs {d_ref:.4e}")
    if mode == "fp32":
        # north_star: fused image and perturbation within 1e-3 max-abs, on identical outcomes.  The perturbation lives in [0,1]; the
        # random-init generator's image spans +-scale, so its bound is 1e-3 of that range (measured: see DESIGN.md section 6)
        assert ref_err < 1e-3 * max(1.0, scale) and loss_rel < 2e-3 and cos > 0.9995 and agree > 0.995 and step1 > 0.999 and within > 0.95
        assert abs(d_got - d_ref) <= 0.03 * d_ref
    else:
        # bf16 storage at 1024^2.  At the random start the true loss (the fusion moved by an eps/4 average perturbation after 4x4
        # pooling) is BELOW the bf16 noise floor of an image difference, (0.5 % of the range)^2: the measured loss is 4-7x the true
        # one, so the first loss is bounded absolutely, and the first gradient is the true one plus the back-projection of that
        # rounding noise -- a draw that depends on the pair AND on the rounding pattern of every kernel.  Measured over four pairs
        # (seeds 61-64, SFK_TEST_SEED) with the first conv on the CUDA cores / on the tensor cores, two kernels of equal accuracy
        # (tools/diag_c3b.py, diag_c3c.py: both within 0.4 % of one bf16 ulp of the fp64 conv): gradient cosine 0.94/0.88, 0.80/0.78,
        # 0.80/0.83, 0.87/0.81; sign agreement 0.79-0.90; pixels within 1e-3 after two steps 0.54-0.65; outcome MSE -35 %..+12 %
        # of the oracle's (profiles/logs_r2/s4e_seeds.log).  The bounds below hold that whole range; what bf16 storage is HELD to is the
        # success criterion of test_identical_attack_success_outcomes_on_16_fixed_seed_pairs -- north_star's 1e-3 is the fp32 mode's.
        assert ref_err < 0.015 * scale and loss_abs < (0.01 * scale) ** 2 and cos > 0.72 and agree > 0.75 and within > 0.5
        assert abs(d_got - d_ref) <= 0.4 * d_ref


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_512_patch_three_steps_vs_oracle(mode):
    """BASELINE configs[3] at its real size: 161x161 centred patch on StyleGAN2-512, raw-gradient steps, one pair."""
    from oracle.pipeline import AttackCfg as OCfg, LossCfg as OLoss, OraclePipeline, run_attack as oracle_run
    from sfattack.attack_loop import AttackCfg, run_attack
    from sfattack.engine import AttackEngine, LossCfg
    S = 512
    spec, GP, es, EP, vsd, FP = _models(S)
    xa, xb, g = _pairs(1, S, 71)
    side = int(math.sqrt(0.1) * S)
    o = (S - side) // 2
    mask = torch.zeros(1, 3, S, S)
    mask[:, :, o:o + side, o:o + side] = 1.0
    patch0 = torch.rand(1, 3, S, S, generator=g)
    pipe = OraclePipeline(spec, _to(GP, DEV), es, _to(EP, DEV), _to(vsd, DEV), None)
    # raw-gradient steps (adversarial_patch.py:133) sized so that a step moves the patch by <= 0.05: with a much larger step the
    # patch saturates at the clean-range clamp after one iteration and the comparison turns into a comparison of sign(g) ties
    probe = []
    oracle_run(pipe, xa.to(DEV), xb.to(DEV), OCfg(kind="patch", steps=1, lr=0.0, loss=OLoss(1.0, 1.0)), mask=mask.to(DEV),
               patch0=patch0.expand(2, -1, -1, -1).contiguous().to(DEV), record=probe)
    lr = 0.05 / probe[0]["grad"].abs().max().item()
    cfg = dict(kind="patch", steps=3, lr=lr)
    want = oracle_run(pipe, xa.to(DEV), xb.to(DEV), OCfg(loss=OLoss(1.0, 1.0), **cfg), mask=mask.to(DEV),
                      patch0=patch0.expand(2, -1, -1, -1).contiguous().to(DEV))
    torch.cuda.empty_cache()
    _set_mode(mode)
    try:
        eng = AttackEngine(spec, GP, es, EP, vsd, None, batch=1, device=DEV, loss=LossCfg(1.0, 1.0))
        got = run_attack(eng, xa.to(DEV), xb.to(DEV), AttackCfg(**cfg), mask=mask, patch0=patch0)
    finally:
        _set_mode("bf16")
    m = mask.to(DEV).expand(2, -1, -1, -1) == 1
    dp, dp_o = (got["patch"] - patch0.to(DEV))[m], (want["patch"] - patch0.to(DEV))[m]
    assert dp_o.abs().max() > 1e-5, "degenerate: the oracle's patch did not move"
    l_rel = _rel(got["losses"], want["losses"])
    print(f"[512 patch {mode}] patch update rel err {_rel(dp, dp_o):.3e}, losses rel {l_rel:.2e}, x_adv rel {_rel(got['x_adv'], want['x_adv']):.2e}")
    assert torch.equal(got["x_adv"][~m], torch.cat([xa, xb]).to(DEV)[~m])
    close = ((got["x_adv"] - want["x_adv"]).abs() < 2e-3).float().mean().item()
    print(f"[512 patch {mode}] cos(patch update) {_cos(dp, dp_o):.5f}, x_adv within 2e-3: {close:.5f}")
    if mode == "fp32":      # measured: update 1.8 %, losses 3e-4, every step's loss and 99.9 % of the pixels within 2e-3
        assert _rel(dp, dp_o) < 3e-2 and l_rel < 2e-3 and close > 0.995
    else:                   # bf16 storage, measured: update 44 % (gradient cosine 0.96-0.98 per step, summed over 3 steps), losses 2 %
        assert _rel(dp, dp_o) < 0.6 and _cos(dp, dp_o) > 0.88 and l_rel < 0.06 and close > 0.9


def _small(size=64, seed=0):
    from sfattack.params import (EncSpec, gen_spec, make_encoder_params, make_fusion_params, make_generator_params,
                                 make_vgg_state_dict)
    ch = {4: 64, 8: 64, 16: 32, 32: 32, 64: 16}
    spec = gen_spec(size, style_dim=64, n_mlp=2, channels=ch)
    es = EncSpec(n_latent=spec.n_latent, style_dim=64, widths=(16, 32, 64), in_res=size)
    return (spec, make_generator_params(spec, seed), es, make_encoder_params(es, seed + 1), make_vgg_state_dict(seed + 2, width_div=4),
            make_fusion_params(spec.s_dim, seed + 3))


def test_identical_attack_success_outcomes_on_16_fixed_seed_pairs():
    from oracle.pipeline import AttackCfg as OCfg, LossCfg as OLoss, OraclePipeline, run_attack as oracle_run
    from sfattack.attack_loop import AttackCfg, run_attack
    from sfattack.engine import AttackEngine, LossCfg
    S, B, steps = 64, 4, 10
    spec, GP, es, EP, vsd, FP = _small(S)
    budgets = [1 / 255, 2 / 255, 4 / 255, 8 / 255]
    pipe = OraclePipeline(spec, GP, es, EP, vsd, FP, fusion="spatial", vgg_res=S)
    cases, mse_o = [], []
    for bi, eps in enumerate(budgets):              # 4 budgets x 4 pairs = 16 pairs
        g = torch.Generator().manual_seed(1234 + bi)
        xa = F.avg_pool2d(torch.rand(B, 3, S + 4, S + 4, generator=g), 5, 1)
        xb = F.avg_pool2d(torch.rand(B, 3, S + 4, S + 4, generator=g), 5, 1)
        noise = torch.rand(2, B, 3, S, S, generator=g) * 2 - 1
        cases.append((eps, xa, xb, noise))
        out = oracle_run(pipe, xa, xb, OCfg(kind="linf", steps=steps, eps=eps, alpha=eps / 4, loss=OLoss(1.0, 1.0)), start_noise=noise)
        mse_o.append(((out["fused_adv"] - out["fused_ref"]) ** 2).flatten(1).mean(1))
    mse_o = torch.cat(mse_o)
    srt = mse_o.sort().values
    # tau: middle of the widest gap (in log space) among the central half of the oracle's sorted outcomes
    lo, hi = len(srt) // 4, 3 * len(srt) // 4
    gaps = (srt[lo + 1:hi + 1].log() - srt[lo:hi].log())
    j = int(gaps.argmax()) + lo
    tau = float((srt[j] * srt[j + 1]).sqrt())
    succ_o = mse_o > tau
    assert 3 <= int(succ_o.sum()) <= 13, "criterion must discriminate"
    margin = float(min((srt[j + 1] / tau), (tau / srt[j])))
    print(f"[success] tau = {tau:.4e} (gap factor {margin:.3f} on either side); oracle outcomes {succ_o.int().tolist()}")
    for mode in ("fp32", "bf16"):
        _set_mode(mode)
        try:
            eng = AttackEngine(spec, GP, es, EP, vsd, FP, fusion="spatial", batch=B, device=DEV, loss=LossCfg(1.0, 1.0), vgg_res=S,
                               vgg_width_div=4)
            mse_g = []
            for eps, xa, xb, noise in cases:
                out = run_attack(eng, xa.to(DEV), xb.to(DEV), AttackCfg(kind="linf", steps=steps, eps=eps, alpha=eps / 4), start_noise=noise)
                mse_g.append(((out["fused_adv"] - out["fused_ref"]) ** 2).flatten(1).mean(1).cpu())
        finally:
            _set_mode("bf16")
        mse_g = torch.cat(mse_g)
        rel = ((mse_g - mse_o).abs() / mse_o).max().item()
        print(f"[success {mode}] outcome MSE max rel dev {rel:.3e}; outcomes {(mse_g > tau).int().tolist()}")
        assert torch.equal(mse_g > tau, succ_o), f"{mode}: success bits differ from the oracle's"
        assert rel < (0.02 if mode == "fp32" else 0.4)      # the criterion is the 16 bits above; measured 0.7 % / 30 %
