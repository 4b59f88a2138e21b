"""N-way fusion ON THE GRADIENT PATH (VERDICT r1 missing #4): the reference fuses 5 / 4 / 3 inputs for ffhq / car /
church (code/attack/attack_main2.py:521-581, style_fusion_simple.py:82-108,163-165).  AttackEngine(fusion="hierarchy", n_inputs=N)
encodes the N inputs, blends their StyleSpace vectors through the hierarchy's gate chain, and back-propagates the fused-output loss
to all N inputs; fusion="arithmetic" with n_inputs=N is interpolation()'s mean of the W+ codes (interpolation.py:658-669).
Each against the oracle (oracle/pipeline.py with the blend of oracle/fusion_ref.py) on shared seeds."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _cos(a, b):
    a, b = a.flatten().double().cpu(), b.flatten().double().cpu()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _small(n_inputs, B, size=32, seed=0):
    from sfattack.params import EncSpec, gen_spec, make_encoder_params, make_generator_params, make_vgg_state_dict
    ch = {4: 64, 8: 64, 16: 32, 32: 32, 64: 16}
    spec = gen_spec(size, style_dim=64, n_mlp=2, channels=ch)
    GP = make_generator_params(spec, seed=seed)
    es = EncSpec(n_latent=spec.n_latent, style_dim=64, widths=(16, 32, 64), in_res=size)
    EP = make_encoder_params(es, seed=seed + 1)
    vsd = make_vgg_state_dict(seed + 2, width_div=4)
    g = torch.Generator().manual_seed(seed + 4)
    xs = [F.avg_pool2d(torch.rand(B, 3, size + 4, size + 4, generator=g), 5, 1) for _ in range(n_inputs)]
    return spec, GP, es, EP, vsd, xs, g


def _hier(name, s_dim, seed=20):
    from sfattack.params import make_fusion_params
    from sfattack.style_fusion_simple import _PARTS, fusion_hierarchy
    gates = {p: make_fusion_params(s_dim, seed + i) for i, p in enumerate(_PARTS[name]) if p != "all"}
    return fusion_hierarchy(name + "_encode", _PARTS[name], gates)


@pytest.fixture(params=["bf16", "fp32"])
def mode(request):
    from sfattack import lib
    lib.set_activation_dtype(torch.float32 if request.param == "fp32" else torch.bfloat16)
    try:
        yield request.param
    finally:
        lib.set_activation_dtype(torch.bfloat16)


@pytest.mark.parametrize("name,fusion", [("ffhq", "hierarchy"), ("car", "hierarchy"), ("church", "hierarchy"), ("ffhq", "arithmetic"),
                                         ("church", "arithmetic")])
def test_nway_gradient_and_pgd_vs_oracle(mode, name, fusion):
    from oracle.pipeline import AttackCfg as OCfg, LossCfg as OLoss, OraclePipeline, run_attack as oracle_run
    from sfattack.attack_loop import AttackCfg, run_attack
    from sfattack.engine import AttackEngine, LossCfg
    n_in = {"ffhq": 5, "car": 4, "church": 3}[name]
    B, S = 2, 32
    spec, GP, es, EP, vsd, xs, g = _small(n_in, B, S)
    hier = _hier(name, spec.s_dim) if fusion == "hierarchy" else None
    assert hier is None or hier["n_inputs"] == n_in
    pipe = OraclePipeline(spec, GP, es, EP, vsd, None, fusion=fusion, vgg_res=S)
    pipe.hier = hier
    eng = AttackEngine(spec, GP, es, EP, vsd, None, fusion=fusion, batch=B, device=DEV, loss=LossCfg(1.0, 1.0), vgg_res=S, vgg_width_div=4,
                       n_inputs=n_in, hierarchy=hier)
    noise = torch.rand(n_in, B, 3, S, S, generator=g) * 2 - 1
    X0 = torch.cat(xs)
    Xs = torch.clamp(X0 + (8 / 255) * noise.reshape(X0.shape), 0, 1)
    split = lambda T: [T[k * B:(k + 1) * B] for k in range(n_in)]
    with torch.no_grad():
        ref_img, ref_feats = pipe.reference_of(pipe.fused(*xs))
    L_ref, _, gl = pipe.input_grads_n(split(Xs), ref_img, ref_feats, OLoss(1.0, 1.0))
    g_ref = torch.cat(gl)
    eng.set_inputs(*[x.to(DEV) for x in xs])
    eng.compute_reference()
    eng.x.copy_(Xs.to(DEV))
    loss = eng.forward_backward()[0].clone()      # (the engine owns the buffer: later iterations overwrite it)
    eng.check()
    gfull = eng.full_res_grad()
    # every input receives gradient, and (hierarchy) the inputs play DIFFERENT roles
    per_in = [gfull[k * B:(k + 1) * B].abs().mean().item() for k in range(n_in)]
    assert min(per_in) > 0
    c = _cos(gfull, g_ref)
    cs = [_cos(gfull[k * B:(k + 1) * B], g_ref[k * B:(k + 1) * B]) for k in range(n_in)]
    steps = 3
    out_ref = oracle_run(pipe, xs, None, OCfg(kind="linf", steps=steps, loss=OLoss(1.0, 1.0)), start_noise=noise)
    out = run_attack(eng, [x.to(DEV) for x in xs], None, AttackCfg(kind="linf", steps=steps, graph=True), start_noise=noise)
    x_adv, x_ref = out["x_adv"].cpu(), out_ref["x_adv"]
    assert x_adv.shape == (n_in * B, 3, S, S) and (x_adv - X0).abs().max() <= 8 / 255 + 1e-6 and x_adv.min() >= 0 and x_adv.max() <= 1
    same = ((x_adv - x_ref).abs() < 1e-3).float().mean().item()
    d_ref = ((out_ref["fused_adv"] - out_ref["fused_ref"]) ** 2).flatten(1).mean(1)
    d_gpu = ((out["fused_adv"] - out["fused_ref"]) ** 2).flatten(1).mean(1).cpu()
    print(f"[{name} {fusion} {mode}] ref img rel {_rel(eng.ref_img, ref_img):.2e} loss rel {_rel(loss, L_ref):.2e} grad cos {c:.5f} per input "
          f"{[round(v, 4) for v in cs]} x_adv within 1e-3: {same:.4f} outcome {d_gpu.tolist()} vs {d_ref.tolist()}")
    if mode == "fp32":
        assert _rel(eng.ref_img, ref_img) < 1e-4 and _rel(loss, L_ref) < 2e-3 and c > 0.9995 and min(cs) > 0.999 and same > 0.97
        assert torch.allclose(d_gpu, d_ref, rtol=0.03)
    else:
        # (measured on the ffhq hierarchy case: loss rel 5e-3, gradient cosine 0.978 overall / 0.949-0.978 per input, 64 % of the
        #  pixels within 1e-3 after three steps; the bounds leave room for the other role assignments)
        assert _rel(eng.ref_img, ref_img) < 2e-2 and _rel(loss, L_ref) < 0.25 and c > 0.95 and min(cs) > 0.88 and same > 0.5
        assert torch.allclose(d_gpu, d_ref, rtol=0.2)


def test_church_fusion_hierarchy_matches_generate_img_roles():
    """StyleFusionSimple.fusion_hierarchy (input indices per part) against the oracle's generate_img (style TENSORS swapped per part,
    code/style_fusion_simple.py:84-104) on the real church geometry (256x256, 3 inputs [bg_top, bg_bottom, body]): the engine's clean
    fusion of three images = oracle fusion() of their three latents."""
    from oracle import fusion_ref, stylegan2 as sg
    from oracle.fusion_ref import OracleFusion
    from oracle.pipeline import get_latents
    from sfattack import lib
    from sfattack.engine import AttackEngine, LossCfg
    from sfattack.params import EncSpec, make_encoder_params, make_vgg_state_dict
    from sfattack.style_fusion_simple import StyleFusionSimple
    drawer = StyleFusionSimple("church", None, None, DEV)
    G = drawer.original_net
    to = lambda d: {k: v.to(DEV) for k, v in d.items()}
    od = OracleFusion("church", sg.OracleGenerator(G.spec, to(G.params)),
                      {p: to(drawer.sf_hierarchy.nodes[p].fusion_net.p) for p in drawer.sf_hierarchy.nodes if p != "all"})
    es = EncSpec(n_latent=G.spec.n_latent)
    EP = make_encoder_params(es, 1)
    vsd = make_vgg_state_dict(2)
    hier = drawer.fusion_hierarchy("church_encode")
    assert hier["n_inputs"] == 3 and hier["source"][0] == 2
    g = torch.Generator().manual_seed(5)
    xs = [F.avg_pool2d(torch.rand(1, 3, 260, 260, generator=g), 5, 1).to(DEV) for _ in range(3)]
    lib.set_activation_dtype(torch.float32)
    try:
        eng = AttackEngine(G.spec, G.params, es, EP, vsd, None, fusion="hierarchy", batch=1, device=DEV, loss=LossCfg(1.0, 1.0),
                           n_inputs=3, hierarchy=hier)
        eng.set_inputs(*xs)
        img = eng.fused_forward().clone()
        eng.check()
    finally:
        lib.set_activation_dtype(torch.bfloat16)
    with torch.no_grad():
        lat = torch.cat([get_latents(to(EP), es, 2 * x - 1) for x in xs], 0)            # (3, 14, 512): [bg_top, bg_bottom, body]
        want, _, _ = fusion_ref.fusion("church_encode", lat, od)
    assert _rel(img, want) < 1e-4, _rel(img, want)
    # a permutation of the inputs is another fusion
    eng2_in = [xs[2], xs[0], xs[1]]
    lib.set_activation_dtype(torch.float32)
    try:
        eng.set_inputs(*eng2_in)
        img2 = eng.fused_forward().clone()
    finally:
        lib.set_activation_dtype(torch.bfloat16)
    assert _rel(img2, want) > 2e-3      # (random-init gates mix about half-half: roles differ by ~1e-2, two orders above the match)


def test_ffhq_5way_hierarchy_at_1024_vs_oracle():
    """The reference's ffhq fusion at its real geometry: five 1024x1024 inputs [mouth, background, hair, eyes, global]
    (attack_main2.py:526), StyleGAN2-1024, hierarchy gate chain -- clean fusion, first gradient and one PGD step against the oracle
    (fp32 on the same GPU, TF32 off) in the fp32 parity mode (north_star's 1e-3 tolerance), and the bf16 product path on the same
    inputs held to the bf16 bounds of tests/test_fullsize_gpu.py."""
    from oracle.pipeline import AttackCfg as OCfg, LossCfg as OLoss, OraclePipeline, run_attack as oracle_run
    from sfattack import lib
    from sfattack.attack_loop import AttackCfg, run_attack
    from sfattack.engine import AttackEngine, LossCfg
    from sfattack.params import EncSpec, gen_spec, make_encoder_params, make_generator_params, make_vgg_state_dict
    S, n_in = 1024, 5
    spec = gen_spec(S)
    GP, es = make_generator_params(spec, 0), EncSpec(n_latent=spec.n_latent)
    EP, vsd = make_encoder_params(es, 1), make_vgg_state_dict(2)
    hier = _hier("ffhq", spec.s_dim)
    to = lambda d: {k: v.to(DEV) for k, v in d.items()}
    g = torch.Generator().manual_seed(71)
    xs = [F.avg_pool2d(torch.rand(1, 3, S + 4, S + 4, generator=g), 5, 1) for _ in range(n_in)]
    noise = torch.rand(n_in, 1, 3, S, S, generator=g) * 2 - 1
    eps, alpha = 8 / 255, 2 / 255
    pipe = OraclePipeline(spec, to(GP), es, to(EP), to(vsd), None, fusion="hierarchy")
    pipe.hier = dict(hier, gates={p: to(v) for p, v in hier["gates"].items()})
    rec = []
    want = oracle_run(pipe, [x.to(DEV) for x in xs], None, OCfg(kind="linf", steps=1, eps=eps, alpha=alpha, loss=OLoss(1.0, 1.0)),
                      start_noise=noise.to(DEV), record=rec)
    torch.cuda.empty_cache()
    scale = want["fused_ref"].abs().max().item()
    for mode in ("fp32", "bf16"):
        lib.set_activation_dtype(torch.float32 if mode == "fp32" else torch.bfloat16)
        try:
            eng = AttackEngine(spec, GP, es, EP, vsd, None, fusion="hierarchy", batch=1, device=DEV, loss=LossCfg(1.0, 1.0), n_inputs=n_in,
                               hierarchy=hier)
            rec_g = []
            got = run_attack(eng, [x.to(DEV) for x in xs], None, AttackCfg(kind="linf", steps=1, eps=eps, alpha=alpha), start_noise=noise,
                             record=rec_g)
        finally:
            lib.set_activation_dtype(torch.bfloat16)
        ref_err = (got["fused_ref"] - want["fused_ref"]).abs().max().item()
        g_ref, g_got = rec[0]["grad"], rec_g[0]["grad"]
        c = _cos(g_got, g_ref)
        per_in = [_cos(g_got[k:k + 1], g_ref[k:k + 1]) for k in range(n_in)]
        band = g_ref.abs() > (1e-3 if mode == "fp32" else 5e-2) * g_ref.abs().mean()
        step1 = ((got["x_adv"] - want["x_adv"]).abs() < 1e-3)[band].float().mean().item()
        print(f"[ffhq 5-way 1024 {mode}] range +-{scale:.2f}: clean fusion max-abs err {ref_err:.2e}, grad cos {c:.6f} per input "
              f"{[round(v, 4) for v in per_in]}, x_adv after the step within 1e-3 outside the tie band: {step1:.5f}")
        del eng
        torch.cuda.empty_cache()
        if mode == "fp32":
            assert ref_err < 1e-3 * max(1.0, scale) and c > 0.9995 and min(per_in) > 0.999 and step1 > 0.999
        else:
            assert ref_err < 0.015 * scale and c > 0.72 and step1 > 0.75
