"""Pin the oracle: VGG restatement vs golden vectors produced by the reference's own code/vgg.py
(oracle/gen_golden.py), plus self-consistency checks of the builder-written (unpinned) pieces."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import stylegan2 as sg
from oracle.pipeline import (AttackCfg, LossCfg, OraclePipeline, l2_step, linf_step, patch_apply, run_attack)
from oracle.vgg_ref import vgg_forward
from sfattack.params import (EncSpec, gen_spec, make_encoder_params, make_fusion_params, make_generator_params,
                             make_vgg_state_dict)

GOLD = os.path.join(os.path.dirname(__file__), "golden", "vgg_golden.pt")


def test_vgg_matches_reference_golden():
    g = torch.load(GOLD)
    sd = make_vgg_state_dict(g["vgg_seed"])
    assert len(sd) == 26
    for case in g["cases"]:
        x = case["x"].clone().requires_grad_(True)
        taps = vgg_forward(sd, x)
        assert [tuple(t.shape) for t in taps] == [tuple(t.shape) for t in case["taps"]]
        for t, r in zip(taps, case["taps"]):
            torch.testing.assert_close(t, r, rtol=1e-4, atol=1e-5)
        loss = sum((t ** 2).mean() for t in taps)
        (gx,) = torch.autograd.grad(loss, x)
        torch.testing.assert_close(gx, case["grad"], rtol=1e-3, atol=1e-6)


def test_vgg_tap_shapes_at_256():
    sd = make_vgg_state_dict(0)
    with torch.no_grad():
        taps = vgg_forward(sd, torch.zeros(1, 3, 64, 64))
    assert [tuple(t.shape) for t in taps] == [(1, 64, 64, 64), (1, 64, 64, 64), (1, 128, 16, 16), (1, 512, 8, 8)]


def test_upfirdn2d_sizes_and_dc_gain():
    k = sg.make_kernel_2d()
    x = torch.ones(1, 2, 5, 5)
    assert sg.upfirdn2d(torch.ones(1, 2, 11, 11), k * 4, pad=(1, 1)).shape[-1] == 10
    up = sg.upfirdn2d(x, k * 4, up=2, pad=(2, 1))
    assert up.shape[-1] == 10
    torch.testing.assert_close(up[:, :, 2:-2, 2:-2], torch.ones(1, 2, 6, 6))


def test_generator_spec_tables():
    for size, nl, sdim in [(1024, 18, 9088), (512, 16, None), (256, 14, None)]:
        spec = gen_spec(size)
        assert spec.n_latent == nl          # code/style_fusion_simple.py:31,35,39
        if sdim:
            assert spec.s_dim == sdim and len(spec.layers) == 26
        assert max(l.w_idx for l in spec.layers) == nl - 1


def _tiny():
    ch = {4: 16, 8: 16, 16: 8, 32: 8}
    spec = gen_spec(16, style_dim=32, n_mlp=2, channels=ch)
    GP = make_generator_params(spec, seed=3)
    return spec, GP


def test_generator_call_surface():
    spec, GP = _tiny()
    G = sg.OracleGenerator(spec, GP)
    w = torch.randn(2, spec.n_latent, spec.style_dim)
    img, lat = G([w], input_is_latent=True, randomize_noise=False, return_latents=True)
    assert img.shape == (2, 3, 16, 16) and lat.shape == w.shape
    s = G([w], input_is_latent=True, randomize_noise=False, return_style_vector=True)
    assert len(s) == len(spec.layers)
    img2, feats, _ = G([torch.zeros(1, spec.style_dim)], randomize_noise=False, style_vector=s)
    torch.testing.assert_close(img2, img)
    z = torch.randn(1, spec.style_dim)
    ml = G.mean_latent(64)
    img3, _ = G([z], truncation=0.7, truncation_latent=ml, randomize_noise=False)
    assert img3.shape == (1, 3, 16, 16)
    # style mixing of two latents
    img4, lat4 = G([z, torch.randn(1, spec.style_dim)], inject_index=3, randomize_noise=False, return_latents=True)
    assert lat4.shape == (1, spec.n_latent, spec.style_dim)


def test_generator_gradcheck_fp64():
    spec, GP = _tiny()
    GP = {k: v.double() for k, v in GP.items()}
    w = torch.randn(1, spec.n_latent, spec.style_dim, dtype=torch.double, requires_grad=True)

    def f(w):
        return sg.synthesis_from_styles(GP, spec, sg.styles_from_wplus(GP, spec, w))[:, :, ::5, ::5]

    assert torch.autograd.gradcheck(f, (w,), eps=1e-6, atol=1e-5, nondet_tol=1e-8)


def test_update_rules():
    x = torch.rand(2, 3, 4, 4)
    g = torch.randn(2, 3, 4, 4)
    y = linf_step(x, x, g, 2 / 255, 8 / 255)
    assert (y - x).abs().max() <= 2 / 255 + 1e-7 and y.min() >= 0 and y.max() <= 1
    y = x.clone()
    for _ in range(10):
        y = linf_step(y, x, g, 2 / 255, 8 / 255)
    assert (y - x).abs().max() <= 8 / 255 + 1e-7
    z = l2_step(x, x, g, 0.5, 0.25)
    assert ((z - x).flatten(1).norm(dim=1) <= 0.25 + 1e-5).all()
    m = torch.zeros(1, 3, 4, 4); m[..., 1:3, 1:3] = 1
    p = torch.full((1, 3, 4, 4), 5.0)
    a = patch_apply(x, m, p)
    assert torch.equal(a[..., 0, :], x[..., 0, :])
    assert (a.flatten(1).max(1)[0] <= x.flatten(1).max(1)[0] + 1e-7).all()


def test_fgsm_on_tiny_pipeline_moves_fusion():
    spec, GP = _tiny()
    es = EncSpec(n_latent=spec.n_latent, style_dim=32, widths=(8, 8), in_res=16)
    EP = make_encoder_params(es)
    vsd = make_vgg_state_dict(0, width_div=8)
    for fusion in ("arithmetic", "spatial"):
        pipe = OraclePipeline(spec, GP, es, EP, vsd, make_fusion_params(spec.s_dim), fusion=fusion, vgg_res=16)
        g = torch.Generator().manual_seed(5)
        xa, xb = torch.rand(2, 3, 16, 16, generator=g), torch.rand(2, 3, 16, 16, generator=g)
        noise = torch.rand(2, 2, 3, 16, 16, generator=g) * 2 - 1
        out = run_attack(pipe, xa, xb, AttackCfg(kind="linf", steps=3, loss=LossCfg(1.0, 1.0)), start_noise=noise)
        assert out["x_adv"].shape == (4, 3, 16, 16)
        assert (out["x_adv"] - torch.cat([xa, xb])).abs().max() <= 8 / 255 + 1e-6
        assert out["losses"].shape == (3, 2)
        assert (out["losses"][-1] >= out["losses"][0]).all()      # untargeted ascent
