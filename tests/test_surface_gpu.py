"""The rest of the drop-in surface, each entry point against an oracle restatement
on shared seeds: StyleFusionSimple (constructor + every method), fusion() / interpolation() with the reference's latent order for
ffhq / car / church, cal_result / cal_rec_loss / cal_SSMI, patch.attack / train / main, white_box(which_adv), main_optimize,
setup_model, the three loss menus of optimize_vgg, and the shared (universal) patch loop.

Large generators (church 256, car 512, ffhq 1024) run the oracle on the GPU in fp32 with TF32 disabled (it is plain torch code);
tolerances are the bf16-storage ones of the product path unless a test switches to the fp32 parity mode."""
import argparse
import json
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _to(d, dev):
    return {k: v.to(dev) for k, v in d.items()}


def _small_net():
    from sfattack.models import PSPNet
    ch = {4: 64, 8: 64, 16: 32, 32: 32}
    return PSPNet(32, DEV, channels=ch, style_dim=64, n_mlp=2, enc_widths=(16, 32, 64), enc_res=32)


def _smooth(g, *shape):
    n, c, h, w = shape
    return F.avg_pool2d(torch.rand(n, c, h + 4, w + 4, generator=g), 5, 1) * 2 - 1


@pytest.fixture
def fp32_mode():
    from sfattack import lib
    lib.set_activation_dtype(torch.float32)
    try:
        yield
    finally:
        lib.set_activation_dtype(torch.bfloat16)


def _oracle_drawer(drawer, dev=DEV):
    """OracleFusion on the SAME weights and gates as a product StyleFusionSimple (oracle arithmetic on `dev`, fp32)."""
    from oracle import stylegan2 as sg
    from oracle.fusion_ref import OracleFusion
    G = drawer.original_net
    O = sg.OracleGenerator(G.spec, _to(G.params, dev))
    gates = {p: _to(drawer.sf_hierarchy.nodes[p].fusion_net.p, dev) for p in drawer.sf_hierarchy.nodes if p != "all"}
    return OracleFusion(drawer.stylegan_type, O, gates)


# ---------------------------------------------------------------------------------------------------------------------------
def test_style_fusion_simple_constructor_and_methods_church(tmp_path):
    """code/style_fusion_simple.py:26-177 on the 256x256 'church' generator: attribute table, fusion-net loading from the JSON
    index, and all 12 methods against the oracle."""
    from oracle.fusion_ref import blend
    from sfattack.params import make_fusion_params
    from sfattack.style_fusion_simple import StyleFusionSimple
    # a fusion-net index as the reference reads it (:73-80): {part: path}; two parts get gates from files, the rest keep their init
    g = torch.Generator().manual_seed(5)
    drawer0 = StyleFusionSimple("church", None, None, DEV)
    s_dim = drawer0.original_net.spec.s_dim
    files = {}
    for i, part in enumerate(["background_top", "bg"]):
        path = tmp_path / f"{part}.pt"
        torch.save(make_fusion_params(s_dim, 100 + i), path)
        files[part] = path.name
    idx = tmp_path / "fusion_nets.json"
    idx.write_text(json.dumps(files))
    drawer = StyleFusionSimple("church", None, str(idx), DEV, GAN=drawer0.original_net)
    assert (drawer.truncation, drawer.stylegan_size, drawer.stylegan_layers) == (0.5, 256, 14)               # :36-39
    assert drawer.original_net is drawer0.original_net and drawer.base_blender is drawer.sf_hierarchy.nodes["all"]
    assert torch.equal(drawer.sf_hierarchy.nodes["bg"].fusion_net.p["alpha"].cpu(), make_fusion_params(s_dim, 101)["alpha"])
    with pytest.raises(KeyError):
        bad = tmp_path / "bad.pt"
        torch.save({"alpha": torch.zeros(s_dim)}, bad)
        drawer.sf_hierarchy.nodes["bg"].load_fusion_net(str(bad), DEV)
    od = _oracle_drawer(drawer)
    assert _rel(drawer.mean_latent, od.mean_latent) < 1e-4                                                    # :60
    # --- latent conversions
    z = drawer.seed_to_z((7, 2))                                                                              # :110-113
    torch.manual_seed(7)
    assert torch.equal(z.cpu(), torch.randn((3, 1, 512), device=DEV)[2].cpu()) and z.shape == (1, 512)
    s = drawer.z_to_s(z)                                                                                      # :115-118
    assert len(s) == 20 and _rel(torch.cat(s, 1), od.z_to_s(z)) < 1e-4
    wp = drawer.z_to_w_plus(z)                                                                                # :120-124
    assert wp.shape == (14, 512)
    w = torch.randn(1, 14, 512, generator=g).to(DEV)
    assert _rel(torch.cat(drawer.w_plus_to_s(w, 1), 1), od.w_plus_to_s(w)) < 1e-5                             # :126-129
    assert _rel(torch.cat(drawer.general_latent_to_s(w[:, 0], "w"), 1), od.general_latent_to_s(w[:, 0], "w")) < 1e-5   # :131-144, (1,512) form
    with pytest.raises(AssertionError):
        drawer.general_latent_to_s(w[:, :3], "w")
    # --- images (bf16 activations through 14 layers: 2 % of the norm)
    img, feats = drawer.w_plus_to_image(w)                                                                    # :155-157
    img_o, feats_o = od.s_to_image(od.w_plus_to_s(w))
    assert img.shape == (1, 3, 256, 256) and len(feats) == len(feats_o) == 7
    assert _rel(img, img_o) < 2e-2 and _rel(feats[-1], feats_o[-1]) < 2e-2
    img_z, _ = drawer.z_to_image(z)                                                                           # :159-161
    assert _rel(img_z, od.s_to_image(od.z_to_s(z))[0]) < 2e-2
    w2 = torch.randn(1, 14, 512, generator=g).to(DEV)
    w3 = torch.randn(1, 14, 512, generator=g).to(DEV)
    got, _ = drawer.generate_img(w, latents_type="w", bg_top=w2, bg_bottom=w3)                                # :82-108
    want, _ = od.generate_img(w, latents_type="w", bg_top=w2, bg_bottom=w3)
    assert _rel(got, want) < 2e-2
    assert _rel(got, img) > 0.05, "the swapped parts must change the image"
    got_b, _ = drawer.generate_img(w, latents_type="w", background=w2)
    assert _rel(got_b, od.generate_img(w, latents_type="w", background=w2)[0]) < 2e-2
    d_w = {"all": w, "background_top": w2, "bg": w3}
    got_d, _ = drawer.w_plus_dict_to_image(d_w)                                                               # :167-171
    want_d, _ = od.s_to_image(blend(["all", "background_top", "bg"], od.gates, {k: od.w_plus_to_s(v) for k, v in d_w.items()}))
    assert _rel(got_d, want_d) < 2e-2
    z2 = torch.randn(1, 512, generator=g).to(DEV)
    got_zd, _ = drawer.z_dict_to_image({"all": z, "background": z2})                                          # :173-177
    assert got_zd.shape == (1, 3, 256, 256) and torch.isfinite(got_zd).all()
    s_list = drawer.z_to_s(z)
    got_s, _ = drawer.s_dict_to_image({"all": s_list})                                                        # :163-165 (list-of-styles form)
    assert _rel(got_s, img_z) < 1e-6


@pytest.mark.parametrize("name,n_lat", [("ffhq", 5), ("car", 4), ("church", 3)])
def test_fusion_and_interpolation_reference_latent_order(name, n_lat):
    """fusion() (attack_main2.py:521-581) and interpolation() (interpolation.py:658-669): N-way inputs in the reference's order
    (base = LAST latent), substring dataset matching, (I_fused, I_all, cat of the singles' inner features)."""
    from oracle import fusion_ref
    from sfattack.attack.attack_main2 import fusion, interpolation
    from sfattack.style_fusion_simple import StyleFusionSimple
    drawer = StyleFusionSimple(name, None, None, DEV)
    od = _oracle_drawer(drawer)
    L, S = drawer.stylegan_layers, drawer.stylegan_size
    g = torch.Generator().manual_seed(3)
    lat = (0.6 * torch.randn(n_lat, L, 512, generator=g)).to(DEV)
    I_f, I_all, feats = fusion(name + "_encode", lat, drawer, feature_idx=2)
    I_f_o, I_all_o, feats_o = fusion_ref.fusion(name + "_encode", lat, od, feature_idx=2)
    assert I_f.shape == (1, 3, S, S) and I_all.shape == (n_lat, 3, S, S) and feats.shape == feats_o.shape and feats.shape[0] == n_lat
    assert _rel(I_f, I_f_o) < 2e-2 and _rel(I_all, I_all_o) < 2e-2 and _rel(feats, feats_o) < 2e-2
    # roles matter: the same latents in another order give another fusion (the oracle above follows the reference's order)
    I_perm, _, _ = fusion(name, lat.flip(0), drawer, feature_idx=2)
    assert _rel(I_perm, I_f) > 0.05
    J_f, J_all, jf = interpolation(drawer, lat)
    J_f_o, J_all_o, jf_o = fusion_ref.interpolation(od, lat)
    assert J_f.shape == (1, 3, S, S) and jf.shape == jf_o.shape
    assert _rel(J_f, J_f_o) < 2e-2 and _rel(J_all, J_all_o) < 2e-2 and _rel(jf, jf_o) < 2e-2
    with pytest.raises(ValueError):
        fusion("bedroom", lat, drawer)


def test_cal_result_metrics_at_full_resolution():
    """cal_rec_loss / cal_SSMI / cal_result (interpolation.py:848-855, 903-919, 1076-1091): VGG runs on the un-pooled images."""
    from oracle import metrics_ref
    from sfattack import metrics
    from sfattack.params import make_vgg_state_dict
    from sfattack.vgg import vgg16
    vsd = make_vgg_state_dict(4)
    vgg = vgg16(vsd, DEV)
    g = torch.Generator().manual_seed(8)
    S = 256
    f0 = _smooth(g, 1, 3, S, S)
    advs = torch.cat([f0 + 0.05 * s * _smooth(g, 1, 3, S, S) for s in (1.0, 3.0, 9.0)])
    mse, vg, ss = metrics.cal_result(f0.to(DEV), advs.to(DEV), vgg=vgg)
    mse_o, vg_o, ss_o = metrics_ref.cal_result(_to(vsd, DEV), f0.to(DEV), advs.to(DEV))
    assert list(mse) == [0, 1, 2] and list(vg) == [0, 1, 2] and list(ss) == [0, 1, 2]
    for i in range(3):
        assert abs(mse[i] - mse_o[i]) <= 1e-5 * mse_o[i]
        # bf16 feature maps: a feature MSE is a difference of two bf16-rounded tensors, accurate once the perturbation is well above
        # the 2^-9 rounding step (5 % of the image range and up here; the fp32 storage mode below has no such floor)
        assert abs(vg[i] - vg_o[i]) <= 5e-2 * vg_o[i], (i, vg[i], vg_o[i])
        assert abs(ss[i] - ss_o[i]) <= 1e-4, (ss[i], ss_o[i])
    assert ss[0] > ss[1] > ss[2] and mse[0] < mse[1] < mse[2]
    from sfattack import lib
    lib.set_activation_dtype(torch.float32)
    try:
        tiny = f0 + 0.002 * _smooth(g, 1, 3, S, S)          # a perturbation far below the bf16 step
        _, vg32, _ = metrics.cal_result(f0.to(DEV), tiny.to(DEV), vgg=vgg)
        _, vg32_o, _ = metrics_ref.cal_result(_to(vsd, DEV), f0.to(DEV), tiny.to(DEV))
        assert abs(vg32[0] - vg32_o[0]) <= 1e-2 * vg32_o[0], (vg32[0], vg32_o[0])
    finally:
        lib.set_activation_dtype(torch.bfloat16)
    r = metrics.cal_rec_loss(f0.expand(3, -1, -1, -1).to(DEV), advs.to(DEV))
    assert r.shape == (3,) and _rel(r, metrics_ref.cal_rec_loss(f0.expand(3, -1, -1, -1), advs)) < 1e-5
    assert abs(metrics.cal_SSMI(f0[0], advs[1]) - ss_o[1]) < 1e-4
    assert abs(metrics.cal_SSMI(f0[0], f0[0]) - 1.0) < 1e-5
    with pytest.raises(ValueError):
        metrics.cal_SSMI(f0[0], advs[1][:, :100])
    # ragged sizes: windows that straddle tile borders of the kernel
    a, b = _smooth(g, 2, 3, 37, 53), _smooth(g, 2, 3, 37, 53)
    assert _rel(metrics.ssim(a.to(DEV), b.to(DEV)), metrics_ref.ssim(a, b)) < 1e-4


# ---------------------------------------------------------------------------------------------------------------------------
def _patch_setup(seed=3, n=1, size=32):
    from sfattack.params import make_vgg_state_dict
    from sfattack.vgg import vgg16
    net = _small_net()
    vsd = make_vgg_state_dict(5, width_div=4)
    vgg = vgg16(vsd, DEV)
    g = torch.Generator().manual_seed(seed)
    img, tgt = _smooth(g, n, 3, size, size), _smooth(g, 1, 3, size, size)
    mask = torch.zeros(1, 3, size, size)
    mask[..., 10:22, 8:20] = 1
    p0 = torch.rand(1, 3, size, size, generator=g) * 1.2 - 0.6
    return net, vsd, vgg, img, tgt, mask, p0


def _oracle_models(net):
    return net.decoder.spec, net.decoder.params, net.encoder.spec, net.encoder.params


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_patch_attack_trajectory_vs_oracle(mode):
    """patch.attack (adversarial_patch.py:94-160) against the autograd restatement: first loss/gradient, the patch after 4 raw-gradient
    steps, the mask/clamp invariants, the reconstruction it returns and the loss lines it writes."""
    from oracle.pipeline import patch_attack_oracle
    from sfattack import lib
    from sfattack.attack.patch import adversarial_patch as patch
    net, vsd, vgg, img, tgt, mask, p0 = _patch_setup(n=2)
    args = argparse.Namespace(max_count=4, save_img=False)
    rec = []
    adv_o, _, p_o, rec_o = patch_attack_oracle(*_oracle_models(net), vsd, img, p0.expand_as(img), mask.expand_as(img), tgt, 4, record=rec)
    if mode == "fp32":
        lib.set_activation_dtype(torch.float32)
    try:
        adv, m, p, rec_img = patch.attack(img.to(DEV), p0.to(DEV), mask.to(DEV), net.decoder, net.encoder, vgg, DEV, args, tgt.to(DEV))
    finally:
        lib.set_activation_dtype(torch.bfloat16)
    assert adv.shape == img.shape and rec_img.shape == img.shape and m.shape == img.shape
    out = (mask.expand_as(img) == 0)
    assert torch.equal(adv.cpu()[out], img[out])                                                     # :137
    assert (adv.cpu().flatten(1).min(1).values >= img.flatten(1).min(1).values - 1e-6).all()          # :138
    assert (adv.cpu().flatten(1).max(1).values <= img.flatten(1).max(1).values + 1e-6).all()
    d, d_o = (p.cpu() - p0)[mask.expand_as(img) == 1], (p_o - p0)[mask.expand_as(img) == 1]
    assert d_o.abs().max() > 1e-4, "degenerate case: the oracle's patch did not move"
    tol = 2e-2 if mode == "fp32" else 0.3
    assert _rel(d, d_o) < tol, (mode, _rel(d, d_o))
    assert _rel(adv, adv_o) < tol and _rel(rec_img, rec_o) < (1e-3 if mode == "fp32" else 3e-2)


def _list_loader(g, n, size):
    return [_smooth(g, 1, 3, size, size) for _ in range(n)]


def _train_oracle(net, vsd, loader, tgt, args, seed):
    """patch.train / patch.main restated (adversarial_patch.py:26-74, 216-236) on the oracle attack; same numpy seed -> same placements"""
    from oracle.pipeline import patch_attack_oracle
    from sfattack.attack.patch import adversarial_patch_util as U
    rng = np.random.RandomState(seed)
    patch, shape = U.init_patch_square(args.image_size, args.patch_size, rng=rng)
    for data in loader:
        p_np, m_np = U.square_transform(patch, tuple(data.shape), shape, args.image_size, rng=rng)
        p_t, m_t = torch.from_numpy(p_np).float(), torch.from_numpy(m_np).float()
        _, m_t, p_t, _ = patch_attack_oracle(*_oracle_models(net), vsd, data, p_t, m_t, tgt, int(args.max_count))
        masked = (m_t * p_t).numpy()
        new = np.zeros(shape)
        for i in range(shape[0]):
            for j in range(shape[1]):
                new[i][j] = U.submatrix(masked[i][j])
        patch = new
    full, m = U.square_transform(patch, (1, 3, args.image_size, args.image_size), shape, args.image_size, rng=rng)
    return torch.from_numpy(full).float(), torch.from_numpy(m).float()


def test_patch_train_and_main_carry_the_patch_across_images(fp32_mode, tmp_path):
    """patch.main -> train -> attack (adversarial_patch.py:163-243, 26-74): the patch is initialised, placed at a random position and
    rotation per image, optimised, cropped back and carried to the next image; final placement, saved file names."""
    from sfattack.attack.patch import adversarial_patch as patch
    net, vsd, vgg, _, tgt, _, _ = _patch_setup()
    g = torch.Generator().manual_seed(9)
    loader = _list_loader(g, 3, 32)
    args = argparse.Namespace(max_count=3, save_img=False, patch_type="square", patch_size=0.1, image_size=32, epochs=1,
                              dataset_name="ffhq", train_size=3)
    want_p, want_m = _train_oracle(net, vsd, loader, tgt, args, seed=77)
    got_p, got_m = patch.main(None, net, vgg, loader, DEV, str(tmp_path), args, tgt.to(DEV), rng=np.random.RandomState(77))
    assert got_p.shape == (1, 3, 32, 32) and got_m.shape == (1, 3, 32, 32)
    assert torch.equal(got_m.cpu(), want_m), "same seed -> same placement"
    d = int(math.sqrt(32 * 32 * 0.1))
    assert got_m.sum().item() == 3 * d * d
    assert _rel(got_p, want_p) < 2e-2, _rel(got_p, want_p)
    assert os.path.exists(tmp_path / "patch" / "ffhq_3_0.100_mask.npz") and os.path.exists(tmp_path / "patch" / "ffhq_3_0.100_patch.npz")
    assert torch.equal(torch.load(tmp_path / "patch" / "ffhq_3_0.100_patch.npz").cpu(), got_p.cpu())
    assert os.path.exists(tmp_path / "output_file_loss.txt")
    # circle patches go through the same loop (:39-40, 203-205)
    args.patch_type = "circle"
    cp, cm = patch.main(None, net, vgg, loader[:1], DEV, None, args, tgt.to(DEV), rng=np.random.RandomState(5))
    r = int(math.sqrt(int(32 * 32 * 0.1) / math.pi))
    assert cp.shape == (1, 3, 32, 32) and 0 < cm[0, 0].sum().item() <= (2 * r + 1) ** 2 and torch.isfinite(cp).all()


def test_universal_patch_shared_over_a_batch(fp32_mode):
    """one patch for all images of a shard (the in-loop all-reduce is skipped at world size 1)."""
    from oracle.pipeline import universal_patch_oracle
    from sfattack.attack.patch import adversarial_patch as patch
    net, vsd, vgg, img, tgt, mask, p0 = _patch_setup(n=3)
    args = argparse.Namespace(max_count=4)
    want_p, want_x, want_l = universal_patch_oracle(*_oracle_models(net), vsd, img, p0, mask, tgt, 4, lr=2.0)
    got_p, got_x, got_l = patch.train_universal(p0, mask, img, net, vgg, DEV, args, tgt, lr=2.0)
    assert got_p.shape == (1, 3, 32, 32) and got_x.shape == img.shape and got_l.shape == (4, 3)
    assert _rel(got_l, want_l) < 1e-3
    assert _rel((got_p.cpu() - p0)[mask == 1], (want_p - p0)[mask == 1]) < 2e-2
    assert _rel(got_x, want_x) < 1e-3
    assert torch.equal((got_p.cpu() - p0)[mask == 0], torch.zeros_like(p0)[mask == 0])


# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("menu", ["attack_main2", "interpolation", "inter_copy"])
def test_optimize_vgg_loss_menus(fp32_mode, menu, tmp_path):
    """the three weightings of `inversion_loss` (attack_main2.py:649, interpolation.py:818, inter_copy.py:658) incl. the VGG term on
    the reconstruction, first-iteration loss/gradient and the 6-step Adam trajectory; the log lines of :657-666."""
    from oracle.pipeline import ReconLossCfg as OCfg, optimize_vgg_oracle
    from sfattack.attack.attack_main2 import LOSS_MENUS, _recon_engine, optimize_vgg
    from sfattack.params import make_vgg_state_dict
    from sfattack.vgg import vgg16
    net = _small_net()
    vsd = make_vgg_state_dict(5, width_div=4)
    vgg = vgg16(vsd, DEV)
    g = torch.Generator().manual_seed(2)
    img, tgt = _smooth(g, 2, 3, 32, 32), _smooth(g, 2, 3, 32, 32)
    cfg = LOSS_MENUS[menu]
    ocfg = OCfg(**vars(cfg))
    args = argparse.Namespace(lr=5e-3, save_img=True)
    rec = []
    want = optimize_vgg_oracle(*_oracle_models(net), vsd, img, tgt, ocfg, 6, args.lr, record=rec)
    got = optimize_vgg(0, net, vgg, img.to(DEV), tgt.to(DEV), str(tmp_path), DEV, "x", args, n_iters=6, loss=cfg)
    assert _rel(got.cpu() - img, want - img) < 5e-2, _rel(got.cpu() - img, want - img)
    eng = _recon_engine(net, vgg, 2, DEV, cfg)
    eng.set_inputs(img.to(DEV), tgt.to(DEV))
    loss, _, _ = eng.forward_backward()
    eng.check()
    assert _rel(loss, rec[0]["loss"]) < 1e-3
    gf = eng.full_res_grad().cpu()
    cos = float((gf.flatten() @ rec[0]["grad"].flatten()) / (gf.norm() * rec[0]["grad"].norm()))
    assert cos > 0.9999, cos
    lines = open(tmp_path / "optimize_output.txt").read().strip().split("\n")
    assert len(lines) == 2 and lines[0].startswith("0th img iter: 5 l_latent_target:") and "l_latent_org:" in lines[0] and "l_img_org:" in lines[0]
    assert lines[1].startswith("1th img iter: 5 ")


def test_white_box_which_adv_main_optimize_and_setup_model(tmp_path):
    """white_box (attack_main2.py:465-498): only args.which_adv rows are attacked, the rest pass through; one target per input in
    the white_box_patch form; main_optimize (:299-404) dispatch incl. patch_white_box; setup_model (utils/model_utils.py:7-18)."""
    from sfattack.attack.attack_main2 import main_optimize, white_box
    from sfattack.models import setup_model
    from sfattack.params import make_vgg_state_dict
    from sfattack.vgg import vgg16
    net = _small_net()
    vgg = vgg16(make_vgg_state_dict(5, width_div=4), DEV)
    g = torch.Generator().manual_seed(4)
    x = _smooth(g, 4, 3, 32, 32).to(DEV)
    tgt = _smooth(g, 1, 3, 32, 32).to(DEV)
    args = argparse.Namespace(lr=5e-3, save_img=False, which_adv=[1, 3])
    out = white_box(x, tgt, None, net, vgg, args, 3)
    assert out.shape == x.shape and torch.equal(out[0], x[0]) and torch.equal(out[2], x[2])
    assert (out[1] - x[1]).abs().max() > 1e-3 and (out[3] - x[3]).abs().max() > 1e-3
    # batch-of-2 result == the two images attacked one at a time (the reference's loop), up to bf16 atomics order
    args1 = argparse.Namespace(lr=5e-3, save_img=False, which_adv=[3])
    one = white_box(x, tgt, None, net, vgg, args1, 3)
    assert ((one[3] - out[3]).abs() < 2e-3).float().mean() > 0.99
    # empty list = everything (:469-470); per-input targets (:479)
    args2 = argparse.Namespace(lr=5e-3, save_img=False, which_adv=[])
    tg4 = _smooth(g, 4, 3, 32, 32).to(DEV)
    allv = white_box(x, tg4, None, net, vgg, args2, 2)
    assert args2.which_adv == [0, 1, 2, 3] and all((allv[i] - x[i]).abs().max() > 1e-3 for i in range(4))
    # main_optimize: list with one tensor per attack
    a3 = argparse.Namespace(lr=5e-3, save_img=False, which_adv=[0], adversarial="white_box_target")
    res = main_optimize(x, None, net, tgt, a3, DEV, {32: 2}, None, None, vgg=vgg)
    assert isinstance(res, list) and len(res) == 1 and res[0].shape == x.shape and torch.equal(res[0][1:], x[1:])
    a4 = argparse.Namespace(max_count=2, save_img=False, patch_type="square", patch_size=0.1, epochs=1, adversarial="patch_white_box",
                            dataset_name="ffhq", train_size=2, patch_centre=True)
    loader = [x[i:i + 1].cpu() for i in range(2)]
    res = main_optimize(x, None, net, tgt, a4, DEV, {32: 2}, loader, str(tmp_path), vgg=vgg)
    assert a4.image_size == 32 and res[0].shape == x.shape                                       # :323
    d = int(math.sqrt(32 * 32 * 0.1))
    moved = (res[0] != x).flatten(1).any(0).view(3, 32, 32)
    o = (32 - d) // 2
    assert not moved[:, :o].any() and not moved[:, o + d:].any() and moved[:, o:o + d, o:o + d].float().mean() > 0.9
    with pytest.raises(NotImplementedError):
        main_optimize(x, None, net, tgt, argparse.Namespace(adversarial="dp_noise"), DEV, {32: 2})
    # setup_model: checkpoint round trip (decoder.* keys + latent_avg), random-init without one
    ck = tmp_path / "e4e.pt"
    sd = {"decoder." + k: v for k, v in net.decoder.state_dict().items()}
    torch.save({"state_dict": sd, "latent_avg": net.latent_avg.cpu() + 1.0}, ck)
    net2, opts = setup_model(str(ck), DEV, size=32, channels={4: 64, 8: 64, 16: 32, 32: 32}, style_dim=64, n_mlp=2, enc_widths=(16, 32, 64),
                             enc_res=32, seed=123)
    assert opts.start_from_latent_avg and net2.decoder.size == 32
    assert torch.equal(net2.latent_avg.cpu(), net.latent_avg.cpu() + 1.0)
    w = torch.randn(1, net.decoder.n_latent, 64, generator=g).to(DEV)
    a, _ = net.decoder([w], input_is_latent=True, randomize_noise=False)
    b, _ = net2.decoder([w], input_is_latent=True, randomize_noise=False)
    assert torch.equal(a, b), "loaded decoder weights must reproduce the source decoder"


def test_engine_cache_follows_weights_and_library_mode():
    """a cached engine must not outlive the storage dtype it was built for, nor the weights it copied."""
    from sfattack import lib
    from sfattack.attack.attack_main2 import _ENGINES, _recon_engine
    from sfattack.engine import ReconLossCfg
    from sfattack.params import make_generator_params, make_vgg_state_dict
    from sfattack.vgg import vgg16
    net = _small_net()
    vgg = vgg16(make_vgg_state_dict(5, width_div=4), DEV)
    cfg = ReconLossCfg()
    e1 = _recon_engine(net, vgg, 1, DEV, cfg)
    assert _recon_engine(net, vgg, 1, DEV, cfg) is e1
    lib.set_activation_dtype(torch.float32)
    try:
        e2 = _recon_engine(net, vgg, 1, DEV, cfg)
        assert e2 is not e1 and e2.x.dtype == torch.float32 and e2.enc.out[0].dtype == torch.float32
    finally:
        lib.set_activation_dtype(torch.bfloat16)
    assert _recon_engine(net, vgg, 1, DEV, cfg) is e1
    with pytest.raises(AssertionError):        # an engine driven under the wrong mode refuses instead of reading bf16 buffers as fp32
        lib.set_activation_dtype(torch.float32)
        try:
            e1.syn.forward()
        finally:
            lib.set_activation_dtype(torch.bfloat16)
    net.decoder.load_state_dict(make_generator_params(net.decoder.spec, seed=9))
    e3 = _recon_engine(net, vgg, 1, DEV, cfg)
    assert e3 is not e1
    assert len(_ENGINES) <= 4


class _ToyE4E(torch.nn.Module):
    """`net.encoder` as a real torch module with the layer types of the reference's e4e encoder (`Encoder4Editing(50,'ir_se')`,
    code/utils/model_utils.py:24, un-vendored): strided convs, eval-mode BatchNorm, PReLU, an SE gate, a residual shortcut."""

    def __init__(self, n_latent, style_dim, seed=0):
        super().__init__()
        torch.manual_seed(seed)
        nn = torch.nn
        self.stem = nn.Sequential(nn.Conv2d(3, 16, 3, 1, 1, bias=False), nn.BatchNorm2d(16), nn.PReLU(16))
        self.res = nn.Sequential(nn.BatchNorm2d(16), nn.Conv2d(16, 32, 3, 1, 1, bias=False), nn.PReLU(32),
                                 nn.Conv2d(32, 32, 3, 2, 1, bias=False), nn.BatchNorm2d(32))
        self.short = nn.Sequential(nn.Conv2d(16, 32, 1, 2, bias=False), nn.BatchNorm2d(32))
        self.se = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(32, 8, 1, bias=False), nn.ReLU(), nn.Conv2d(8, 32, 1, bias=False),
                                nn.Sigmoid())
        self.head = nn.Linear(32 * 16, n_latent * style_dim)
        self.head.weight.data.mul_(4.0)
        self.n_latent, self.style_dim = n_latent, style_dim
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1)
                m.running_var.uniform_(0.5, 1.5)
        self.eval()

    def forward(self, x):
        x = self.stem(x)
        r = self.res(x)
        x = r * self.se(r) + self.short(x)
        return self.head(F.adaptive_avg_pool2d(x, 4).flatten(1)).view(x.shape[0], self.n_latent, self.style_dim)


def test_optimize_vgg_with_a_torch_module_encoder(fp32_mode, tmp_path):
    """The reference's live loop (attack_main2.py:584-671) with `Model.encoder` a real torch module: optimize_vgg
    keeps it on the gradient path through autograd; first loss / gradient and a 5-step Adam trajectory against the oracle holding
    the same module on the CPU."""
    from oracle.pipeline import ReconLossCfg as OCfg, optimize_vgg_oracle
    from sfattack.attack.attack_main2 import LOSS_MENUS, _recon_engine, optimize_vgg
    from sfattack.params import EncSpec, make_vgg_state_dict
    from sfattack.vgg import vgg16
    net = _small_net()
    dec = net.decoder
    enc_cpu = _ToyE4E(dec.spec.n_latent, dec.spec.style_dim, seed=7)
    enc_gpu = _ToyE4E(dec.spec.n_latent, dec.spec.style_dim, seed=7)
    enc_gpu.load_state_dict(enc_cpu.state_dict())
    net.encoder = enc_gpu.to(DEV)
    vsd = make_vgg_state_dict(5, width_div=4)
    vgg = vgg16(vsd, DEV)
    g = torch.Generator().manual_seed(2)
    img, tgt = _smooth(g, 2, 3, 32, 32), _smooth(g, 2, 3, 32, 32)
    cfg = LOSS_MENUS["interpolation"]
    es = EncSpec(n_latent=dec.spec.n_latent, style_dim=dec.spec.style_dim, in_res=32)
    args = argparse.Namespace(lr=5e-3, save_img=False)
    rec = []
    want = optimize_vgg_oracle(dec.spec, dec.params, es, None, vsd, img, tgt, OCfg(**vars(cfg)), 5, args.lr, record=rec,
                               encoder_module=enc_cpu)
    got = optimize_vgg(0, net, vgg, img.to(DEV), tgt.to(DEV), str(tmp_path), DEV, "x", args, n_iters=5, loss=cfg)
    assert _rel(got.cpu() - img, want - img) < 5e-2, _rel(got.cpu() - img, want - img)
    eng = _recon_engine(net, vgg, 2, DEV, cfg)
    assert eng.enc_module is net.encoder and not eng.graph_ok
    eng.set_inputs(img.to(DEV), tgt.to(DEV))
    loss, _, _ = eng.forward_backward()
    eng.check()
    assert _rel(loss, rec[0]["loss"]) < 1e-3
    gf = eng.full_res_grad().cpu()
    cos = float((gf.flatten() @ rec[0]["grad"].flatten()) / (gf.norm() * rec[0]["grad"].norm()))
    assert cos > 0.9999, cos
