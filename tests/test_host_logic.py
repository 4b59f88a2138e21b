"""CPU tests of the host-side logic that feeds the kernels: tap lists (emulated on CPU against torch convs), model tables,
sharding over ranks with a real 2-process gloo group."""
import math
import os

import pytest
import torch
import torch.nn.functional as F


def emulate_igemm(A, Wt, taps, out_hw, num_acc, out_c):
    """CPU model of the sfk_igemm contract: A (N,P,H,W,C), Wt (rows, C): out[n,acc,h,w,:] += A[n,plane,h+dy,w+dx,:] @ Wt[brow:brow+out_c].T"""
    N, P, H, W, C = A.shape
    oh, ow = out_hw
    out = torch.zeros(N, num_acc, oh, ow, out_c, dtype=A.dtype)
    Ap = F.pad(A, [0, 0, 4, 4, 4, 4])
    for dy, dx, plane, acc, brow in taps:
        win = Ap[:, plane, 4 + dy:4 + dy + oh, 4 + dx:4 + dx + ow, :]
        if win.shape[1] < oh or win.shape[2] < ow:
            win = F.pad(win, [0, 0, 0, ow - win.shape[2], 0, oh - win.shape[1]])
        out[:, acc] += torch.einsum("nhwc,oc->nhwo", win, Wt[brow:brow + out_c])
    return out


def test_tap_lists_reproduce_conv_dgrad_tconv_and_its_transpose():
    from sfattack import lib
    g = torch.Generator().manual_seed(0)
    N, H, Cin, Cout = 2, 6, 5, 4
    x = torch.randn(N, Cin, H, H, generator=g, dtype=torch.double)
    w = torch.randn(Cout, Cin, 3, 3, generator=g, dtype=torch.double)
    A = x.permute(0, 2, 3, 1)[:, None]
    Wf = w.permute(2, 3, 0, 1).reshape(9 * Cout, Cin)
    got = emulate_igemm(A, Wf, lib.conv3x3_taps(Cout), (H, H), 1, Cout)[:, 0].permute(0, 3, 1, 2)
    torch.testing.assert_close(got, F.conv2d(x, w, padding=1))
    gz = torch.randn(N, Cout, H, H, generator=g, dtype=torch.double)
    Wb = w.permute(2, 3, 1, 0).reshape(9 * Cin, Cout)
    got = emulate_igemm(gz.permute(0, 2, 3, 1)[:, None], Wb, lib.conv3x3_dgrad_taps(Cin), (H, H), 1, Cin)[:, 0].permute(0, 3, 1, 2)
    torch.testing.assert_close(got, F.conv_transpose2d(gz, w, padding=1))
    # stride-2 transposed conv as 4 phase planes of (H+1)^2
    T = emulate_igemm(A, Wf, lib.tconv_taps(Cout), (H + 1, H + 1), 4, Cout)
    full = torch.zeros(N, 2 * H + 2, 2 * H + 2, Cout, dtype=torch.double)
    for a in (0, 1):
        for b in (0, 1):
            full[:, a::2, b::2] = T[:, 2 * a + b]
    ref = F.conv_transpose2d(x, w.permute(1, 0, 2, 3), stride=2)
    torch.testing.assert_close(full[:, :2 * H + 1, :2 * H + 1].permute(0, 3, 1, 2), ref)
    # ... and its transpose, read back from the phase planes
    gT = torch.randn(N, Cout, 2 * H + 1, 2 * H + 1, generator=g, dtype=torch.double)
    gTp = F.pad(gT, [0, 1, 0, 1]).permute(0, 2, 3, 1)
    planes = torch.stack([gTp[:, a::2, b::2] for a in (0, 1) for b in (0, 1)], 1)
    got = emulate_igemm(planes, Wb, lib.tconv_dgrad_taps(Cin), (H, H), 1, Cin)[:, 0].permute(0, 3, 1, 2)
    torch.testing.assert_close(got, F.conv2d(gT, w.permute(1, 0, 2, 3), stride=2))


def test_model_tables_and_flop_accounting():
    import bench
    from sfattack.params import EncSpec, gen_spec
    spec = gen_spec(1024)
    fl = bench.flops_per_iter_image(spec, EncSpec(n_latent=18))
    assert abs(fl["generator_fwd"] / 1e9 - 148.52) < 0.5
    assert abs(fl["vgg_fwd"] / 1e9 - 31.63) < 0.1
    assert abs(fl["attack"] / 1e9 - 360.3) < 1.0
    assert [gen_spec(s).n_latent for s in (256, 512, 1024)] == [14, 16, 18]
    from sfattack.engine import vgg_layers
    taps = [l.tap for l in vgg_layers() if l.tap >= 0]
    assert taps == [0, 1, 2, 3] and [l.kind for l in vgg_layers()].count("pool") == 3


def test_block_n_and_patch_utils():
    from sfattack import lib
    from sfattack.attack.patch.adversarial_patch_util import init_patch_square, square_transform, submatrix
    assert lib.pick_block_n(512) == 128 and lib.pick_block_n(32) == 32 and lib.pick_block_n(256, 4) == 64 and 2 * 4 * lib.pick_block_n(512, 4) <= 512
    patch, shape = init_patch_square(512, 0.1)
    assert shape == (1, 3, 161, 161)                            # floor(sqrt(0.1)*512) = 161
    canvas, mask = square_transform(patch, (2, 3, 512, 512), shape, 512)
    assert mask.sum() == 2 * 3 * 161 * 161 and canvas.shape == (2, 3, 512, 512)
    assert submatrix(canvas[0, 0]).shape == (161, 161)


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from sfattack.parallel import gather_results, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(7, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float32)[:, None] * torch.ones(1, 3)
    full = gather_results(local, 7)
    out.put((rank, lo, hi, full[:, 0].tolist()))
    dist.destroy_process_group()


def test_sharding_and_final_gather_two_ranks_gloo():
    import torch.multiprocessing as mp
    from sfattack.parallel import shard_range
    assert [shard_range(7, r, 2) for r in range(2)] == [(0, 4), (4, 7)]
    assert [shard_range(64, r, 8) for r in range(8)][-1] == (56, 64)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, 29533, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
    assert res[0][1:3] == (0, 4) and res[1][1:3] == (4, 7)
    assert res[0][3] == res[1][3] == [float(i) for i in range(7)]


@pytest.mark.parametrize("name,n_in", [("ffhq", 5), ("car", 4), ("church", 3)])
def test_fusion_hierarchy_roles_match_generate_img_and_the_engine_chain(name, n_in):
    """Host logic of the N-way gradient path: `fusion_hierarchy` (input INDEX per part) against the oracle's generate_img, which
    swaps style TENSORS per part as the reference does (code/style_fusion_simple.py:84-104, roles of fusion() at
    code/attack/attack_main2.py:526-566); and the engine's static gate chain (skip a base-assigned part only while nothing has
    been gated in) against the oracle's value-based blend.  Tiny CPU generator; styles are passed as latents_type='s'."""
    from oracle import stylegan2 as sg
    from oracle.fusion_ref import PARTS, OracleFusion, blend, gate
    from sfattack.params import gen_spec, make_fusion_params, make_generator_params
    from sfattack.style_fusion_simple import _PARTS, fusion_hierarchy
    assert _PARTS[name] == PARTS[name]
    spec = gen_spec(16, style_dim=32, n_mlp=2, channels={4: 16, 8: 16, 16: 8})
    GP = make_generator_params(spec, seed=0)
    gates = {p: make_fusion_params(spec.s_dim, 30 + i) for i, p in enumerate(PARTS[name]) if p != "all"}
    od = OracleFusion(name, sg.OracleGenerator(spec, GP), gates, mean_latent=torch.zeros(1, 32))
    g = torch.Generator().manual_seed(1)
    s = [torch.randn(1, spec.s_dim, generator=g) for _ in range(n_in)]
    if name == "ffhq":      # [mouth, background, hair, eyes, global]
        want, _ = od.generate_img(s[4], latents_type="s", hair=s[2], eyes=s[3], background=s[1], mouth=s[0])
    elif name == "car":     # [wheel, bg_top, bg_bottom, body]
        want, _ = od.generate_img(s[3], latents_type="s", wheels=s[0], bg_top=s[1], bg_bottom=s[2])
    else:                   # [bg_top, bg_bottom, body]
        want, _ = od.generate_img(s[2], latents_type="s", bg_top=s[0], bg_bottom=s[1])
    h = fusion_hierarchy(name + "_encode", _PARTS[name], gates)
    assert h["n_inputs"] == n_in and h["source"][0] == n_in - 1
    got, _ = od.s_to_image(blend(h["parts"], gates, {p: s[k] for p, k in zip(h["parts"], h["source"])}))
    torch.testing.assert_close(got, want)
    # the engine's chain (AttackEngine.__init__, fusion="hierarchy")
    base, chain = h["source"][0], []
    for p, k in zip(h["parts"][1:], h["source"][1:]):
        if not chain and k == base:
            continue
        chain.append((p, k))
    out = s[base]
    for p, k in chain:
        out = gate(gates[p], out, s[k])
    torch.testing.assert_close(od.s_to_image(out)[0], want)
    assert len({k for _, k in chain} | {base}) == n_in        # every input takes part
    with pytest.raises(ValueError):
        fusion_hierarchy("bedroom", _PARTS[name], gates)


def test_run_artifacts_in_the_reference_formats(tmp_path):
    """parameters.txt lines (attack_main2.py:975-988), the `.npz`-named torch pickles in benign/ and adversarial/ (:1098-1111) and the
    results header / row order (interpolation.py:1256-1258)."""
    import argparse
    from sfattack import artifacts as A
    root = A.new_run_folder(str(tmp_path / "run"))
    assert A.new_run_folder(root) == root
    args = argparse.Namespace(dataset_name="ffhq_encode", epochs=1, max_count=50, patch_size=0.1, train_size=2000, patch_type="square",
                              lr=0.01, use_generate_img=False)
    pf = A.write_parameters(root, "white_box", args, 1024, 100)
    lines = open(pf).read().splitlines()
    assert lines == ["adversarial attack white_box", "dataset ffhq_encode", "dataset size 1024", "epochs 1", "max_count 50",
                     "patch_size 0.1", "train_size 2000", "patch_type square", "white-box max_iter 100", "white-box lr 0.01",
                     "use_generate_img False"]
    A.write_parameters(root, "patch", args, 1024, 100)                         # the reference opens the file with 'a'
    assert len(open(pf).read().splitlines()) == 22
    rec = A.RunRecorder(root)
    x = torch.rand(5, 3, 8, 8)
    rec.add(all_inputs=x, all_adv_inputs=x + 0.01)
    rec.add(all_inputs=x * 0.5, all_adv_inputs=x * 0.5 + 0.01, all_rec_loss=torch.rand(5))
    with pytest.raises(KeyError):
        rec.add(bogus=x)
    paths = rec.save()
    assert sorted(os.path.relpath(p, root) for p in paths) == ["adversarial/all_adv_inputs.npz", "benign/all_inputs.npz",
                                                              "benign/all_rec_loss.npz"]
    back = torch.load(os.path.join(root, "benign", "all_inputs.npz"))
    assert back.shape == (10, 3, 8, 8) and torch.equal(back[:5], x)
    cols = A.result_columns(A.DATASET_N["car"])
    assert len(cols) == 4 + 6 * 5 and cols[:4] == ["noise"] * 4 and cols[4:9] == ["cri_spati"] * 5 and cols[-5:] == ["ssmi_arith"] * 5
    d = {i: float(i) for i in range(5)}
    row = A.result_row([0.1, 0.2, 0.3, 0.4], d, d, d, d, d, d)
    assert len(row) == len(cols) and row[4:9] == [0.0, 1.0, 2.0, 3.0, 4.0]
    with pytest.raises(ValueError):
        A.result_row([0.1], d, d, d, d, d, d)


def test_first_conv_hi_lo_operand_layout_on_cpu():
    """The packed operand of the first conv (sfk_c3_pack) and its weight layouts (lib.c3_pack_weights), emulated on the CPU with the
    tap lists the kernel gets: forward = x.hi*W.hi + x.lo*W.hi + x.hi*W.lo within 2^-16 of the fp64 conv (a plain bf16 copy of x
    would be 2^-9), data gradient = channels [0:3] + [3:6] of the packed gradient."""
    from sfattack import lib
    g = torch.Generator().manual_seed(3)
    N, H, W_, Cout = 2, 7, 9, 16
    x = torch.rand(N, 3, H, W_, generator=g) * 2 - 1
    w = torch.randn(Cout, 3, 3, 3, generator=g) * 0.3
    wf, wb = lib.c3_pack_weights(w)
    assert wf.shape == (9 * Cout, 16) and wb.shape == (9 * 16, Cout) and wf.dtype == torch.bfloat16
    hi = x.bfloat16().float()
    lo = (x - hi).bfloat16().float()
    xp = torch.zeros(N, H, W_, 16)
    xp[..., 0:3], xp[..., 3:6], xp[..., 6:9] = hi.permute(0, 2, 3, 1), lo.permute(0, 2, 3, 1), hi.permute(0, 2, 3, 1)
    got = emulate_igemm(xp.double()[:, None], wf.double(), lib.conv3x3_taps(Cout), (H, W_), 1, Cout)[:, 0].permute(0, 3, 1, 2)
    ref = F.conv2d(x.double(), w.double(), padding=1)
    scale = ref.abs().max().item()
    assert (got - ref).abs().max().item() < 2 ** -15 * scale
    plain = F.conv2d(hi.double(), w.bfloat16().double(), padding=1)           # what a bf16 copy of image and weights would give
    assert (plain - ref).abs().max().item() > 20 * (got - ref).abs().max().item()
    gz = torch.randn(N, Cout, H, W_, generator=g).bfloat16().double()
    gp = emulate_igemm(gz.permute(0, 2, 3, 1)[:, None], wb.double(), lib.conv3x3_dgrad_taps(16), (H, W_), 1, 16)[:, 0]
    gx = (gp[..., 0:3] + gp[..., 3:6]).permute(0, 3, 1, 2)
    gref = F.conv_transpose2d(gz, w.double(), padding=1)
    assert (gx - gref).abs().max().item() < 2 ** -15 * gref.abs().max().item()
    assert gp[..., 6:].abs().max().item() == 0
