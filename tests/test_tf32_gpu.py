"""The tensor-core conv under fp32 storage (the parity mode): `kind::tf32` MMAs on the same TMA/TMEM pipeline as the bf16
product path, as one pass (plain tf32) or three passes over hi/lo-split operands (split tf32, the default), against fp64
references.  Per-kernel tolerances (max-abs error / max-abs of the reference), stated here and in DESIGN.md section 6:

    split tf32 (tf32x3)   2e-5     measured 0.8e-6 .. 2.0e-6: products accurate to ~2^-21, k-blocks spread over partial accumulators
                                   (the tensor core's fp32 accumulation truncates: one accumulator costs 1.2e-8 per element of K)
    plain tf32            3e-3     measured 5.1e-4 .. 8.2e-4: operands truncated to 10 mantissa bits
    CUDA cores            1e-5     measured 4.4e-7 .. 1.1e-6: fp32 FMA loops (the third opinion)

Every launch variant the product path uses is driven through the fp32 kernel here: resident weights + halo loads, streamed
weights with two M tiles per stage, per-sample weights, the 4-accumulator transposed conv and its transpose, the depth-to-space
forward and space-to-depth data gradient of the fused upsample conv, and every epilogue flag.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = {"tf32x3": 2e-5, "tf32": 3e-3, "cuda_cores": 1e-5}
MODES = ["tf32x3", "tf32", "cuda_cores"]


def _dev():
    return torch.device("cuda:0")


def _gen(seed):
    return torch.Generator(device=_dev()).manual_seed(seed)


def _rn(*shape, g, scale=1.0):
    return torch.randn(*shape, generator=g, device=_dev()) * scale


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().float()


def _nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def _check(got, want, tol, what):
    want = want.double()
    scale = want.abs().max().item() + 1e-30
    err = (got.double() - want).abs().max().item()
    print(f"  {what}: rel err {err / scale:.2e} (tol {tol:.0e})")
    assert math.isfinite(err) and err <= tol * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e} (rel {err / scale:.2e} > {tol:.0e})"
    return err / scale


@pytest.fixture(params=MODES)
def f32mode(request):
    from sfattack import lib
    lib.set_activation_dtype(torch.float32)
    lib.set_conv_math(request.param)
    try:
        yield request.param
    finally:
        lib.set_conv_math("auto")
        lib.set_activation_dtype(torch.bfloat16)


def _w9(w):   # (Cout,Cin,3,3) -> [9][Cout][Cin] fp32
    return w.permute(2, 3, 0, 1).contiguous().reshape(9, w.shape[0], w.shape[1]).float()


@pytest.mark.parametrize("n,h,w,cin,cout,per_sample", [
    (2, 40, 36, 32, 32, True),      # SW128 (32 fp32 = 128 B), resident weights + halo loads, ragged tiles
    (1, 16, 16, 16, 16, False),     # SW64
    (1, 8, 8, 8 * 2, 32, False),    # Cin = 16: SW64, TW = 8
    (1, 4, 4, 64, 32, False),       # TW = 4
    (2, 32, 32, 128, 128, False),   # streamed weights, 4 channel blocks, two M tiles per stage
    (2, 16, 16, 64, 256, True),     # per-sample streamed weights, two N blocks
    (3, 33, 17, 64, 64, False),     # odd sizes
    (1, 96, 80, 32, 64, True),      # many tiles per CTA
])
def test_conv3x3_fwd_fp32_storage(f32mode, n, h, w, cin, cout, per_sample):
    from sfattack import lib
    g = _gen(1)
    x = _rn(n, cin, h, w, g=g)
    S = n if per_sample else 1
    wt = _rn(S, cout, cin, 3, 3, g=g, scale=1.0 / math.sqrt(cin * 9))
    bias = _rn(cout, g=g)
    ref = torch.cat([F.conv2d(x[i:i + 1].double(), wt[i if per_sample else 0].double(), bias.double(), padding=1) for i in range(n)]).relu()
    xb = _nhwc(x)
    wb = torch.stack([_w9(wt[s]) for s in range(S)]).contiguous()
    err = torch.zeros(1, dtype=torch.int32, device=_dev())
    out = torch.full((n, h, w, cout), float("nan"), device=_dev())
    d = lib.make_igemm_desc(xb, n, h, w, cin, 1, wb, S, 9 * cout, out, h, w, cout, 1, lib.pick_block_n(cout),
                            lib.conv3x3_taps(cout), flags=lib.EP_BIAS | lib.EP_RELU, bias=bias, err=err)
    lib.igemm(d)
    lib.igemm(d)      # second launch through the cached plan
    torch.cuda.synchronize()
    assert err.item() == 0, "kernel reported an internal timeout"
    _check(_nchw(out), ref, TOL[f32mode], f"conv fwd [{f32mode}]")
    # the one-shot entry point plans and launches in one call and must agree bit for bit
    out2 = torch.full_like(out, float("nan"))
    d2 = lib.make_igemm_desc(xb, n, h, w, cin, 1, wb, S, 9 * cout, out2, h, w, cout, 1, lib.pick_block_n(cout),
                             lib.conv3x3_taps(cout), flags=lib.EP_BIAS | lib.EP_RELU, bias=bias, err=err)
    lib.igemm(d2, oneshot=True)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)


def test_dgrad_all_flags_fp32_storage(f32mode):
    from sfattack import lib
    g = _gen(2)
    n, h, w, cin, cout = 2, 24, 24, 64, 128
    gz = _rn(n, cout, h, w, g=g)
    wt = _rn(cout, cin, 3, 3, g=g, scale=1.0 / math.sqrt(cin * 9))
    xin = _rn(n, cin, h, w, g=g)
    s = _rn(n, cin, g=g) + 1
    prev = _rn(n, cin, h, w, g=g)
    gxt = F.conv_transpose2d(gz.double(), wt.double(), padding=1)
    ref_gs = (xin.double() * gxt).sum((2, 3))
    ref_out = prev.double() + s.double()[:, :, None, None] * gxt * (xin > 0)
    wT = wt.permute(2, 3, 1, 0).contiguous().reshape(9 * cin, cout).float().contiguous()
    out = _nhwc(prev).clone()
    gs = torch.zeros(n, cin, device=_dev())
    err = torch.zeros(1, dtype=torch.int32, device=_dev())
    d = lib.make_igemm_desc(_nhwc(gz), n, h, w, cout, 1, wT, 1, 9 * cin, out, h, w, cin, 1, lib.pick_block_n(cin),
                            lib.conv3x3_dgrad_taps(cin), flags=lib.EP_XMASK | lib.EP_GSDOT | lib.EP_COLSCALE | lib.EP_ACCUM,
                            xin=_nhwc(xin), colscale=s, gs=gs, err=err)
    lib.igemm(d)
    torch.cuda.synchronize()
    assert err.item() == 0
    _check(_nchw(out), ref_out, TOL[f32mode], f"dgrad out [{f32mode}]")
    _check(gs, ref_gs, max(TOL[f32mode], 1e-4), f"dgrad gs [{f32mode}]")     # + fp32 atomics over 576 partial sums


@pytest.mark.parametrize("n,h,cin,cout", [(2, 8, 64, 32), (1, 16, 128, 128)])
def test_tconv_phases_and_transpose_fp32_storage(f32mode, n, h, cin, cout):
    """stride-2 transposed conv as 4 phase accumulators and its transpose (the unfused up-layers)."""
    from sfattack import lib
    g = _gen(3)
    x = _rn(n, cin, h, h, g=g)
    wt = _rn(n, cout, cin, 3, 3, g=g, scale=1.0 / math.sqrt(cin * 9))
    t = torch.cat([F.conv_transpose2d(x[i:i + 1].double(), wt[i].double().transpose(0, 1), stride=2) for i in range(n)])
    wb = torch.stack([_w9(wt[s]) for s in range(n)]).contiguous()
    err = torch.zeros(1, dtype=torch.int32, device=_dev())
    T = torch.full((n, 4, h + 1, h + 1, cout), float("nan"), device=_dev())
    d = lib.make_igemm_desc(_nhwc(x), n, h, h, cin, 1, wb, n, 9 * cout, T, h + 1, h + 1, cout, 4, lib.pick_block_n(cout, 4),
                            lib.tconv_taps(cout), err=err)
    lib.igemm(d)
    torch.cuda.synchronize()
    assert err.item() == 0
    full = torch.zeros(n, 2 * h + 2, 2 * h + 2, cout, device=_dev())
    for a in (0, 1):
        for b in (0, 1):
            full[:, a::2, b::2] = T[:, a * 2 + b]
    _check(full[:, :2 * h + 1, :2 * h + 1].permute(0, 3, 1, 2), t, TOL[f32mode], f"tconv phases [{f32mode}]")
    # transpose: gx~ = tconv^T(W, gT) with gT phase-planar
    gfull = _rn(n, cout, 2 * h + 2, 2 * h + 2, g=g)
    gfull[:, :, 2 * h + 1:] = 0
    gfull[:, :, :, 2 * h + 1:] = 0
    gT = torch.stack([gfull[:, :, a::2, b::2] for a in (0, 1) for b in (0, 1)], 1).permute(0, 1, 3, 4, 2).contiguous()
    ref = torch.cat([F.conv2d(gfull[i:i + 1, :, :2 * h + 1, :2 * h + 1].double(), wt[i].double().transpose(0, 1), stride=2)
                     for i in range(n)])
    wT = torch.stack([wt[s].permute(2, 3, 1, 0).contiguous().reshape(9 * cin, cout) for s in range(n)]).float().contiguous()
    gx = torch.full((n, h, h, cin), float("nan"), device=_dev())
    d2 = lib.make_igemm_desc(gT, n, h + 1, h + 1, cout, 4, wT, n, 9 * cin, gx, h, h, cin, 1, lib.pick_block_n(cin),
                             lib.tconv_dgrad_taps(cin), err=err)
    lib.igemm(d2)
    torch.cuda.synchronize()
    assert err.item() == 0
    _check(_nchw(gx), ref, TOL[f32mode], f"tconv transpose [{f32mode}]")


@pytest.mark.parametrize("n,h,cin,cout", [(2, 16, 64, 32), (1, 24, 128, 64), (1, 8, 64, 16)])
def test_fused_upsample_conv_fp32_storage(f32mode, n, h, cin, cout):
    """depth-to-space forward and space-to-depth data gradient of the fused upsample conv."""
    from sfattack import lib
    from sfattack.params import fused_up_base_weights
    g = _gen(7)
    x = _rn(n, cin, h, h, g=g)
    w0 = _rn(cout, cin, 3, 3, g=g, scale=1.0 / math.sqrt(cin * 9))
    sty = torch.rand(n, cin, generator=g, device=_dev()) + 0.5
    noise = _rn(2 * h, 2 * h, g=g)
    bias = _rn(cout, g=g, scale=0.1)
    nw = 0.3
    weff = fused_up_base_weights(w0)                                   # (3,3,4,Cout,Cin)
    wmod = (weff[None] * sty[:, None, None, None, None, :]).float()
    xd = x.double().requires_grad_(True)
    ys = []
    for i in range(n):
        wi = wmod[i].double().reshape(3, 3, 4 * cout, cin).permute(2, 3, 0, 1)
        y = F.conv2d(xd[i:i + 1], wi, padding=1).view(1, 2, 2, cout, h, h)
        ys.append(y.permute(0, 3, 4, 1, 5, 2).reshape(1, cout, 2 * h, 2 * h))
    z = torch.cat(ys)
    u = z + math.sqrt(2) * (nw * noise.double() + bias.double().view(1, -1, 1, 1))
    ref = torch.maximum(u, 0.2 * u)                                    # the gain-folded epilogue (SFK_EP_LRELU_RAW)
    gz = _rn(n, cout, 2 * h, 2 * h, g=g)
    (gx_ref,) = torch.autograd.grad((z * gz.double()).sum(), xd)
    err = torch.zeros(1, dtype=torch.int32, device=_dev())
    out = torch.full((n, 2 * h, 2 * h, cout), float("nan"), device=_dev())
    wb = wmod.reshape(n, 9 * 4 * cout, cin).contiguous()
    d = lib.make_igemm_desc(_nhwc(x), n, h, h, cin, 1, wb, n, 9 * 4 * cout, out, h, h, 4 * cout, 1, lib.pick_block_n(4 * cout),
                            lib.conv3x3_taps(4 * cout), flags=lib.EP_NOISE | lib.EP_BIAS | lib.EP_LRELU_RAW,
                            bias=bias * math.sqrt(2), noise=noise, noise_w=nw * math.sqrt(2), err=err, out_d2s=1)
    lib.igemm(d)
    wT = wmod.reshape(n, 3, 3, 4 * cout, cin).permute(0, 1, 2, 4, 3).reshape(n, 9 * cin, 4 * cout).contiguous()
    gx = torch.full((n, h, h, cin), float("nan"), device=_dev())
    d2 = lib.make_igemm_desc(_nhwc(gz), n, h, h, 4 * cout, 1, wT, n, 9 * cin, gx, h, h, cin, 1, lib.pick_block_n(cin),
                             lib.conv3x3_dgrad_taps(cin), err=err, a_s2d=1)
    lib.igemm(d2)
    torch.cuda.synchronize()
    assert err.item() == 0
    _check(_nchw(out), ref.detach(), TOL[f32mode], f"fused up fwd [{f32mode}]")
    _check(_nchw(gx), gx_ref, TOL[f32mode], f"fused up dgrad [{f32mode}]")


def test_parity_mode_launches_the_tensor_core_kernel():
    """the parity mode must exercise the tcgen05 pipeline, not the CUDA-core loops.  torch's profiler lists
    the kernels a parity-mode conv launches: igemm_tc2_kernel<..., float, ...> (+ the operand split), never igemm_ref_kernel."""
    from torch.profiler import ProfilerActivity, profile
    from sfattack import lib
    lib.set_activation_dtype(torch.float32)
    try:
        g = _gen(9)
        n, h, cin, cout = 1, 32, 32, 32
        x, wt = _rn(n, h, h, cin, g=g), _rn(1, 9 * cout, cin, g=g)
        out = torch.empty(n, h, h, cout, device=_dev())
        err = torch.zeros(1, dtype=torch.int32, device=_dev())
        seen = {}
        for mode in ("tf32x3", "tf32"):
            lib.set_conv_math(mode)
            d = lib.make_igemm_desc(x, n, h, h, cin, 1, wt, 1, 9 * cout, out, h, h, cout, 1, 32, lib.conv3x3_taps(cout), err=err)
            lib.igemm(d)
            torch.cuda.synchronize()
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                lib.igemm(d)
                torch.cuda.synchronize()
            seen[mode] = [e.key for e in prof.key_averages()]
        for mode, names in seen.items():
            assert any("igemm_tc2_kernel" in k and "float" in k for k in names), (mode, names)
            assert not any("igemm_ref_kernel" in k for k in names), (mode, names)
        assert any("tf32_split_kernel" in k for k in seen["tf32x3"]) and not any("tf32_split_kernel" in k for k in seen["tf32"])
    finally:
        lib.set_conv_math("auto")
        lib.set_activation_dtype(torch.bfloat16)
