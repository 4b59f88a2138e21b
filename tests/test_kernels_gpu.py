"""GPU parity tests of every kernel behind the C ABI against plain PyTorch fp32 references of the same op
(inputs are bf16-representable so the only differences are fp32 accumulation order and the final bf16
rounding: tolerance 2^-7 relative for bf16 outputs, 1e-4 for fp32 outputs)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF16_RTOL = 2.0 ** -7


def _dev():
    return torch.device("cuda:0")


def _rb(*shape, g, scale=1.0):
    """random bf16-representable fp32 tensor on the GPU"""
    return (torch.randn(*shape, generator=g, device=_dev()) * scale).bfloat16().float()


def _gen(seed):
    return torch.Generator(device=_dev()).manual_seed(seed)


def _nhwc(x):  # NCHW fp32 -> NHWC bf16
    return x.permute(0, 2, 3, 1).contiguous().bfloat16()


def _nchw(x):  # NHWC bf16 -> NCHW fp32
    return x.float().permute(0, 3, 1, 2).contiguous()


def _close(a, b, rtol=BF16_RTOL, what=""):
    a, b = a.float(), b.float()
    scale = b.abs().max().item() + 1e-12
    err = (a - b).abs().max().item()
    assert err <= rtol * scale + 1e-6, f"{what}: max err {err:.4e} vs scale {scale:.4e} (rel {err / scale:.3e})"


def _conv_weights(w):  # (Cout,Cin,3,3) fp32 -> [9][Cout][Cin] bf16
    return w.permute(2, 3, 0, 1).contiguous().reshape(9, w.shape[0], w.shape[1]).bfloat16()


@pytest.mark.parametrize("n,h,w,cin,cout,per_sample", [
    (2, 32, 32, 64, 64, False),     # SW128, one k-block per tap
    (1, 16, 16, 128, 64, False),    # two channel blocks
    (2, 20, 12, 32, 32, False),     # SW64, ragged tiles
    (1, 8, 8, 16, 16, False),       # SW32, TW=8
    (1, 4, 4, 64, 32, False),       # TW=4, mostly out-of-bounds tile
    (2, 16, 16, 64, 256, True),     # per-sample weights, two N blocks
    (3, 33, 17, 64, 128, False),    # odd sizes
    (2, 128, 128, 32, 32, True),    # many tiles per CTA, resident weights, double-buffered TMEM
    (1, 96, 80, 64, 64, True),      # resident weights, SW128
    (2, 64, 64, 256, 128, False),   # streamed weights, 4 channel blocks
])
def test_igemm_conv3x3_fwd(n, h, w, cin, cout, per_sample):
    from sfattack import lib
    g = _gen(1)
    x = _rb(n, cin, h, w, g=g)
    S = n if per_sample else 1
    wt = _rb(S, cout, cin, 3, 3, g=g, scale=1.0 / math.sqrt(cin * 9))
    bias = torch.randn(cout, generator=g, device=_dev())
    ref = torch.cat([F.conv2d(x[i:i + 1], wt[i if per_sample else 0], bias, padding=1) for i in range(n)]).relu()
    xb = _nhwc(x)
    wb = torch.stack([_conv_weights(wt[s]) for s in range(S)]).contiguous()
    err = torch.zeros(1, dtype=torch.int32, device=_dev())
    for use_ref in ("ref", "v1", "v2"):
        out = torch.full((n, h, w, cout), float("nan"), device=_dev(), dtype=torch.bfloat16)
        d = lib.make_igemm_desc(xb, n, h, w, cin, 1, wb, S, 9 * cout, out, h, w, cout, 1, lib.pick_block_n(cout),
                                lib.conv3x3_taps(cout), flags=lib.EP_BIAS | lib.EP_RELU, bias=bias, err=err)
        lib.igemm(d, ref=use_ref == "ref", v1=use_ref == "v1")
        torch.cuda.synchronize()
        assert err.item() == 0, "kernel reported an internal timeout"
        _close(_nchw(out), ref, what=f"igemm fwd ref={use_ref}")


def test_igemm_dgrad_flags():
    """XMASK + GSDOT + COLSCALE + ACCUM epilogue against torch."""
    from sfattack import lib
    g = _gen(2)
    n, h, w, cin, cout = 2, 24, 24, 64, 128   # original conv: cin -> cout ; dgrad: K = cout, N = cin
    gz = _rb(n, cout, h, w, g=g)
    wt = _rb(cout, cin, 3, 3, g=g, scale=1.0 / math.sqrt(cin * 9))
    xin = _rb(n, cin, h, w, g=g)
    s = torch.randn(n, cin, generator=g, device=_dev()) + 1
    prev = _rb(n, cin, h, w, g=g)
    gxt = F.conv_transpose2d(gz, wt, padding=1)
    ref_gs = (xin * gxt).sum((2, 3))
    ref_out = prev + s[:, :, None, None] * gxt * (xin > 0)
    wT = wt.permute(2, 3, 1, 0).contiguous().reshape(9 * cin, cout).bfloat16()   # [tap][cin][cout]
    for use_ref in ("ref", "v1", "v2"):
        out = _nhwc(prev).clone()
        gs = torch.zeros(n, cin, device=_dev())
        err = torch.zeros(1, dtype=torch.int32, device=_dev())
        d = lib.make_igemm_desc(_nhwc(gz), n, h, w, cout, 1, wT, 1, 9 * cin, out, h, w, cin, 1, lib.pick_block_n(cin),
                                lib.conv3x3_dgrad_taps(cin), flags=lib.EP_XMASK | lib.EP_GSDOT | lib.EP_COLSCALE | lib.EP_ACCUM,
                                xin=_nhwc(xin), colscale=s, gs=gs, err=err)
        lib.igemm(d, ref=use_ref == "ref", v1=use_ref == "v1")
        torch.cuda.synchronize()
        assert err.item() == 0
        _close(_nchw(out), ref_out, what=f"dgrad out ref={use_ref}")
        _close(gs, ref_gs, rtol=2e-3, what=f"dgrad gs ref={use_ref}")


@pytest.mark.parametrize("n,h,cin,cout", [
    (2, 8, 64, 32), (1, 4, 32, 64), (1, 16, 128, 128),   # streaming blur (one strip / two strips)
    (2, 64, 32, 64),                                      # streaming blur: 4(+1) strips x 8 row blocks
    (1, 16, 32, 256),                                     # streaming blur, 512 consumer threads
    (1, 12, 64, 128), (1, 8, 32, 192),                    # outside its domain (ragged strip / channel count): the tiled kernel
])
def test_tconv_blur_act_fwd_bwd(n, h, cin, cout):
    """stride-2 transposed conv (4 phase accumulators) + blur/noise/bias/lrelu, forward and backward,
    against autograd through conv_transpose2d + upfirdn2d; the backward both with a finished incoming gradient and with the
    consumer-side finishing (s_in, gs_in)."""
    from oracle import stylegan2 as sg
    from sfattack import lib
    g = _gen(3)
    x = _rb(n, cin, h, h, g=g).requires_grad_(True)
    wt = _rb(n, cout, cin, 3, 3, g=g, scale=1.0 / math.sqrt(cin * 9))      # already-modulated per-sample weights
    dsc = torch.rand(n, cout, generator=g, device=_dev()) + 0.5
    noise = torch.randn(2 * h, 2 * h, generator=g, device=_dev())
    bias = torch.randn(cout, generator=g, device=_dev()) * 0.1
    nw = 0.3
    k2 = (sg.make_kernel_2d() * 4).to(_dev())
    t = F.conv_transpose2d(x.reshape(1, n * cin, h, h), wt.transpose(1, 2).reshape(n * cin, cout, 3, 3), stride=2, groups=n)
    t = t.view(n, cout, 2 * h + 1, 2 * h + 1)
    z = sg.upfirdn2d(t, k2, pad=(1, 1))
    ref = F.leaky_relu(z * dsc[:, :, None, None] + nw * noise + bias.view(1, -1, 1, 1), 0.2) * math.sqrt(2)
    gout = _rb(n, cout, 2 * h, 2 * h, g=g)

    xb = _nhwc(x.detach())
    wb = torch.stack([_conv_weights(wt[s]) for s in range(n)]).contiguous()
    err = torch.zeros(1, dtype=torch.int32, device=_dev())
    T = torch.full((n, 4, h + 1, h + 1, cout), float("nan"), device=_dev(), dtype=torch.bfloat16)
    bn = lib.pick_block_n(cout, 4)
    d = lib.make_igemm_desc(xb, n, h, h, cin, 1, wb, n, 9 * cout, T, h + 1, h + 1, cout, 4, bn, lib.tconv_taps(cout), err=err)
    lib.igemm(d)
    out = torch.full((n, 2 * h, 2 * h, cout), float("nan"), device=_dev(), dtype=torch.bfloat16)
    lib.blur_act_fwd(T, out, dsc, noise, nw, bias)
    torch.cuda.synchronize()
    assert err.item() == 0
    # phase planes vs the interleaved transposed conv
    Tf = T.float()
    full = torch.zeros(n, 2 * h + 2, 2 * h + 2, cout, device=_dev())
    for a in (0, 1):
        for b in (0, 1):
            full[:, a::2, b::2] = Tf[:, a * 2 + b]
    _close(full[:, :2 * h + 1, :2 * h + 1].permute(0, 3, 1, 2), t.detach(), what="tconv phases")
    _close(_nchw(out), ref.detach(), rtol=2 ** -6, what="blur_act_fwd")
    # the slots of the phase planes outside the (2h+1)^2 grid must not reach the output: poison them and run again
    Tp = T.clone()
    Tp[:, 1, :, h] = float("nan"); Tp[:, 3, :, h] = float("nan"); Tp[:, 2, h] = float("nan"); Tp[:, 3, h] = float("nan")
    out2 = torch.empty_like(out)
    lib.blur_act_fwd(Tp, out2, dsc, noise, nw, bias)
    assert torch.equal(out2, out), "blur_act_fwd read a slot outside the transposed-conv grid"

    out_ref_b = _nhwc(ref.detach())
    o = out_ref_b.float().permute(0, 3, 1, 2)
    wT = torch.stack([wt[s].permute(2, 3, 1, 0).contiguous().reshape(9 * cin, cout) for s in range(n)]).bfloat16().contiguous()
    s_in = (torch.rand(n, 3 * cout, generator=g, device=_dev()) + 0.5)     # a wider [n][s_dim] array: this layer's slice starts at cout
    for with_s in (False, True):
        # backward: gT = blur^T(d*act'(out)*g) (phase planar), gdacc, then gx~ = tconv^T(W, gT);  g = gout or s_in * gout
        sv = s_in[:, cout:2 * cout] if with_s else torch.ones(n, cout, device=_dev())
        g_eff = gout * sv[:, :, None, None]
        (gx_ref,) = torch.autograd.grad((ref * g_eff).sum(), x, retain_graph=True)
        gT = torch.full_like(T, float("nan"))
        gdacc = torch.zeros(n, cout, device=_dev())
        gs = torch.zeros(n, 3 * cout, device=_dev())
        if with_s:
            lib.blur_act_bwd(out_ref_b, _nhwc(gout), gT, dsc, noise, nw, bias, gdacc, s_in=s_in, gs_in=gs, in_off=cout)
        else:
            lib.blur_act_bwd(out_ref_b, _nhwc(gout), gT, dsc, noise, nw, bias, gdacc)
        torch.cuda.synchronize()
        assert torch.isfinite(gT.float()).all(), "blur_act_bwd must write every slot of every phase plane"
        # dgrad with per-sample (already modulated) weights, transposed to [tap][cin][cout]
        gx = torch.empty(n, h, h, cin, device=_dev(), dtype=torch.bfloat16)
        d2 = lib.make_igemm_desc(gT, n, h + 1, h + 1, cout, 4, wT, n, 9 * cin, gx, h, h, cin, 1, lib.pick_block_n(cin),
                                 lib.tconv_dgrad_taps(cin), err=err)
        lib.igemm(d2)
        torch.cuda.synchronize()
        assert err.item() == 0
        _close(_nchw(gx), gx_ref, rtol=2 ** -5, what=f"tconv dgrad (s_in={with_s})")
        # demod reduction: gdacc = sum gy*y, y = d*z
        gy = g_eff * math.sqrt(2) * torch.where(o > 0, 1.0, 0.2)
        y = torch.where(o > 0, o / math.sqrt(2), o / (0.2 * math.sqrt(2))) - nw * noise - bias.view(1, -1, 1, 1)
        _close(gdacc, (gy * y).sum((2, 3)), rtol=2e-3, what=f"gdacc (s_in={with_s})")
        if with_s:
            _close(gs[:, cout:2 * cout], (o * gout).sum((2, 3)), rtol=2e-3, what="gs_in")
            assert float(gs[:, :cout].abs().max()) == 0 and float(gs[:, 2 * cout:].abs().max()) == 0


@pytest.mark.parametrize("n,h,cin,cout", [(2, 16, 64, 32), (1, 24, 64, 16), (2, 32, 32, 32), (1, 8, 128, 64),
                                          (1, 24, 128, 64)])   # last: streamed weights, two M tiles per stage, ragged rows
def test_fused_upsample_conv_fwd_bwd(n, h, cin, cout):
    """blur(tconv(x)) as ONE launch: four 3x3 phase convs with a depth-to-space epilogue (forward) and the 3x3 data gradient
    reading the fine-grid gradient through a space-to-depth view (backward), against autograd through conv_transpose2d + upfirdn2d."""
    from oracle import stylegan2 as sg
    from sfattack import lib
    from sfattack.params import fused_up_base_weights
    g = _gen(7)
    x = _rb(n, cin, h, h, g=g).requires_grad_(True)
    w0 = _rb(cout, cin, 3, 3, g=g, scale=1.0 / math.sqrt(cin * 9))
    sty = torch.rand(n, cin, generator=g, device=_dev()) + 0.5
    dsc = torch.rand(n, cout, generator=g, device=_dev()) + 0.5
    noise = torch.randn(2 * h, 2 * h, generator=g, device=_dev())
    bias = torch.randn(cout, generator=g, device=_dev()) * 0.1
    nw = 0.3
    weff = fused_up_base_weights(w0)                                   # (3,3,4,Cout,Cin)
    wmod = (weff[None] * sty[:, None, None, None, None, :]).bfloat16()  # per-sample modulated, rounded as the kernel sees them
    # reference from the SAME rounded phase weights: out[2m+a][2n+b] = sum W[dy][dx][2a+b] x[m+dy-1][n+dx-1]
    ys = []
    for i in range(n):
        wi = wmod[i].float().reshape(3, 3, 4 * cout, cin).permute(2, 3, 0, 1)
        y = F.conv2d(x[i:i + 1], wi, padding=1).view(1, 2, 2, cout, h, h)          # (1,a,b,co,m,n)
        ys.append(y.permute(0, 3, 4, 1, 5, 2).reshape(1, cout, 2 * h, 2 * h))
    z = torch.cat(ys)
    ref = F.leaky_relu(z * dsc[:, :, None, None] + nw * noise + bias.view(1, -1, 1, 1), 0.2) * math.sqrt(2)
    # and the phase weights really are tconv + blur (fp32 weights, looser tolerance: bf16 rounding of W_eff vs of W)
    k2 = (sg.make_kernel_2d() * 4).to(_dev())
    t = F.conv_transpose2d(x.detach(), (w0[None] * sty[:, None, :, None, None])[0].transpose(0, 1), stride=2)
    z0 = sg.upfirdn2d(t[:1], k2, pad=(1, 1))
    _close(z[:1].detach(), z0, rtol=2 ** -6, what="phase weights vs tconv+blur")

    gz = _rb(n, cout, 2 * h, 2 * h, g=g)
    (gx_ref,) = torch.autograd.grad((z * gz).sum(), x)
    xin = _rb(n, cin, h, h, g=g)
    prev = _rb(n, cin, h, h, g=g)
    ref_gs = (xin * gx_ref).sum((2, 3))
    ref_gx = prev + sty[:, :, None, None] * gx_ref

    xb = _nhwc(x.detach())
    wb = wmod.reshape(n, 9 * 4 * cout, cin).contiguous()
    wT = wmod.reshape(n, 3, 3, 4 * cout, cin).permute(0, 1, 2, 4, 3).reshape(n, 9 * cin, 4 * cout).contiguous()
    for use_ref in ("ref", "v2"):
        err = torch.zeros(1, dtype=torch.int32, device=_dev())
        out = torch.full((n, 2 * h, 2 * h, cout), float("nan"), device=_dev(), dtype=torch.bfloat16)
        d = lib.make_igemm_desc(xb, n, h, h, cin, 1, wb, n, 9 * 4 * cout, out, h, h, 4 * cout, 1, lib.pick_block_n(4 * cout),
                                lib.conv3x3_taps(4 * cout), flags=lib.EP_DSCALE | lib.EP_NOISE | lib.EP_BIAS | lib.EP_LRELU,
                                dscale=dsc, bias=bias, noise=noise, noise_w=nw, err=err, out_d2s=1)
        lib.igemm(d, ref=use_ref == "ref")
        gx = _nhwc(prev).clone()
        gs = torch.zeros(n, cin, device=_dev())
        d2 = lib.make_igemm_desc(_nhwc(gz), n, h, h, 4 * cout, 1, wT, n, 9 * cin, gx, h, h, cin, 1, lib.pick_block_n(cin),
                                 lib.conv3x3_dgrad_taps(cin), flags=lib.EP_GSDOT | lib.EP_COLSCALE | lib.EP_ACCUM,
                                 xin=_nhwc(xin), colscale=sty, gs=gs, err=err, a_s2d=1)
        lib.igemm(d2, ref=use_ref == "ref")
        torch.cuda.synchronize()
        assert err.item() == 0
        _close(_nchw(out), ref.detach(), what=f"fused up fwd {use_ref}")
        _close(_nchw(gx), ref_gx, what=f"fused up dgrad {use_ref}")
        _close(gs, ref_gs, rtol=2e-3, what=f"fused up gs {use_ref}")


@pytest.mark.parametrize("n,h,w,cout", [(2, 20, 24, 64), (1, 9, 7, 32), (3, 5, 18, 8)])
def test_conv_c3_fwd_bwd(n, h, w, cout):
    from sfattack import lib
    g = _gen(4)
    x = torch.randn(n, 3, h, w, generator=g, device=_dev()).requires_grad_(True)
    wt = torch.randn(cout, 3, 3, 3, generator=g, device=_dev()) * 0.2
    b = torch.randn(cout, generator=g, device=_dev()) * 0.1
    ref = F.relu(F.conv2d(x, wt, b, padding=1))
    out = torch.empty(n, h, w, cout, device=_dev(), dtype=torch.bfloat16)
    lib.conv_c3_fwd(x.detach(), wt, b, out, relu=True)
    _close(_nchw(out), ref.detach(), what="conv_c3_fwd")
    gpre = _rb(n, cout, h, w, g=g)
    (gx_ref,) = torch.autograd.grad(F.conv2d(x, wt, b, padding=1), x, gpre)
    gx = torch.empty(n, 3, h, w, device=_dev())
    lib.conv_c3_bwd(_nhwc(gpre), wt, gx)
    _close(gx, gx_ref, rtol=1e-4, what="conv_c3_bwd")


@pytest.mark.parametrize("n,h,w,cout", [(2, 20, 24, 64), (1, 33, 17, 32), (2, 64, 64, 16)])
def test_conv_c3_tensor_core_path(n, h, w, cout):
    """First conv on the tensor cores (sfk_c3_pack -> sfk_igemm with hi/lo-split image and weights -> sfk_c3_unpack) against
    torch fp32: the hi/lo split must keep a 1/255-scale perturbation of the image visible (a plain bf16 copy would not)."""
    from sfattack import lib
    g = _gen(14)
    x = (torch.rand(n, 3, h, w, generator=g, device=_dev()) * 2 - 1).requires_grad_(True)
    wt = torch.randn(cout, 3, 3, 3, generator=g, device=_dev()) * 0.2
    b = torch.randn(cout, generator=g, device=_dev()) * 0.1
    wf, wb = lib.c3_pack_weights(wt)
    err = torch.zeros(1, dtype=torch.int32, device=_dev())

    def fwd(xin, relu):
        xp = torch.empty(n, h, w, 16, device=_dev(), dtype=torch.bfloat16)
        out = torch.full((n, h, w, cout), float("nan"), device=_dev(), dtype=torch.bfloat16)
        lib.c3_pack(xin, xp)
        d = lib.make_igemm_desc(xp, n, h, w, 16, 1, wf, 1, 9 * cout, out, h, w, cout, 1, lib.pick_block_n(cout), lib.conv3x3_taps(cout),
                                flags=lib.EP_BIAS | (lib.EP_RELU if relu else 0), bias=b, err=err)
        lib.igemm(d)
        torch.cuda.synchronize()
        assert err.item() == 0
        return out

    ref = F.conv2d(x, wt, b, padding=1)
    _close(_nchw(fwd(x.detach(), True)), ref.detach().relu(), what="c3 tensor-core fwd")
    # a perturbation of 1/255 per pixel: the packed operand must carry it to fp32 accuracy (checked on the pre-activation, which the
    # bf16 OUTPUT rounds to 2^-9 -- so compare the operand itself)
    delta = (torch.rand(n, 3, h, w, generator=g, device=_dev()) - 0.5) * (2.0 / 255)
    xp = torch.empty(n, h, w, 16, device=_dev(), dtype=torch.bfloat16)
    lib.c3_pack((x.detach() + delta).contiguous(), xp)
    rec = (xp[..., 0:3].float() + xp[..., 3:6].float()).permute(0, 3, 1, 2)
    assert (rec - (x.detach() + delta)).abs().max().item() < 2e-5
    assert torch.equal(xp[..., 0:3], xp[..., 6:9]) and xp[..., 9:].abs().max().item() == 0
    # data gradient
    gpre = _rb(n, cout, h, w, g=g)
    (gx_ref,) = torch.autograd.grad(ref, x, gpre)
    gp = torch.full((n, h, w, 16), float("nan"), device=_dev(), dtype=torch.bfloat16)
    d = lib.make_igemm_desc(_nhwc(gpre), n, h, w, cout, 1, wb, 1, 9 * 16, gp, h, w, 16, 1, 16, lib.conv3x3_dgrad_taps(16), err=err)
    lib.igemm(d)
    gx = torch.empty(n, 3, h, w, device=_dev())
    lib.c3_unpack(gp, gx)
    torch.cuda.synchronize()
    assert err.item() == 0
    _close(gx, gx_ref, rtol=BF16_RTOL, what="c3 tensor-core dgrad")     # the gradient leaves the conv kernel in bf16


@pytest.mark.parametrize("h,w", [(16, 16), (9, 7)])
def test_maxpool_fwd_bwd(h, w):
    from sfattack import lib
    g = _gen(5)
    n, c = 2, 32
    x = _rb(n, c, h, w, g=g).relu().requires_grad_(True)
    ref = F.max_pool2d(x, 2, 2, ceil_mode=True)
    y = torch.empty(n, (h + 1) // 2, (w + 1) // 2, c, device=_dev(), dtype=torch.bfloat16)
    lib.maxpool2_fwd(_nhwc(x.detach()), y)
    _close(_nchw(y), ref.detach(), rtol=0, what="maxpool fwd")
    gy = _rb(*ref.shape, g=g)
    (gx_ref,) = torch.autograd.grad(ref, x, gy)
    tap = _rb(n, c, h, w, g=g)
    gx = torch.empty(n, h, w, c, device=_dev(), dtype=torch.bfloat16)
    lib.maxpool2_bwd(_nhwc(x.detach()), _nhwc(gy), gx, tap_ref=_nhwc(tap), tap_coef=0.5, relu_mask=True)
    want = (gx_ref + 0.5 * (x.detach() - tap)) * (x.detach() > 0)
    _close(_nchw(gx), want, what="maxpool bwd")


def test_pools_gap_linear():
    from sfattack import lib
    g = _gen(6)
    x = torch.rand(2, 3, 32, 32, generator=g, device=_dev())
    y = torch.empty(2, 3, 8, 8, device=_dev())
    lib.avgpool_affine_fwd(x, y, 4, 2.0, -1.0)
    _close(y, 2 * F.avg_pool2d(x, 4, 4) - 1, rtol=1e-5, what="avgpool_affine")
    f = _rb(2, 64, 6, 6, g=g).relu()
    gp = torch.empty(2, 64, device=_dev())
    lib.gap_fwd(_nhwc(f), gp)
    _close(gp, f.mean((2, 3)), rtol=1e-5, what="gap_fwd")
    gy = torch.randn(2, 64, generator=g, device=_dev())
    gx = torch.empty(2, 6, 6, 64, device=_dev(), dtype=torch.bfloat16)
    lib.gap_bwd(_nhwc(f), gy, gx)
    _close(_nchw(gx), (gy[:, :, None, None] / 36.0) * (f > 0), what="gap_bwd")
    W = torch.randn(96, 64, generator=g, device=_dev())
    b = torch.randn(96, generator=g, device=_dev())
    out = torch.empty(2, 96, device=_dev())
    lib.linear_fwd(gp, W, b, out)
    _close(out, F.linear(gp, W, b), rtol=1e-5, what="linear_fwd")
    go = torch.randn(2, 96, generator=g, device=_dev())
    gi = torch.empty(2, 64, device=_dev())
    lib.linear_bwd(go, W, gi)
    _close(gi, go @ W, rtol=1e-5, what="linear_bwd")


def test_losses():
    from sfattack import lib
    g = _gen(7)
    n, c, h = 2, 32, 8
    f, r = _rb(n, c, h, h, g=g).relu(), _rb(n, c, h, h, g=g).relu()
    gbuf = _nhwc(_rb(n, c, h, h, g=g))
    g0 = gbuf.clone()
    loss = torch.zeros(n, device=_dev())
    lib.mse_tap(_nhwc(f), _nhwc(r), gbuf, loss, 0.25, 0.5, accumulate=True, relu_mask=True)
    _close(loss, 0.25 * ((f - r) ** 2).flatten(1).sum(1), rtol=1e-4, what="mse loss")
    _close(_nchw(gbuf), _nchw(g0) + 0.5 * (f - r) * (f > 0), what="mse grad")
    img, ref = torch.randn(n, 3, 16, 16, generator=g, device=_dev()), torch.randn(n, 3, 16, 16, generator=g, device=_dev())
    gpool = torch.randn(n, 3, 8, 8, generator=g, device=_dev())
    gi = torch.empty_like(img)
    loss.zero_()
    lib.image_loss_grad(img, ref, gpool, gi, loss, 2.0, 0.1, 2)
    _close(gi, 0.1 * (img - ref) + F.interpolate(gpool, scale_factor=2, mode="nearest") / 4, rtol=1e-5, what="image grad")
    _close(loss, 2.0 * ((img - ref) ** 2).flatten(1).sum(1), rtol=1e-4, what="image loss")


def test_style_space_kernels():
    from sfattack import lib
    g = _gen(8)
    n, L, D = 3, 6, 64
    cins = [32, 32, 16, 48]
    widx = [0, 1, 1, 3]
    SD = sum(cins)
    A = torch.randn(SD, D, generator=g, device=_dev())
    bias = torch.ones(SD, device=_dev())
    row_widx = torch.cat([torch.full((c,), wi, dtype=torch.int32) for c, wi in zip(cins, widx)]).to(_dev())
    w = torch.randn(n, L, D, generator=g, device=_dev(), requires_grad=True)
    scale = 1 / math.sqrt(D)
    ref = torch.cat([F.linear(w[:, wi], A[sum(cins[:i]):sum(cins[:i + 1])] * scale) for i, wi in enumerate(widx)], 1) + bias
    s = torch.empty(n, SD, device=_dev())
    lib.style_affine_fwd(w.detach(), A, bias, row_widx, s, scale)
    _close(s, ref.detach(), rtol=1e-5, what="affine fwd")
    gs = torch.randn(n, SD, generator=g, device=_dev())
    (gw_ref,) = torch.autograd.grad(ref, w, gs)
    starts = torch.tensor([0] + [sum(cins[:i + 1]) for i in range(len(cins))], dtype=torch.int32, device=_dev())
    gw = torch.empty(n, L, D, device=_dev())
    lib.style_affine_bwd(gs, A, starts, torch.tensor(widx, dtype=torch.int32, device=_dev()), gw, scale)
    _close(gw, gw_ref, rtol=1e-5, what="affine bwd")
    # demod
    cin, cout = 32, 48
    Q = torch.rand(cout, cin, generator=g, device=_dev())
    sv = s.detach().clone().requires_grad_(True)
    dref = torch.rsqrt((sv[:, :cin] ** 2) @ Q.t() + 1e-8)
    d = torch.empty(n, cout, device=_dev())
    lib.demod_fwd(s, 0, Q, d)
    _close(d, dref.detach(), rtol=1e-4, what="demod fwd")
    gdacc = torch.randn(n, cout, generator=g, device=_dev())
    # gd = gdacc/d ; gs -= s * sum_j gd_j d_j^3 Q_ji
    (gs_ref,) = torch.autograd.grad(dref, sv, gdacc / dref.detach())
    gs2 = torch.zeros(n, SD, device=_dev())
    lib.demod_bwd(s, 0, Q, d, gdacc, gs2)
    _close(gs2[:, :cin], gs_ref[:, :cin], rtol=1e-4, what="demod bwd")
    # weight modulation
    wbase = torch.randn(9, cout, cin, generator=g, device=_dev())
    wmod = torch.empty(n, 9, cout, cin, device=_dev(), dtype=torch.bfloat16)
    lib.modulate_weights(wbase, s, 0, wmod)
    _close(wmod, wbase[None] * s[:, None, None, :cin], what="modulate")
    lib.modulate_weights(wbase, s, 0, wmod, d)       # demodulation folded in
    _close(wmod, wbase[None] * s[:, None, None, :cin] * d[:, None, :, None], what="modulate*demod")
    if cout % 4 == 0:                                # four phases of cout/4 channels share d (fused upsample conv)
        d4 = d[:, :cout // 4].contiguous()
        lib.modulate_weights(wbase, s, 0, wmod, d4)
        _close(wmod, wbase[None] * s[:, None, None, :cin] * d4.repeat(1, 4)[:, None, :, None], what="modulate*demod (4 phases)")
    # spatial fusion gate
    sa, sb = torch.randn(n, SD, generator=g, device=_dev(), requires_grad=True), torch.randn(n, SD, generator=g, device=_dev(), requires_grad=True)
    al, be, cc = (torch.randn(SD, generator=g, device=_dev()) for _ in range(3))
    q = torch.sigmoid(al * sa + be * sb + cc)
    fref = q * sa + (1 - q) * sb
    fs = torch.empty(n, SD, device=_dev())
    lib.fuse_spatial_fwd(sa.detach(), sb.detach(), al, be, cc, fs)
    _close(fs, fref.detach(), rtol=1e-5, what="fuse fwd")
    ga_ref, gb_ref = torch.autograd.grad(fref, [sa, sb], gs)
    ga, gb = torch.empty_like(fs), torch.empty_like(fs)
    lib.fuse_spatial_bwd(sa.detach(), sb.detach(), al, be, cc, gs, ga, gb)
    _close(ga, ga_ref, rtol=1e-4, what="fuse bwd a")
    _close(gb, gb_ref, rtol=1e-4, what="fuse bwd b")


def test_style_space_batched_matches_per_layer():
    """the one-launch-for-all-layers variants against the per-layer kernels (two layers, one of them 4-phase with folded d)"""
    from sfattack import lib
    g = _gen(21)
    n = 3
    layers = [dict(cin=32, cout=16, rows=9 * 16, fold=0, s_off=0), dict(cin=16, cout=8, rows=9 * 4 * 8, fold=1, s_off=32),
              dict(cin=16, cout=16, rows=9 * 16, fold=2, s_off=48)]
    SD = 64
    s = torch.randn(n, SD, generator=g, device=_dev()) + 1
    tab, qo, do, wo = [], 0, 0, 0
    for L in layers:
        L["Q"] = torch.rand(L["cout"], L["cin"], generator=g, device=_dev())
        L["wb"] = torch.randn(9, L["rows"] // 9, L["cin"], generator=g, device=_dev())
        L["gd"] = torch.randn(n, L["cout"], generator=g, device=_dev())
        tab.append([L["s_off"], L["cin"], L["cout"], qo, do, L["rows"], wo, wo, L["cout"], L["fold"]])
        L["qo"], L["do"], L["wo"] = qo, do, wo
        qo += L["cout"] * L["cin"]; do += n * L["cout"]; wo += L["rows"] * L["cin"]
    q_cat = torch.cat([L["Q"].reshape(-1) for L in layers])
    wb_cat = torch.cat([L["wb"].reshape(-1) for L in layers])
    gd_cat = torch.cat([L["gd"].reshape(-1) for L in layers])
    d_cat = torch.empty(do, device=_dev())
    wm_cat = torch.empty(n * wo, device=_dev(), dtype=torch.bfloat16)
    tabt = torch.tensor(tab, dtype=torch.int64, device=_dev())
    lib.demod_fwd_batched(s, q_cat, d_cat, tabt, 16)
    lib.modulate_weights_batched(wb_cat, s, wm_cat, d_cat, tabt)
    gs = torch.zeros(n, SD, device=_dev())
    lib.demod_bwd_batched(s, q_cat, d_cat, gd_cat, gs, tabt, 32)
    gs_ref = torch.zeros(n, SD, device=_dev())
    for L in layers:
        d = torch.empty(n, L["cout"], device=_dev())
        lib.demod_fwd(s, L["s_off"], L["Q"], d)
        assert torch.equal(d, d_cat[L["do"]:L["do"] + n * L["cout"]].view(n, L["cout"]))
        wm = torch.empty(n, L["rows"], L["cin"], device=_dev(), dtype=torch.bfloat16)
        lib.modulate_weights(L["wb"], s, L["s_off"], wm, d if L["fold"] else None)
        got = wm_cat[n * L["wo"]:n * (L["wo"] + L["rows"] * L["cin"])].view(n, L["rows"], L["cin"])
        if L["fold"] == 2:      # demodulation AND the activation gain sqrt(2) folded in
            ref = L["wb"].reshape(1, L["rows"], L["cin"]) * s[:, None, L["s_off"]:L["s_off"] + L["cin"]] * \
                d.repeat(1, L["rows"] // L["cout"])[:, :, None] * math.sqrt(2)
            _close(got, ref, what="modulate * demod * sqrt2")
        else:
            assert torch.equal(wm, got)
        lib.demod_bwd(s, L["s_off"], L["Q"], d, L["gd"], gs_ref)
    assert torch.equal(gs, gs_ref)


@pytest.mark.parametrize("n,c,h", [(2, 64, 12), (2, 32, 128), (1, 64, 128), (2, 128, 64), (1, 512, 32)])
def test_act_bwd_and_torgb(n, c, h):
    """(2,64,12) runs the register kernels; the larger shapes the cp.async.bulk ring kernels of csrc/sfk_stream.cu."""
    from oracle import stylegan2 as sg
    from sfattack import lib
    g = _gen(9)
    z = _rb(n, c, h, h, g=g)
    dsc = torch.rand(n, c, generator=g, device=_dev()) + 0.5
    noise = torch.randn(h, h, generator=g, device=_dev())
    bias = torch.randn(c, generator=g, device=_dev()) * 0.1
    nw = 0.2
    out = (F.leaky_relu(z * dsc[:, :, None, None] + nw * noise + bias.view(1, -1, 1, 1), 0.2) * math.sqrt(2)).bfloat16().float()
    gout = _rb(n, c, h, h, g=g)
    gz = torch.empty(n, h, h, c, device=_dev(), dtype=torch.bfloat16)
    gd = torch.zeros(n, c, device=_dev())
    lib.act_bwd(_nhwc(out), _nhwc(gout), gz, dsc, noise, nw, bias, gd)
    gy = gout * math.sqrt(2) * torch.where(out > 0, 1.0, 0.2)
    y = torch.where(out > 0, out / math.sqrt(2), out / (0.2 * math.sqrt(2))) - nw * noise - bias.view(1, -1, 1, 1)
    _close(_nchw(gz), gy * dsc[:, :, None, None], what="act_bwd gz")
    _close(gd, (gy * y).sum((2, 3)), rtol=2e-3, what="act_bwd gdacc")
    # ... finishing a plain (flags 0) data gradient of the consumer conv: gout <- s_in * gout, gs_in += sum out * gout
    SD, off = 3 * c, c
    s_all = torch.randn(n, SD, generator=g, device=_dev()) + 1
    gs_all = torch.zeros(n, SD, device=_dev())
    gd.zero_()
    lib.act_bwd(_nhwc(out), _nhwc(gout), gz, dsc, noise, nw, bias, gd, s_in=s_all, gs_in=gs_all, in_off=off)
    sin = s_all[:, off:off + c]
    gy_s = gout * sin[:, :, None, None] * math.sqrt(2) * torch.where(out > 0, 1.0, 0.2)
    _close(_nchw(gz), gy_s * dsc[:, :, None, None], what="act_bwd(s_in) gz")
    _close(gd, (gy_s * y).sum((2, 3)), rtol=2e-3, what="act_bwd(s_in) gdacc")
    _close(gs_all[:, off:off + c], (out * gout).sum((2, 3)), rtol=1e-3, what="act_bwd gs_in")
    assert gs_all[:, :off].abs().max() == 0 and gs_all[:, off + c:].abs().max() == 0
    # ToRGB forward / backward with skip upsample
    x = _rb(n, c, h, h, g=g).requires_grad_(True)
    wr = torch.randn(3, c, generator=g, device=_dev()) / math.sqrt(c)
    s = (torch.randn(n, c, generator=g, device=_dev()) + 1).requires_grad_(True)
    rb = torch.randn(3, generator=g, device=_dev()) * 0.1
    skip = torch.randn(n, 3, h // 2, h // 2, generator=g, device=_dev(), requires_grad=True)
    k2 = (sg.make_kernel_2d() * 4).to(_dev())
    ref = torch.einsum("ci,bi,bihw->bchw", wr, s, x) + rb.view(1, 3, 1, 1) + sg.upfirdn2d(skip, k2, up=2, pad=(2, 1))
    rgb = torch.empty(n, 3, h, h, device=_dev())
    lib.torgb_fwd(_nhwc(x.detach()), wr, s.detach(), 0, rb, skip.detach(), rgb)
    _close(rgb, ref.detach(), rtol=1e-4, what="torgb fwd")
    grgb = torch.randn(n, 3, h, h, generator=g, device=_dev())
    gx_ref, gs_ref, gsk_ref = torch.autograd.grad(ref, [x, s, skip], grgb)
    gx = torch.empty(n, h, h, c, device=_dev(), dtype=torch.bfloat16)
    gs = torch.zeros(n, c, device=_dev())
    lib.torgb_bwd(_nhwc(x.detach()), wr, s.detach(), 0, grgb, gx, gs)
    _close(_nchw(gx), gx_ref, what="torgb bwd gx")
    _close(gs, gs_ref, rtol=1e-3, what="torgb bwd gs")
    # fused: ToRGB backward + activation backward of the conv feeding it, with and without an incoming gradient
    gt = torch.einsum("ci,bchw->bihw", wr, grgb)
    for have_gin in (True, False):
        gin = gout if have_gin else torch.zeros_like(gout)
        gtot = gin + s.detach()[:, :, None, None] * gt
        gy2 = gtot * math.sqrt(2) * torch.where(out > 0, 1.0, 0.2)
        buf = _nhwc(gin).clone()
        gd2 = torch.zeros(n, c, device=_dev())
        gs2 = torch.zeros(n, c, device=_dev())
        lib.act_torgb_bwd(_nhwc(out), buf if have_gin else None, buf, dsc, noise, nw, bias, gd2, wr, s.detach(), 0, grgb, gs2)
        _close(_nchw(buf), gy2 * dsc[:, :, None, None], what=f"act_torgb_bwd gz gin={have_gin}")
        _close(gd2, (gy2 * y).sum((2, 3)), rtol=2e-3, what="act_torgb_bwd gdacc")
        _close(gs2, (out * gt).sum((2, 3)), rtol=1e-3, what="act_torgb_bwd gs_rgb")
    # and with the producer's modulation / style gradient finished in the same pass (styles live in one (n, SD) vector)
    s3 = torch.randn(n, SD, generator=g, device=_dev()) + 1
    gs3 = torch.zeros(n, SD, device=_dev())
    s_rgb, s_prod = s3[:, :c], s3[:, 2 * c:]
    gtot = gout * s_prod[:, :, None, None] + s_rgb[:, :, None, None] * gt
    gy3 = gtot * math.sqrt(2) * torch.where(out > 0, 1.0, 0.2)
    buf = _nhwc(gout).clone()
    gd3 = torch.zeros(n, c, device=_dev())
    lib.act_torgb_bwd(_nhwc(out), buf, buf, dsc, noise, nw, bias, gd3, wr, s3, 0, grgb, gs3, in_off=2 * c)
    _close(_nchw(buf), gy3 * dsc[:, :, None, None], what="act_torgb_bwd(s_in) gz")
    _close(gd3, (gy3 * y).sum((2, 3)), rtol=2e-3, what="act_torgb_bwd(s_in) gdacc")
    _close(gs3[:, :c], (out * gt).sum((2, 3)), rtol=1e-3, what="act_torgb_bwd(s_in) gs_rgb")
    _close(gs3[:, 2 * c:], (out * gout).sum((2, 3)), rtol=1e-3, what="act_torgb_bwd gs_in")
    assert gs3[:, c:2 * c].abs().max() == 0
    gsk = torch.empty(n, 3, h // 2, h // 2, device=_dev())
    lib.rgb_down(grgb, gsk)
    _close(gsk, gsk_ref, rtol=1e-4, what="rgb_down")


def test_layout_converters_and_updates():
    from oracle.pipeline import adam_step, l2_step, linf_step, patch_apply, patch_step
    from sfattack import lib
    g = _gen(10)
    x = torch.randn(2, 16, 5, 7, generator=g, device=_dev())
    y = torch.empty(2, 5, 7, 16, device=_dev(), dtype=torch.bfloat16)
    lib.nchw_to_nhwc_bf16(x, y)
    assert torch.equal(y, _nhwc(x))
    z = torch.empty(2, 16, 5, 7, device=_dev())
    lib.nhwc_bf16_to_nchw(y, z)
    assert torch.equal(z, _nchw(y))
    n, S, k = 3, 16, 2
    x0 = torch.rand(n, 3, S, S, generator=g, device=_dev())
    xa = (x0 + 0.02 * torch.randn(n, 3, S, S, generator=g, device=_dev())).clamp(0, 1)
    gp = torch.randn(n, 3, S // k, S // k, generator=g, device=_dev())
    gfull = F.interpolate(gp, scale_factor=k, mode="nearest")
    stats = torch.zeros(n, device=_dev())
    xl = xa.clone()
    lib.attack_update_linf(xl, x0, gp, 2 / 255, 8 / 255, 1.0, 0.0, 1.0, stats, k)
    want = linf_step(xa, x0, gfull, 2 / 255, 8 / 255, 1.0)
    _close(xl, want, rtol=1e-6, what="linf")
    _close(stats, (want - x0).abs().flatten(1).sum(1), rtol=1e-4, what="linf stats")
    # l2 (three phases)
    x2 = xa.clone()
    norms, dn = torch.zeros(n, device=_dev()), torch.zeros(n, device=_dev())
    for ph in range(3):
        lib.attack_update_l2(x2, x0, gp, norms, dn, 0.5, 0.3, 1.0, 0.0, 1.0, ph, k)
    _close(x2, l2_step(xa, x0, gfull, 0.5, 0.3, 1.0), rtol=1e-5, what="l2")
    # patch
    mask = torch.zeros(n, 3, S, S, device=_dev())
    mask[..., 4:10, 4:10] = 1
    patch = torch.rand(n, 3, S, S, generator=g, device=_dev())
    lo, hi = torch.empty(n, device=_dev()), torch.empty(n, device=_dev())
    lib.minmax_per_sample(x0, lo, hi)
    _close(lo, x0.flatten(1).min(1)[0], rtol=0, what="min")
    _close(hi, x0.flatten(1).max(1)[0], rtol=0, what="max")
    p2, xp = patch.clone(), torch.empty_like(x0)
    lib.attack_update_patch(xp, x0, p2, mask, gp, 1.0, -1.0, False, lo, hi, 1.0 / (k * k), None, k)
    pw = patch_step(patch, gfull / (k * k), 1.0, False, -1.0)
    _close(p2, pw, rtol=1e-6, what="patch")
    _close(xp, patch_apply(x0, mask, pw), rtol=1e-6, what="patch apply")
    # adam
    m, v = torch.zeros_like(x0), torch.zeros_like(x0)
    xad = xa.clone()
    lib.attack_update_adam(xad, gp, m, v, 1e-2, 1, 1.0, k)
    wx, wm, wv = adam_step(xa, gfull, torch.zeros_like(x0), torch.zeros_like(x0), 1, 1e-2)
    _close(xad, wx, rtol=1e-5, what="adam")
