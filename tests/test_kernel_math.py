"""The algebra the CUDA kernels implement, written out in torch on CPU and
checked against autograd through the oracle.  Keeps the kernel design honest before any GPU time:
  * shared-weight backward of the modulated conv: gx = s*convT(W, d*gy), gs = sum(x*gx~) - demod term
  * stride-2 transposed conv as 4 phase planes (+ its transpose), blur read from / written to phase planes
  * activation inverse used to recover the pre-activation from the stored output
"""
import math

import torch
import torch.nn.functional as F

from oracle import stylegan2 as sg
from sfattack.params import ModLayer, blur_kernel_1d

SQ2 = math.sqrt(2.0)


def _layer(kind, cin, cout, seed=0):
    g = torch.Generator().manual_seed(seed)
    l = ModLayer("L", kind, 8, cin, cout, 0, noise_idx=0)
    P = {"L.conv.weight": torch.randn(1, cout, cin, 3, 3, generator=g, dtype=torch.double),
         "L.conv.modulation.weight": torch.randn(cin, 16, generator=g, dtype=torch.double),
         "L.conv.modulation.bias": torch.ones(cin, dtype=torch.double),
         "L.noise.weight": torch.tensor([0.3], dtype=torch.double),
         "L.activate.bias": torch.randn(cout, generator=g, dtype=torch.double) * 0.2}
    return l, P, g


def tconv_phases(x, wmod):
    """x (B,Cin,H,W), wmod (B,Cout,Cin,3,3) -> phase planes T[a][b] (B,Cout,H+1,W+1), T[2m+a,2n+b]."""
    B, Cin, H, W = x.shape
    xp = F.pad(x, [1, 1, 1, 1])                      # xp[i+1] = x[i]; index -1 and H read zero
    T = [[None, None], [None, None]]
    for a in (0, 1):
        for b in (0, 1):
            acc = 0
            for ky in ((0, 2) if a == 0 else (1,)):
                for kx in ((0, 2) if b == 0 else (1,)):
                    dy, dx = -(ky // 2), -(kx // 2)
                    xs = xp[:, :, 1 + dy:1 + dy + H + 1, 1 + dx:1 + dx + W + 1]      # x[m+dy, n+dx], m in [0,H]
                    acc = acc + torch.einsum("boi,bihw->bohw", wmod[:, :, :, ky, kx], xs)
            T[a][b] = acc
    return T


def interleave(T, H, W):
    B, C = T[0][0].shape[:2]
    full = T[0][0].new_zeros(B, C, 2 * H + 2, 2 * W + 2)
    for a in (0, 1):
        for b in (0, 1):
            full[:, :, a::2, b::2] = T[a][b]
    return full[:, :, :2 * H + 1, :2 * W + 1]


def test_tconv_phase_decomposition_and_blur():
    l, P, g = _layer("up", 6, 4)
    B, H = 2, 5
    x = torch.randn(B, 6, H, H, generator=g, dtype=torch.double)
    s = torch.randn(B, 6, generator=g, dtype=torch.double) + 1
    scale = 1 / math.sqrt(6 * 9)
    wmod = scale * P["L.conv.weight"] * s.view(B, 1, 6, 1, 1)
    T = interleave(tconv_phases(x, wmod), H, H)
    ref = F.conv_transpose2d(x.reshape(1, B * 6, H, H), wmod.transpose(1, 2).reshape(B * 6, 4, 3, 3), stride=2,
                             groups=B).view(B, 4, 2 * H + 1, 2 * H + 1)
    torch.testing.assert_close(T, ref)
    # blur with pad (1,1):  out[o] = sum_t T[o+t-1] * k[t] * 2 per axis
    k = torch.tensor(blur_kernel_1d(), dtype=torch.double) * 2
    Tp = F.pad(T, [1, 1, 1, 1])   # T[-1] and T[2H+1] are the zero pads
    out = 0
    for t in range(4):
        for u in range(4):
            out = out + Tp[:, :, t:t + 2 * H, u:u + 2 * H] * k[t] * k[u]
    ref_b = sg.upfirdn2d(ref, (sg.make_kernel_2d() * 4).double(), pad=(1, 1))
    torch.testing.assert_close(out, ref_b)


def _act_inverse(out):
    return torch.where(out > 0, out / SQ2, out / (0.2 * SQ2))


def _modlayer_backward_model(l, P, x, s, noise, gout, up):
    """What the kernels compute, in the order they compute it.  Returns gx, gs."""
    B, Cin = s.shape
    W = P["L.conv.weight"][0]
    Cout = W.shape[0]
    scale = 1 / math.sqrt(Cin * 9)
    Q = (scale ** 2) * (W ** 2).sum((2, 3))                     # (Cout,Cin)
    d = torch.rsqrt((s ** 2) @ Q.t() + 1e-8)                    # (B,Cout)
    wmod = scale * W[None] * s.view(B, 1, Cin, 1, 1)
    H = x.shape[-1]
    if up:
        z = sg.upfirdn2d(interleave(tconv_phases(x, wmod), H, H), (sg.make_kernel_2d() * 4).double(), pad=(1, 1))
    else:
        z = torch.stack([F.conv2d(x[b:b + 1], wmod[b], padding=1)[0] for b in range(B)])
    nw, bias = P["L.noise.weight"], P["L.activate.bias"].view(1, -1, 1, 1)
    out = F.leaky_relu(d[:, :, None, None] * z + nw * noise + bias, 0.2) * SQ2
    # ---- backward, from `out` and `gout` only (z is NOT kept)
    gy = gout * SQ2 * torch.where(out > 0, 1.0, 0.2)
    y = _act_inverse(out) - nw * noise - bias
    gd_acc = (gy * y).sum((2, 3))                                # (B,Cout) = d * sum(gy*z)
    gz = d[:, :, None, None] * gy
    Ws = scale * W                                               # shared, un-modulated
    if up:
        k2 = (sg.make_kernel_2d() * 4).double()
        gT = sg.upfirdn2d(gz, k2, pad=(2, 2))                    # blur^T : (2H)^2 -> (2H+1)^2
        gxt = F.conv2d(gT, Ws.transpose(0, 1).flip([]) if False else Ws.permute(1, 0, 2, 3), stride=2)
    else:
        gxt = F.conv_transpose2d(gz, Ws, padding=1)
    gs = (x * gxt).sum((2, 3)) - s * ((gd_acc * d * d) @ Q)
    gx = s[:, :, None, None] * gxt
    return out, gx, gs


def _check_layer(kind):
    up = kind == "up"
    l, P, g = _layer(kind, 6, 4, seed=1)
    B, H = 2, 5
    Ho = 2 * H if up else H
    x = torch.randn(B, 6, H, H, generator=g, dtype=torch.double, requires_grad=True)
    s = (torch.randn(B, 6, generator=g, dtype=torch.double) + 1).requires_grad_(True)
    noise = torch.randn(1, 1, Ho, Ho, generator=g, dtype=torch.double)
    gout = torch.randn(B, 4, Ho, Ho, generator=g, dtype=torch.double)
    P2 = dict(P)
    ref = sg.styled_conv(P2, l, x, s, noise)
    gx_ref, gs_ref = torch.autograd.grad((ref * gout).sum(), [x, s])
    out, gx, gs = _modlayer_backward_model(l, P, x.detach(), s.detach(), noise, gout, up)
    torch.testing.assert_close(out, ref.detach())
    torch.testing.assert_close(gx, gx_ref)
    torch.testing.assert_close(gs, gs_ref)


def test_modconv_backward_algebra():
    _check_layer("conv")


def test_upconv_backward_algebra():
    _check_layer("up")


def test_tconv_transpose_from_phase_planes():
    """gx~[i] = sum_k gT[2i+k] w[k]  ==  sum over planes (k%2) shifted by k//2 (the dgrad tap list)."""
    g = torch.Generator().manual_seed(2)
    B, Cin, Cout, H = 1, 3, 2, 4
    gT = torch.randn(B, Cout, 2 * H + 1, 2 * H + 1, generator=g, dtype=torch.double)
    W = torch.randn(Cout, Cin, 3, 3, generator=g, dtype=torch.double)
    ref = F.conv2d(gT, W.permute(1, 0, 2, 3), stride=2)
    gTp = F.pad(gT, [0, 1, 0, 1])
    planes = [[gTp[:, :, a::2, b::2] for b in (0, 1)] for a in (0, 1)]       # (H+1)x(W+1) each
    acc = 0
    for ky in range(3):
        for kx in range(3):
            pl = planes[ky % 2][kx % 2][:, :, ky // 2:ky // 2 + H, kx // 2:kx // 2 + H]
            acc = acc + torch.einsum("oi,bohw->bihw", W[:, :, ky, kx], pl)
    torch.testing.assert_close(acc, ref)


def test_torgb_and_skip_upsample_algebra():
    g = torch.Generator().manual_seed(3)
    B, C, H = 2, 5, 4
    l = ModLayer("R", "rgb", H, C, 3, 0)
    P = {"R.conv.weight": torch.randn(1, 3, C, 1, 1, generator=g, dtype=torch.double),
         "R.conv.modulation.weight": torch.randn(C, 8, generator=g, dtype=torch.double),
         "R.conv.modulation.bias": torch.ones(C, dtype=torch.double), "R.bias": torch.zeros(1, 3, 1, 1, dtype=torch.double)}
    x = torch.randn(B, C, H, H, generator=g, dtype=torch.double, requires_grad=True)
    s = (torch.randn(B, C, generator=g, dtype=torch.double) + 1).requires_grad_(True)
    skip = torch.randn(B, 3, H // 2, H // 2, generator=g, dtype=torch.double, requires_grad=True)
    grgb = torch.randn(B, 3, H, H, generator=g, dtype=torch.double)
    ref = sg.to_rgb(P, l, x, s, skip)
    gx_r, gs_r, gsk_r = torch.autograd.grad((ref * grgb).sum(), [x, s, skip])
    Wr = P["R.conv.weight"][0, :, :, 0, 0] / math.sqrt(C)
    gxt = torch.einsum("ci,bchw->bihw", Wr, grgb)
    torch.testing.assert_close(s.detach()[:, :, None, None] * gxt, gx_r)
    torch.testing.assert_close((x.detach() * gxt).sum((2, 3)), gs_r)
    # skip-upsample transpose: up=2 pad (2,1)  ->  down=2 pad (1,2) with the same (symmetric) kernel
    gsk = sg.upfirdn2d(grgb, (sg.make_kernel_2d() * 4).double(), down=2, pad=(1, 2))
    torch.testing.assert_close(gsk, gsk_r)


def fused_up_weights(w):
    """(Cout,Cin,3,3) -> phase weights (3,3,4,Cout,Cin): blur([1,3,3,1]*2 per axis, pad (1,1)) o conv_transpose(stride 2) collapses to
    four 3x3 convolutions over the INPUT grid, one per output phase (a,b):  out[2m+a][2n+b] = sum_{dy,dx} Weff[dy][dx][2a+b] x[m+dy-1][n+dx-1]."""
    k = torch.tensor(blur_kernel_1d(), dtype=w.dtype) * 2
    Cout, Cin = w.shape[:2]
    W = torch.zeros(3, 3, 4, Cout, Cin, dtype=w.dtype)
    for a in (0, 1):
        for b in (0, 1):
            for t in range(4):
                for ky in range(3):
                    if (a + t - 1 - ky) % 2:
                        continue
                    dy = (a + t - 1 - ky) // 2
                    if dy < -1 or dy > 1:
                        continue
                    for u in range(4):
                        for kx in range(3):
                            if (b + u - 1 - kx) % 2:
                                continue
                            dx = (b + u - 1 - kx) // 2
                            if dx < -1 or dx > 1:
                                continue
                            W[dy + 1, dx + 1, 2 * a + b] += w[:, :, ky, kx] * k[t] * k[u]
    return W


def test_fused_upsample_conv_equals_tconv_plus_blur():
    from sfattack.params import fused_up_base_weights
    g = torch.Generator().manual_seed(4)
    B, Cin, Cout, H = 2, 5, 3, 6
    x = torch.randn(B, Cin, H, H, generator=g, dtype=torch.double)
    w = torch.randn(Cout, Cin, 3, 3, generator=g, dtype=torch.double)
    ref = sg.upfirdn2d(F.conv_transpose2d(x, w.permute(1, 0, 2, 3), stride=2), (sg.make_kernel_2d() * 4).double(), pad=(1, 1))
    We = fused_up_weights(w)
    torch.testing.assert_close(fused_up_base_weights(w), We)
    out = torch.zeros(B, Cout, 2 * H, 2 * H, dtype=torch.double)
    xp = F.pad(x, [1, 1, 1, 1])
    for a in (0, 1):
        for b in (0, 1):
            acc = 0
            for dy in range(3):
                for dx in range(3):
                    acc = acc + torch.einsum("oi,bihw->bohw", We[dy, dx, 2 * a + b], xp[:, :, dy:dy + H, dx:dx + H])
            out[:, :, a::2, b::2] = acc
    torch.testing.assert_close(out, ref)
    # transpose: gx[i] = sum_{dy,dx,phase} Weff^T g[2(i-dy+1)+a][...]  == conv3x3^T over the space-to-depth view of g
    gz = torch.randn(B, Cout, 2 * H, 2 * H, generator=g, dtype=torch.double)
    xr = x.clone().requires_grad_(True)
    r2 = sg.upfirdn2d(F.conv_transpose2d(xr, w.permute(1, 0, 2, 3), stride=2), (sg.make_kernel_2d() * 4).double(), pad=(1, 1))
    (gx_ref,) = torch.autograd.grad((r2 * gz).sum(), xr)
    s2d = torch.stack([gz[:, :, a::2, b::2] for a in (0, 1) for b in (0, 1)], 1)          # (B,4,Cout,H,H)
    s2dp = F.pad(s2d, [1, 1, 1, 1])
    gx = 0
    for dy in range(3):
        for dx in range(3):
            # dgrad tap: gx[p] += W[dy][dx]^T g[p - (d - 1)]
            win = s2dp[:, :, :, 2 - dy:2 - dy + H, 2 - dx:2 - dx + H]
            gx = gx + torch.einsum("poi,bpohw->bihw", We[dy, dx], win)
    torch.testing.assert_close(gx, gx_ref)
