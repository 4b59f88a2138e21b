"""ORACLE (test infrastructure only -- never imported by the product path).

CPU/PyTorch fp32 restatement of the StyleGAN2 config-f generator in the rosinality conventions that
the reference's un-vendored ``stylefusion.sf_stylegan2[_hook].SFGenerator`` follows.

PARITY UNPINNED: /root/reference does not contain this arithmetic; the only facts the
reference fixes are the constructor (code/style_fusion_simple.py:51), the forward kwargs
(code/style_fusion_simple.py:116-129,151-153; code/attack/attack_main2.py:619-621), ``.size``,
``.mean_latent(n)`` (code/style_fusion_simple.py:60) and the size / latent-count / truncation table
(code/style_fusion_simple.py:28-39).  Everything else restates the public architecture
 and is the specification the CUDA path is held to.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

from sfattack.params import GenSpec, ModLayer, blur_kernel_1d


def make_kernel_2d() -> torch.Tensor:
    k = torch.tensor(blur_kernel_1d(), dtype=torch.float32)
    return k[:, None] * k[None, :]


def upfirdn2d(x: torch.Tensor, k: torch.Tensor, up: int = 1, down: int = 1, pad=(0, 0)) -> torch.Tensor:
    """zero-insert x`up`, pad (p0 before, p1 after), correlate with flipped k, decimate."""
    B, C, H, W = x.shape
    if up > 1:
        z = x.new_zeros(B, C, H, up, W, up)
        z[:, :, :, 0, :, 0] = x
        x = z.reshape(B, C, H * up, W * up)
    p0, p1 = pad
    x = F.pad(x, [max(p0, 0), max(p1, 0), max(p0, 0), max(p1, 0)])
    if p0 < 0 or p1 < 0:
        x = x[:, :, max(-p0, 0): x.shape[2] - max(-p1, 0), max(-p0, 0): x.shape[3] - max(-p1, 0)]
    w = torch.flip(k, [0, 1])[None, None].to(x).repeat(C, 1, 1, 1)
    x = F.conv2d(x, w, groups=C)
    return x[:, :, ::down, ::down]


def pixel_norm(x: torch.Tensor) -> torch.Tensor:
    return x * torch.rsqrt(torch.mean(x * x, dim=1, keepdim=True) + 1e-8)


def equal_linear(x, weight, bias, lr_mul=1.0, act=False):
    scale = (1.0 / math.sqrt(weight.shape[1])) * lr_mul
    y = F.linear(x, weight * scale)
    if act:
        return F.leaky_relu(y + bias * lr_mul, 0.2) * math.sqrt(2.0)
    return y + bias * lr_mul


def mapping(P: Dict[str, torch.Tensor], spec: GenSpec, z: torch.Tensor) -> torch.Tensor:
    x = pixel_norm(z)
    for i in range(spec.n_mlp):
        x = equal_linear(x, P[f"style.{i + 1}.weight"], P[f"style.{i + 1}.bias"], lr_mul=0.01, act=True)
    return x


def layer_style(P, l: ModLayer, w_row: torch.Tensor) -> torch.Tensor:
    """s = modulation(w): EqualLinear(style_dim, cin, bias_init=1)."""
    return equal_linear(w_row, P[f"{l.name}.conv.modulation.weight"], P[f"{l.name}.conv.modulation.bias"])


def modulated_conv(P, l: ModLayer, x: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
    """Per-sample modulated (and, except ToRGB, demodulated) grouped conv."""
    B, Cin, H, W = x.shape
    weight = P[f"{l.name}.conv.weight"]            # (1, Cout, Cin, k, k)
    _, Cout, _, k, _ = weight.shape
    scale = 1.0 / math.sqrt(Cin * k * k)
    wgt = scale * weight * s.view(B, 1, Cin, 1, 1)
    if l.kind != "rgb":
        d = torch.rsqrt(wgt.pow(2).sum([2, 3, 4]) + 1e-8)
        wgt = wgt * d.view(B, Cout, 1, 1, 1)
    if l.kind == "up":
        xx = x.reshape(1, B * Cin, H, W)
        wt = wgt.transpose(1, 2).reshape(B * Cin, Cout, k, k)
        out = F.conv_transpose2d(xx, wt, padding=0, stride=2, groups=B)
        out = out.view(B, Cout, out.shape[2], out.shape[3])
        return upfirdn2d(out, make_kernel_2d() * 4.0, pad=(1, 1))
    xx = x.reshape(1, B * Cin, H, W)
    out = F.conv2d(xx, wgt.view(B * Cout, Cin, k, k), padding=k // 2, groups=B)
    return out.view(B, Cout, H, W)


def styled_conv(P, l: ModLayer, x, s, noise):
    out = modulated_conv(P, l, x, s)
    out = out + P[f"{l.name}.noise.weight"] * noise
    return F.leaky_relu(out + P[f"{l.name}.activate.bias"].view(1, -1, 1, 1), 0.2) * math.sqrt(2.0)


def to_rgb(P, l: ModLayer, x, s, skip):
    out = modulated_conv(P, l, x, s) + P[f"{l.name}.bias"]
    if skip is not None:
        out = out + upfirdn2d(skip, make_kernel_2d() * 4.0, up=2, pad=(2, 1))
    return out


def styles_from_wplus(P, spec: GenSpec, wplus: torch.Tensor) -> List[torch.Tensor]:
    """W+ (B, n_latent, 512) -> StyleSpace list (conv1, to_rgb1, then up/conv/to_rgb per resolution)."""
    return [layer_style(P, l, wplus[:, l.w_idx]) for l in spec.layers]


def synthesis_from_styles(P, spec: GenSpec, styles: Sequence[torch.Tensor], return_features=False):
    B = styles[0].shape[0]
    x = P["input.input"].repeat(B, 1, 1, 1)
    skip = None
    feats = []
    for l, s in zip(spec.layers, styles):
        if l.kind == "rgb":
            skip = to_rgb(P, l, x, s, skip)
            feats.append(x)
        else:
            x = styled_conv(P, l, x, s, P[f"noises.noise_{l.noise_idx}"])
    if return_features:
        return skip, feats
    return skip


def mean_latent(P, spec: GenSpec, n: int, seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(n, spec.style_dim, generator=g).to(P["input.input"].device)   # the oracle may be run on a GPU (fp32, TF32 off) for the large cases
    return mapping(P, spec, z).mean(0, keepdim=True)


class OracleGenerator:
    """Callable with the rosinality forward signature used by the reference call sites."""

    def __init__(self, spec: GenSpec, params: Dict[str, torch.Tensor]):
        self.spec, self.P = spec, params
        self.size = spec.size
        self.n_latent = spec.n_latent

    def mean_latent(self, n):
        return mean_latent(self.P, self.spec, n)

    def get_latent(self, z):
        return mapping(self.P, self.spec, z)

    def __call__(self, styles, return_latents=False, inject_index=None, truncation=1, truncation_latent=None,
                 input_is_latent=False, noise=None, randomize_noise=True, return_style_vector=False,
                 style_vector=None):
        assert not randomize_noise, "the attack path always passes randomize_noise=False"
        spec, P = self.spec, self.P
        if style_vector is not None:
            img, feats = synthesis_from_styles(P, spec, style_vector, return_features=True)
            return img, feats, None
        if not input_is_latent:
            styles = [mapping(P, spec, s) for s in styles]
        if truncation < 1:
            styles = [truncation_latent + truncation * (s - truncation_latent) for s in styles]
        if len(styles) < 2:
            inject_index = spec.n_latent
            latent = styles[0].unsqueeze(1).repeat(1, inject_index, 1) if styles[0].ndim < 3 else styles[0]
        else:
            if inject_index is None:
                inject_index = spec.n_latent // 2
            la = styles[0].unsqueeze(1).repeat(1, inject_index, 1)
            lb = styles[1].unsqueeze(1).repeat(1, spec.n_latent - inject_index, 1)
            latent = torch.cat([la, lb], 1)
        svec = styles_from_wplus(P, spec, latent)
        if return_style_vector:
            return svec
        img = synthesis_from_styles(P, spec, svec)
        if return_latents:
            return img, latent
        return img, None
