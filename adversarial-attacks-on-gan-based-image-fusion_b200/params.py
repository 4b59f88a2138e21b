"""Model specifications and deterministic random-init parameters.

No checkpoint of the reference's models can be fetched (no network) and the reference ships none,
so every weight on the attack path is random-init from a seed.  The SAME dictionaries built here
are handed to the CUDA path and to the CPU oracle, so both sides see identical bits.

Key names follow the public rosinality StyleGAN2 ``g_ema`` state-dict convention used by the
reference's un-vendored generator (call sites: code/style_fusion_simple.py:51,116-129,151-153;
code/attack/attack_main2.py:619-621), so a real checkpoint can be dropped in later.
VGG keys are positional, as code/vgg.py:66-77 copies the first 26 tensors of the .pth by position.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

# config-f channel table (channel_multiplier=2), SURVEY Appendix A.3
CHANNELS_F = {4: 512, 8: 512, 16: 512, 32: 512, 64: 512, 128: 256, 256: 128, 512: 64, 1024: 32}


@dataclass
class ModLayer:
    """One ModulatedConv2d of the synthesis network, in StyleSpace order."""
    name: str          # state-dict prefix, e.g. "convs.3" / "to_rgbs.1" / "conv1" / "to_rgb1"
    kind: str          # "conv" | "up" | "rgb"
    res: int           # output resolution
    cin: int
    cout: int
    w_idx: int         # which W+ row drives this layer
    s_off: int = 0     # offset of this layer's style inside the concatenated S vector
    noise_idx: int = -1


@dataclass
class GenSpec:
    size: int
    style_dim: int = 512
    n_mlp: int = 8
    channels: Dict[int, int] = field(default_factory=lambda: dict(CHANNELS_F))
    layers: List[ModLayer] = field(default_factory=list)

    @property
    def log_size(self) -> int:
        return int(math.log2(self.size))

    @property
    def n_latent(self) -> int:
        return self.log_size * 2 - 2

    @property
    def s_dim(self) -> int:
        return sum(l.cin for l in self.layers)

    @property
    def num_noises(self) -> int:
        return (self.log_size - 2) * 2 + 1


def gen_spec(size: int, style_dim: int = 512, n_mlp: int = 8,
             channels: Optional[Dict[int, int]] = None) -> GenSpec:
    ch = dict(CHANNELS_F if channels is None else channels)
    spec = GenSpec(size=size, style_dim=style_dim, n_mlp=n_mlp, channels=ch)
    L: List[ModLayer] = []
    L.append(ModLayer("conv1", "conv", 4, ch[4], ch[4], 0, noise_idx=0))
    L.append(ModLayer("to_rgb1", "rgb", 4, ch[4], 3, 1))
    cin = ch[4]
    i = 1
    for k, lr in enumerate(range(3, spec.log_size + 1)):
        res = 2 ** lr
        cout = ch[res]
        L.append(ModLayer(f"convs.{2 * k}", "up", res, cin, cout, i, noise_idx=2 * k + 1))
        L.append(ModLayer(f"convs.{2 * k + 1}", "conv", res, cout, cout, i + 1, noise_idx=2 * k + 2))
        L.append(ModLayer(f"to_rgbs.{k}", "rgb", res, cout, 3, i + 2))
        cin = cout
        i += 2
    off = 0
    for l in L:
        l.s_off = off
        off += l.cin
    spec.layers = L
    return spec


def make_generator_params(spec: GenSpec, seed: int = 0, trained_like: bool = True) -> Dict[str, torch.Tensor]:
    """Random-init in the rosinality convention.

    trained_like=True gives the noise strengths / biases small non-zero values (the public init is 0,
    which would leave those code paths unexercised by parity tests)."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    P: Dict[str, torch.Tensor] = {}
    lr_mlp = 0.01
    for i in range(spec.n_mlp):
        P[f"style.{i + 1}.weight"] = rn(spec.style_dim, spec.style_dim) / lr_mlp
        P[f"style.{i + 1}.bias"] = torch.zeros(spec.style_dim)
    P["input.input"] = rn(1, spec.channels[4], 4, 4)
    small = 0.1 if trained_like else 0.0
    for l in spec.layers:
        if l.kind == "rgb":
            P[f"{l.name}.conv.weight"] = rn(1, 3, l.cin, 1, 1)
            P[f"{l.name}.conv.modulation.weight"] = rn(l.cin, spec.style_dim)
            P[f"{l.name}.conv.modulation.bias"] = torch.ones(l.cin)
            P[f"{l.name}.bias"] = small * rn(1, 3, 1, 1)
        else:
            P[f"{l.name}.conv.weight"] = rn(1, l.cout, l.cin, 3, 3)
            P[f"{l.name}.conv.modulation.weight"] = rn(l.cin, spec.style_dim)
            P[f"{l.name}.conv.modulation.bias"] = torch.ones(l.cin)
            P[f"{l.name}.noise.weight"] = small * rn(1)
            P[f"{l.name}.activate.bias"] = small * rn(l.cout)
    for i in range(spec.num_noises):
        r = 2 ** ((i + 5) // 2)
        P[f"noises.noise_{i}"] = rn(1, 1, r, r)
    return P


# ----------------------------------------------------------------------------------------------
# VGG (code/vgg.py:12-39 layer table; only conv1_1..conv4_2 are executed, code/vgg.py:44-64)
VGG_CONVS = [("conv1_1", 3, 64), ("conv1_2", 64, 64), ("conv2_1", 64, 128), ("conv2_2", 128, 128),
             ("conv3_1", 128, 256), ("conv3_2", 256, 256), ("conv3_3", 256, 256),
             ("conv4_1", 256, 512), ("conv4_2", 512, 512), ("conv4_3", 512, 512),
             ("conv5_1", 512, 512), ("conv5_2", 512, 512), ("conv5_3", 512, 512)]
VGG_EXECUTED = 9  # conv1_1 .. conv4_2


def make_vgg_state_dict(seed: int = 0, width_div: int = 1) -> Dict[str, torch.Tensor]:
    """26 tensors in torchvision ``vgg16().features`` order (what code/vgg.py:70-74 expects in the .pth).

    He-normal init so that post-ReLU activations keep O(1) scale through 9 layers (torch's default
    Conv2d init shrinks them ~3x per layer, which would make the feature loss degenerate)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    tv_idx = [0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28]
    for (name, cin, cout), k in zip(VGG_CONVS, tv_idx):
        ci = cin if cin == 3 else cin // width_div
        co = cout // width_div
        std = math.sqrt(2.0 / (ci * 9))
        sd[f"{k}.weight"] = torch.randn(co, ci, 3, 3, generator=g) * std
        sd[f"{k}.bias"] = torch.randn(co, generator=g) * 0.05
    return sd


# ----------------------------------------------------------------------------------------------
# Encoder stand-in: image (B,3,256,256) in [-1,1] -> W+ codes (B,n_latent,512).
@dataclass
class EncSpec:
    n_latent: int
    style_dim: int = 512
    widths: tuple = (32, 64, 128, 256, 512)   # conv3x3+ReLU, each followed by maxpool2 except the last
    in_res: int = 256


def make_encoder_params(spec: EncSpec, seed: int = 1) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    P: Dict[str, torch.Tensor] = {}
    cin = 3
    for i, co in enumerate(spec.widths):
        P[f"convs.{i}.weight"] = torch.randn(co, cin, 3, 3, generator=g) * math.sqrt(2.0 / (cin * 9))
        P[f"convs.{i}.bias"] = torch.randn(co, generator=g) * 0.05
        cin = co
    out = spec.n_latent * spec.style_dim
    P["head.weight"] = torch.randn(out, cin, generator=g) * (1.0 / math.sqrt(cin))
    P["head.bias"] = torch.zeros(out)
    P["latent_avg"] = torch.randn(spec.n_latent, spec.style_dim, generator=g) * 0.1
    return P


# ----------------------------------------------------------------------------------------------
# Spatial pair-fusion stand-in: per-S-dimension gate
#   q = sigmoid(alpha*s_a + beta*s_b + c),  s = q*s_a + (1-q)*s_b
def make_fusion_params(s_dim: int, seed: int = 2) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    return {"alpha": torch.randn(s_dim, generator=g) * 0.5,
            "beta": torch.randn(s_dim, generator=g) * 0.5,
            "c": torch.randn(s_dim, generator=g) * 2.0}


def blur_kernel_1d() -> List[float]:
    """[1,3,3,1] normalised to sum 1 (make_kernel); 2-D kernel is the outer product."""
    return [0.125, 0.375, 0.375, 0.125]


def fused_up_base_weights(w: torch.Tensor) -> torch.Tensor:
    """(Cout,Cin,3,3) -> (3,3,4,Cout,Cin): `upfirdn2d([1,3,3,1]*4, pad (1,1)) o conv_transpose2d(stride 2)` (upsample=True)
    collapsed into four 3x3 convolutions over the INPUT grid, one per output phase p = 2a+b:
        out[2m+a][2n+b] = sum_{dy,dx} Weff[dy][dx][p] . x[m+dy-1][n+dx-1]
    (tests/test_kernel_math.py::test_fused_upsample_conv_equals_tconv_plus_blur).  The blur is linear and per-channel, so style
    modulation still acts on the input-channel axis of Weff and demodulation coefficients are those of the original weights."""
    k = torch.tensor(blur_kernel_1d(), dtype=w.dtype, device=w.device) * 2
    cout, cin = w.shape[:2]
    W = torch.zeros(3, 3, 4, cout, cin, dtype=w.dtype, device=w.device)
    for a in (0, 1):
        for b in (0, 1):
            for t in range(4):
                for ky in range(3):
                    if (a + t - 1 - ky) % 2 or not (-1 <= (a + t - 1 - ky) // 2 <= 1):
                        continue
                    dy = (a + t - 1 - ky) // 2
                    for u in range(4):
                        for kx in range(3):
                            if (b + u - 1 - kx) % 2 or not (-1 <= (b + u - 1 - kx) // 2 <= 1):
                                continue
                            dx = (b + u - 1 - kx) // 2
                            W[dy + 1, dx + 1, 2 * a + b] += w[:, :, ky, kx] * k[t] * k[u]
    return W
