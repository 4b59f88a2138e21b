"""StyleGAN2 generator with the rosinality call surface the reference relies on
(`SFGenerator_hook(size, 512, 8, GAN, weights)`, code/style_fusion_simple.py:51; `decoder([w], input_is_latent=True,
randomize_noise=False, return_latents=True)`, code/attack/attack_main2.py:619-621; `.size`, `.mean_latent(n)`,
`return_style_vector=` / `style_vector=`, code/style_fusion_simple.py:60,116-129,151-153).
Synthesis runs on the CUDA kernels (engine.SynthesisEngine); this class only marshals latents."""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch

from . import lib
from .engine import SynthesisEngine
from .params import GenSpec, gen_spec, make_generator_params


class Generator:
    def __init__(self, size: int, style_dim: int = 512, n_mlp: int = 8, channel_multiplier: int = 2, device="cuda:0", seed: int = 0,
                 channels: Optional[Dict[int, int]] = None, params: Optional[Dict[str, torch.Tensor]] = None):
        assert channel_multiplier == 2, "config-f only"
        self.spec: GenSpec = gen_spec(size, style_dim, n_mlp, channels)
        self.size, self.style_dim, self.n_latent = size, style_dim, self.spec.n_latent
        self.device = torch.device(device)
        self.params = params if params is not None else make_generator_params(self.spec, seed)
        self._engines: Dict[tuple, SynthesisEngine] = {}
        self._err = None
        self._map = None
        self.version = 0          # bumped whenever the weights or the device change (engine caches key on it)

    # --- torch.nn.Module-ish conveniences used by the reference scripts
    def to(self, device):
        if torch.device(device) != self.device:
            self.device = torch.device(device)
            self._engines.clear()
            self._err = None
            self._map = None
            self.version += 1
        return self

    def eval(self):
        return self

    def state_dict(self):
        return self.params

    def load_state_dict(self, sd, strict=True):
        missing = [k for k in self.params if k not in sd]
        if strict and missing:
            raise KeyError(f"missing keys: {missing[:5]}")
        self.params = {k: sd[k].detach().clone() for k in self.params}
        self._engines.clear()
        self._map = None
        self.version += 1

    def engine(self, batch: int) -> SynthesisEngine:
        key = (batch, lib.mode_key())
        if key not in self._engines:
            if self._err is None:
                self._err = torch.zeros(1, dtype=torch.int32, device=self.device)
            self._engines[key] = SynthesisEngine(self.spec, self.params, batch, self.device, self._err)
        return self._engines[key]

    # --- mapping network z -> w  (8 x EqualLinear(lr_mul=0.01) + fused lrelu; off the hot path)
    def get_latent(self, z: torch.Tensor) -> torch.Tensor:
        if self._map is None:
            sc = (1.0 / math.sqrt(self.style_dim)) * 0.01
            self._map = [((self.params[f"style.{i + 1}.weight"] * sc).to(self.device).contiguous(),
                          (self.params[f"style.{i + 1}.bias"] * 0.01).to(self.device).contiguous()) for i in range(self.spec.n_mlp)]
        x = z.to(self.device, torch.float32)
        x = (x * torch.rsqrt(torch.mean(x * x, dim=1, keepdim=True) + 1e-8)).contiguous()
        for W, b in self._map:
            y = torch.empty(x.shape[0], W.shape[0], device=self.device)
            lib.linear_fwd(x, W, b, y)
            x = torch.nn.functional.leaky_relu(y, 0.2) * math.sqrt(2.0)
        return x

    def mean_latent(self, n_latent: int) -> torch.Tensor:
        g = torch.Generator().manual_seed(0)
        z = torch.randn(n_latent, self.style_dim, generator=g)
        return self.get_latent(z).mean(0, keepdim=True)

    def __call__(self, styles, return_latents=False, inject_index=None, truncation=1, truncation_latent=None, input_is_latent=False,
                 noise=None, randomize_noise=True, return_style_vector=False, style_vector=None):
        assert not randomize_noise, "the attack path always passes randomize_noise=False (attack_main2.py:620)"
        spec = self.spec
        if style_vector is not None:
            s = torch.cat([t.to(self.device, torch.float32) for t in style_vector], 1).contiguous()
            eng = self.engine(s.shape[0])
            img = eng.forward(s).clone()
            feats = [e["out"].float().permute(0, 3, 1, 2) for e in eng.L if e["l"].kind == "conv"]
            return img, feats, None
        styles = [s.to(self.device, torch.float32) for s in styles]
        if not input_is_latent:
            styles = [self.get_latent(s) for s in styles]
        if truncation < 1:
            styles = [truncation_latent + truncation * (s - truncation_latent) for s in styles]
        if len(styles) < 2:
            latent = styles[0].unsqueeze(1).repeat(1, spec.n_latent, 1) if styles[0].ndim < 3 else styles[0]
        else:
            if inject_index is None:
                inject_index = spec.n_latent // 2
            latent = torch.cat([styles[0].unsqueeze(1).repeat(1, inject_index, 1),
                                styles[1].unsqueeze(1).repeat(1, spec.n_latent - inject_index, 1)], 1)
        latent = latent.contiguous()
        eng = self.engine(latent.shape[0])
        s = eng.styles_from_wplus(latent)
        if return_style_vector:
            return [s[:, l.s_off:l.s_off + l.cin].clone() for l in spec.layers]
        img = eng.forward().clone()
        return (img, latent) if return_latents else (img, None)
