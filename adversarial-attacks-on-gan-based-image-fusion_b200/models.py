"""`net` object of the reference (pSp / e4e, un-vendored: code/utils/model_utils.py:7-18): `.encoder`, `.decoder`,
`.latent_avg`, `.opts.start_from_latent_avg`.  The encoder is the documented stand-in."""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict

import torch

from . import lib
from .engine import ConvStack, encoder_layers
from .generator import Generator
from .params import EncSpec, make_encoder_params


class Encoder:
    """image (B,3,256,256) in [-1,1] -> W+ codes (B, n_latent, 512), forward only (the attack engines own the backward)."""

    def __init__(self, spec: EncSpec, params=None, device="cuda:0", seed: int = 1):
        self.spec = spec
        self.params = params if params is not None else make_encoder_params(spec, seed)
        self.device = torch.device(device)
        self._stacks: Dict[tuple, ConvStack] = {}
        self._err = None
        self.version = 0

    def to(self, device):
        if torch.device(device) != self.device:
            self.device = torch.device(device)
            self._stacks.clear()
            self._err = None
            self.version += 1
        return self

    def eval(self):
        return self

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        n = x.shape[0]
        key = (n, lib.mode_key())
        if key not in self._stacks:
            if self._err is None:
                self._err = torch.zeros(1, dtype=torch.int32, device=self.device)
            w = [(self.params[f"convs.{i}.weight"], self.params[f"convs.{i}.bias"]) for i in range(len(self.spec.widths))]
            self._stacks[key] = ConvStack(encoder_layers(self.spec), w, n, self.spec.in_res, self.device, self._err)
            self._hw = self.params["head.weight"].to(self.device).contiguous()
            self._hb = self.params["head.bias"].to(self.device).contiguous()
        st = self._stacks[key]
        top = st.forward(x.to(self.device, torch.float32).contiguous())
        feat = torch.empty(n, self.spec.widths[-1], device=self.device)
        lib.gap_fwd(top, feat)
        codes = torch.empty(n, self.spec.n_latent * self.spec.style_dim, device=self.device)
        lib.linear_fwd(feat, self._hw, self._hb, codes)
        return codes.view(n, self.spec.n_latent, self.spec.style_dim)


class PSPNet:
    def __init__(self, size: int = 1024, device="cuda:0", gen_params=None, enc_params=None, seed: int = 0, channels=None, style_dim=512,
                 n_mlp=8, enc_widths=(32, 64, 128, 256, 512), enc_res=256):
        self.decoder = Generator(size, style_dim, n_mlp, device=device, seed=seed, channels=channels, params=gen_params)
        espec = EncSpec(n_latent=self.decoder.n_latent, style_dim=style_dim, widths=tuple(enc_widths), in_res=enc_res)
        self.encoder = Encoder(espec, enc_params, device, seed + 1)
        self.latent_avg = self.encoder.params["latent_avg"].to(device)
        self.opts = SimpleNamespace(start_from_latent_avg=True, device=device, stylegan_size=size)

    def to(self, device):
        self.decoder.to(device)
        self.encoder.to(device)
        self.latent_avg = self.latent_avg.to(device)
        return self

    def eval(self):
        return self


def setup_model(checkpoint_path=None, device="cuda:0", size: int = 1024, **kw):
    """code/utils/model_utils.py:7-18 -> (net, opts).  Without a checkpoint the model is random-init (no network here)."""
    net = PSPNet(size, device, **kw)
    if checkpoint_path:
        ckpt = torch.load(checkpoint_path, map_location="cpu")
        net.decoder.load_state_dict({k[len("decoder."):]: v for k, v in ckpt["state_dict"].items() if k.startswith("decoder.")})
        if "latent_avg" in ckpt:
            net.latent_avg = ckpt["latent_avg"].to(device)
    return net, net.opts
