"""Drop-in for the reference's feature extractor (code/vgg.py): `vgg16(pth) -> VGGBase`,
`vgg(x) -> (conv1_1, conv1_2, conv3_2, conv4_2)` (code/vgg.py:44-64, 79-81).  The .pth holds torchvision-ordered tensors that
are copied by POSITION into conv1_1..conv5_3 (code/vgg.py:66-77); only conv1_1..conv4_2 are ever executed."""
from __future__ import annotations

from typing import Dict

import torch

from . import lib
from .engine import ConvStack, vgg_layers
from .params import VGG_EXECUTED


class VGGBase:
    def __init__(self, pth, device="cuda:0"):
        self.pth = pth
        sd = pth if isinstance(pth, dict) else torch.load(pth, map_location="cpu")
        vals = list(sd.values())
        assert len(vals) >= 2 * VGG_EXECUTED, "state dict too short (code/vgg.py:73-74 copies the first 26 tensors)"
        self.sd = sd
        self.weights = [(vals[2 * i], vals[2 * i + 1]) for i in range(VGG_EXECUTED)]
        self.width_div = 64 // self.weights[0][0].shape[0]
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

    def stack(self, n: int, res: int) -> ConvStack:
        key = (n, res, lib.mode_key())
        if key not in self._stacks:
            if self._err is None:
                self._err = torch.zeros(1, dtype=torch.int32, device=self.device)
            self._stacks[key] = ConvStack(vgg_layers(self.width_div), self.weights, n, res, self.device, self._err)
        return self._stacks[key]

    def forward(self, image: torch.Tensor):
        n, c, h, w = image.shape
        assert c == 3 and h == w, "square RGB input (the reference always feeds 256x256, attack_main2.py:590-591)"
        st = self.stack(n, h)
        st.forward(image.to(self.device, torch.float32).contiguous())
        outs = []
        for t in st.tap_outputs():
            y = torch.empty(t.shape[0], t.shape[3], t.shape[1], t.shape[2], device=self.device)
            lib.nhwc_bf16_to_nchw(t, y)
            outs.append(y)
        return tuple(outs)

    __call__ = forward


def vgg16(pth, device="cuda:0"):
    return VGGBase(pth, device)
