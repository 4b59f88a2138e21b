"""ORACLE (test infrastructure only -- the product path never imports this).

CPU/PyTorch restatement of the fusion-model facade and the two fusion entry points:
  * `generate_img` part-swap lists and latent conversions: code/style_fusion_simple.py:82-108, 126-144 (the reference's own code:
    the swap lists ARE the contract);
  * `fusion()` latent order / roles per dataset: code/attack/attack_main2.py:521-581;  `interpolation()`: code/attack/interpolation.py:658-669;
  * the hierarchy blender `base_blender.forward(s_dict)` (code/style_fusion_simple.py:164) is the un-vendored StyleFusion
    hierarchy: restated as the documented stand-in of SURVEY A.4 (a chain of per-dimension gates in hierarchy order) ->
    PARITY UNPINNED for that piece.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import stylegan2 as sg

PARTS = {   # get_all_active_parts() order of the three hierarchies (swap lists at code/style_fusion_simple.py:93-102)
    "ffhq": ["all", "bg_hair_clothes", "hair", "face", "eyes", "skin_mouth", "mouth", "skin", "shirt", "background", "background_top",
             "background_bottom", "bg"],
    "car": ["all", "wheels", "car", "body", "car_body", "background", "background_top", "background_bottom", "bg"],
    "church": ["all", "background", "background_top", "background_bottom", "bg"],
}
TABLE = {"ffhq": (0.7, 1024, 18), "car": (0.5, 512, 16), "church": (0.5, 256, 14)}    # code/style_fusion_simple.py:28-39


def gate(p, s_a, s_b):
    q = torch.sigmoid(p["alpha"] * s_a + p["beta"] * s_b + p["c"])
    return q * s_a + (1.0 - q) * s_b


def blend(parts: List[str], gates: Dict[str, dict], s_dict: Dict[str, torch.Tensor]) -> torch.Tensor:
    out = s_dict[parts[0]]
    for p in parts[1:]:
        if p not in s_dict or torch.equal(s_dict[p], out):
            continue
        out = gate(gates[p], out, s_dict[p])
    return out


class OracleFusion:
    """StyleFusionSimple restated on the oracle generator; gates are handed in so both sides hold the same bits."""

    def __init__(self, stylegan_type: str, O: sg.OracleGenerator, gates: Dict[str, dict], mean_latent: Optional[torch.Tensor] = None):
        self.type = stylegan_type
        self.truncation, self.size, self.layers = TABLE[stylegan_type]
        self.O, self.gates, self.parts = O, gates, PARTS[stylegan_type]
        self.mean_latent = O.mean_latent(4096) if mean_latent is None else mean_latent

    def w_plus_to_s(self, w_plus, truncation=1):                                               # :126-129
        return torch.cat(self.O([w_plus], input_is_latent=True, truncation=truncation, truncation_latent=self.mean_latent,
                                randomize_noise=False, return_style_vector=True), 1)

    def z_to_s(self, z):                                                                       # :115-118
        return torch.cat(self.O([z], truncation=self.truncation, truncation_latent=self.mean_latent, randomize_noise=False,
                                return_style_vector=True), 1)

    def general_latent_to_s(self, l, latent_type):                                             # :131-144
        assert latent_type in ["z", "w", "w+", "s"]
        if latent_type == "z":
            assert l.size() == (1, 512)
            return self.z_to_s(l)
        if latent_type in ("w", "w+"):
            assert l.size() == (1, 512) or l.size() == (1, self.layers, 512)
            if l.dim() == 2:
                return self.w_plus_to_s(l.unsqueeze(0).repeat(1, self.layers, 1))
            return self.w_plus_to_s(l)
        return l

    def s_to_image(self, s):                                                                   # :146-153
        styles = [s[:, l.s_off:l.s_off + l.cin] for l in self.O.spec.layers]
        img, feats, _ = self.O([torch.zeros(1, 512, device=s.device)], randomize_noise=False, style_vector=styles)
        return img, feats

    def generate_img(self, base_latent, latents_type="z", hair=None, face=None, background=None, all=None, mouth=None, eyes=None,
                     wheels=None, car=None, bg_top=None, bg_bottom=None):                      # :82-108
        s_dict = {part: self.general_latent_to_s(base_latent, latents_type) for part in self.parts}

        def swap(value, keys):
            if value is not None:
                for k in keys:
                    s_dict[k] = self.general_latent_to_s(value, latents_type)

        swap(hair, ["bg_hair_clothes", "hair"])
        swap(face, ["face", "eyes", "skin_mouth", "mouth", "skin", "shirt"])
        swap(background, ["background", "background_top", "background_bottom", "bg"])
        swap(all, ["all"])
        swap(mouth, ["skin_mouth", "face"])
        swap(eyes, ["eyes", "face"])
        swap(wheels, ["wheels"])
        swap(car, ["car", "body", "wheels", "car_body"])
        swap(bg_top, ["background_top"])
        swap(bg_bottom, ["background_bottom"])
        return self.s_to_image(blend(self.parts, self.gates, s_dict))


def fusion(dataset_name, all_latents, drawer: OracleFusion, feature_idx=-1):                   # attack_main2.py:521-581
    lat = list(all_latents.unsqueeze(1))
    if "ffhq" in dataset_name:
        z_mouth, z_background, z_hair, z_eyes, z_global = lat
        I_fused, _ = drawer.generate_img(z_global, hair=z_hair, eyes=z_eyes, background=z_background, mouth=z_mouth, latents_type="w")
        singles = [z_mouth, z_background, z_hair, z_eyes, z_global]
    elif "car" in dataset_name:
        z_wheel, z_bg_top, z_bg_bottom, z_body = lat
        I_fused, _ = drawer.generate_img(z_body, wheels=z_wheel, bg_top=z_bg_top, bg_bottom=z_bg_bottom, latents_type="w")
        singles = [z_body, z_wheel, z_bg_top, z_bg_bottom]
    else:
        z_bg_top, z_bg_bottom, z_body = lat
        I_fused, _ = drawer.generate_img(z_body, bg_top=z_bg_top, bg_bottom=z_bg_bottom, latents_type="w")
        singles = [z_body, z_bg_top, z_bg_bottom]
    imgs, feats = zip(*[drawer.generate_img(z, latents_type="w") for z in singles])
    return I_fused, torch.cat(imgs, 0), torch.cat([f[feature_idx] for f in feats], 0)


def interpolation(drawer: OracleFusion, all_latents, feature_idx=-1):                          # interpolation.py:658-669
    I_fused, _ = drawer.generate_img(torch.mean(all_latents, dim=0, keepdim=True), latents_type="w")
    imgs, feats = zip(*[drawer.generate_img(all_latents[i].unsqueeze(0), latents_type="w") for i in range(all_latents.size(0))])
    return I_fused, torch.cat(imgs, 0), torch.cat([f[feature_idx] for f in feats], 0)
