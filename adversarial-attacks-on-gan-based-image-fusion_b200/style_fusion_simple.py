"""Drop-in for code/style_fusion_simple.py: same constructor, attributes and methods (file:line cited per method).
The StyleFusion hierarchy of FusionNets (`stylefusion.sf_hierarchy`, un-vendored) is replaced by the documented stand-in of
SURVEY A.4: parts are blended pairwise in StyleSpace with a learned (here random-init) per-dimension gate, keeping the
`s_dict` / part-name interface so a real hierarchy can be dropped in later."""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional

import torch

from . import lib
from .generator import Generator
from .params import make_fusion_params

_PARTS = {   # get_all_active_parts() of the three hierarchies (code/style_fusion_simple.py:62-71, swap lists :89-104)
    "ffhq": ["all", "bg_hair_clothes", "hair", "face", "eyes", "skin_mouth", "mouth", "skin", "shirt", "background", "background_top",
             "background_bottom", "bg"],
    "car": ["all", "wheels", "car", "body", "car_body", "background", "background_top", "background_bottom", "bg"],
    "church": ["all", "background", "background_top", "background_bottom", "bg"],
}


def part_sources(parts: List[str], base: int, hair=None, face=None, background=None, all=None, mouth=None, eyes=None, wheels=None,
                 car=None, bg_top=None, bg_bottom=None) -> List[int]:
    """The s_dict of generate_img (code/style_fusion_simple.py:84-104) with input indices in place of style tensors: every part
    starts on `base`, then the keyword roles overwrite their part lists in the reference's order (later swaps win)."""
    src = {p_: base for p_ in parts}
    for value, keys in ((hair, ["bg_hair_clothes", "hair"]), (face, ["face", "eyes", "skin_mouth", "mouth", "skin", "shirt"]),
                        (background, ["background", "background_top", "background_bottom", "bg"]), (all, ["all"]),
                        (mouth, ["skin_mouth", "face"]), (eyes, ["eyes", "face"]), (wheels, ["wheels"]),
                        (car, ["car", "body", "wheels", "car_body"]), (bg_top, ["background_top"]), (bg_bottom, ["background_bottom"])):
        if value is not None:
            for k in keys:
                if k in src:
                    src[k] = value
    return [src[p_] for p_ in parts]


def fusion_hierarchy(dataset_name: str, parts: List[str], gates: Dict[str, dict]) -> dict:
    """Role assignment of `fusion()` (code/attack/attack_main2.py:521-581) as the attack engine's `hierarchy` argument: ffhq inputs
    [mouth, background, hair, eyes, global] (:526), car [wheel, bg_top, bg_bottom, body] (:547), church [bg_top, bg_bottom, body]
    (:566); the LAST input is the base; names are matched by substring as upstream."""
    if "ffhq" in dataset_name:
        source, n = part_sources(parts, 4, hair=2, eyes=3, background=1, mouth=0), 5
    elif "car" in dataset_name:
        source, n = part_sources(parts, 3, wheels=0, bg_top=1, bg_bottom=2), 4
    elif "church" in dataset_name:
        source, n = part_sources(parts, 2, bg_top=0, bg_bottom=1), 3
    else:
        raise ValueError(f"unknown dataset {dataset_name!r} (expected a name containing ffhq / car / church)")
    return dict(parts=list(parts), source=source, gates=gates, n_inputs=n)


class _Node:
    """One node of the hierarchy (stand-in for a `stylefusion.sf_hierarchy` node): its FusionNet is the per-dimension
    gate q = sigmoid(alpha*s_a + beta*s_b + c), s = q*s_a + (1-q)*s_b that blends this part's StyleSpace vector into the running
    result.  `load_fusion_net(path, device)` (code/style_fusion_simple.py:77) loads the gate {"alpha","beta","c"} from a torch file."""

    def __init__(self, name: str, s_dim: int, device, seed: int):
        self.name, self.device, self.s_dim = name, device, s_dim
        self.fusion_net = _Gate({k: v.to(device) for k, v in make_fusion_params(s_dim, seed).items()})

    def load_fusion_net(self, path, device):
        sd = torch.load(path, map_location="cpu")
        missing = [k for k in ("alpha", "beta", "c") if k not in sd]
        if missing:
            raise KeyError(f"{path}: fusion-net file lacks {missing} (expected the gate tensors alpha/beta/c of length {self.s_dim})")
        for k in ("alpha", "beta", "c"):
            if tuple(sd[k].shape) != (self.s_dim,):
                raise ValueError(f"{path}: {k} has shape {tuple(sd[k].shape)}, expected ({self.s_dim},)")
        self.fusion_net = _Gate({k: sd[k].float().to(device).contiguous() for k in ("alpha", "beta", "c")})
        return self.fusion_net


class _Gate:
    def __init__(self, p):
        self.p = p

    def to(self, device):
        self.p = {k: v.to(device) for k, v in self.p.items()}
        return self

    def eval(self):
        return self

    def __call__(self, s_a: torch.Tensor, s_b: torch.Tensor) -> torch.Tensor:
        out = torch.empty_like(s_a)
        lib.fuse_spatial_fwd(s_a.contiguous(), s_b.contiguous(), self.p["alpha"], self.p["beta"], self.p["c"], out)
        return out


class _Root(_Node):
    """nodes["all"]: walks the parts in hierarchy order and gates every part whose style differs from the running result"""

    def __init__(self, hierarchy, parts: List[str], s_dim: int, device, seed: int):
        super().__init__("all", s_dim, device, seed)
        self.hierarchy, self.parts = hierarchy, parts

    def get_all_active_parts(self):
        return list(self.parts)

    def forward(self, s_dict: Dict[str, list]):
        cat = lambda s: torch.cat(list(s), 1).contiguous() if isinstance(s, (list, tuple)) else s
        out = cat(s_dict[self.parts[0]])
        for p in self.parts[1:]:
            if p not in s_dict:
                continue
            b = cat(s_dict[p])
            if torch.equal(b, out):
                continue
            out = self.hierarchy.nodes[p].fusion_net(out, b)
        return out


class _Hierarchy:
    """stand-in for SFHierarchyFFHQ / SFHierarchyCar / SFHierarchyChurch: `.nodes[name]`, root `nodes["all"]`"""

    def __init__(self, parts: List[str], s_dim: int, device, seed: int = 2):
        self.nodes = {p: _Node(p, s_dim, device, seed + i) for i, p in enumerate(parts) if p != "all"}
        self.nodes["all"] = _Root(self, parts, s_dim, device, seed + parts.index("all"))


class StyleFusionSimple:
    def __init__(self, stylegan_type, stylegan_weights, fusion_nets_weights, device, GAN=None):          # style_fusion_simple.py:26
        self.stylegan_type = stylegan_type
        self.truncation, self.stylegan_size, self.stylegan_layers = {"ffhq": (0.7, 1024, 18), "car": (0.5, 512, 16),
                                                                     "church": (0.5, 256, 14)}[stylegan_type]   # :28-39
        self.device = device
        if GAN is not None and isinstance(GAN, Generator) and GAN.size == self.stylegan_size:
            self.original_net = GAN                                                                        # :51 reuses the decoder
        else:
            self.original_net = Generator(self.stylegan_size, 512, 8, device=device)
            if stylegan_weights:
                self.original_net.load_state_dict(torch.load(stylegan_weights, map_location="cpu")["g_ema"], strict=True)
        self.original_net.to(self.device)
        self.mean_latent = self.original_net.mean_latent(4096)                                             # :60
        self.sf_hierarchy = _Hierarchy(_PARTS[stylegan_type], self.original_net.spec.s_dim, self.device)   # :62-71
        self.base_blender = self.sf_hierarchy.nodes["all"]
        if fusion_nets_weights:
            with open(fusion_nets_weights, "r") as f:                                                      # :73-80
                fusion_nets_paths = json.load(f)
            base = os.path.dirname(os.path.abspath(fusion_nets_weights))
            for key in fusion_nets_paths.keys():
                path = fusion_nets_paths[key]
                self.sf_hierarchy.nodes[key].load_fusion_net(path if os.path.isabs(path) else os.path.join(base, path), self.device)
                self.sf_hierarchy.nodes[key].fusion_net.to(self.device)
                self.sf_hierarchy.nodes[key].fusion_net.eval()

    def generate_img(self, base_latent, latents_type="z", hair=None, face=None, background=None, all=None, mouth=None, eyes=None,
                     wheels=None, car=None, bg_top=None, bg_bottom=None):                                  # :82-108
        s_dict = dict()
        for part in self.sf_hierarchy.nodes["all"].get_all_active_parts():
            s_dict[part] = self.general_latent_to_s(base_latent, latents_type)

        def swap(value, keys):
            if value is None:
                return
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
        return self.s_dict_to_image(s_dict)

    # ---- the same part assignment with input INDICES instead of style tensors: what the attack engine's N-way gradient path needs
    def part_sources(self, base: int, **roles) -> List[int]:
        """generate_img's s_dict (:84-104) as a list: for every active part, the index of the input whose style vector it holds."""
        return part_sources(self.sf_hierarchy.nodes["all"].get_all_active_parts(), base, **roles)

    def fusion_hierarchy(self, dataset_name: str) -> dict:
        """`AttackEngine(..., fusion="hierarchy", n_inputs=N, hierarchy=...)` argument for the role assignment of `fusion()`
        (code/attack/attack_main2.py:521-581); the gates are this model's FusionNet stand-ins."""
        parts = self.sf_hierarchy.nodes["all"].get_all_active_parts()
        gates = {p_: self.sf_hierarchy.nodes[p_].fusion_net.p for p_ in parts if p_ != "all"}
        return fusion_hierarchy(dataset_name, parts, gates)

    def seed_to_z(self, seed):                                                                             # :110-113
        torch.manual_seed(seed[0])
        z_regular = torch.randn((seed[1] + 1, 1, 512), device=self.device)
        return z_regular[seed[1]]

    def z_to_s(self, z):                                                                                   # :115-118
        return self.original_net([z], truncation=self.truncation, truncation_latent=self.mean_latent, randomize_noise=False,
                                 return_style_vector=True)

    def z_to_w_plus(self, z):                                                                              # :120-124
        _, res = self.original_net([z], truncation=self.truncation, truncation_latent=self.mean_latent, randomize_noise=False,
                                   return_latents=True)
        return res[0]

    def w_plus_to_s(self, w_plus, truncation):                                                             # :126-129
        return self.original_net([w_plus], input_is_latent=True, truncation=truncation, truncation_latent=self.mean_latent,
                                 randomize_noise=False, return_style_vector=True)

    def general_latent_to_s(self, l, latent_type):                                                         # :131-144
        assert latent_type in ["z", "w", "w+", "s"]
        if latent_type == "z":
            assert l.size() == (1, 512)
            return self.z_to_s(l)
        elif latent_type == "w" or latent_type == "w+":
            assert l.size() == (1, 512) or l.size() == (1, self.stylegan_layers, 512)
            if l.dim() == 2:
                return self.w_plus_to_s(l.unsqueeze(0).repeat(1, self.stylegan_layers, 1), truncation=1)
            return self.w_plus_to_s(l, truncation=1)
        return l

    def s_to_image(self, s):                                                                               # :146-153
        if torch.is_tensor(s):
            s = [s[:, l.s_off:l.s_off + l.cin] for l in self.original_net.spec.layers]
        img, features, _ = self.original_net([torch.zeros(1, 512, device=self.device)], randomize_noise=False, style_vector=s)
        return img, features

    def w_plus_to_image(self, w_plus):                                                                     # :155-157
        return self.s_to_image(self.w_plus_to_s(w_plus, truncation=1))

    def z_to_image(self, z):                                                                               # :159-161
        return self.s_to_image(self.z_to_s(z))

    def s_dict_to_image(self, s_dict):                                                                     # :163-165
        return self.s_to_image(self.base_blender.forward(s_dict))

    def w_plus_dict_to_image(self, w_plus_dict, truncation=1):                                             # :167-171
        return self.s_dict_to_image({k: self.w_plus_to_s(v, truncation=truncation) for k, v in w_plus_dict.items()})

    def z_dict_to_image(self, z_dict):                                                                     # :173-177
        return self.s_dict_to_image({k: self.z_to_s(v) for k, v in z_dict.items()})
