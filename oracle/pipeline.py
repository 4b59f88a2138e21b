"""ORACLE (test infrastructure only -- the product path never imports this).

CPU/PyTorch fp32 restatement of the attack hot path: encoder stand-in -> pair fusion -> StyleGAN2
synthesis -> VGG feature loss -> autograd -> projected sign-gradient step.

What is pinned / unpinned:
  * VGG: pinned to the reference's code/vgg.py through tests/golden (see oracle/vgg_ref.py).
  * PGD/FGSM step: follows the only sign-gradient step in the reference, the commented torchattacks
    restatement at code/attack/interpolation.py:54-96 (random start :74-76, step :92, projection :93,
    clamp :94).
  * patch update: code/attack/patch/adversarial_patch.py:106,131-138 (raw-gradient step, re-mask,
    clamp to the clean image's [min,max]); mask apply: code/attack/attack_main2.py:413-433.
  * Adam-on-pixels: code/attack/attack_main2.py:606,614-653 (torch.optim.Adam defaults).
  * L2 variant, encoder stand-in and pair fusion: NOT in the reference;
    builder-defined here -> PARITY UNPINNED for those pieces.
  * N-way fusion (`fusion="hierarchy"`, OraclePipeline.hier): the part assignment is the reference's
    (generate_img's swap lists, code/style_fusion_simple.py:84-104; roles of fusion(),
    code/attack/attack_main2.py:526-566); the blender itself is the gate-chain stand-in of
    oracle/fusion_ref.py (un-vendored FusionNets) -> PARITY UNPINNED for the gates.
  * `encoder_module`: any torch module in `net.encoder`'s place (code/utils/model_utils.py:24).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from sfattack.params import EncSpec, GenSpec
from . import stylegan2 as sg
from .vgg_ref import vgg_forward


# ---------------------------------------------------------------------------------------------
def encoder_forward(P: Dict[str, torch.Tensor], spec: EncSpec, x: torch.Tensor) -> torch.Tensor:
    """x (N,3,256,256) in [-1,1] -> codes (N, n_latent, 512) WITHOUT latent_avg
    (the reference adds it in get_latents, code/attack/attack_main2.py:137-146)."""
    n = len(spec.widths)
    for i in range(n):
        x = F.relu(F.conv2d(x, P[f"convs.{i}.weight"], P[f"convs.{i}.bias"], padding=1))
        if i < n - 1:
            x = F.max_pool2d(x, 2, 2)
    x = x.mean(dim=(2, 3))
    y = F.linear(x, P["head.weight"], P["head.bias"])
    return y.view(x.shape[0], spec.n_latent, spec.style_dim)


def get_latents(P, spec: EncSpec, x: torch.Tensor) -> torch.Tensor:
    return encoder_forward(P, spec, x) + P["latent_avg"][None]


def fuse_arithmetic(w_a, w_b):
    """mean of the inputs' W+ codes (code/attack/interpolation.py:661)."""
    return 0.5 * (w_a + w_b)


def fuse_spatial(FP, s_a: torch.Tensor, s_b: torch.Tensor) -> torch.Tensor:
    """Stand-in for base_blender.forward(s_dict) (code/style_fusion_simple.py:164), on concatenated S vectors."""
    q = torch.sigmoid(FP["alpha"] * s_a + FP["beta"] * s_b + FP["c"])
    return q * s_a + (1.0 - q) * s_b


def cat_styles(styles: List[torch.Tensor]) -> torch.Tensor:
    return torch.cat(styles, dim=1)


def split_styles(spec: GenSpec, s: torch.Tensor) -> List[torch.Tensor]:
    return [s[:, l.s_off:l.s_off + l.cin] for l in spec.layers]


# ---------------------------------------------------------------------------------------------
@dataclass
class LossCfg:
    c_pix: float = 1.0        # weight of per-sample mean-squared error on the fused image (full res)
    c_feat: float = 1.0       # weight of the sum over the 4 VGG taps of per-sample MSE (pooled to 256)
    c_reg: float = 0.0        # weight of the VGG perceptual regulariser on the adversarial INPUTS (config 5)


def per_sample_mse(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """nn.MSELoss(reduction='mean') applied per sample (the reference runs at batch 1,
    code/attack/attack_main2.py:475,605, so 'mean' never mixes samples)."""
    return ((a - b) ** 2).flatten(1).mean(1)


class OraclePipeline:
    """Everything the hot loop evaluates, as plain autograd-able torch code."""

    def __init__(self, gspec: GenSpec, GP, espec: EncSpec, EP, vgg_sd, FP=None, fusion: str = "arithmetic",
                 vgg_res: int = 256, encoder_module=None, latent_avg=None):
        self.gspec, self.GP, self.espec, self.EP, self.vgg_sd, self.FP = gspec, GP, espec, EP, vgg_sd, FP
        self.fusion = fusion
        self.vgg_res = vgg_res
        # N-way hierarchy fusion: hier = dict(parts=[...], source=[input index per part], gates={part: {alpha,beta,c}})
        self.hier = None
        # any torch module in the encoder's place (`net.encoder` of the reference, code/utils/model_utils.py:24) + latent_avg
        self.encoder_module, self.latent_avg = encoder_module, latent_avg

    # pixels in [0,1] -> [-1,1], box-pool to the encoder resolution (attack_main2.py:590-591,619)
    def pool_in(self, x01: torch.Tensor) -> torch.Tensor:
        k = x01.shape[-1] // self.espec.in_res
        x = 2.0 * x01 - 1.0
        return F.avg_pool2d(x, k, k) if k > 1 else x

    def latents(self, x01):
        if self.encoder_module is not None:                     # get_latents, code/attack/attack_main2.py:137-146
            codes = self.encoder_module(self.pool_in(x01))
            if codes.ndim == 2:
                codes = codes.unsqueeze(1).expand(-1, self.espec.n_latent, -1)
            return codes + self.latent_avg[None] if self.latent_avg is not None else codes
        return get_latents(self.EP, self.espec, self.pool_in(x01))

    def styles(self, *ws):
        if self.fusion == "arithmetic":      # mean of the N inputs' W+ codes (interpolation.py:661)
            return sg.styles_from_wplus(self.GP, self.gspec, sum(ws) / float(len(ws)))
        ss = [cat_styles(sg.styles_from_wplus(self.GP, self.gspec, w)) for w in ws]
        if self.fusion == "hierarchy":
            # generate_img's s_dict (code/style_fusion_simple.py:84-104: every part holds the style vector of the input assigned to
            # it) -> base_blender.forward(s_dict) (:164), restated as the chain of per-part gates of oracle/fusion_ref.py
            from .fusion_ref import blend
            h = self.hier
            return split_styles(self.gspec, blend(h["parts"], h["gates"], {p: ss[k] for p, k in zip(h["parts"], h["source"])}))
        return split_styles(self.gspec, fuse_spatial(self.FP, ss[0], ss[1]))

    def fused(self, *xs01):
        """fused image (B,3,S,S), roughly [-1,1], of the N inputs (each (B,3,S,S) in [0,1])"""
        return sg.synthesis_from_styles(self.GP, self.gspec, self.styles(*[self.latents(x) for x in xs01]))

    def pool_vgg(self, img):
        k = img.shape[-1] // self.vgg_res
        return F.avg_pool2d(img, k, k) if k > 1 else img

    def features(self, img):
        return vgg_forward(self.vgg_sd, self.pool_vgg(img))

    def reference_of(self, img) -> Tuple[torch.Tensor, tuple]:
        with torch.no_grad():
            return img.detach(), tuple(f.detach() for f in self.features(img))

    def loss_n(self, xs01, ref_img, ref_feats, cfg: LossCfg, reg_refs=None):
        """N-input form of loss(): per-sample loss (B,) and the fused image."""
        img = self.fused(*xs01)
        L = cfg.c_pix * per_sample_mse(img, ref_img)
        if cfg.c_feat != 0.0:
            for f, r in zip(self.features(img), ref_feats):
                L = L + cfg.c_feat * per_sample_mse(f, r)
        if cfg.c_reg != 0.0:
            for x, refs in zip(xs01, reg_refs):
                for f, r in zip(vgg_forward(self.vgg_sd, self.pool_in(x)), refs):
                    L = L - cfg.c_reg * per_sample_mse(f, r)
        return L, img

    def input_grads_n(self, xs01, ref_img, ref_feats, cfg: LossCfg, reg_refs=None):
        xs = [x.detach().clone().requires_grad_(True) for x in xs01]
        L, img = self.loss_n(xs, ref_img, ref_feats, cfg, reg_refs=reg_refs)
        return L.detach(), img.detach(), list(torch.autograd.grad(L.sum(), xs))

    def loss(self, xa01, xb01, ref_img, ref_feats, cfg: LossCfg, xa_clean=None, xb_clean=None, reg_refs=None):
        """per-sample loss (B,) and the fused image."""
        img = self.fused(xa01, xb01)
        L = cfg.c_pix * per_sample_mse(img, ref_img)
        if cfg.c_feat != 0.0:
            for f, r in zip(self.features(img), ref_feats):
                L = L + cfg.c_feat * per_sample_mse(f, r)
        if cfg.c_reg != 0.0:
            for x, refs in ((xa01, reg_refs[0]), (xb01, reg_refs[1])):
                for f, r in zip(vgg_forward(self.vgg_sd, self.pool_in(x)), refs):
                    L = L - cfg.c_reg * per_sample_mse(f, r)
        return L, img

    def input_grads(self, xa01, xb01, ref_img, ref_feats, cfg: LossCfg, reg_refs=None):
        xa = xa01.detach().clone().requires_grad_(True)
        xb = xb01.detach().clone().requires_grad_(True)
        L, img = self.loss(xa, xb, ref_img, ref_feats, cfg, reg_refs=reg_refs)
        ga, gb = torch.autograd.grad(L.sum(), [xa, xb])
        return L.detach(), img.detach(), ga, gb


# ---------------------------------------------------------------------------------------------
# update rules (each operates on one tensor of images; the pair is just the concatenation)
def linf_step(x_adv, x_clean, g, alpha, eps, direction=1.0, lo=0.0, hi=1.0):
    """interpolation.py:92-94"""
    x = x_adv + direction * alpha * torch.sign(g)
    delta = torch.clamp(x - x_clean, min=-eps, max=eps)
    return torch.clamp(x_clean + delta, min=lo, max=hi)


def l2_step(x_adv, x_clean, g, alpha, eps, direction=1.0, lo=0.0, hi=1.0):
    """builder-defined: normalised-gradient step, projection onto the eps L2 ball, clamp."""
    gn = g.flatten(1).norm(dim=1).clamp_min(1e-12).view(-1, 1, 1, 1)
    x = x_adv + direction * alpha * g / gn
    delta = x - x_clean
    dn = delta.flatten(1).norm(dim=1).view(-1, 1, 1, 1)
    delta = delta * torch.clamp(eps / dn.clamp_min(1e-12), max=1.0)
    return torch.clamp(x_clean + delta, min=lo, max=hi)


def patch_apply(x_clean, mask, patch, lo=None, hi=None):
    """(1-mask)*x + mask*patch, clamp to the clean image's own range
    (attack_main2.py:416-418; adversarial_patch.py:137-138)."""
    adv = (1 - mask) * x_clean + mask * patch
    if lo is None:
        lo = x_clean.flatten(1).min(1)[0].view(-1, 1, 1, 1)
        hi = x_clean.flatten(1).max(1)[0].view(-1, 1, 1, 1)
    return torch.maximum(torch.minimum(adv, hi), lo)


def patch_step(patch, g, lr, sign=False, direction=-1.0):
    """adversarial_patch.py:133 `patch -= adv_grad` (direction=-1, lr=1, raw gradient)."""
    return patch + direction * lr * (torch.sign(g) if sign else g)


def adam_step(x, g, m, v, t, lr, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam defaults as used at attack_main2.py:606 (minimises)."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    mh = m / (1 - b1 ** t)
    vh = v / (1 - b2 ** t)
    return x - lr * mh / (vh.sqrt() + eps), m, v


# ---------------------------------------------------------------------------------------------
@dataclass
class AttackCfg:
    kind: str = "linf"            # linf | l2 | patch | adam
    steps: int = 10
    eps: float = 8.0 / 255.0
    alpha: float = 2.0 / 255.0
    random_start: bool = True
    targeted: bool = False
    patch_sign: bool = False
    lr: float = 1.0
    loss: LossCfg = None

    def __post_init__(self):
        if self.loss is None:
            self.loss = LossCfg()


def run_attack(pipe: OraclePipeline, xa, xb, cfg: AttackCfg, start_noise=None, target=None, mask=None,
               patch0=None, record=None):
    """The whole iterative attack on a batch of pairs.  xa, xb in [0,1], (B,3,S,S).

    untargeted: reference = clean fusion, ascend.   targeted: reference = fusion of (target,target), descend.
    start_noise (2,B,3,S,S) in [-1,1] scales to U(-eps,eps) (pre-generated so both sides share bits)."""
    direction = -1.0 if cfg.targeted else 1.0
    # N-way form: xa is a list of the N inputs and xb is None; start_noise is then (N,B,3,S,S)
    inputs = list(xa) if xb is None else [xa, xb]
    NI, B = len(inputs), inputs[0].shape[0]
    split = lambda T: [T[k * B:(k + 1) * B] for k in range(NI)]
    with torch.no_grad():
        ref_src = pipe.fused(*target) if cfg.targeted else pipe.fused(*inputs)
    ref_img, ref_feats = pipe.reference_of(ref_src)
    reg_refs = None
    if cfg.loss.c_reg != 0.0:
        with torch.no_grad():
            reg_refs = tuple(tuple(vgg_forward(pipe.vgg_sd, pipe.pool_in(x))) for x in inputs)
    X0 = torch.cat(inputs, 0)
    if cfg.kind == "patch":
        patch = patch0.clone()
        lo = X0.flatten(1).min(1)[0].view(-1, 1, 1, 1)
        hi = X0.flatten(1).max(1)[0].view(-1, 1, 1, 1)
        X = patch_apply(X0, mask, patch, lo, hi)
    else:
        X = X0.clone()
        if cfg.random_start and start_noise is not None and cfg.kind in ("linf", "l2"):
            X = torch.clamp(X0 + cfg.eps * start_noise.reshape(X0.shape), 0.0, 1.0)   # interpolation.py:74-76
    m = torch.zeros_like(X); v = torch.zeros_like(X)
    losses = []
    for it in range(cfg.steps):
        L, img, gl = pipe.input_grads_n(split(X), ref_img, ref_feats, cfg.loss, reg_refs=reg_refs)
        g = torch.cat(gl, 0)
        losses.append(L)
        if record is not None:
            record.append(dict(loss=L.clone(), grad=g.clone(), img=img.clone(), x=X.clone()))
        if cfg.kind == "linf":
            X = linf_step(X, X0, g, cfg.alpha, cfg.eps, direction)
        elif cfg.kind == "l2":
            X = l2_step(X, X0, g, cfg.alpha, cfg.eps, direction)
        elif cfg.kind == "patch":
            patch = patch_step(patch, g, cfg.lr, cfg.patch_sign, direction)
            X = patch_apply(X0, mask, patch, lo, hi)
        elif cfg.kind == "adam":
            X, m, v = adam_step(X, -direction * g, m, v, it + 1, cfg.lr)
        else:
            raise ValueError(cfg.kind)
    with torch.no_grad():
        final = pipe.fused(*split(X))
    out = dict(x_adv=X, fused_adv=final, fused_ref=ref_img, losses=torch.stack(losses, 0) if losses else None)
    if cfg.kind == "patch":
        out["patch"] = patch
    return out


def success_metrics(fused_adv, fused_clean, pipe: OraclePipeline):
    """cal_result's 'or_f_ad_f' pixel MSE and the summed VGG-tap MSE (interpolation.py:1076-1091, :848-855);
    SSIM is omitted (skimage absent)."""
    with torch.no_grad():
        mse = per_sample_mse(fused_adv, fused_clean)
        fa, fc = pipe.features(fused_adv), pipe.features(fused_clean)
        vg = sum(per_sample_mse(a, c) for a, c in zip(fa, fc))
    return mse, vg


# ---------------------------------------------------------------------------------------------
@dataclass
class ReconLossCfg:
    """weights of optimize_vgg's `inversion_loss` (attack_main2.py:649; variants interpolation.py:818, inter_copy.py:658)"""
    w_latent_target: float = 10.0
    w_latent_org: float = -1.0
    w_img_rec_target: float = 1.0
    w_img_org: float = 20.0
    w_lpips_img: float = 1.0
    w_lpips_rec: float = 0.0
    lpips_rec_ref: str = "target"


def optimize_vgg_oracle(gspec, GP, espec, EP, vgg_sd, img, img_target, cfg: ReconLossCfg, n_iters: int, lr: float,
                        record=None, encoder_module=None):
    """Restatement of the reference's live loop (attack_main2.py:584-671): Adam on the pixels of `img` ([-1,1]) against the
    encoder->decoder reconstruction; per-sample means (the reference runs batch 1).  No file I/O."""
    k = gspec.size // espec.in_res
    pool = (lambda t: F.avg_pool2d(t, k, k)) if k > 1 else (lambda t: t)
    if encoder_module is not None:           # `Model.encoder` as any torch module (the reference's e4e encoder, model_utils.py:24)
        def encoder_forward(_EP, _spec, t):  # noqa: F811  (shadows the stand-in for this call)
            c = encoder_module(t)
            return c.unsqueeze(1).expand(-1, espec.n_latent, -1) if c.ndim == 2 else c
    else:
        encoder_forward = globals()["encoder_forward"]
    img_org = img.clone().detach()
    with torch.no_grad():                                                    # :597-603
        latent_target = encoder_forward(EP, espec, pool(img_target))
        latent_org = encoder_forward(EP, espec, pool(img_org))
        f_target = vgg_forward(vgg_sd, pool(img_target))
        f_org = vgg_forward(vgg_sd, pool(img_org))
    x = img.clone().detach().requires_grad_(True)
    opt = torch.optim.Adam([x], lr=lr)                                       # :606
    for it in range(n_iters):
        opt.zero_grad()
        lat = encoder_forward(EP, espec, pool(x))                            # :619,622 (evaluated once; same value)
        img_rec = sg.synthesis_from_styles(GP, gspec, sg.styles_from_wplus(GP, gspec, lat))
        L = cfg.w_latent_target * per_sample_mse(lat, latent_target) + cfg.w_latent_org * per_sample_mse(lat, latent_org)
        L = L + cfg.w_img_rec_target * per_sample_mse(img_rec, img_target) + cfg.w_img_org * per_sample_mse(x, img_org)
        if cfg.w_lpips_img != 0.0:
            L = L + cfg.w_lpips_img * sum(per_sample_mse(a, b) for a, b in zip(vgg_forward(vgg_sd, pool(x)), f_org))
        if cfg.w_lpips_rec != 0.0:
            refs = f_target if cfg.lpips_rec_ref == "target" else f_org
            L = L + cfg.w_lpips_rec * sum(per_sample_mse(a, b) for a, b in zip(vgg_forward(vgg_sd, pool(img_rec)), refs))
        L.sum().backward()
        if record is not None:
            record.append(dict(loss=L.detach().clone(), grad=x.grad.detach().clone(), img_rec=img_rec.detach().clone()))
        opt.step()
    return x.detach()


# ---------------------------------------------------------------------------------------------
def patch_attack_oracle(gspec, GP, espec, EP, vgg_sd, img, patch, mask, target_img, max_count: int, cfg: ReconLossCfg = None,
                        record=None):
    """Restatement of patch.attack (code/attack/patch/adversarial_patch.py:94-160): raw-gradient descent on the patch through
    encoder -> generator -> VGG, `Loss = -l_latent_org` by default (:126), re-mask and clamp to the clean range every step
    (:131-138).  img/patch/mask (b,3,S,S) in [-1,1]; per-sample means (the reference runs batch 1).  -> (adv_x, mask, patch, rec)."""
    cfg = cfg or ReconLossCfg(0.0, -1.0, 0.0, 0.0, 0.0, 0.0)
    k = gspec.size // espec.in_res
    pool = (lambda t: F.avg_pool2d(t, k, k)) if k > 1 else (lambda t: t)
    with torch.no_grad():                                                                   # :98-104
        latent_org = encoder_forward(EP, espec, pool(img))
        latent_target = encoder_forward(EP, espec, pool(target_img))
        f_target = vgg_forward(vgg_sd, pool(target_img))
    lo = img.flatten(1).min(1)[0].view(-1, 1, 1, 1)
    hi = img.flatten(1).max(1)[0].view(-1, 1, 1, 1)
    patch = patch.clone()
    adv_x = (1 - mask) * img + mask * patch                                                 # :106 (no clamp before the first step)
    adv_x = torch.maximum(torch.minimum(adv_x, hi), lo)      # the CUDA path clamps here too; a no-op for patches inside the range
    rec = None
    for count in range(max_count):                                                          # :111-158
        x = adv_x.detach().clone().requires_grad_(True)
        lat = encoder_forward(EP, espec, pool(x))
        rec = sg.synthesis_from_styles(GP, gspec, sg.styles_from_wplus(GP, gspec, lat))
        L = cfg.w_latent_target * per_sample_mse(lat, latent_target) + cfg.w_latent_org * per_sample_mse(lat, latent_org)
        L = L + cfg.w_img_rec_target * per_sample_mse(rec, target_img)
        if cfg.w_lpips_rec != 0.0:
            L = L + cfg.w_lpips_rec * sum(per_sample_mse(a, b) for a, b in zip(vgg_forward(vgg_sd, pool(rec)), f_target))
        (g,) = torch.autograd.grad(L.sum(), x)
        if record is not None:
            record.append(dict(loss=L.detach().clone(), grad=g.clone()))
        patch = patch - g                                                                   # :133
        adv_x = (1 - mask) * img + mask * patch                                             # :137
        adv_x = torch.maximum(torch.minimum(adv_x, hi), lo)                                 # :138
    return adv_x.detach(), mask, patch, rec.detach()


def universal_patch_oracle(gspec, GP, espec, EP, vgg_sd, imgs, patch, mask, target_img, max_count: int, lr: float = 1.0):
    """batch form of the above with ONE shared patch: patch -= lr * sum_n mask * dL_n/dx_n."""
    cfg = ReconLossCfg(0.0, -1.0, 0.0, 0.0, 0.0, 0.0)
    k = gspec.size // espec.in_res
    pool = (lambda t: F.avg_pool2d(t, k, k)) if k > 1 else (lambda t: t)
    with torch.no_grad():
        latent_org = encoder_forward(EP, espec, pool(imgs))
    lo = imgs.flatten(1).min(1)[0].view(-1, 1, 1, 1)
    hi = imgs.flatten(1).max(1)[0].view(-1, 1, 1, 1)
    patch = patch.clone()
    apply = lambda p: torch.maximum(torch.minimum((1 - mask) * imgs + mask * p, hi), lo)
    adv_x = apply(patch)
    losses = []
    for _ in range(max_count):
        x = adv_x.detach().clone().requires_grad_(True)
        lat = encoder_forward(EP, espec, pool(x))
        L = cfg.w_latent_org * per_sample_mse(lat, latent_org)
        (g,) = torch.autograd.grad(L.sum(), x)
        losses.append(L.detach())
        patch = patch - lr * (mask * g).sum(0, keepdim=True)
        adv_x = apply(patch)
    return patch, adv_x.detach(), torch.stack(losses)
