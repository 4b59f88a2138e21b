"""Attack entry points with the reference's names and signatures (code/attack/attack_main2.py, file:line per function).
Each one is a thin host wrapper over the CUDA schedules in engine.py; no arithmetic happens in torch.
Globals `vgg`, `device`, `param_file`, `is_cars` that the reference reads at module scope (:315-352) are explicit kwargs."""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

from .. import lib
from ..engine import ReconAttackEngine, ReconLossCfg

_ENGINES: Dict[tuple, ReconAttackEngine] = {}


def get_latents(net, x, is_cars=False):                                                   # attack_main2.py:137-146
    codes = net.encoder(x)
    if net.opts.start_from_latent_avg:
        if codes.ndim == 2:
            codes = codes + net.latent_avg.repeat(codes.shape[0], 1, 1)[:, 0, :]
        else:
            codes = codes + net.latent_avg.repeat(codes.shape[0], 1, 1)
    if codes.shape[1] == 18 and is_cars:
        codes = codes[:, :16, :]
    return codes


def patch_white_box(inputs, mask, adv_patch):                                             # attack_main2.py:413-433
    """(1-mask)*x + mask*patch, clamped to each clean image's own [min, max] -- one fused kernel (zero-step patch update)."""
    n, _, s, _ = inputs.shape
    dev = inputs.device
    x0 = inputs.contiguous().float()
    out = torch.empty_like(x0)
    patch = adv_patch.to(dev).float().expand_as(x0).contiguous().clone()
    m = mask.to(dev).float().expand_as(x0).contiguous()
    lo, hi = torch.empty(n, device=dev), torch.empty(n, device=dev)
    lib.minmax_per_sample(x0, lo, hi)
    zero = torch.zeros(n, 3, s, s, device=dev)
    lib.attack_update_patch(out, x0, patch, m, zero, 0.0, 1.0, False, lo, hi, 1.0, None, 1)
    return out


def fusion(dataset_name, all_latents, drawer, save_dir="/", file_name="filename", feature_idx=-1):   # attack_main2.py:521-581
    """spatial fusion of N inputs' W+ codes through the StyleSpace blender + the N single reconstructions."""
    lat = [all_latents[i][None] for i in range(all_latents.shape[0])]
    if dataset_name == "ffhq":
        kw = dict(hair=lat[1], eyes=lat[2] if len(lat) > 2 else None, background=lat[3] if len(lat) > 3 else None,
                  mouth=lat[4] if len(lat) > 4 else None)
    elif dataset_name == "car":
        kw = dict(wheels=lat[1], bg_top=lat[2] if len(lat) > 2 else None, bg_bottom=lat[3] if len(lat) > 3 else None)
    else:
        kw = dict(bg_top=lat[1], bg_bottom=lat[2] if len(lat) > 2 else None)
    I_fused, feats = drawer.generate_img(lat[0], latents_type="w", **kw)
    singles, inner = [], [feats[feature_idx]]
    for l in lat:
        img, f = drawer.generate_img(l, latents_type="w")
        singles.append(img)
        inner.append(f[feature_idx])
    return I_fused, torch.cat(singles, 0), inner


def interpolation(drawer, all_latents, feature_idx=-1):                                    # interpolation.py:658-669
    """arithmetic fusion: mean of the inputs' W+ codes -> generator, plus the N single reconstructions."""
    avg_latent = torch.mean(all_latents, dim=0, keepdim=True)
    I_fused, feats = drawer.generate_img(avg_latent, latents_type="w")
    singles = [drawer.generate_img(all_latents[i][None], latents_type="w")[0] for i in range(all_latents.shape[0])]
    return I_fused, torch.cat(singles, 0), feats[feature_idx]


def _recon_engine(Model, vgg, batch, device, loss: ReconLossCfg) -> ReconAttackEngine:
    key = (id(Model), id(vgg), batch, str(device), tuple(vars(loss).values()))
    if key not in _ENGINES:
        dec, enc = Model.decoder, Model.encoder
        _ENGINES[key] = ReconAttackEngine(dec.spec, dec.params, enc.spec, enc.params, vgg.sd, batch=batch, device=str(device), loss=loss,
                                          vgg_res=enc.spec.in_res, vgg_width_div=vgg.width_div)
    return _ENGINES[key]


def optimize_vgg(i_th_img, Model, vgg, img, img_target, run_dir, device, file_name, args, n_iters=1000,
                 loss: Optional[ReconLossCfg] = None):                                      # attack_main2.py:584-671
    """Adam on the pixels of `img` ([-1,1]) against the encoder->decoder reconstruction, loss menu of :649.
    Losses stay on the device; if args.save_img the per-5-iteration lines of :657-666 are written once after the loop."""
    loss = loss or ReconLossCfg()
    eng = _recon_engine(Model, vgg, img.shape[0], device, loss)
    eng.set_inputs(img.to(device).float().contiguous(), img_target.to(device).float().expand_as(img).contiguous())
    log = torch.zeros(n_iters, img.shape[0], device=eng.dev)
    for it in range(n_iters):
        l, _, _ = eng.forward_backward()
        log[it].copy_(l)
        eng.adam_step(it + 1, float(args.lr))
    eng.check()
    if getattr(args, "save_img", False) and run_dir:
        os.makedirs(run_dir, exist_ok=True)
        with open(os.path.join(run_dir, "optimize_output.txt"), "a") as f:
            for it in range(5, n_iters, 5):
                f.write("%dth img iter: %d loss:%.5f\n" % (i_th_img, it, float(log[it].sum())))
    return eng.x.detach().clone()


def white_box(inputs, target_img, drawer, net, vgg, args, n_iters, is_cars=False, save_dir=None, device=None,
              loss: Optional[ReconLossCfg] = None):                                          # attack_main2.py:465-498
    """The reference loops over the batch at batch size 1 (:472-483); here the whole batch is one launch sequence."""
    device = device or inputs.device
    return optimize_vgg(0, net, vgg, inputs.clone(), target_img, save_dir, device, "adv", args, n_iters=n_iters, loss=loss)


def main_optimize(inputs, drawer, net, target_img, args, device, iter_dict, train_dataloader=None, save_dir=None, vgg=None,
                  is_cars=False):                                                           # attack_main2.py:299-404
    """dispatch on args.adversarial; only the gradient attacks of SURVEY 8a are served (the one-shot corruptions are out of scope)."""
    out = []
    n_iters = iter_dict[net.decoder.size]
    for kind in (args.adversarial if isinstance(args.adversarial, (list, tuple)) else [args.adversarial]):
        if kind in ("white_box", "white_box_target"):
            out.append(white_box(inputs, target_img, drawer, net, vgg, args, n_iters, is_cars, save_dir, device))
        elif kind == "patch_white_box":
            from .patch import adversarial_patch as patch
            p, m = patch.main(drawer, net, vgg, train_dataloader, device, save_dir, args, target_img)
            out.append(patch_white_box(inputs, m, p))
        else:
            raise NotImplementedError(f"{kind}: not a gradient attack (out of scope, SURVEY section 2)")
    return out
