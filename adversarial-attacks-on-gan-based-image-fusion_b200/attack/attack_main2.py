"""Attack entry points with the reference's names and signatures (code/attack/attack_main2.py, file:line per function).
Each one is a thin host wrapper over the CUDA schedules in engine.py; no arithmetic happens in torch.
Globals `vgg`, `device`, `param_file`, `is_cars` that the reference reads at module scope (:315-352) are explicit kwargs."""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Optional

import torch

from .. import lib
from ..engine import ReconAttackEngine, ReconLossCfg

# Engines own ~0.6 GB of buffers per sample at 1024^2, so the cache is small and LRU.  The key carries everything an engine's
# buffers and plans depend on: the model objects AND their weight versions (load_state_dict / .to() bump them), the batch, the
# device, the loss menu and the library's process-global storage dtype / conv math.
_ENGINES: "OrderedDict[tuple, ReconAttackEngine]" = OrderedDict()
_MAX_ENGINES = 4

# the three weightings of `inversion_loss` in the reference's three copies of optimize_vgg
LOSS_MENUS = {
    "attack_main2": ReconLossCfg(10.0, -1.0, 1.0, 20.0, 1.0, 0.0, "target"),    # attack_main2.py:649
    "interpolation": ReconLossCfg(10.0, -1.0, 1.0, 10.0, 1.0, 0.1, "target"),   # interpolation.py:818
    "inter_copy": ReconLossCfg(10.0, -1.0, 10.0, 5.0, 0.0, 0.5, "org"),         # inter_copy.py:658
}


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
    with torch.cuda.device(dev):
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
    """Spatial fusion of the N inputs' W+ codes + the N single reconstructions -> (I_fused, I_all, cat of the singles' inner
    features).  Latent order and roles are the reference's: ffhq [mouth, background, hair, eyes, global] (:526), car
    [wheel, bg_top, bg_bottom, body] (:547), church [bg_top, bg_bottom, body] (:566); the LAST one is the base; dataset names are
    matched by substring as upstream ('ffhq' in dataset_name)."""
    lat = list(all_latents.unsqueeze(1))
    singles = None
    if "ffhq" in dataset_name:
        z_mouth, z_background, z_hair, z_eyes, z_global = lat
        I_fused, _ = drawer.generate_img(z_global, hair=z_hair, eyes=z_eyes, background=z_background, mouth=z_mouth, latents_type="w")
        singles = [z_mouth, z_background, z_hair, z_eyes, z_global]                                 # :533-543
    if "car" in dataset_name:
        z_wheel, z_bg_top, z_bg_bottom, z_body = lat
        I_fused, _ = drawer.generate_img(z_body, wheels=z_wheel, bg_top=z_bg_top, bg_bottom=z_bg_bottom, latents_type="w")
        singles = [z_body, z_wheel, z_bg_top, z_bg_bottom]                                          # :553-561 (body first)
    if "church" in dataset_name:
        z_bg_top, z_bg_bottom, z_body = lat
        I_fused, _ = drawer.generate_img(z_body, bg_top=z_bg_top, bg_bottom=z_bg_bottom, latents_type="w")
        singles = [z_body, z_bg_top, z_bg_bottom]                                                   # :571-577
    if singles is None:
        raise ValueError(f"unknown dataset {dataset_name!r} (expected a name containing ffhq / car / church)")
    imgs, feats = [], []
    for z in singles:
        img, f = drawer.generate_img(z, latents_type="w")
        imgs.append(img)
        feats.append(f[feature_idx])
    return I_fused, torch.cat(imgs, 0), torch.cat(feats, 0)


def interpolation(drawer, all_latents, feature_idx=-1):                                    # interpolation.py:658-669
    """arithmetic fusion: mean of the inputs' W+ codes -> generator, plus the N single reconstructions and their inner features."""
    avg_latent = torch.mean(all_latents, dim=0, keepdim=True)
    I_fused, _ = drawer.generate_img(avg_latent, latents_type="w")
    imgs, feats = [], []
    for i in range(all_latents.size(0)):
        img, f = drawer.generate_img(all_latents[i].unsqueeze(0), latents_type="w")
        imgs.append(img)
        feats.append(f[feature_idx])
    return I_fused, torch.cat(imgs, 0), torch.cat(feats, 0)


def _version(obj) -> int:
    return int(getattr(obj, "version", 0))


def _recon_engine(Model, vgg, batch, device, loss: ReconLossCfg) -> ReconAttackEngine:
    dec, enc = Model.decoder, Model.encoder
    key = (id(dec), _version(dec), id(enc), _version(enc), id(vgg), _version(vgg), batch, str(torch.device(device)),
           tuple(vars(loss).values()), lib.mode_key())
    eng = _ENGINES.get(key)
    if eng is None and isinstance(enc, torch.nn.Module):
        # `Model.encoder` is a real torch encoder (the reference's e4e network, code/utils/model_utils.py:24): it stays a torch
        # module on the gradient path; the reference feeds it the image pooled to 256x256 (attack_main2.py:590-591,619)
        from ..params import EncSpec
        es = EncSpec(n_latent=dec.spec.n_latent, style_dim=dec.spec.style_dim, in_res=min(256, dec.spec.size))
        eng = ReconAttackEngine(dec.spec, dec.params, es, None, vgg.sd, batch=batch, device=str(device), loss=loss,
                                vgg_res=es.in_res, vgg_width_div=vgg.width_div, encoder_module=enc)
    elif eng is None:
        eng = ReconAttackEngine(dec.spec, dec.params, enc.spec, enc.params, vgg.sd, batch=batch, device=str(device), loss=loss,
                                vgg_res=enc.spec.in_res, vgg_width_div=vgg.width_div)
        eng._owners = (dec, enc, vgg)        # the ids in the key stay valid as long as the engine lives
        _ENGINES[key] = eng
        while len(_ENGINES) > _MAX_ENGINES:
            _ENGINES.popitem(last=False)
    else:
        _ENGINES.move_to_end(key)
    return eng


def optimize_vgg(i_th_img, Model, vgg, img, img_target, run_dir, device, file_name, args, n_iters=1000,
                 loss: Optional[ReconLossCfg] = None):                                      # attack_main2.py:584-671
    """Adam on the pixels of `img` ([-1,1]) against the encoder->decoder reconstruction, loss menu of :649 (other menus:
    LOSS_MENUS).  The loss terms stay on the device; if args.save_img, the lines the reference prints and appends to
    optimize_output.txt every 5 iterations (:657-666: l_latent_target, l_latent_org, l_img_org) are written once after the loop."""
    loss = loss or LOSS_MENUS["attack_main2"]
    eng = _recon_engine(Model, vgg, img.shape[0], device, loss)
    eng.set_inputs(img.to(device).float().contiguous(), img_target.to(device).float().expand_as(img).contiguous())
    log = torch.zeros(n_iters, 3, img.shape[0], device=eng.dev)
    for it in range(n_iters):
        eng.forward_backward()
        log[it].copy_(eng.terms)
        eng.adam_step(it + 1, float(args.lr))
    eng.check()
    if getattr(args, "save_img", False) and run_dir:
        os.makedirs(run_dir, exist_ok=True)
        rows = log.cpu()
        with open(os.path.join(run_dir, "optimize_output.txt"), "a") as f:
            for it in range(5, n_iters, 5):
                for b in range(rows.shape[2]):
                    f.write("%dth img iter: %d l_latent_target:%.5f;   l_latent_org:%.5f;     l_img_org:%f \n" % (
                        i_th_img + b, it, rows[it, 0, b], rows[it, 1, b], rows[it, 2, b]))
    return eng.x.detach().clone()


def white_box(inputs, target_img, drawer, net, vgg, args, n_iters, is_cars=False, save_dir=None, device=None,
              loss: Optional[ReconLossCfg] = None):                                          # attack_main2.py:465-498
    """Attack the images listed in args.which_adv (all of them if the list is empty, :469-470) and pass the others through
    unchanged (:485-486).  The reference loops over the selected images at batch size 1 (:472-483); here they are one batch.
    target_img: (1,3,S,S) shared target ('white_box_target', :474) or one target per input ('white_box_patch', :479)."""
    device = device or inputs.device
    which = list(getattr(args, "which_adv", None) or [])
    if len(which) == 0:
        which = list(range(inputs.size(0)))
        args.which_adv = which
    sel = [i for i in range(inputs.size(0)) if i in which]
    out = inputs.clone()
    if sel:
        idx = torch.tensor(sel, device=inputs.device)
        tgt = target_img if target_img.size(0) == 1 else target_img.index_select(0, idx.to(target_img.device))
        adv = optimize_vgg(sel[0], net, vgg, inputs.index_select(0, idx), tgt, save_dir, device, "optimize", args, n_iters=n_iters,
                           loss=loss)
        out[idx] = adv.to(out.device, out.dtype)
    return out


def main_optimize(inputs, drawer, net, target_img, args, device, iter_dict, train_dataloader=None, save_dir=None, vgg=None,
                  is_cars=False):                                                           # attack_main2.py:299-404
    """dispatch on args.adversarial; only the gradient attacks of SURVEY 8a are served (the one-shot corruptions are out of scope).
    Returns a list with one tensor of adversarial inputs per attack, as the reference does (:306,404)."""
    out = []
    n_iters = iter_dict[net.decoder.size]
    for kind in (args.adversarial if isinstance(args.adversarial, (list, tuple)) else [args.adversarial]):
        if kind in ("white_box", "white_box_target"):
            out.append(white_box(inputs, target_img, drawer, net, vgg, args, n_iters, is_cars, save_dir, device))
        elif kind == "patch_white_box":
            from .patch import adversarial_patch as patch
            args.image_size = inputs.size(3)                                               # :323
            p, m = patch.main(drawer, net, vgg, train_dataloader, device, save_dir, args, target_img)
            out.append(patch_white_box(inputs, m.to(device), p.to(device)))                 # :327-328
        else:
            raise NotImplementedError(f"{kind}: not a gradient attack (out of scope, SURVEY section 2)")
    return out
