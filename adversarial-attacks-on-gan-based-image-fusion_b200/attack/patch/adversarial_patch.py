"""Patch attack with the reference's entry points (code/attack/patch/adversarial_patch.py: attack :94-160, train :26-74,
main :163-243).  The reference's loss has only the encoder term live (`Loss = -1*l_latent_org_adv`, :126); the weights are a
ReconLossCfg so the other terms can be switched on.

`train_universal` is the data-parallel form of `train`: ONE patch shared by every image of a batch that is
sharded over ranks; each iteration sums the masked image gradients over the local batch in one kernel, all-reduces the 3 x S x S
sum over NCCL (the only collective inside any loop of this package) and applies the same raw-gradient step on every rank."""
from __future__ import annotations

import os

import numpy as np
import torch

from ... import lib
from ...engine import ReconLossCfg
from ..attack_main2 import _recon_engine
from .adversarial_patch_util import (circle_transform, init_patch_circle, init_patch_square, square_transform, submatrix)

PATCH_LOSS = ReconLossCfg(w_latent_target=0.0, w_latent_org=-1.0, w_img_rec_target=0.0, w_img_org=0.0, w_lpips_img=0.0)   # :126


class _Pair:
    """the reference passes generator and encoder separately (:94); the engine cache wants the pair"""

    def __init__(self, generator, encoder):
        self.decoder, self.encoder = generator, encoder


def attack(img, patch, mask, generator, encoder, vgg, device, args, target_img, save_dir=None, epoch=0, batch_idx=0, loss=None,
           Model=None):                                                                     # adversarial_patch.py:94-160
    """raw-gradient patch descent: patch -= grad; adv_x = clamp((1-mask)*img + mask*patch, min(img), max(img)) (:131-138).
    -> (adv_x, mask, patch, adv_img_rec) like the reference.  The per-iteration `Loss` lines of :141-156 are written once after the
    loop (they force a host sync per iteration upstream)."""
    loss = loss or PATCH_LOSS
    assert loss.w_img_org == 0.0, "the patch loss has no pixel term on adv_x (adversarial_patch.py:126); the update takes the pooled gradient only"
    Model = Model or _Pair(generator, encoder)
    eng = _recon_engine(Model, vgg, img.shape[0], device, loss)
    dev = eng.dev
    with torch.cuda.device(dev):
        img = img.to(dev).float().contiguous()
        eng.set_inputs(img, target_img.to(dev).float().expand_as(img).contiguous())
        patch = patch.to(dev).float().expand_as(img).contiguous().clone()
        mask = mask.to(dev).float().expand_as(img).contiguous()
        n = img.shape[0]
        lo, hi = torch.empty(n, device=dev), torch.empty(n, device=dev)
        lib.minmax_per_sample(img, lo, hi)
        k = eng.k_in
        zero = torch.zeros_like(eng.g_xin)
        lib.attack_update_patch(eng.x, img, patch, mask, zero, 0.0, -1.0, False, lo, hi, 1.0, None, k)       # :106
        n_it = int(args.max_count)
        log = torch.zeros(n_it, n, device=dev)
        for count in range(n_it):                                                                           # :111-158
            l, _, _ = eng.forward_backward()
            log[count].copy_(l)
            lib.attack_update_patch(eng.x, img, patch, mask, eng.g_xin, 1.0, -1.0, False, lo, hi, 1.0 / (k * k), None, k)
        eng.check()
        adv_img_rec = eng.reconstruct().clone()
    if save_dir:
        os.makedirs(save_dir, exist_ok=True)
        rows = log.cpu()
        with open(os.path.join(save_dir, "output_file_loss.txt"), "a") as f:                                # :155-156
            for count in range(n_it):
                for b in range(n):
                    f.write("%dth img count: %d loss:%.5f\n" % (batch_idx + b, count + 1, rows[count, b]))
    return eng.x.clone(), mask, patch, adv_img_rec


def _transform(args, patch, data_shape, patch_shape, rng=None):
    centre = bool(getattr(args, "patch_centre", False))
    if getattr(args, "patch_type", "square") == "circle":                                                    # :39-42
        p, m, patch_shape = circle_transform(patch, data_shape, patch_shape, args.image_size, centre=centre, rng=rng)
    else:
        p, m = square_transform(patch, data_shape, patch_shape, args.image_size, centre=centre, rng=rng)
    return p, m, patch_shape


def _crop_back(masked: np.ndarray, patch_shape) -> np.ndarray:
    """masked patch canvas -> (1,3,d,d) array: every plane cropped to the bounding box of its non-zero entries (:61-69)"""
    new_patch = np.zeros(patch_shape)
    for i in range(new_patch.shape[0]):
        for j in range(new_patch.shape[1]):
            sm = submatrix(masked[i][j])
            new_patch[i][j] = sm if sm.shape == new_patch[i][j].shape else np.resize(sm, new_patch[i][j].shape)
    return new_patch


def train(epoch, patch, patch_shape, net, drawer, vgg, train_loader, device, save_dir, args, target_img, rng=None):   # :26-74
    """carry one patch over the images of train_loader (batch 1 in the reference): place it (random position / rotation), run the
    iterated attack on the GPU, crop the optimised patch back out on the host (:61-69)."""
    mask = None
    for batch_idx, data in enumerate(train_loader):
        data = data.to(device)
        p_np, m_np, patch_shape = _transform(args, patch, tuple(data.shape), patch_shape, rng)               # :38-42
        p_t, m_t = torch.from_numpy(p_np).float().to(device), torch.from_numpy(m_np).float().to(device)      # :43
        adv_x, mask, p_t, _ = attack(data, p_t, m_t, net.decoder, net.encoder, vgg, device, args, target_img, save_dir, epoch, batch_idx,
                                     Model=net)
        masked = (mask * p_t)[:1].cpu().numpy()                                                             # :61-63
        patch = _crop_back(masked, patch_shape)
    return patch, mask


def main(drawer, net, vgg, train_dataloader, device, save_dir, args, target_img, rng=None):                  # :163-243
    """init (:203-207 / :216-219) -> epochs of train (:223-225) -> final placement (:231-236) -> (patch, mask) tensors, saved under
    save_dir/patch with the reference's file names (:238-239; torch pickles despite the .npz suffix).
    (As checked in, the reference short-circuits at :211-213 and loads tensors from the author's disk; this is the intended path.)"""
    if getattr(args, "patch_type", "square") == "circle":
        patch, patch_shape = init_patch_circle(args.image_size, args.patch_size, rng=rng)
    elif getattr(args, "patch_type", "square") == "square":
        patch, patch_shape = init_patch_square(args.image_size, args.patch_size, rng=rng)
    else:
        raise SystemExit("Please choose a square or circle patch")                                          # :208
    for epoch in range(1, int(getattr(args, "epochs", 1)) + 1):
        patch, _ = train(epoch, patch, patch_shape, net, drawer, vgg, train_dataloader, device, save_dir, args, target_img, rng=rng)
    full, m, _ = _transform(args, patch, (1, 3, args.image_size, args.image_size), patch_shape, rng)
    patch_t, mask_t = torch.from_numpy(full).float().to(device), torch.from_numpy(m).float().to(device)
    if save_dir:
        d = os.path.join(save_dir, "patch")
        os.makedirs(d, exist_ok=True)
        tag = "%s_%d_%.3f" % (getattr(args, "dataset_name", "data"), int(getattr(args, "train_size", 0)), args.patch_size)
        torch.save(mask_t, os.path.join(d, tag + "_mask.npz"))
        torch.save(patch_t, os.path.join(d, tag + "_patch.npz"))
    return patch_t, mask_t


def train_universal(patch, mask, images, net, vgg, device, args, target_img, loss=None, lr: float = 1.0, group=None):
    """One patch for ALL images: `images` (b,3,S,S) is this rank's shard, `patch`/`mask` (1,3,S,S) are shared.
    Per iteration: forward/backward of the local batch, sum_n mask * dL_n/dx_n in one kernel, all-reduce of that 3 x S x S sum
    (NCCL / gloo; skipped when torch.distributed is not initialised), `patch -= lr * sum` (the batch form of :133), re-mask and
    clamp every image to its own clean range (:137-138).  -> (patch (1,3,S,S), adv_x (b,3,S,S), loss log (iters, b))."""
    import torch.distributed as dist
    loss = loss or PATCH_LOSS
    assert loss.w_img_org == 0.0
    eng = _recon_engine(net, vgg, images.shape[0], device, loss)
    dev = eng.dev
    use_dist = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    with torch.cuda.device(dev):
        img = images.to(dev).float().contiguous()
        n = img.shape[0]
        eng.set_inputs(img, target_img.to(dev).float().expand_as(img).contiguous())
        patch = patch.to(dev).float().reshape(1, 3, eng.S, eng.S).contiguous().clone()
        mask = mask.to(dev).float().reshape(1, 3, eng.S, eng.S).contiguous()
        lo, hi = torch.empty(n, device=dev), torch.empty(n, device=dev)
        lib.minmax_per_sample(img, lo, hi)
        k = eng.k_in
        gsum = torch.zeros_like(patch)
        lib.patch_apply_shared(eng.x, img, patch, mask, lo, hi)                                             # :106
        n_it = int(args.max_count)
        log = torch.zeros(n_it, n, device=dev)
        for count in range(n_it):
            l, _, _ = eng.forward_backward()
            log[count].copy_(l)
            lib.patch_grad_reduce(eng.g_xin, mask, gsum, 1.0 / (k * k), k)
            if use_dist:
                dist.all_reduce(gsum, group=group)
            lib.axpby(patch, gsum, patch, 1.0, -lr)                                                         # patch -= lr * sum_n grad_n
            lib.patch_apply_shared(eng.x, img, patch, mask, lo, hi)                                         # :137-138
        eng.check()
    return patch, eng.x.clone(), log
