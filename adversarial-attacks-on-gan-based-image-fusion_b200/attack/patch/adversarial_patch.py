"""Patch attack with the reference's entry points (code/attack/patch/adversarial_patch.py: attack :94-160, train :26-74,
main :163-243).  The reference's loss has only the encoder term live (`Loss = -1*l_latent_org_adv`, :126); the weights are a
ReconLossCfg so the other terms can be switched on."""
from __future__ import annotations

import numpy as np
import torch

from ... import lib
from ...engine import ReconLossCfg
from ..attack_main2 import _recon_engine
from .adversarial_patch_util import init_patch_square, square_transform, submatrix

PATCH_LOSS = ReconLossCfg(w_latent_target=0.0, w_latent_org=-1.0, w_img_rec_target=0.0, w_img_org=0.0, w_lpips_img=0.0)   # :126


def attack(img, patch, mask, generator, encoder, vgg, device, args, target_img, save_dir=None, epoch=0, batch_idx=0, loss=None,
           Model=None):                                                                     # adversarial_patch.py:94-160
    """raw-gradient patch descent: patch -= grad; adv_x = clamp((1-mask)*img + mask*patch, min(img), max(img)) (:131-138)."""
    class _M:  # the reference passes generator/encoder separately; the engine wants both
        pass
    if Model is None:
        Model = _M()
        Model.decoder, Model.encoder = generator, encoder
    eng = _recon_engine(Model, vgg, img.shape[0], device, loss or PATCH_LOSS)
    dev = eng.dev
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
    for count in range(int(args.max_count)):                                                            # :111-158
        eng.forward_backward()
        lib.attack_update_patch(eng.x, img, patch, mask, eng.g_xin, 1.0, -1.0, False, lo, hi, 1.0 / (k * k), None, k)
    eng.check()
    adv_img_rec = eng.reconstruct().clone()
    return eng.x.clone(), mask, patch, adv_img_rec


def train(epoch, patch, patch_shape, net, drawer, vgg, train_loader, device, save_dir, args, target_img):   # :26-74
    """carry one patch over the images of train_loader (batch 1 in the reference); the crop to the patch's bounding box
    (:62-69) is done once per image on the host, as in the reference."""
    mask = None
    for batch_idx, data in enumerate(train_loader):
        data = data.to(device)
        data_shape = tuple(data.shape)
        p_np, m_np = square_transform(patch, data_shape, patch_shape, args.image_size)                   # :38-42
        p_t, m_t = torch.from_numpy(p_np).float().to(device), torch.from_numpy(m_np).float().to(device)
        adv_x, mask, p_t, _ = attack(data, p_t, m_t, net.decoder, net.encoder, vgg, device, args, target_img, save_dir, epoch, batch_idx,
                                     Model=net)
        masked = (mask * p_t).cpu().numpy()                                                             # :61-63
        new_patch = np.zeros(patch_shape)
        for i in range(new_patch.shape[0]):
            for j in range(new_patch.shape[1]):
                sm = submatrix(masked[i][j])
                new_patch[i][j] = sm if sm.shape == new_patch[i][j].shape else np.resize(sm, new_patch[i][j].shape)
        patch = new_patch
    return patch, mask


def main(drawer, net, vgg, train_dataloader, device, save_dir, args, target_img):                          # :163-243
    patch, patch_shape = init_patch_square(args.image_size, args.patch_size)                              # :216-219
    mask = None
    for epoch in range(1, int(getattr(args, "epochs", 1)) + 1):
        patch, mask = train(epoch, patch, patch_shape, net, drawer, vgg, train_dataloader, device, save_dir, args, target_img)
    full, m = square_transform(patch, (1, 3, args.image_size, args.image_size), patch_shape, args.image_size)
    return torch.from_numpy(full).float().to(device), torch.from_numpy(m).float().to(device)
