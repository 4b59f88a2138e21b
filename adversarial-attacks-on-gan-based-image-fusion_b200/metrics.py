"""Outcome metrics of the reference: cal_result / cal_rec_loss (code/attack/interpolation.py:1076-1091, 848-855):
pixel MSE and the summed 4-tap VGG feature MSE between a benign and an adversarial fusion.  SSIM (skimage, :903-919) is not
available in this image and is left out."""
import torch

from . import lib


def cal_rec_loss(vgg, a: torch.Tensor, b: torch.Tensor):
    """-> (pixel MSE per sample, sum over taps of feature MSE per sample); VGG runs at the images' own resolution (:1083-1084)."""
    n = a.shape[0]
    dev = a.device
    a, b = a.float().contiguous(), b.float().contiguous()
    mse = torch.zeros(n, device=dev)
    g = torch.empty_like(a)
    lib.image_loss_grad(a, b, None, g, mse, 1.0 / (a.numel() // n), 0.0, 1)
    st_a = vgg.stack(n, a.shape[-1])
    st_a.forward(a)
    fa = [t.clone() for t in st_a.tap_outputs()]
    st_a.forward(b)
    feat = torch.zeros(n, device=dev)
    for x, y in zip(fa, st_a.tap_outputs()):
        lib.mse_tap(x, y, None, feat, 1.0 / (x.numel() // n), 0.0)
    return mse, feat


def cal_result(vgg, benign_fused: torch.Tensor, adv_fused_list):
    rows = []
    for adv in adv_fused_list:
        m, f = cal_rec_loss(vgg, benign_fused.expand_as(adv), adv)
        rows.append(dict(or_f_ad_f=m.tolist(), vgg=f.tolist()))
    return rows
