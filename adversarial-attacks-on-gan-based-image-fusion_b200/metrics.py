"""Outcome metrics of the reference, same names and return shapes (code/attack/interpolation.py):
  cal_rec_loss(img, rec_img) -> per-sample pixel MSE                                   (:848-855)
  cal_SSMI(original_image, distorted_image) -> SSIM of the gray images                 (:903-919)
  cal_result(original_f, adv_f_all) -> (or_f_ad_f_all, vg_all, ssmi_all) dicts by index (:1076-1091): pixel MSE, sum of the four
      VGG-tap MSEs (VGG at the images' own resolution, no pooling: :1083-1084) and SSIM between the benign fusion and each
      adversarial fusion -- the reference's notion of attack success.
Every number is reduced on the GPU by the kernels of libsfattack; the reference's module globals `vgg`, `device` are kwargs."""
from typing import Dict, Tuple

import torch

from . import lib

_VGG = None


def set_vgg(vgg):
    """module-level VGG, as the reference's scripts set a global `vgg` (interpolation.py:1117)"""
    global _VGG
    _VGG = vgg


def cal_rec_loss(img: torch.Tensor, rec_img: torch.Tensor) -> torch.Tensor:
    n = img.shape[0]
    with torch.cuda.device(img.device):
        a, b = img.float().contiguous(), rec_img.float().contiguous()
        mse = torch.zeros(n, device=a.device)
        g = torch.empty_like(a)
        lib.image_loss_grad(a, b, None, g, mse, 1.0 / (a.numel() // n), 0.0, 1)
    return mse


def ssim(a: torch.Tensor, b: torch.Tensor, data_range: float = 2.0) -> torch.Tensor:
    """(n,3,H,W) x2 -> (n,) mean SSIM; data_range 2 = float images in [-1,1] (what the library assumes for floats)"""
    with torch.cuda.device(a.device):
        a, b = a.float().contiguous(), b.float().contiguous()
        out = torch.empty(a.shape[0], device=a.device)
        lib.ssim_gray7(a, b, out, data_range)
    return out


def cal_SSMI(original_image: torch.Tensor, distorted_image: torch.Tensor) -> float:
    if original_image.shape != distorted_image.shape:
        raise ValueError("Both images must have the same dimensions and shape.")                  # :908-909
    dev = original_image.device if original_image.is_cuda else torch.device("cuda:0")
    return float(ssim(original_image[None].to(dev), distorted_image[None].to(dev))[0])


def vgg_tap_mse(vgg, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """sum over the 4 taps of per-sample feature MSE"""
    n = a.shape[0]
    with torch.cuda.device(a.device):
        st = vgg.stack(n, a.shape[-1])
        st.forward(a.float().contiguous())
        fa = [t.clone() for t in st.tap_outputs()]
        st.forward(b.float().contiguous())
        feat = torch.zeros(n, device=a.device)
        for x, y in zip(fa, st.tap_outputs()):
            lib.mse_tap(x, y, None, feat, 1.0 / (x.numel() // n), 0.0)
    return feat


def cal_result(original_f: torch.Tensor, adv_f_all: torch.Tensor, vgg=None) -> Tuple[Dict[int, float], Dict[int, float], Dict[int, float]]:
    vgg = vgg or _VGG
    assert vgg is not None, "cal_result needs a VGG (pass vgg= or call metrics.set_vgg)"
    n = adv_f_all.size(0)
    ref = original_f.expand(n, -1, -1, -1).contiguous()
    adv = adv_f_all.contiguous()
    mse, vg, ss = cal_rec_loss(ref, adv), vgg_tap_mse(vgg, ref, adv), ssim(ref, adv)
    mse, vg, ss = mse.tolist(), vg.tolist(), ss.tolist()            # one device-to-host copy per metric, after all kernels
    return ({i: mse[i] for i in range(n)}, {i: vg[i] for i in range(n)}, {i: ss[i] for i in range(n)})
