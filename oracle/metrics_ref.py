"""ORACLE (test infrastructure only).  Outcome metrics of the reference restated in plain torch:
  cal_rec_loss   code/attack/interpolation.py:848-855   per-sample pixel MSE
  cal_result     code/attack/interpolation.py:1076-1091 pixel MSE, sum of the 4 VGG-tap MSEs (VGG at full resolution), SSIM
  cal_SSMI       code/attack/interpolation.py:903-919   skimage.color.rgb2gray + skimage.metrics.structural_similarity(defaults)
skimage is absent from this image, so the SSIM is a restatement of the library's published algorithm (Wang et al. 2004
as implemented by skimage: 7x7 uniform window, sample covariance NP/(NP-1), K1 0.01, K2 0.03, data_range 2 for float images in
[-1,1], mean over the region whose window lies inside the image) -> PARITY UNPINNED for SSIM."""
import torch
import torch.nn.functional as F

from .vgg_ref import vgg_forward


def cal_rec_loss(img, rec_img):
    return ((img - rec_img) ** 2).mean(dim=[1, 2, 3])


def rgb2gray(x):
    return 0.2125 * x[:, 0:1] + 0.7154 * x[:, 1:2] + 0.0721 * x[:, 2:3]


def ssim(a, b, data_range=2.0, win=7):
    ga, gb = rgb2gray(a.double()), rgb2gray(b.double())
    NP = win * win
    cov_norm = NP / (NP - 1.0)
    f = lambda t: F.avg_pool2d(t, win, 1)                      # uniform filter, 'valid' region only
    ux, uy = f(ga), f(gb)
    uxx, uyy, uxy = f(ga * ga), f(gb * gb), f(ga * gb)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2))
    return S.mean(dim=[1, 2, 3])


def cal_result(vgg_sd, original_f, adv_f_all):
    mse, vg, ss = {}, {}, {}
    fo = vgg_forward(vgg_sd, original_f)
    for i in range(adv_f_all.size(0)):
        adv = adv_f_all[i:i + 1]
        mse[i] = F.mse_loss(original_f, adv).item()
        vg[i] = sum(F.mse_loss(x, y) for x, y in zip(fo, vgg_forward(vgg_sd, adv))).item()
        ss[i] = ssim(original_f, adv)[0].item()
    return mse, vg, ss
