"""Weight-gradient GEMMs (include/sfk.h: sfk_conv3x3_wgrad, sfk_bias_grad, sfk_modconv_wgrad_finish) against torch.autograd.grad.

The reference never freezes parameters (code/attack/attack_main2.py:301-304), so its autograd evaluates dL/dW of every VGG conv
(code/vgg.py:45-63) and of every ModulatedConv2d (decoder call, attack_main2.py:619-621) on each iteration; north_star lists
"forward and both backward GEMMs".  Inputs are bf16-representable, products accumulate in fp32 (tensor core / fmaf), the splits are
reduced with fp32 atomics: tolerance 2e-3 of the gradient's scale."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _gen(seed):
    return torch.Generator(device=DEV).manual_seed(seed)


def _rb(*shape, g, scale=1.0):
    return (torch.randn(*shape, generator=g, device=DEV) * scale).bfloat16().float()


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().bfloat16()


def _close(a, b, rtol, what):
    scale = b.abs().max().item() + 1e-12
    err = (a.float() - b.float()).abs().max().item()
    assert err <= rtol * scale, f"{what}: max err {err:.4e} vs scale {scale:.4e} (rel {err / scale:.3e})"


def _ref_dw(x, gz, w_shape):
    w = torch.zeros(w_shape, device=DEV, requires_grad=True)
    y = F.conv2d(x, w, padding=1)
    (gw,) = torch.autograd.grad((y * gz).sum(), w)
    return gw.permute(2, 3, 0, 1).reshape(9, w_shape[0], w_shape[1])     # [tap = ky*3+kx][cout][cin]


@pytest.mark.parametrize("n,h,w,cin,cout", [
    (2, 32, 32, 64, 64),      # one box per operand
    (3, 20, 12, 32, 128),     # ragged tiles, channel boxes zero-filled past cin, two M boxes
    (1, 16, 16, 128, 256),    # two output-channel blocks, two N boxes
    (2, 8, 8, 512, 512),      # TW = 8, four input-channel blocks
    (2, 33, 17, 64, 96),      # odd sizes, second M box partly out of range
    (1, 64, 64, 192, 64),     # last input-channel block half out of range
    (8, 64, 64, 64, 64),      # many splits (atomics across CTAs)
])
def test_conv3x3_wgrad_shared_weights(n, h, w, cin, cout):
    from sfattack import lib
    g = _gen(11)
    x = _rb(n, cin, h, w, g=g)
    gz = _rb(n, cout, h, w, g=g)
    want = _ref_dw(x, gz, (cout, cin, 3, 3))
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    for ref in (False, True):
        dw = lib.conv3x3_wgrad(_nhwc(x), _nhwc(gz), per_sample=False, ref=ref, err=err)
        torch.cuda.synchronize()
        assert err.item() == 0, "kernel reported an internal timeout"
        _close(dw[0], want, 2e-3, f"wgrad ref={ref}")
    # accumulation contract: a second call adds
    dw2 = lib.conv3x3_wgrad(_nhwc(x), _nhwc(gz), dw=dw.clone(), per_sample=False, err=err)
    _close(dw2[0], 2 * want, 2e-3, "wgrad accumulates")


@pytest.mark.parametrize("n,h,w,cin,cout", [(4, 16, 16, 64, 64), (2, 32, 32, 128, 64), (8, 8, 8, 256, 256)])
def test_conv3x3_wgrad_per_sample(n, h, w, cin, cout):
    from sfattack import lib
    g = _gen(12)
    x = _rb(n, cin, h, w, g=g)
    gz = _rb(n, cout, h, w, g=g)
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    dw = lib.conv3x3_wgrad(_nhwc(x), _nhwc(gz), per_sample=True, err=err)
    dwr = lib.conv3x3_wgrad(_nhwc(x), _nhwc(gz), per_sample=True, ref=True)
    torch.cuda.synchronize()
    assert err.item() == 0
    for i in range(n):
        want = _ref_dw(x[i:i + 1], gz[i:i + 1], (cout, cin, 3, 3))
        _close(dw[i], want, 2e-3, f"per-sample wgrad image {i}")
        _close(dwr[i], want, 2e-3, f"per-sample wgrad (CUDA cores) image {i}")


def test_conv3x3_wgrad_small_shapes_use_cuda_cores():
    """4x4 maps (width < 8) and channel counts that are not multiples of 8 are outside the tensor-core kernel's domain."""
    from sfattack import lib
    g = _gen(13)
    for (n, h, w, cin, cout) in [(2, 4, 4, 64, 64), (1, 8, 8, 12, 20)]:
        x = _rb(n, cin, h, w, g=g)
        gz = _rb(n, cout, h, w, g=g)
        dw = lib.conv3x3_wgrad(_nhwc(x), _nhwc(gz))
        _close(dw[0], _ref_dw(x, gz, (cout, cin, 3, 3)), 2e-3, "small wgrad")


def test_vgg_conv_bias_relu_weight_and_bias_grads():
    """conv3x3+bias+ReLU (code/vgg.py:45-63): dW and db from the masked output gradient, against autograd through the layer."""
    from sfattack import lib
    g = _gen(14)
    n, h, w, cin, cout = 2, 32, 32, 64, 128
    x = _rb(n, cin, h, w, g=g)
    wt = (_rb(cout, cin, 3, 3, g=g, scale=1 / math.sqrt(9 * cin))).requires_grad_(True)
    b = torch.randn(cout, generator=g, device=DEV).requires_grad_(True)
    y = F.relu(F.conv2d(x, wt, b, padding=1))
    gy = _rb(n, cout, h, w, g=g)
    gw_ref, gb_ref = torch.autograd.grad((y * gy).sum(), (wt, b))
    gz = (gy * (y > 0)).detach()                                   # what the engines' dgrad chain leaves in the gradient buffer
    dw = lib.conv3x3_wgrad(_nhwc(x), _nhwc(gz))
    db = lib.bias_grad(_nhwc(gz))
    _close(dw[0], gw_ref.permute(2, 3, 0, 1).reshape(9, cout, cin), 2e-3, "vgg dW")
    _close(db, gb_ref, 2e-3, "vgg db")


@pytest.mark.parametrize("n,h,cin,cout", [(2, 16, 64, 64), (3, 32, 128, 64)])
def test_modulated_conv_weight_grad(n, h, cin, cout):
    """ModulatedConv2d (oracle/stylegan2.py:modulated_conv): dL/dW of the shared weight through modulation AND
    demodulation = per-sample GEMM (sfk_conv3x3_wgrad) + sfk_modconv_wgrad_finish, against autograd through the oracle's layer."""
    from oracle import stylegan2 as sg
    from sfattack import lib
    from sfattack.params import ModLayer
    g = _gen(15)
    l = ModLayer(name="L", kind="conv", res=h, cin=cin, cout=cout, w_idx=0)
    W = torch.randn(1, cout, cin, 3, 3, generator=g, device=DEV).requires_grad_(True)
    x = _rb(n, cin, h, h, g=g)
    s = (torch.rand(n, cin, generator=g, device=DEV) + 0.5)
    P = {"L.conv.weight": W}
    y = sg.modulated_conv(P, l, x, s)
    gy = _rb(n, cout, h, h, g=g, scale=0.1)
    (gW_ref,) = torch.autograd.grad((y * gy).sum(), W)
    # what the engine holds after its backward: d, gz = d*gy (bf16 buffer), gdacc = sum gy*y, the UNmodulated input x
    scale = 1.0 / math.sqrt(cin * 9)
    wb = (W.detach()[0] * scale).permute(2, 3, 0, 1).reshape(9, cout, cin).contiguous()
    wmod = scale * W.detach() * s.view(n, 1, cin, 1, 1)
    d = torch.rsqrt(wmod.pow(2).sum([2, 3, 4]) + 1e-8)
    gz = gy * d[:, :, None, None]
    gdacc = (gy * y.detach()).sum((2, 3))
    G = lib.conv3x3_wgrad(_nhwc(x), _nhwc(gz), per_sample=True)
    dwb = lib.modconv_wgrad_finish(G, wb, s.contiguous(), 0, d.contiguous(), gdacc.contiguous())
    got = (dwb * scale).reshape(3, 3, cout, cin).permute(2, 3, 0, 1)                  # dL/dW = scale * dL/dWb
    _close(got, gW_ref[0], 1e-2, "modulated conv dW")      # gz passes through a bf16 buffer (2^-9 relative per element)
