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


def _cos(a, b):
    a, b = a.double().cpu().flatten(), b.double().cpu().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def _relerr(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(params=["bf16", "fp32"])
def storage(request):
    """bf16 storage (product path: tensor-core wgrad) and the fp32 parity mode (CUDA-core wgrad on fp32 buffers)"""
    from sfattack import lib
    if request.param == "fp32":
        lib.set_activation_dtype(torch.float32)
    try:
        yield request.param
    finally:
        lib.set_activation_dtype(torch.bfloat16)


def test_vgg_stack_weight_grads_vs_reference_autograd(storage):
    """ConvStack.weight_grads(): every conv of the executed VGG prefix (code/vgg.py:44-64), weights AND biases, against autograd
    through the pinned oracle (oracle/vgg_ref.py) with the parameters left trainable, as the reference leaves them.
    fp32 storage (split-tf32 convs): 5e-3 (measured 0.9e-3 weights, 2.9e-3 biases).  bf16 storage: the loss is the small difference of two bf16-rounded feature maps, so the stored gradients
    carry ~2 % noise (cos 0.98 against the oracle, same as the input gradient of this stack) and the weight gradient, a
    cancelling sum of them, 0.87-0.95 -- torch evaluating dW from the SAME buffers agrees with the kernel to 1.000000
    (tools/diag_wgrad_stack.py), i.e. the floor is the storage, not the GEMM."""
    from oracle.vgg_ref import vgg_forward
    from sfattack.engine import ConvStack, vgg_layers
    from sfattack.params import make_vgg_state_dict
    sd = {k: v.clone().requires_grad_(True) for k, v in make_vgg_state_dict(3, width_div=2).items()}
    n, res = 2, 64
    g = torch.Generator().manual_seed(1)
    x = torch.rand(n, 3, res, res, generator=g) * 2 - 1
    with torch.no_grad():
        refs = [t.detach() for t in vgg_forward(sd, (x + 0.05 * torch.randn(x.shape, generator=g)).clamp(-1, 1))]
    taps = vgg_forward(sd, x)
    L = sum(((t - r) ** 2).flatten(1).mean(1) for t, r in zip(taps, refs))
    vals = list(sd.values())[:18]
    grads = torch.autograd.grad(L.sum(), vals)
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    st = ConvStack(vgg_layers(2), [(vals[2 * i].detach(), vals[2 * i + 1].detach()) for i in range(9)], n, res, torch.device(DEV), err)
    xd = x.to(DEV)
    st.forward(xd)
    loss = torch.zeros(n, device=DEV)
    act = torch.float32 if storage == "fp32" else torch.bfloat16
    st.backward([r.permute(0, 2, 3, 1).contiguous().to(DEV).to(act) for r in refs], 1.0, loss)
    wg = st.weight_grads(xd)
    torch.cuda.synchronize()
    assert err.item() == 0
    conv_idx = [i for i, l in enumerate(st.layers) if l.kind != "pool"]
    assert sorted(wg) == conv_idx and len(conv_idx) == 9
    for k, i in enumerate(conv_idx):
        dw, db = wg[i]
        assert dw.shape == grads[2 * k].shape and db.shape == grads[2 * k + 1].shape
        if storage == "fp32":
            assert _relerr(dw, grads[2 * k]) < 5e-3 and _relerr(db, grads[2 * k + 1]) < 5e-3, (i, _relerr(dw, grads[2 * k]), _relerr(db, grads[2 * k + 1]))
        else:
            assert _cos(dw, grads[2 * k]) > 0.85 and _cos(db, grads[2 * k + 1]) > 0.85, (i, _cos(dw, grads[2 * k]), _cos(db, grads[2 * k + 1]))


def test_synthesis_weight_grads_vs_oracle_autograd(storage):
    """SynthesisEngine.weight_grads(): dL/dW of the non-upsampling 3x3 ModulatedConv2d layers against autograd through the oracle
    generator with its conv weights left trainable."""
    from oracle import stylegan2 as sg
    from sfattack.engine import SynthesisEngine
    from sfattack.params import gen_spec, make_generator_params
    spec = gen_spec(64, style_dim=64, n_mlp=2, channels={4: 64, 8: 64, 16: 32, 32: 32, 64: 16})
    GP = make_generator_params(spec, seed=0)
    B = 2
    g = torch.Generator().manual_seed(7)
    w = torch.randn(B, spec.n_latent, spec.style_dim, generator=g)
    names = [l.name for l in spec.layers if l.kind == "conv"]
    GPt = dict(GP)
    for nme in names:
        GPt[f"{nme}.conv.weight"] = GP[f"{nme}.conv.weight"].clone().requires_grad_(True)
    styles = sg.styles_from_wplus(GPt, spec, w)
    img_ref = sg.synthesis_from_styles(GPt, spec, styles)
    gimg = torch.randn(img_ref.shape, generator=g)
    grads = torch.autograd.grad((img_ref * gimg).sum(), [GPt[f"{nme}.conv.weight"] for nme in names])
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    syn = SynthesisEngine(spec, GP, B, torch.device(DEV), err)
    syn.styles_from_wplus(w.to(DEV))
    syn.forward()
    syn.backward(gimg.to(DEV))
    wg = syn.weight_grads()
    torch.cuda.synchronize()
    assert err.item() == 0
    assert sorted(wg) == sorted(names)
    for nme, gr in zip(names, grads):
        got = wg[nme]
        assert got.shape == gr[0].shape
        if storage == "fp32":
            assert _relerr(got, gr[0]) < 2e-3, (nme, _relerr(got, gr[0]))
        else:   # same bf16 floor as the style gradient of this engine (tests/test_engine_gpu.py): nearly-cancelling demodulation terms
            assert _cos(got, gr[0]) > 0.99 and _relerr(got, gr[0]) < 0.15, (nme, _cos(got, gr[0]), _relerr(got, gr[0]))
