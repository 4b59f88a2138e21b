"""Layer-by-layer diagnostic of the kernel schedules against the oracle (prints, never asserts).
Run on a GPU box:  python tests/diag_engine.py > gpurun_out/diag.log"""
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import stylegan2 as sg  # noqa: E402
from oracle.vgg_ref import vgg_forward, executed_convs  # noqa: E402
from sfattack.engine import ConvStack, SynthesisEngine, vgg_layers  # noqa: E402
from sfattack.params import gen_spec, make_generator_params, make_vgg_state_dict  # noqa: E402

DEV = torch.device("cuda:0")


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def nchw(t):
    return t.float().permute(0, 3, 1, 2)


def diag_synthesis(size=64):
    ch = {4: 64, 8: 64, 16: 32, 32: 32, 64: 16}
    spec = gen_spec(size, style_dim=64, n_mlp=2, channels=ch)
    GP = make_generator_params(spec, seed=0)
    B = 2
    g = torch.Generator().manual_seed(7)
    w = torch.randn(B, spec.n_latent, spec.style_dim, generator=g)
    styles = [s.detach().requires_grad_(True) for s in sg.styles_from_wplus(GP, spec, w)]
    # oracle with intermediates
    x = GP["input.input"].repeat(B, 1, 1, 1)
    skip = None
    inter = []
    for l, s in zip(spec.layers, styles):
        if l.kind == "rgb":
            skip = sg.to_rgb(GP, l, x, s, skip)
            inter.append(skip)
        else:
            x = sg.styled_conv(GP, l, x, s, GP[f"noises.noise_{l.noise_idx}"])
            x.retain_grad()
            inter.append(x)
    gimg = torch.randn(skip.shape, generator=g)
    (skip * gimg).sum().backward()
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    syn = SynthesisEngine(spec, GP, B, DEV, err)
    syn.styles_from_wplus(w.to(DEV))
    print("styles rel", rel(syn.s, torch.cat([s.detach() for s in styles], 1)))
    syn.forward()
    torch.cuda.synchronize()
    print("err flag", err.item())
    for e, ref in zip(syn.L, inter):
        l = e["l"]
        got = e["rgb"] if l.kind == "rgb" else nchw(e["out"])
        print(f"fwd {l.name:10s} {l.kind:4s} res={l.res:4d} cin={l.cin:4d} cout={l.cout:4d} rel={rel(got, ref.detach()):.3e}")
    syn.backward(gimg.to(DEV))
    torch.cuda.synchronize()
    print("err flag", err.item())
    for e, ref, s in zip(syn.L, inter, styles):
        l = e["l"]
        gs_got = syn.gs[:, l.s_off:l.s_off + l.cin]
        line = f"bwd {l.name:10s} {l.kind:4s} gs rel={rel(gs_got, s.grad):.3e}"
        print(line)


def diag_vgg():
    sd = make_vgg_state_dict(3, width_div=2)
    n, res = 2, 64
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(n, 3, res, res, generator=g) * 2 - 1).requires_grad_(True)
    # oracle with all intermediates
    c = executed_convs(sd)
    outs = []
    o = x
    plan = ["c", "c", "p", "c", "c", "p", "c", "c", "c", "pc", "c", "c"]
    ci = 0
    for p in plan:
        if p == "c":
            o = F.relu(F.conv2d(o, c[ci][0], c[ci][1], padding=1)); ci += 1
        else:
            o = F.max_pool2d(o, 2, 2, ceil_mode=(p == "pc"))
        o.retain_grad()
        outs.append(o)
    taps = [outs[0], outs[1], outs[5], outs[11]]
    refs = [torch.randn_like(t) * t.std() for t in taps]
    L = sum(((t - r.detach()) ** 2).flatten(1).mean(1) for t, r in zip(taps, refs))
    L.sum().backward()
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    vals = list(sd.values())
    st = ConvStack(vgg_layers(2), [(vals[2 * i], vals[2 * i + 1]) for i in range(9)], n, res, DEV, err)
    st.forward(x.detach().to(DEV))
    for i, (got, ref) in enumerate(zip(st.out, outs)):
        print(f"vgg fwd layer {i:2d} {st.layers[i].kind:4s} rel={rel(nchw(got), ref.detach()):.3e}")
    loss = torch.zeros(n, device=DEV)
    refs_d = [r.detach().permute(0, 2, 3, 1).contiguous().to(DEV).bfloat16() for r in refs]
    gx = st.backward(refs_d, 1.0, loss)
    torch.cuda.synchronize()
    print("err flag", err.item(), "loss", loss.tolist(), "ref", L.tolist())
    for i, (got, ref) in enumerate(zip(st.g, outs)):
        l = st.layers[i]
        want = ref.grad * (ref.detach() > 0) if l.kind != "pool" else ref.grad
        print(f"vgg bwd layer {i:2d} {l.kind:4s} g rel={rel(nchw(got), want):.3e}  (|want|={want.norm():.3e})")
    print("vgg g_in rel", rel(gx, x.grad))


if __name__ == "__main__":
    torch.manual_seed(0)
    diag_vgg()
    diag_synthesis()
