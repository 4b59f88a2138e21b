"""Gradient of the VGG tap loss at a SMALL input difference (the attack's regime): ConvStack with the first conv on the tensor
cores vs on CUDA cores, both against torch fp32 autograd of the same stack."""
import os, sys, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchvision
from sfattack import lib
from sfattack.engine import ConvStack, vgg_layers
from sfattack.params import VGG_EXECUTED

dev = torch.device("cuda:0")
lib.load()
torch.manual_seed(0)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
sd = torchvision.models.vgg16(weights=None).features.state_dict()
vals = [v.to(dev) for v in sd.values()]
vgg_w = [(vals[2 * i], vals[2 * i + 1]) for i in range(VGG_EXECUTED)]
err = torch.zeros(1, dtype=torch.int32, device=dev)

def torch_taps(x):
    L = vgg_layers(1)
    taps, wi = [], 0
    for i, l in enumerate(L):
        if l.kind == "pool":
            x = F.max_pool2d(x, 2, 2, ceil_mode=True)
        else:
            W, b = vgg_w[wi]; wi += 1
            x = F.relu(F.conv2d(x, W, b, padding=1))
        if l.tap >= 0:
            taps.append(x)
    return taps

def cos(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return (a @ b / (a.norm() * b.norm())).item()

n, res = 2, 256
base = F.avg_pool2d(torch.rand(n, 3, 4 * res, 4 * res, device=dev), 5, 1, 2)      # smooth image in [0,1]
for amp in (8 / 255, 2 / 255, 0.5 / 255):
    delta = (torch.rand(n, 3, 4 * res, 4 * res, device=dev) * 2 - 1) * amp
    x0 = F.avg_pool2d(2 * base - 1, 4, 4).contiguous()
    x1 = F.avg_pool2d(2 * (base + delta).clamp(0, 1) - 1, 4, 4).contiguous()
    xr = x1.clone().requires_grad_(True)
    with torch.no_grad():
        refs32 = torch_taps(x0)
    loss = sum(((a - b) ** 2).flatten(1).mean(1) for a, b in zip(torch_taps(xr), refs32)).sum()
    (g_true,) = torch.autograd.grad(loss, xr)
    row = []
    for tc in ("0", "1"):
        os.environ["SFK_C3_TC"] = tc
        s = ConvStack(vgg_layers(1), vgg_w, n, res, dev, err)
        s.forward(x0)
        refs = [t.clone() for t in s.tap_outputs()]
        s.forward(x1)
        l = torch.zeros(n, device=dev)
        g = s.backward(refs, 1.0, l).clone()
        torch.cuda.synchronize()
        row.append((cos(g, g_true), l.sum().item(), g))
    print(f"amp {amp*255:.1f}/255: true loss {loss.item():.4e} | cuda-core c3: cos {row[0][0]:.5f} loss {row[0][1]:.4e} | tensor-core c3: cos {row[1][0]:.5f} loss {row[1][1]:.4e} | "
          f"between them cos {cos(row[0][2], row[1][2]):.5f}  err={err.item()}")
