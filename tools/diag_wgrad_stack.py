import os, sys, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.vgg_ref import executed_convs
from sfattack.engine import ConvStack, vgg_layers
from sfattack.params import make_vgg_state_dict
from sfattack import lib
DEV = "cuda:0"
def cos(a, b):
    a, b = a.double().cpu().flatten(), b.double().cpu().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))
sd = make_vgg_state_dict(3, width_div=2)
n, res = 2, 64
g = torch.Generator().manual_seed(1)
x = torch.rand(n, 3, res, res, generator=g) * 2 - 1
c = [(w.clone().requires_grad_(True), b.clone().requires_grad_(True)) for w, b in executed_convs(sd)]
def fwd(image, keep):
    z0 = F.conv2d(image, c[0][0], c[0][1], padding=1); keep.append(z0); o = F.relu(z0); t0 = o
    z1 = F.conv2d(o, c[1][0], c[1][1], padding=1); keep.append(z1); o = F.relu(z1); t1 = o
    o = F.max_pool2d(o, 2, 2)
    z2 = F.conv2d(o, c[2][0], c[2][1], padding=1); keep.append(z2); o = F.relu(z2)
    z3 = F.conv2d(o, c[3][0], c[3][1], padding=1); keep.append(z3); o = F.relu(z3)
    o = F.max_pool2d(o, 2, 2); t2 = o
    for k in (4, 5, 6):
        z = F.conv2d(o, c[k][0], c[k][1], padding=1); keep.append(z); o = F.relu(z)
    o = F.max_pool2d(o, 2, 2, ceil_mode=True)
    for k in (7, 8):
        z = F.conv2d(o, c[k][0], c[k][1], padding=1); keep.append(z); o = F.relu(z)
    return t0, t1, t2, o
with torch.no_grad():
    refs = [t.detach() for t in fwd((x + 0.05 * torch.randn(x.shape, generator=g)).clamp(-1, 1), [])]
zs = []
taps = fwd(x, zs)
for z in zs: z.retain_grad()
L = sum(((t - r) ** 2).flatten(1).mean(1) for t, r in zip(taps, refs))
L.sum().backward()
err = torch.zeros(1, dtype=torch.int32, device=DEV)
st = ConvStack(vgg_layers(2), [(w.detach(), b.detach()) for w, b in c], n, res, torch.device(DEV), err)
xd = x.to(DEV)
st.forward(xd)
loss = torch.zeros(n, device=DEV)
st.backward([r.permute(0, 2, 3, 1).contiguous().to(DEV).bfloat16() for r in refs], 1.0, loss)
wg = st.weight_grads(xd)
conv_idx = [i for i, l in enumerate(st.layers) if l.kind != "pool"]
for k, i in enumerate(conv_idx):
    gz_gpu = st.g[i].float().permute(0, 3, 1, 2)
    print(f"layer {i}: cos(g buffer, oracle dL/dz) = {cos(gz_gpu, zs[k].grad):.4f}   cos(dW) = {cos(wg[i][0], c[k][0].grad):.4f}   cos(db) = {cos(wg[i][1], c[k][1].grad):.4f}", end="")
    if i > 0:
        xin = st.out[i - 1].float().permute(0, 3, 1, 2)
        w0 = torch.zeros_like(c[k][0].detach()).to(DEV).requires_grad_(True)
        (gw_t,) = torch.autograd.grad((F.conv2d(xin, w0, padding=1) * gz_gpu).sum(), w0)
        print(f"   cos(kernel dW, torch dW from the same buffers) = {cos(wg[i][0], gw_t):.6f}   cos(torch dW from buffers, oracle) = {cos(gw_t, c[k][0].grad):.4f}")
    else:
        print()
