"""per-layer / per-sample style-gradient error of the all-fused synthesis in fp32 mode: python tests/diag_fused.py [B] [fused_res]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 3
os.environ["SFK_FUSED_UP_RES"] = sys.argv[2] if len(sys.argv) > 2 else "8"
import torch
from oracle import stylegan2 as sg
from sfattack import lib
from sfattack.engine import SynthesisEngine
from test_engine_gpu import _small_setup
spec, GP, *_ = _small_setup(size=64)
g = torch.Generator().manual_seed(17)
w = torch.randn(B, spec.n_latent, spec.style_dim, generator=g)
if os.environ.get("PERM"):
    w = w[[int(c) for c in os.environ["PERM"]]]
wr = w.clone().requires_grad_(True)
styles = sg.styles_from_wplus(GP, spec, wr)
img_ref = sg.synthesis_from_styles(GP, spec, styles)
gimg = torch.randn(img_ref.shape, generator=g)
gs_ref = torch.cat(torch.autograd.grad((img_ref * gimg).sum(), styles), 1)
lib.set_activation_dtype(torch.float32)
err = torch.zeros(1, dtype=torch.int32, device="cuda:0")
syn = SynthesisEngine(spec, GP, B, torch.device("cuda:0"), err)
syn.styles_from_wplus(w.cuda())
img = syn.forward()
gs = syn.backward(gimg.cuda()).cpu()
print("img err per sample", (img.cpu() - img_ref.detach()).abs().flatten(1).max(1).values.tolist())
for l in spec.layers:
    a, b = gs[:, l.s_off:l.s_off + l.cin], gs_ref[:, l.s_off:l.s_off + l.cin]
    print(f"{l.name:12s} {l.kind:4s} res {l.res:3d} cin {l.cin:3d} cout {l.cout:3d}  relerr per sample", [f"{float((a[i]-b[i]).norm()/b[i].norm()):.2e}" for i in range(B)])

s_cat = torch.cat(styles, 1).detach()
for l in spec.layers:
    if l.kind == "rgb": continue
    a, b = gs[1, l.s_off:l.s_off + l.cin], gs_ref[1, l.s_off:l.s_off + l.cin]
    e = (a - b).abs()
    i = int(e.argmax())
    srt = e.sort(descending=True).values[:4].tolist()
    print(f"{l.name:10s} worst ch {i}: got {a[i]:.5f} ref {b[i]:.5f} s={s_cat[1, l.s_off + i]:.5f}  top errs {[f'{v:.2e}' for v in srt]}  |gs| mean {b.abs().mean():.3f}")
