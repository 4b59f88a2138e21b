"""Does the tensor-core first conv carry the lo terms?  Fraction of outputs that differ from bf16(round(fp64 conv))."""
import os, sys, math, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sfattack import lib
dev = torch.device("cuda:0")
lib.load()
g = torch.Generator(device=dev).manual_seed(3)
err = torch.zeros(1, dtype=torch.int32, device=dev)
for n, h, w, cout in ((2, 64, 64, 16), (2, 64, 64, 32), (2, 64, 64, 64), (2, 256, 256, 32), (1, 256, 256, 64), (2, 64, 64, 128)):
    x = torch.rand(n, 3, h, w, generator=g, device=dev) * 2 - 1
    wt = torch.randn(cout, 3, 3, 3, generator=g, device=dev) * math.sqrt(2.0 / 27)
    b = torch.randn(cout, generator=g, device=dev) * 0.05
    ref = F.conv2d(x.double(), wt.double(), b.double(), padding=1).relu()
    want = ref.float().bfloat16().permute(0, 2, 3, 1).contiguous()
    wf, wb = lib.c3_pack_weights(wt)
    xp = torch.empty(n, h, w, 16, device=dev, dtype=torch.bfloat16)
    out = torch.empty(n, h, w, cout, device=dev, dtype=torch.bfloat16)
    lib.c3_pack(x, xp)
    d = lib.make_igemm_desc(xp, n, h, w, 16, 1, wf, 1, 9 * cout, out, h, w, cout, 1, lib.pick_block_n(cout), lib.conv3x3_taps(cout),
                            flags=lib.EP_BIAS | lib.EP_RELU, bias=b, err=err)
    lib.igemm(d)
    out2 = torch.empty_like(out)
    lib.conv_c3_fwd(x, wt, b, out2, relu=True)
    out3 = torch.empty_like(out)
    lib.igemm(d, ref=True)
    torch.cuda.synchronize()
    nz = want != 0
    mm = lambda o: ((o != want) & nz).float().sum().item() / nz.float().sum().item()
    # per-channel mismatch of the tensor-core path
    pc = ((out != want) & nz).float().sum((0, 1, 2)) / nz.float().sum((0, 1, 2)).clamp_min(1)
    print(f"n={n} {h}x{w} cout={cout}: mismatch vs bf16(fp64 conv): tensor-core {mm(out):.4f}  cuda-core c3 {mm(out2):.4f}  igemm_ref on the packed operand {mm(out3):.4f}   per-channel max {pc.max().item():.4f} min {pc.min().item():.4f}  info={lib.plan_info(d)}")
