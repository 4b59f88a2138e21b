"""diagnostic: error of the fp32-storage conv modes vs fp64 as a function of K (is the tensor-core accumulation biased?)"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from sfattack import lib
dev = torch.device("cuda:0")
lib.set_activation_dtype(torch.float32)
g = torch.Generator(device=dev).manual_seed(0)
for positive in (False, True):
    for cin in (16, 64, 256, 1024):
        n, h, w, cout = 1, 32, 32, 64
        x = torch.randn(n, cin, h, w, generator=g, device=dev)
        wt = torch.randn(cout, cin, 3, 3, generator=g, device=dev) / math.sqrt(cin * 9)
        if positive:
            x, wt = x.abs(), wt.abs()
        ref = F.conv2d(x.double(), wt.double(), padding=1)
        xb = x.permute(0, 2, 3, 1).contiguous()
        wb = wt.permute(2, 3, 0, 1).contiguous().reshape(1, 9 * cout, cin).contiguous()
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        row = []
        for mode in ("cuda_cores", "tf32x3", "tf32"):
            lib.set_conv_math(mode)
            out = torch.empty(n, h, w, cout, device=dev)
            d = lib.make_igemm_desc(xb, n, h, w, cin, 1, wb, 1, 9 * cout, out, h, w, cout, 1, 64, lib.conv3x3_taps(cout), err=err)
            lib.igemm(d)
            torch.cuda.synchronize()
            e = out.permute(0, 3, 1, 2).double() - ref
            row.append(f"{mode}: max {e.abs().max().item() / ref.abs().max().item():.2e} rms {e.pow(2).mean().sqrt().item() / ref.pow(2).mean().sqrt().item():.2e} "
                       f"bias {(e.mean() / ref.abs().mean()).item():+.2e}")
        t = F.conv2d(x, wt, padding=1)          # torch fp32 (TF32 off by default for conv? printed for reference)
        e = t.double() - ref
        row.append(f"torch-fp32: max {e.abs().max().item() / ref.abs().max().item():.2e}")
        print(f"positive={positive} K={9 * cin:5d} | " + " | ".join(row))
lib.set_conv_math("auto")
lib.set_activation_dtype(torch.bfloat16)
