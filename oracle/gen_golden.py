"""Generate tests/golden/vgg_golden.pt by running the REFERENCE's own code/vgg.py (imported unmodified
from /root/reference, which exists only in the build container).  The committed fixture pins
oracle/vgg_ref.py to the reference (code/vgg.py is the only hot-path arithmetic that is
importable here).  Run:  python -m oracle.gen_golden
"""
import importlib.util
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sfattack.params import make_vgg_state_dict  # noqa: E402

REF_VGG = "/root/reference/code/vgg.py"


def main():
    spec = importlib.util.spec_from_file_location("ref_vgg", REF_VGG)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    seed = 11
    sd = make_vgg_state_dict(seed)
    with tempfile.TemporaryDirectory() as d:
        pth = os.path.join(d, "vgg_synth.pth")
        torch.save(sd, pth)
        net = ref.vgg16(pth).eval()          # code/vgg.py:79-81
    cases = []
    for (n, hw, s) in [(1, 16, 0), (2, 20, 1), (1, 36, 2)]:   # 36 -> pool3 input 9x9 exercises ceil_mode
        g = torch.Generator().manual_seed(100 + s)
        x = (torch.rand(n, 3, hw, hw, generator=g) * 2 - 1).requires_grad_(True)
        taps = net(x)
        loss = sum((t ** 2).mean() for t in taps)
        (gx,) = torch.autograd.grad(loss, x)
        cases.append(dict(n=n, hw=hw, x=x.detach().clone(), taps=[t.detach().clone() for t in taps],
                          grad=gx.clone()))
    out = os.path.join(ROOT, "tests", "golden", "vgg_golden.pt")
    torch.save(dict(vgg_seed=seed, cases=cases, torch=str(torch.__version__)), out)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
