"""Time the bandwidth-bound kernels at one synthesis resolution:  python tests/prof_elem.py n res c  (res = OUTPUT resolution)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sfattack import lib

def timeit(fn, reps=5):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def main():
    n, res, c = map(int, sys.argv[1:4])
    dev = torch.device("cuda:0")
    h = res // 2
    bf = torch.bfloat16
    T = torch.randn(n, 4, h + 1, h + 1, c, device=dev).to(bf)
    out = torch.randn(n, res, res, c, device=dev).to(bf)
    gout = torch.randn(n, res, res, c, device=dev).to(bf)
    gz = torch.empty_like(gout)
    d = torch.rand(n, c, device=dev) + 0.5
    noise = torch.randn(res, res, device=dev)
    bias = torch.randn(c, device=dev) * 0.1
    gd = torch.zeros(n, c, device=dev)
    act_bytes = n * res * res * c * 2
    rows = []
    ms = timeit(lambda: lib.blur_act_fwd(T, out, d, noise, 0.1, bias)); rows.append(("blur_act_fwd", ms, 2 * act_bytes))
    ms = timeit(lambda: lib.blur_act_bwd(out, gout, T, d, noise, 0.1, bias, gd)); rows.append(("blur_act_bwd", ms, 3 * act_bytes))
    ms = timeit(lambda: lib.act_bwd(out, gout, gz, d, noise, 0.1, bias, gd)); rows.append(("act_bwd", ms, 3 * act_bytes))
    wr = torch.randn(3, c, device=dev); s = torch.rand(n, c, device=dev) + 0.5; rb = torch.zeros(3, device=dev)
    skip = torch.randn(n, 3, h, h, device=dev); rgb = torch.empty(n, 3, res, res, device=dev)
    ms = timeit(lambda: lib.torgb_fwd(out, wr, s, 0, rb, skip, rgb)); rows.append(("torgb_fwd", ms, act_bytes + n * 3 * res * res * 4))
    gs = torch.zeros(n, c, device=dev)
    ms = timeit(lambda: lib.torgb_bwd(out, wr, s, 0, rgb, gz, gs)); rows.append(("torgb_bwd", ms, 2 * act_bytes + n * 3 * res * res * 4))
    x = torch.rand(2 * n, 3, res, res, device=dev); x0 = x.clone(); gp = torch.randn(2 * n, 3, 256, 256, device=dev); st = torch.zeros(2 * n, device=dev)
    k = res // 256
    if k >= 1:
        ms = timeit(lambda: lib.attack_update_linf(x, x0, gp, 0.01, 0.03, 1.0, 0.0, 1.0, st, k)); rows.append(("update_linf", ms, 3 * x.numel() * 4))
        img = torch.randn(n, 3, res, res, device=dev); gpool = torch.randn(n, 3, 256, 256, device=dev); g = torch.empty_like(img); loss = torch.zeros(n, device=dev)
        ms = timeit(lambda: lib.image_loss_grad(img, img, gpool, g, loss, 1.0, 1.0, k)); rows.append(("image_loss_grad", ms, 3 * img.numel() * 4))
    for name, ms, b in rows:
        print(f"{name:16s} n={n} res={res} c={c}: {ms*1e3:8.1f} us  {b/ms/1e6:8.1f} GB/s algorithmic")

if __name__ == "__main__":
    main()
