"""Stand-alone launcher of one conv layer shape for ncu / timing:  python tests/prof_igemm.py n h w cin cout [mode] [reps]
mode: fwd (modulated conv epilogue) | dgrad (style-gradient epilogue) | tconv"""
import math, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sfattack import lib

def main():
    n, h, w, cin, cout = map(int, sys.argv[1:6])
    mode = sys.argv[6] if len(sys.argv) > 6 else "fwd"
    reps = int(sys.argv[7]) if len(sys.argv) > 7 else 5
    stages = int(sys.argv[8]) if len(sys.argv) > 8 else 0
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(n, h, w, cin, generator=g, device=dev).bfloat16()
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    if mode == "fwd":
        wt = (torch.randn(n, 9 * cout, cin, generator=g, device=dev) / math.sqrt(9 * cin)).bfloat16()
        out = torch.empty(n, h, w, cout, device=dev, dtype=torch.bfloat16)
        d = lib.make_igemm_desc(x, n, h, w, cin, 1, wt, n, 9 * cout, out, h, w, cout, 1, lib.pick_block_n(cout), lib.conv3x3_taps(cout),
                                flags=lib.EP_DSCALE | lib.EP_NOISE | lib.EP_BIAS | lib.EP_LRELU, dscale=torch.ones(n, cout, device=dev),
                                bias=torch.zeros(cout, device=dev), noise=torch.randn(h, w, device=dev), noise_w=0.1, err=err)
        bytes_alg = 2 * n * h * w * (cin + cout)
    elif mode == "dgrad":
        wt = (torch.randn(9 * cout, cin, generator=g, device=dev) / math.sqrt(9 * cin)).bfloat16()
        out = torch.empty(n, h, w, cout, device=dev, dtype=torch.bfloat16)
        xin = torch.randn(n, h, w, cout, generator=g, device=dev).bfloat16()
        gs = torch.zeros(n, cout, device=dev)
        d = lib.make_igemm_desc(x, n, h, w, cin, 1, wt, 1, 9 * cout, out, h, w, cout, 1, lib.pick_block_n(cout), lib.conv3x3_dgrad_taps(cout),
                                flags=lib.EP_GSDOT | lib.EP_COLSCALE, xin=xin, colscale=torch.ones(n, cout, device=dev), gs=gs, err=err)
        bytes_alg = 2 * n * h * w * (cin + 2 * cout)
    elif mode == "fup":      # fused upsample conv forward: x (h,w,cin) -> (2h,2w,cout)
        wt = (torch.randn(n, 36 * cout, cin, generator=g, device=dev) / math.sqrt(9 * cin)).bfloat16()
        out = torch.empty(n, 2 * h, 2 * w, cout, device=dev, dtype=torch.bfloat16)
        d = lib.make_igemm_desc(x, n, h, w, cin, 1, wt, n, 36 * cout, out, h, w, 4 * cout, 1, lib.pick_block_n(4 * cout), lib.conv3x3_taps(4 * cout),
                                flags=lib.EP_DSCALE | lib.EP_NOISE | lib.EP_BIAS | lib.EP_LRELU, dscale=torch.ones(n, cout, device=dev),
                                bias=torch.zeros(cout, device=dev), noise=torch.randn(2 * h, 2 * w, device=dev), noise_w=0.1, err=err, out_d2s=1)
        bytes_alg = 2 * n * h * w * (cin + 4 * cout)
    elif mode == "fupb":     # its data gradient: g (2h,2w,cin) -> (h,w,cout) ; here cin = channels of the FINE gradient
        x = torch.randn(n, 2 * h, 2 * w, cin, generator=g, device=dev).bfloat16()
        wt = (torch.randn(9 * cout, 4 * cin, generator=g, device=dev) / math.sqrt(9 * cin)).bfloat16()
        out = torch.zeros(n, h, w, cout, device=dev, dtype=torch.bfloat16)
        xin = torch.randn(n, h, w, cout, generator=g, device=dev).bfloat16()
        gs = torch.zeros(n, cout, device=dev)
        d = lib.make_igemm_desc(x, n, h, w, 4 * cin, 1, wt, 1, 9 * cout, out, h, w, cout, 1, lib.pick_block_n(cout), lib.conv3x3_dgrad_taps(cout),
                                flags=lib.EP_GSDOT | lib.EP_COLSCALE | lib.EP_ACCUM, xin=xin, colscale=torch.ones(n, cout, device=dev), gs=gs,
                                err=err, a_s2d=1)
        bytes_alg = 2 * n * h * w * (4 * cin + 3 * cout)
    else:
        wt = (torch.randn(n, 9 * cout, cin, generator=g, device=dev) / math.sqrt(9 * cin)).bfloat16()
        out = torch.empty(n, 4, h + 1, w + 1, cout, device=dev, dtype=torch.bfloat16)
        d = lib.make_igemm_desc(x, n, h, w, cin, 1, wt, n, 9 * cout, out, h + 1, w + 1, cout, 4, lib.pick_block_n(cout, 4), lib.tconv_taps(cout), err=err)
        bytes_alg = 2 * n * h * w * (cin + 4 * cout)
    d.stages = stages
    if os.environ.get('SFK_BN'):
        d.block_n = int(os.environ['SFK_BN'])
    if os.environ.get('SFK_FLAGS') is not None:
        d.flags = int(os.environ['SFK_FLAGS'])
    if os.environ.get('SFK_ROLES'):
        d.flags |= lib.EP_PROFILE
        lib.role_cycles(True)
    for _ in range(2):
        lib.igemm(d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        lib.igemm(d)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = lib.igemm_flops(d)
    if os.environ.get('SFK_ROLES'):
        rc = lib.role_cycles(True)
        tiles = max(rc[7], 1)
        names = ['prod_wait', 'prod_total', 'mma_wait_data', 'mma_wait_acc', 'mma_total', 'epi_wait', 'epi_total']
        print('  role cycles per tile: ' + ', '.join(f'{n}={v / tiles:.0f}' for n, v in zip(names, rc[:7])) + f'  (tiles={tiles})')
    print(f"stages={stages} {mode} n={n} {h}x{w} cin={cin} cout={cout}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s  {bytes_alg/ms/1e6:.1f} GB/s (algorithmic)  err={err.item()}")

if __name__ == "__main__":
    main()
