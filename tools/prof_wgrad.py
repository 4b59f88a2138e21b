"""Time sfk_conv3x3_wgrad on the conv shapes of the path (8 pairs): TFLOP/s per launch."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sfattack import lib
dev = torch.device("cuda:0")
def timeit(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
shapes = [("vgg conv1_2 (16 img)", 16, 256, 64, 64, False), ("vgg conv2_2", 16, 128, 128, 128, False), ("vgg conv3_2", 16, 64, 256, 256, False),
          ("vgg conv4_2", 16, 32, 512, 512, False), ("gen 64^2 512->512 per-sample", 8, 64, 512, 512, True),
          ("gen 256^2 128->128 per-sample", 8, 256, 128, 128, True), ("gen 1024^2 32->32 per-sample", 8, 1024, 32, 32, True)]
for name, n, h, cin, cout, ps in shapes:
    x = torch.randn(n, h, h, cin, device=dev).bfloat16(); gz = torch.randn(n, h, h, cout, device=dev).bfloat16()
    dw = torch.zeros(n if ps else 1, 9, cout, cin, device=dev)
    ms = timeit(lambda: lib.conv3x3_wgrad(x, gz, dw=dw, per_sample=ps))
    fl = 2.0 * n * h * h * 9 * cin * cout
    print(f"{name:34s} {ms*1e3:8.1f} us  {fl/ms/1e9:8.1f} TFLOP/s  ({(x.numel()+gz.numel())*2/ms/1e6:7.1f} GB/s of operands)")
