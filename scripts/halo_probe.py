"""GPU box: the halo-reuse 3x3 kernel vs the per-tap TMA kernel on the K-poor full-resolution layers (CUDA events); --once for ncu."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import gpu_ops as K

once = "--once" in sys.argv
dev = "cuda"
for (cin, cout, n, half) in ((32, 32, 2, False), (32, 64, 1, False), (64, 64, 1, True)):
    H, W = 1088, 1920
    x = torch.randn(n, H, W, cin, device=dev)
    w = torch.randn(cout, cin, 3, 3, device=dev) / (cin * 9) ** 0.5
    b = torch.zeros(cout, device=dev)
    xin = x.half() if half else x
    for name, fn in (("halo", lambda: K.conv2d_halo(xin, w, b, 2, out_half=half)),
                     ("per-tap", (lambda: K.conv2d_tc_f16(xin, w, b, 2, None, out_half=True)) if half else (lambda: K.conv2d_tc(xin, w, b, 2)))):
        fn(); torch.cuda.synchronize()
        if once:
            continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("3x3 c%d>%d @%dx%dx%d %s %-8s %.3f ms  %.1f TFLOP/s" % (cin, cout, n, H, W, "f16" if half else "tf32", name, ms, 2.0 * n * H * W * cin * cout * 9 / ms / 1e9), flush=True)
