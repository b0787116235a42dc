"""GPU micro-benchmark of the tcgen05 conv kernel on the dominant layer shapes (CUDA events)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import gpu_ops as K

dev = "cuda"
shapes = [(256, 256, 3, 1088, 1920, 1), (128, 128, 1, 1088, 1920, 1), (32, 32, 3, 1088, 1920, 2), (64, 64, 3, 1088, 1920, 1), (128, 128, 3, 272, 480, 1)]
if len(sys.argv) > 1:
    shapes = shapes[: int(sys.argv[1])]
for cin, cout, k, H, W, n in shapes:
    x = torch.randn(n, H, W, cin, device=dev)
    w = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
    b = torch.randn(cout, device=dev)
    for _ in range(2):
        K.conv2d_tc(x, w, b)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # weights are re-packed by the helper each call: time only the kernel by pre-packing
    pw = K.pack_weight_tc(w)
    import ctypes as C
    from gimmvfi_b200._lib import default_lib, view_of
    lib = default_lib()
    out = torch.empty(n, H, W, cout, device=dev)
    bb = torch.zeros((cout + 31) // 32 * 32 + 256, device=dev); bb[:cout] = b
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    def call():
        lib.check(lib.dll.gimmvfi_op_conv2d_tc(C.byref(view_of(x)), None, C.c_void_p(pw.data_ptr()), C.c_void_p(bb.data_ptr()), cin, cout, k, k, 0, None, None, 0, None, None, None, None, 0, C.byref(view_of(out)), s))
    call(); torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        call()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 2.0 * n * H * W * cout * cin * k * k
    byt = 4.0 * n * H * W * (cin + cout)
    print("conv_tc c%d>%d k%d @%dx%dx%d: %.3f ms  %.1f TFLOP/s  %.0f GB/s(algorithmic act bytes)" % (cin, cout, k, n, H, W, ms, fl / ms / 1e9, byt / ms / 1e6), flush=True)
