"""GPU micro-benchmark of the tcgen05 conv kernel on the dominant layer shapes (CUDA events).
usage: tc_microbench.py [plain|split|split16|all] [max_shapes]     (split = 3xTF32, split16 = 3xF16 form of the split kernel)"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import gpu_ops as K
from gimmvfi_b200._lib import default_lib, view_of

dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "all"
limit = int(sys.argv[2]) if len(sys.argv) > 2 else 99
PLAIN = [(256, 256, 3, 3, 1088, 1920, 1), (128, 128, 1, 1, 1088, 1920, 1), (32, 32, 3, 3, 1088, 1920, 2), (64, 64, 3, 3, 1088, 1920, 1)]
SPLIT = [(384, 128, 1, 5, 136, 240, 2), (256, 192, 3, 3, 136, 240, 2), (64, 64, 3, 3, 544, 960, 2), (324, 256, 1, 1, 136, 240, 2)]
jobs = ([(s, 0) for s in PLAIN] if which in ("plain", "all") else []) + ([(s, 1) for s in SPLIT] if which in ("split", "all") else []) + \
       ([(s, 2) for s in SPLIT] if which in ("split16", "all") else [])
lib = default_lib()
for (cin, cout, kh, kw, H, W, n), split in jobs[:limit]:
    x = torch.randn(n, H, W, cin, device=dev)
    w = torch.randn(cout, cin, kh, kw, device=dev) / (cin * kh * kw) ** 0.5
    pw = K.pack_weight_tc(w)
    pws, wsc = K.pack_weight_tc_split_f16(w) if split == 2 else (None, 1.0)
    out = torch.empty(n, H, W, cout, device=dev)
    bb = torch.zeros((cout + 31) // 32 * 32 + 256, device=dev)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def call():
        lib.check(lib.dll.gimmvfi_op_conv2d_tc(C.byref(view_of(x)), None, C.c_void_p(pw.data_ptr()), C.c_void_p(bb.data_ptr()), cin, cout, kh, kw,
                                               0, None, None, 0, None, None, None, None, int(split != 0), C.byref(view_of(out)),
                                               C.c_void_p(pws.data_ptr()) if pws is not None else None, wsc, s))

    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2.0 * n * H * W * cout * cin * kh * kw
    print("conv_tc%s c%d>%d k%dx%d @%dx%dx%d: %.3f ms  %.1f TFLOP/s" % (["", " 3xTF32", " 3xF16"][split], cin, cout, kh, kw, n, H, W, ms, fl / ms / 1e9), flush=True)
