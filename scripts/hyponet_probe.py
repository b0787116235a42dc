"""GPU box: the fused HypoNet kernels at the bench size (one call = 1088x1920 pixels), CUDA-event timed; --once for ncu."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from gimmvfi_b200 import EngineHandle
from gimmvfi_b200._lib import view_of
from gimmvfi_b200.weights import random_state_dict

once = "--once" in sys.argv
dev = "cuda"
H, W = 1088, 1920
eng = EngineHandle(dev)
eng.load_state_dict(random_state_dict(0))
lat = torch.randn(1, H, W, 32, device=dev) * 0.7
ys = -1 + 2 * (torch.arange(H, device=dev) + 0.5) / H
xs = -1 + 2 * (torch.arange(W, device=dev) + 0.5) / W
coord = torch.stack([torch.full((H, W), 0.5, device=dev), ys.view(H, 1).expand(H, W), xs.view(1, W).expand(H, W)], -1).contiguous()
out = torch.empty(1, H, W, 2, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
flops = 2.0 * H * W * (35 * 128 + 3 * 128 * 128 + 128 * 2)
for cls in (1, 0):
    call = lambda: eng.lib.check(eng.lib.dll.gimmvfi_op_hyponet(eng._h, C.byref(view_of(lat)), C.c_void_p(coord.data_ptr()), C.byref(view_of(out)), cls, st), eng._h)
    call()
    torch.cuda.synchronize()
    if once:
        continue
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("hyponet fused %s @%dx%d: %.3f ms  %.1f TFLOP/s algorithmic (x3 on the tensor pipe for the fp32-class kernel)" % ("fp32-class (3xF16)" if cls else "TF32/half", H, W, ms, flops / ms / 1e9), flush=True)
