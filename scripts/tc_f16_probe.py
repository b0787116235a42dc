"""Timing of the fp16-trunk conv (256 -> 256, 3x3, 1088x1920) with the epilogues the network uses (conv1 / conv3 / conv5 of
ResBlock(256, 64), fi_components.py:97-154).  usage: tc_f16_probe.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gpu_ops as K
from gimmvfi_b200._lib import default_lib, view_of

lib = default_lib()
n, H, W = 1, 1088, 1920
x = torch.randn(n, H, W, 256, device="cuda").half()
s1 = torch.randn(n, H, W, 64, device="cuda").half()
res = torch.randn(n, H, W, 256, device="cuda").half()
w = torch.randn(256, 256, 3, 3, device="cuda") / (256 * 9) ** 0.5
pw, pwh = K.pack_weight_tc(w), K.pack_weight_tc_f16(w)
bb = torch.zeros(1024, device="cuda")
slope = torch.full((256,), 0.25, device="cuda")
out = torch.empty(n, H, W, 256, device="cuda", dtype=torch.float16)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
V = lambda v: C.byref(v) if v is not None else None
stall = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
os.environ["GIMMVFI_TC_STALL_BUF"] = str(stall.data_ptr())
res32 = res.float()
CASES = {
    "plain (no act)": dict(a0=view_of(x), a1=None, act1=0, res=None, act2=0),
    "conv1: PReLU": dict(a0=view_of(x), a1=None, act1=3, res=None, act2=0),
    "conv3: 192+64 segments, PReLU": dict(a0=view_of(x, channels=192), a1=view_of(s1), act1=3, res=None, act2=0),
    "conv5: 192+64 segments, + residual, PReLU": dict(a0=view_of(x, channels=192), a1=view_of(s1), act1=0, res=view_of(res), act2=3),
    "256 + residual (half), no act": dict(a0=view_of(x), a1=None, act1=0, res=view_of(res), act2=0),
    "256 + residual (fp32), no act": dict(a0=view_of(x), a1=None, act1=0, res=view_of(res32), act2=0, res32=True),
    "256, act2 PReLU only": dict(a0=view_of(x), a1=None, act1=0, res=None, act2=3),
}
for name, c in CASES.items():
    def call():
        lib.check(lib.dll.gimmvfi_op_conv2d_tc_f16(V(c["a0"]), V(c["a1"]), P(pwh), P(pw), P(bb), 256, 256, 3, 3, c["act1"], P(slope) if c["act1"] == 3 else None,
                                                   V(c["res"]), c["act2"], P(slope) if c["act2"] == 3 else None, 1 | 2 | (4 if (c["res"] is not None and not c.get("res32")) else 0),
                                                   C.byref(view_of(out)), st))
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    stall.zero_(); call(); torch.cuda.synchronize()
    m = stall.view(148, 16).double().mean(0).tolist()
    print("%-46s %.3f ms  %.0f TFLOP/s | cycles/CTA: epi-tfull-wait %.0f epi-output %.0f phase1 %.0f phase2 %.0f chunks %.0f mma-operand-wait(x2) %.0f"
          % (name, ms, 2.0 * n * H * W * 256 * 256 * 9 / ms / 1e9, m[3], m[4], m[8], m[9], m[10], m[1]), flush=True)
