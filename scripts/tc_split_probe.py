"""Timing probe of the 3xTF32 conv kernel on the RAFT shapes under the GIMMVFI_TC_* experiment knobs (one process per
knob set, since the knobs are read once).  usage: tc_split_probe.py  (spawns itself)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIG = [(256, 256, 3, 3, 1088, 1920, 1), (64, 64, 3, 3, 1088, 1920, 1)]
SHAPES = ([] if os.environ.get("PROBE_BIG") else [(384, 128, 1, 5, 136, 240, 2), (256, 192, 3, 3, 136, 240, 2), (128, 256, 3, 3, 136, 240, 2), (64, 64, 3, 3, 544, 960, 2)]) + (BIG if os.environ.get("PROBE_BIG") else [])

if os.environ.get("PROBE_GRU"):   # K sweep of the SepConvGRU gate shape (3xF16 only): what is the per-tile cost that does not scale with K?
    SHAPES = [(384, 128, 1, 5, 136, 240, 2), (256, 128, 1, 5, 136, 240, 2), (128, 128, 1, 5, 136, 240, 2), (256, 256, 1, 5, 136, 240, 2), (128, 128, 1, 1, 136, 240, 2)]

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import ctypes as C
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import gpu_ops as K
    from gimmvfi_b200._lib import default_lib, view_of
    lib = default_lib()
    res = []
    stall = torch.zeros(148 * 16, dtype=torch.int64, device="cuda") if os.environ.get("PROBE_STALL") else None
    if stall is not None:
        os.environ["GIMMVFI_TC_STALL_BUF"] = str(stall.data_ptr())
    for (cin, cout, kh, kw, H, W, n) in SHAPES:
        for split in ((0,) if os.environ.get("PROBE_BIG") else (2,) if os.environ.get("PROBE_GRU") else (0, 1, 2)):   # 0 plain TF32, 1 3xTF32, 2 3xF16
            f16 = bool(os.environ.get("PROBE_F16"))
            x = torch.randn(n, H, W, cin, device="cuda")
            w = torch.randn(cout, cin, kh, kw, device="cuda") / (cin * kh * kw) ** 0.5
            pw = K.pack_weight_tc(w)
            pws, wsc = K.pack_weight_tc_split_f16(w) if split == 2 else (None, 1.0)
            out = torch.empty(n, H, W, cout, device="cuda")
            if f16:
                x = x.half(); out = out.half(); pwh = K.pack_weight_tc_f16(w)
            bb = torch.zeros((cout + 31) // 32 * 32 + 256, device="cuda")
            epi = os.environ.get("PROBE_EPI", "plain")
            side = [torch.rand(n, H, W, cout, device="cuda") for _ in range(3)]
            s = C.c_void_p(torch.cuda.current_stream().cuda_stream)

            def call():
                if f16:
                    lib.check(lib.dll.gimmvfi_op_conv2d_tc_f16(C.byref(view_of(x)), None, C.c_void_p(pwh.data_ptr()), C.c_void_p(pw.data_ptr()), C.c_void_p(bb.data_ptr()),
                                                               cin, cout, kh, kw, 0, None, None, 0, None, 3, C.byref(view_of(out)), s))
                    return
                if epi == "q":     # the GRU candidate's epilogue: + hoisted context term, tanh, h = (1 - z) h + z q   (raft/update.py:58-66)
                    lib.check(lib.dll.gimmvfi_op_conv2d_tc(C.byref(view_of(x)), None, C.c_void_p(pw.data_ptr()), C.c_void_p(bb.data_ptr()), cin, cout, kh, kw,
                                                           0, None, C.byref(view_of(side[0])), 5, None, None, C.byref(view_of(side[1])), C.byref(view_of(side[2])),
                                                           int(split != 0), C.byref(view_of(out)), C.c_void_p(pws.data_ptr()) if pws is not None else None, wsc, s))
                    return
                lib.check(lib.dll.gimmvfi_op_conv2d_tc(C.byref(view_of(x)), None, C.c_void_p(pw.data_ptr()), C.c_void_p(bb.data_ptr()), cin, cout, kh, kw,
                                                       5 if epi == "tanh" else 0, None, None, 0, None, None, None, None, int(split != 0), C.byref(view_of(out)),
                                                       C.c_void_p(pws.data_ptr()) if pws is not None else None, wsc, s))
            for _ in range(3):
                call()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                call()
            e1.record()
            torch.cuda.synchronize()
            res.append("%s%d>%d k%dx%d: %.3f" % (["P ", "S ", "S16 "][split], cin, cout, kh, kw, e0.elapsed_time(e1) / 20))
            if stall is not None:
                stall.zero_(); call(); torch.cuda.synchronize()
                m = stall.view(148, 16).double().mean(0).tolist()
                res.append("\n    stall cycles/CTA: total %.0f | producer-empty %.0f | mma-operands %.0f | mma-acc %.0f | epi-tfull %.0f | epi-output %.0f | splitter-full %.0f | phase1 %.0f phase2 %.0f chunks %.1f\n   "
                           % (m[6], m[0], m[1], m[2], m[3], m[4], m[5], m[8], m[9], m[10]))
    print(" | ".join(res), flush=True)
    sys.exit(0)

KNOBS = ([] if os.environ.get("PROBE_GRU") else [{"PROBE_STALL": "1"}, {"PROBE_STALL": "1", "GIMMVFI_TC_PAIR": "1"}]) + [dict(kv.split("=") for kv in a.split(",")) for a in sys.argv[1:]]
for kn in KNOBS:
    env = dict(os.environ); env.update(kn)
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True, timeout=300)
    print(kn, "->", r.stdout.strip() or r.stderr.strip()[-300:], flush=True)
