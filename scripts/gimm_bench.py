"""Throughput of the standalone motion model (gimmvfi_b200.GIMM == reference GIMM.forward, gimm.py:129-214): flows in,
normalised flow at t out.  CUDA events, inputs resident on the GPU; the CPU oracle port is timed on a bounded sample.
usage: gimm_bench.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch

import gimmvfi_r_oracle as O
from gimmvfi_b200 import GIMM
from gimmvfi_b200.synth import synth_flow_pair

dev = "cuda"
m = GIMM(seed=0).to(dev).eval()
sd = {k: v.cpu() for k, v in m.state_dict().items()}
for (B, H, W, T) in [(1, 256, 448, 1), (8, 256, 448, 1), (1, 1088, 1920, 1), (1, 1088, 1920, 7)]:
    ori = synth_flow_pair(B, H, W, seed=1)
    xs, _ = O.normalize_flow(ori)
    ts = [0.5] if T == 1 else [i / (T + 1) for i in range(1, T + 1)]
    coord = [m.sample_coord_input(B, (H, W), [t], device=dev) for t in ts]
    tt = [t * torch.ones(B, device=dev) for t in ts]
    xs_d, ori_d = xs.to(dev), ori.to(dev)
    for mode in (1, 0):
        m.tensor_cores = mode
        for _ in range(3):
            m(xs_d, coord, True, ori_d, tt)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            out = m(xs_d, coord, True, ori_d, tt)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("GIMM B=%d %dx%d T=%d mode %d (%s): %.2f ms/call, %.1f flow fields/s" % (B, H, W, T, mode, "TF32 tcgen05" if mode else "fp32 CUDA cores", ms,
                                                                              B * T / ms * 1e3), flush=True)
    if (B, H, W, T) == (1, 256, 448, 1):
        torch.set_num_threads(min(32, os.cpu_count()))
        with torch.no_grad():
            O.gimm_forward(sd, xs, [c.cpu() for c in coord], ori, [t.cpu() for t in tt])
            t0 = time.perf_counter()
            for _ in range(3):
                ref = O.gimm_forward(sd, xs, [c.cpu() for c in coord], ori, [t.cpu() for t in tt])
            cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
        err = (out[0].cpu() - ref[0]).abs().max().item()
        print("   CPU oracle port (== reference GIMM, %d threads): %.1f ms/call; GPU fp32 vs oracle max|d| %.2e" % (torch.get_num_threads(), cpu_ms, err), flush=True)
