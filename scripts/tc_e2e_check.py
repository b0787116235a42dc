"""GPU: end-to-end effect of the tcgen05 TF32 path (post-RAFT convs) on parity and speed."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch

import gimmvfi_r_oracle as O
from gimmvfi_b200 import GIMMVFI_R
from gimmvfi_b200.synth import synth_batch

dev = "cuda"
model = GIMMVFI_R(seed=0).to(dev).eval()
model.tensor_cores = False
sd = {k: v.cpu() for k, v in model.state_dict().items()}
eng = model.engine


def run(B, H, W, t=0.5, seed=3):
    xs = synth_batch(B, H, W, seed=seed)
    coord = [(model.sample_coord_input(B, (H, W), [t], device=dev), None)]
    out = model(xs.to(dev), coord, t=[t * torch.ones(B, device=dev)])
    torch.cuda.synchronize()
    return xs, out


for (B, H, W) in [(1, 128, 160), (1, 256, 448)]:
    res = {}
    for tc in (0, 1, 2):
        model.tensor_cores = tc
        xs, out = run(B, H, W)
        res[tc] = out
    with torch.no_grad():
        ref = O.gimmvfi_r_forward(sd, xs, [(O.sample_coord_input(B, (H, W), [0.5]), None)], [0.5 * torch.ones(B)])
    for tc in (0, 1, 2):
        d = (res[tc]["imgt_pred"][0].cpu() - ref["imgt_pred"][0]).abs()
        f = (res[tc]["flowt"][0].cpu() - ref["flowt"][0]).abs()
        print("%dx%d tc=%s imgt_pred max %.3e mean %.3e | flowt max %.3e mean %.3e" % (H, W, tc, d.max(), d.mean(), f.max(), f.mean()), flush=True)

H, W = 1088, 1920
outs = {}
for tc in (0, 1, 2):
    model.tensor_cores = tc
    for _ in range(2):
        xs, out = run(1, H, W, seed=100)
    t0 = time.perf_counter()
    for _ in range(3):
        xs, out = run(1, H, W, seed=100)
    dt = (time.perf_counter() - t0) / 3
    outs[tc] = out["imgt_pred"][0].clone()
    outs[("raft", tc)] = out["raft_flow"].clone()
    outs[("flowt", tc)] = out["flowt"][0].clone()
    eng.set_profile(True)
    run(1, H, W, seed=100)
    prof = eng.profile()
    eng.set_profile(False)
    tot = sum(v["ms"] for v in prof.values())
    print("1080p tc=%s: %.1f ms/forward (incl. synth+H2D), kernels sum %.1f ms" % (tc, dt * 1e3, tot))
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:(34 if tc == 2 else 8)]:
        extra = " %.1f TFLOP/s" % (v["work"] / v["ms"] / 1e9) if k.startswith(("conv2d", "corr_gemm")) else ""
        print("   %-62s %8.2f ms %4d launches%s" % (k, v["ms"], v["launches"], extra))
    with open(os.path.join(ROOT, "gpurun_out", "profile_1080p_tc%d.json" % int(tc)), "w") as f:
        json.dump(prof, f)
for m in (1, 2):
    d = (outs[m] - outs[0]).abs()
    print("1080p tc mode %d vs fp32 imgt_pred: max %.3e mean %.3e p99.99 %.3e" % (m, d.max(), d.mean(), torch.quantile(d.flatten()[::7], 0.9999)))
    r = (outs[("raft", m)] - outs[("raft", 0)]).abs(); f = (outs[("flowt", m)] - outs[("flowt", 0)]).abs()
    print("      raft_flow max %.3e mean %.3e | flowt max %.3e mean %.3e p99.99 %.3e" % (r.max(), r.mean(), f.max(), f.mean(), torch.quantile(f.flatten()[::3], 0.9999)))
