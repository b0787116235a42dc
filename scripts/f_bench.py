"""GIMM-VFI-F on the GPU: BASELINE config 3 (2K pair, ds_factor 0.5, N = 8 -> 7 timesteps) through GIMMVFI_F.forward with the native
FlowFormer estimator; device-timed steps + the per-kernel CUDA-event profile of one forward.
    python scripts/f_bench.py [--h 1088 --w 2048 --ds 0.5 --n 8 --steps 3 --profile-json out.json]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gimmvfi_b200.model_f import GIMMVFI_F  # noqa: E402
from gimmvfi_b200.synth import synth_batch  # noqa: E402
from gimmvfi_b200.weights import random_state_dict_f  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--h", type=int, default=1088)
    ap.add_argument("--w", type=int, default=2048)
    ap.add_argument("--ds", type=float, default=0.5)
    ap.add_argument("--n", type=int, default=8)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--mode", type=int, default=4)
    ap.add_argument("--profile-json", default="")
    a = ap.parse_args()
    dev = "cuda"
    m = GIMMVFI_F(seed=0).to(dev).eval()
    m.load_state_dict(random_state_dict_f(0), strict=True)
    m.tensor_cores = a.mode
    m.aux_outputs = False
    ds = a.ds if a.ds > 0 else None
    ts = [i / a.n for i in range(1, a.n)]
    xs = synth_batch(1, a.h, a.w, seed=1).to(dev)
    coord = [(m.sample_coord_input(1, (a.h, a.w), [t], device=dev, upsample_ratio=ds or 1.0), None) for t in ts]
    tt = [t * torch.ones(1, device=dev) for t in ts]
    for _ in range(a.warmup):
        out = m(xs, coord, t=tt, ds_factor=ds)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out = m(xs, coord, t=tt, ds_factor=ds)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    eng = m.engine
    eng.set_profile(True)
    out = m(xs, coord, t=tt, ds_factor=ds)
    prof = eng.profile()
    eng.set_profile(False)
    tot = sum(v["ms"] for v in prof.values())
    est = sum(v["ms"] for k, v in prof.items())
    line = {"config": "f_%dx%d_ds%s_n%d" % (a.h, a.w, a.ds, a.n), "mode": a.mode, "ms_per_pair": ms, "frames_per_s": (a.n - 1) / (ms / 1e3),
            "launches": eng.last_launches, "profiled_ms": tot, "finite": bool(torch.isfinite(out["imgt_pred"][0]).all()),
            "mem_GB": torch.cuda.max_memory_allocated() / 2**30}
    print(json.dumps(line))
    top = sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:40]
    for k, v in top:
        print("%9.3f ms %6d x  %s" % (v["ms"], v["launches"], k))
    if a.profile_json:
        json.dump({"line": line, "kernels": prof}, open(a.profile_json, "w"), indent=1)


if __name__ == "__main__":
    main()
