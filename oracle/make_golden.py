"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED
reference (imported from /root/reference through oracle/ref_shim.py) on seeded
synthetic inputs and seeded random weights.  Runs only in the build container.

    python oracle/make_golden.py

Each fixture holds the reference's outputs (sub-sampled where large), plus the
max |Δ| of the travelling oracle (oracle/gimmvfi_r_oracle.py) against the
reference on the same case — the "pin" of the oracle.
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
warnings.filterwarnings("ignore")

import gimmvfi_r_oracle as O  # noqa: E402
import ref_shim  # noqa: E402
from gimmvfi_b200.synth import synth_batch  # noqa: E402
from gimmvfi_b200.weights import random_state_dict  # noqa: E402

# name, B, H, W, timesteps, ds_factor, input seed, weight seed, spatial stride for storage
CASES = [
    ("r_128x160_t0.5", 1, 128, 160, [0.5], None, 3, 0, 1),
    ("r_b2_128x192_t0.25_0.75", 2, 128, 192, [0.25, 0.75], None, 4, 0, 2),
    ("r_ds0.5_256x320_t0.5", 1, 256, 320, [0.5], 0.5, 5, 0, 2),
    ("r_256x448_t0.5", 1, 256, 448, [0.5], None, 6, 0, 4),
]


# GIMM standalone: name, B, H, W, timesteps, flow seed
GIMM_CASES = [("gimm_64x96_t0.25_0.75", 1, 64, 96, [0.25, 0.75], 1), ("gimm_b2_72x80_t0.5", 2, 72, 80, [0.5], 2)]


def sub(t, s):
    return t[..., ::s, ::s].contiguous().numpy()


def main():
    torch.set_grad_enabled(False)
    out_dir = os.path.join(ROOT, "tests", "golden")
    manifest = {}
    models = {}
    only_gimm = "--only-gimm" in sys.argv      # keep the GIMM-VFI-R fixtures, (re)generate the GIMM-standalone ones
    if only_gimm:
        with open(os.path.join(out_dir, "manifest.json")) as f:
            manifest = json.load(f)
    for name, B, H, W, ts, ds, iseed, wseed, stride in ([] if only_gimm else CASES):
        if wseed not in models:
            sd = random_state_dict(wseed)
            models[wseed] = (ref_shim.build_reference_model(sd), sd)
        model, sd = models[wseed]
        xs = synth_batch(B, H, W, seed=iseed)
        ratio = 1.0 if ds is None else ds
        coord = [(model.sample_coord_input(B, (H, W), [t], device=xs.device, upsample_ratio=ratio), None) for t in ts]
        tt = [t * torch.ones(B) for t in ts]
        ref = model(xs, coord, t=tt, ds_factor=ds)
        ora = O.gimmvfi_r_forward(sd, xs, [(O.sample_coord_input(B, (H, W), [t], ratio), None) for t in ts], tt, ds_factor=ds)
        pin = 0.0
        for k in ("imgt_pred", "flowt", "ninrflow"):
            for a, b in zip(ref[k], ora[k]):
                pin = max(pin, (a - b).abs().max().item())
        pin = max(pin, (ref["raft_flow"] - ora["raft_flow"]).abs().max().item())
        arrays = {"stride": np.int32(stride)}
        for i in range(len(ts)):
            arrays["imgt_pred_%d" % i] = sub(ref["imgt_pred"][i], stride)
            arrays["flowt_%d" % i] = sub(ref["flowt"][i], stride)
            arrays["img_warp_4_%d" % i] = sub(ref["other_pred"][i][0], stride * 2)
            arrays["flowt0_4_%d" % i] = sub(ref["flowt0_pred"][i][1], stride)
            arrays["imgt_pred_sum_%d" % i] = np.float64(ref["imgt_pred"][i].double().sum().item())
        arrays["raft_flow"] = sub(ref["raft_flow"], stride * 2)
        arrays["raft_flow_absmax"] = np.float32(ref["raft_flow"].abs().max().item())
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrays)
        manifest[name] = dict(B=B, H=H, W=W, timesteps=ts, ds_factor=ds, input_seed=iseed, weight_seed=wseed,
                              stride=stride, oracle_vs_reference_max_abs=pin)
        print(name, "oracle-vs-reference max|Δ| =", pin, flush=True)
    # ---- GIMM standalone (gimm.py:129-214): the unmodified reference GIMM on seeded flows
    from gimmvfi_b200.synth import synth_flow_pair  # noqa: E402
    sd = random_state_dict(0)
    gimm = ref_shim.build_reference_gimm(sd)
    spec = [[k, list(v.shape)] for k, v in gimm.state_dict().items()]
    with open(os.path.join(out_dir, "state_dict_spec_gimm.json"), "w") as f:
        json.dump(spec, f)
    for name, B, H, W, ts, fseed in GIMM_CASES:
        ori = synth_flow_pair(B, H, W, seed=fseed)
        xs, _ = O.normalize_flow(ori)
        coord = [O.sample_coord_input(B, (H, W), [t], 1.0) for t in ts]
        tt = [t * torch.ones(B) for t in ts]
        ref = gimm(xs, coord, True, ori, tt)
        ora = O.gimm_forward(sd, xs, coord, ori, tt)
        pin = max((a - b).abs().max().item() for a, b in zip(ref, ora))
        arrays = {"out_%d" % i: ref[i].numpy() for i in range(len(ts))}
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrays)
        manifest[name] = dict(kind="gimm", B=B, H=H, W=W, timesteps=ts, flow_seed=fseed, weight_seed=0, oracle_vs_reference_max_abs=pin)
        print(name, "oracle-vs-reference max|Δ| =", pin, flush=True)
    with open(os.path.join(out_dir, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
