"""TEST INFRASTRUCTURE — GIMM-VFI-F fixtures from the UNMODIFIED reference (oracle/ref_shim_f.py; build container only).

    python oracle/make_golden_f.py

Writes
  tests/golden/state_dict_spec_f.json      the 639-key state_dict layout of the reference GIMMVFI_F (key, shape, dtype)
  tests/golden/f_<case>.npz                per case: the flow estimator's outputs (`cal_bidirection_flow`, gimmvfi_f.py:114-138: flows,
                                           context features, fnet maps — the inputs of everything downstream) and the model's outputs
The weights are the reference constructors' own seeded random init (torch.manual_seed(0); no checkpoints offline) and are NOT stored:
tests rebuild the synthesis-side weights from the seed through gimmvfi_b200.weights.random_state_dict_f (checked here to load with
strict=True), the flow estimator's outputs travel in the fixture.
PARITY UNPINNED AT THE TIMM BOUNDARY (ref_shim_f.py): the Twins-SVT arithmetic is the reference tree's vendored copy of timm's file."""
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

import ref_shim_f  # noqa: E402
from gimmvfi_b200.synth import synth_batch  # noqa: E402
from gimmvfi_b200.weights import random_state_dict_f  # noqa: E402

# name, B, H, W, timesteps, ds_factor, input seed
CASES = [
    ("f_128x160_t0.5", 1, 128, 160, [0.5], None, 3),
    ("f_ds0.5_256x256_t0.25_0.75", 1, 256, 256, [0.25, 0.75], 0.5, 5),
]


def main():
    torch.set_grad_enabled(False)
    out_dir = os.path.join(ROOT, "tests", "golden")
    model = ref_shim_f.build_reference_model_f(seed=0)
    ref_sd = model.state_dict()
    spec = [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in ref_sd.items()]
    with open(os.path.join(out_dir, "state_dict_spec_f.json"), "w") as f:
        json.dump(spec, f)
    # the synthesis-side weights the tests use: seeded, independent of the reference constructors; the flow estimator keeps its own init
    sd = random_state_dict_f(0, flow_estimator={k: v for k, v in ref_sd.items() if k.startswith("flow_estimator.")})
    model.load_state_dict(sd, strict=True)
    manifest = {}
    for name, B, H, W, ts, ds, seed in CASES:
        xs = synth_batch(B, H, W, seed=seed)
        ratio = 1.0 if ds is None else ds
        coord = [(model.sample_coord_input(B, (H, W), [t], device=xs.device, upsample_ratio=ratio), None) for t in ts]
        tt = [t * torch.ones(B) for t in ts]
        # the estimator's outputs on the (possibly down-scaled) frames, exactly as forward() computes them (gimmvfi_f.py:304-328)
        x_net = xs
        if ds is not None:
            from ref_shim import load_reference_modules

            resize = load_reference_modules()["fi_utils"].resize
            x_net = torch.cat([resize(xs[:, :, 0], scale_factor=ds).unsqueeze(2), resize(xs[:, :, 1], scale_factor=ds).unsqueeze(2)], 2)
        _, flows, _, feats0, feats1, corr_fn, _ = model.cal_bidirection_flow(255 * x_net[:, :, 0], 255 * x_net[:, :, 1])
        # BidirCorrBlock does not keep its inputs: recompute fnet0 / fnet1 the way cal_bidirection_flow obtains them
        f01, features0, fnet0 = model.flow_estimator(255 * x_net[:, :, 0], 255 * x_net[:, :, 1], return_feat=True, iters=None)
        f10, features1, fnet1 = model.flow_estimator(255 * x_net[:, :, 1], 255 * x_net[:, :, 0], return_feat=True, iters=None)
        assert torch.equal(torch.stack([f01[0], f10[0]], 2), flows)
        out = model(xs, coord, t=tt, ds_factor=ds)
        arrays = dict(flows=flows.numpy(), feat4_0=features0[0].numpy(), feat4_1=features1[0].numpy(), feat8_0=features0[1].numpy(),
                      feat8_1=features1[1].numpy(), fnet_0=fnet0.numpy(), fnet_1=fnet1.numpy(), raft_flow=out["raft_flow"].numpy())
        for i in range(len(ts)):
            arrays["imgt_pred_%d" % i] = out["imgt_pred"][i].numpy()
            ft = out["flowt"][i]
            arrays["flowt_%d" % i] = (ft if ft.dim() == 4 else ft[None]).numpy()
            arrays["img_warp_4_%d" % i] = out["other_pred"][i][0].numpy()
            arrays["flowt0_4_%d" % i] = out["flowt0_pred"][i][1].numpy()
            arrays["ninrflow_%d" % i] = out["ninrflow"][i].numpy()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrays)
        manifest[name] = dict(B=B, H=H, W=W, timesteps=ts, ds_factor=ds, input_seed=seed, weight_seed=0)
        print(name, "done; |flow| max %.2f" % float(flows.abs().max()), flush=True)
    with open(os.path.join(out_dir, "manifest_f.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
