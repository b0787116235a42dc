"""TEST INFRASTRUCTURE — fixtures for the NATIVE FlowFormer estimator of GIMM-VFI-F, from the UNMODIFIED reference
(oracle/ref_shim_f.py; build container only, /root/reference does not travel).

    python oracle/make_golden_ff.py

The reference GIMMVFI_F is loaded (strict=True) with gimmvfi_b200.weights.random_state_dict_f(0) — every tensor, FlowFormer included,
comes from a seed, so the GPU tests rebuild the same weights without the reference — and run end to end.  Per case
tests/golden/ff_<case>.npz holds the estimator's products (cal_bidirection_flow, gimmvfi_f.py:114-138: both flows, context features,
fnet maps) and the model's outputs; manifest_ff.json records the case parameters and how far oracle/flowformer_oracle.py (the
restatement) is from the reference on the same inputs (its pin).
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

import flowformer_oracle as FO  # noqa: E402
import ref_shim_f  # noqa: E402
from gimmvfi_b200.synth import synth_batch  # noqa: E402
from gimmvfi_b200.weights import random_state_dict_f  # noqa: E402

# name, B, H, W, timesteps, ds_factor, input seed
CASES = [
    ("ff_128x160_t0.5", 1, 128, 160, [0.5], None, 3),
    ("ff_ds0.5_256x256_t0.25_0.75", 1, 256, 256, [0.25, 0.75], 0.5, 5),
    ("ff_b2_128x128_t0.5", 2, 128, 128, [0.5], None, 7),
]


def main():
    torch.set_grad_enabled(False)
    out_dir = os.path.join(ROOT, "tests", "golden")
    model = ref_shim_f.build_reference_model_f(seed=0)
    sd = random_state_dict_f(0)
    model.load_state_dict(sd, strict=True)
    manifest = {}
    only = sys.argv[1:]
    mpath = os.path.join(out_dir, "manifest_ff.json")
    if only and os.path.exists(mpath):
        manifest = json.load(open(mpath))
    for name, B, H, W, ts, ds, seed in CASES:
        if only and name not in only:
            continue
        xs = synth_batch(B, H, W, seed=seed)
        ratio = 1.0 if ds is None else ds
        coord = [(model.sample_coord_input(B, (H, W), [t], device=xs.device, upsample_ratio=ratio), None) for t in ts]
        tt = [t * torch.ones(B) for t in ts]
        x_net = xs
        if ds is not None:
            from ref_shim import load_reference_modules

            resize = load_reference_modules()["fi_utils"].resize
            x_net = torch.cat([resize(xs[:, :, 0], scale_factor=ds).unsqueeze(2), resize(xs[:, :, 1], scale_factor=ds).unsqueeze(2)], 2)
        im0, im1 = 255 * x_net[:, :, 0], 255 * x_net[:, :, 1]
        f01, features0, fnet0 = model.flow_estimator(im0, im1, return_feat=True, iters=None)
        f10, features1, fnet1 = model.flow_estimator(im1, im0, return_feat=True, iters=None)
        flows = torch.stack([f01[0], f10[0]], 2)
        # pin of the restatement: same weights, same inputs
        (o01, _), ocf, off = FO.flowformer_forward(sd, im0, im1)
        pin = dict(flow=float((o01 - f01[0]).abs().max()), feat8=float((ocf[1] - features0[1]).abs().max()), fnet=float((off - fnet0).abs().max()))
        out = model(xs, coord, t=tt, ds_factor=ds)
        assert torch.equal(out["raft_flow"], flows)
        arrays = dict(flows=flows.numpy(), feat4_0=features0[0].numpy(), feat4_1=features1[0].numpy(), feat8_0=features0[1].numpy(),
                      feat8_1=features1[1].numpy(), fnet_0=fnet0.numpy(), fnet_1=fnet1.numpy(), flow_low_01=f01[1].numpy(), flow_low_10=f10[1].numpy())
        for i in range(len(ts)):
            arrays["imgt_pred_%d" % i] = out["imgt_pred"][i].numpy()
            ft = out["flowt"][i]
            arrays["flowt_%d" % i] = (ft if ft.dim() == 4 else ft[None]).numpy()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrays)
        manifest[name] = dict(B=B, H=H, W=W, timesteps=ts, ds_factor=ds, input_seed=seed, weight_seed=0, flow_absmax=float(flows.abs().max()),
                              oracle_vs_reference=pin)
        print(name, "done; |flow| max %.2f; oracle vs reference %s" % (float(flows.abs().max()), pin), flush=True)
        with open(mpath, "w") as f:
            json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
