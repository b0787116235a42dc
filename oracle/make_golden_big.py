"""TEST INFRASTRUCTURE — reference-generated fixtures at the BENCHMARKED sizes.

    python oracle/make_golden_big.py [case ...]

Runs the UNMODIFIED reference (imported from /root/reference through oracle/ref_shim.py, CPU, fp32) on the
configurations bench.py measures and writes sub-sampled outputs to tests/golden/big_*.npz (stride 8 on the
full-resolution outputs: a 1088x1920 frame becomes 136x240x3 floats).  Build container only (the reference does
not travel); the fixtures do, and `pytest -m gpu` checks the CUDA path against them.

Cases (BASELINE.json configs / SURVEY 8(c)):
  big_r_1088x1920_t0.5        config 2: the bench workload itself (seed 100 = bench.py's rank-0 pair)
  big_r_736x1280_t0.5         config 5: one 1280x720 pair padded to 736x1280
  big_r_ds0.5_1088x2048_T7    the reference's 2K video setting (README.md:87-96, video_Nx.py:164-181): ds_factor 0.5, N = 8
  big_r_demo_736x864_t0.5     the reference's own demo frames (demo/input_frames/000{20,28}.png), replicate-padded by InputPadder(…, 32)
The oracle-vs-reference pin at these sizes is recorded for the 736x1280 case (the restatement is size independent;
the small cases of make_golden.py pin it at 0.0 as well).
"""
import json
import os
import sys
import time
import warnings

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
warnings.filterwarnings("ignore")

import gimmvfi_r_oracle as O  # noqa: E402
import ref_shim  # noqa: E402
from gimmvfi_b200.synth import synth_batch  # noqa: E402
from gimmvfi_b200.weights import random_state_dict  # noqa: E402

STRIDE = 8
# name: (H, W, timesteps, ds_factor, input seed or "demo", pin the oracle too)
CASES = {
    "big_r_736x1280_t0.5": (736, 1280, [0.5], None, 5, True),
    "big_r_demo_736x864_t0.5": (736, 864, [0.5], None, "demo", False),
    "big_r_ds0.5_1088x2048_T7": (1088, 2048, [i / 8 for i in range(1, 8)], 0.5, 7, False),
    "big_r_1088x1920_t0.5": (1088, 1920, [0.5], None, 100, False),
}
DEMO = [os.path.join("/root/reference/demo/input_frames", f) for f in ("00020.png", "00028.png")]


def demo_frames_u8():
    from PIL import Image

    return np.stack([np.array(Image.open(p).convert("RGB")) for p in DEMO], 0)   # (2,720,844,3) uint8 RGB (video_Nx.py:46-50)


def demo_input(frames_u8):
    """video_Nx.py:46-50,151-156: /255, InputPadder(shape, 32).pad (replicate), stack on dim 2."""
    x = torch.from_numpy(frames_u8.copy()).permute(0, 3, 1, 2).float() / 255.0
    ht, wd = x.shape[-2:]
    ph, pw = (((ht // 32) + 1) * 32 - ht) % 32, (((wd // 32) + 1) * 32 - wd) % 32
    x = F.pad(x, [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2], mode="replicate")
    return torch.stack([x[0], x[1]], 1).unsqueeze(0).contiguous()


def sub(t, s):
    return t[..., ::s, ::s].contiguous().numpy()


def main():
    torch.set_grad_enabled(False)
    torch.set_num_threads(os.cpu_count())
    out_dir = os.path.join(ROOT, "tests", "golden")
    mpath = os.path.join(out_dir, "manifest_big.json")
    manifest = json.load(open(mpath)) if os.path.exists(mpath) else {}
    want = [a for a in sys.argv[1:] if not a.startswith("-")] or list(CASES)
    sd = random_state_dict(0)
    model = ref_shim.build_reference_model(sd)
    for name in want:
        H, W, ts, ds, seed, pin_oracle = CASES[name]
        arrays = {"stride": np.int32(STRIDE)}
        if seed == "demo":
            u8 = demo_frames_u8()
            arrays["frames_u8"] = u8
            xs = demo_input(u8)
            assert xs.shape[-2:] == (H, W), xs.shape
        else:
            xs = synth_batch(1, H, W, seed=seed)
        ratio = 1.0 if ds is None else ds
        coord = [(model.sample_coord_input(1, (H, W), [t], device=xs.device, upsample_ratio=ratio), None) for t in ts]
        tt = [t * torch.ones(1) for t in ts]
        t0 = time.time()
        ref = model(xs, coord, t=tt, ds_factor=ds)
        sec = time.time() - t0
        pin = None
        if pin_oracle:
            ora = O.gimmvfi_r_forward(sd, xs, [(O.sample_coord_input(1, (H, W), [t], ratio), None) for t in ts], tt, ds_factor=ds)
            pin = max(max((a - b).abs().max().item() for a, b in zip(ref[k], ora[k])) for k in ("imgt_pred", "flowt"))
            pin = max(pin, (ref["raft_flow"] - ora["raft_flow"]).abs().max().item())
            del ora
        for i in range(len(ts)):
            arrays["imgt_pred_%d" % i] = sub(ref["imgt_pred"][i], STRIDE)
            arrays["imgt_pred_sum_%d" % i] = np.float64(ref["imgt_pred"][i].double().sum().item())
            ft = ref["flowt"][i]
            arrays["flowt_%d" % i] = sub(ft if ft.dim() == 4 else ft[None], STRIDE)
        arrays["raft_flow"] = sub(ref["raft_flow"], STRIDE)
        arrays["raft_flow_absmax"] = np.float32(ref["raft_flow"].abs().max().item())
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrays)
        manifest[name] = dict(B=1, H=H, W=W, timesteps=ts, ds_factor=ds, input=("demo/input_frames/00020.png,00028.png (stored in the fixture)" if seed == "demo" else "synth_batch seed %d" % seed),
                              input_seed=(None if seed == "demo" else seed), weight_seed=0, stride=STRIDE, oracle_vs_reference_max_abs=pin,
                              reference_cpu_seconds=round(sec, 1), cpu_threads=torch.get_num_threads())
        with open(mpath, "w") as f:
            json.dump(manifest, f, indent=1)
        print(name, "reference forward %.1f s on %d threads; oracle pin %s" % (sec, torch.get_num_threads(), pin), flush=True)
        del ref


if __name__ == "__main__":
    main()
