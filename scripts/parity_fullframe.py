"""GPU box: FULL-FRAME parity at the bench configuration (every one of the 6.27 M output values, not the stride-8 fixture grid) —
counts the splat-discontinuity outliers a small flow error can flip.
    python scripts/parity_fullframe.py --make-ref     # CPU oracle (== reference) once, ~100 s on the box's host cores -> /tmp/gv_ref_1080p.pt
    [env knobs] python scripts/parity_fullframe.py    # the CUDA path under the current environment vs that file"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch

from gimmvfi_b200.synth import synth_batch
from gimmvfi_b200.weights import random_state_dict

H, W = 1088, 1920
REF = "/tmp/gv_ref_1080p.pt"
xs = synth_batch(1, H, W, seed=100)
if "--make-ref" in sys.argv:
    import gimmvfi_r_oracle as O

    torch.set_num_threads(min(os.cpu_count(), 64))
    t0 = time.time()
    with torch.no_grad():
        ref = O.gimmvfi_r_forward(random_state_dict(0), xs, [(O.sample_coord_input(1, (H, W), [0.5]), None)], [0.5 * torch.ones(1)])
    torch.save({"imgt_pred": ref["imgt_pred"][0], "raft_flow": ref["raft_flow"], "flowt": ref["flowt"][0]}, REF)
    print("oracle %dx%d on %d threads: %.1f s" % (H, W, torch.get_num_threads(), time.time() - t0), flush=True)
    sys.exit(0)
from gimmvfi_b200 import GIMMVFI_R

ref = torch.load(REF)
dev = "cuda"
model = GIMMVFI_R(seed=0).to(dev).eval()
coord = [(model.sample_coord_input(1, (H, W), [0.5], device=dev), None)]
out = model(xs.to(dev), coord, t=[0.5 * torch.ones(1, device=dev)])
torch.cuda.synchronize()
d = (out["imgt_pred"][0].cpu().double() - ref["imgt_pred"].double()).abs().flatten()
rf = (out["raft_flow"].cpu() - ref["raft_flow"]).abs()
ft = (out["flowt"][0].cpu() - ref["flowt"]).abs()
mse = (d ** 2).mean().item()
q = torch.quantile(d[::5], torch.tensor([0.999, 0.9999], dtype=torch.float64))
knobs = {k: v for k, v in os.environ.items() if k.startswith("GIMMVFI_")}
print("%s: imgt_pred max %.3e p99.9 %.3e p99.99 %.3e PSNR %.1f dB  n(>1e-3) %d of %d | raft_flow max %.3e mean %.3e | flowt max %.3e mean %.3e"
      % (knobs or "default", d.max(), q[0], q[1], -10 * torch.log10(torch.tensor(mse)).item(), int((d > 1e-3).sum()), d.numel(), rf.max(), rf.mean(), ft.max(), ft.mean()), flush=True)
