"""GPU box: parity of every precision mode against the CPU oracle at the bench configuration (1088x1920)."""
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

H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1088, 1920)
dev = "cuda"
model = GIMMVFI_R(seed=0).to(dev).eval()
sd = {k: v.cpu() for k, v in model.state_dict().items()}
xs = synth_batch(1, H, W, seed=100)
torch.set_num_threads(min(os.cpu_count(), 64))
t0 = time.time()
with torch.no_grad():
    ref = O.gimmvfi_r_forward(sd, xs, [(O.sample_coord_input(1, (H, W), [0.5]), None)], [0.5 * torch.ones(1)])
print("oracle %dx%d on %d threads: %.1f s" % (H, W, torch.get_num_threads(), time.time() - t0), flush=True)
coord = [(model.sample_coord_input(1, (H, W), [0.5], device=dev), None)]
for mode in [int(m) for m in os.environ.get("PARITY_MODES", "0,1,2,3").split(",")]:
    model.tensor_cores = mode
    out = model(xs.to(dev), coord, t=[0.5 * torch.ones(1, device=dev)])
    torch.cuda.synchronize()
    d = (out["imgt_pred"][0].cpu().double() - ref["imgt_pred"][0].double()).abs().flatten()
    rf = (out["raft_flow"].cpu() - ref["raft_flow"]).abs()
    ft = (out["flowt"][0].cpu() - ref["flowt"][0]).abs()
    mse = (d ** 2).mean().item()
    q = torch.quantile(d[::5], torch.tensor([0.999, 0.9999], dtype=torch.float64))
    print("mode %d: imgt_pred max %.3e p99.9 %.3e p99.99 %.3e rmse %.3e PSNR %.1f dB  n(>1e-3) %d of %d | raft_flow max %.3e mean %.3e | flowt max %.3e mean %.3e"
          % (mode, d.max(), q[0], q[1], mse ** 0.5, -10 * torch.log10(torch.tensor(mse)).item(), int((d > 1e-3).sum()), d.numel(), rf.max(), rf.mean(), ft.max(), ft.mean()), flush=True)
