"""One small GIMM-VFI-F forward (native FlowFormer, 2 decoder iterations) for compute-sanitizer runs:
    compute-sanitizer --tool memcheck|racecheck|synccheck python scripts/f_sanitize.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gimmvfi_b200 import GIMMVFI_F  # noqa: E402
from gimmvfi_b200.synth import synth_batch  # noqa: E402
from gimmvfi_b200.weights import random_state_dict_f  # noqa: E402

m = GIMMVFI_F(seed=0).to("cuda").eval()
m.load_state_dict(random_state_dict_f(0), strict=True)
m.engine.set_flowformer_iters(2)
xs = synth_batch(1, 128, 160, seed=3).cuda()
coord = [(m.sample_coord_input(1, (128, 160), [0.5], device="cuda"), None)]
out = m(xs, coord, t=[0.5 * torch.ones(1, device="cuda")])
torch.cuda.synchronize()
print("finite:", bool(torch.isfinite(out["imgt_pred"][0]).all()), "launches:", m.engine.last_launches)
