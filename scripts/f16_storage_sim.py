"""CPU experiment behind precision mode 3: the oracle with the final decoder trunk rounded to IEEE half (activations and weights)
vs the fp32 oracle.  usage: f16_storage_sim.py H W   (128 160: max |d imgt_pred| 8.3e-5, trunk activations |x| <= 1.3)"""
import sys, json, os, torch
sys.path.insert(0,'/root/repo/oracle'); sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import gimmvfi_r_oracle as O
from gimmvfi_b200.synth import synth_batch
from gimmvfi_b200.weights import random_state_dict
torch.set_grad_enabled(False); torch.set_num_threads(8)
sd = random_state_dict(seed=0)
H, W = int(sys.argv[1]), int(sys.argv[2])
xs = synth_batch(1, H, W, seed=1)
coord = [(O.sample_coord_input(1, (H, W), [0.5], 1.0), None)]
tt = [0.5 * torch.ones(1)]
ref = O.gimmvfi_r_forward(sd, xs, coord, tt)["imgt_pred"][0]
orig_conv, orig_prelu = O._conv, O._prelu
stats = {}
def h(x): return x.half().float()
mode = {"on": False}
def conv2(sd_, name, x, stride=1, padding=0, reflect=False):
    if mode["on"] and name.startswith("amt_final_decoder.convblock.") and not name.startswith("amt_final_decoder.convblock.0"):
        w = h(sd_[name + ".weight"]); b = sd_.get(name + ".bias")
        stats[name] = max(stats.get(name, 0), float(x.abs().max()))
        return torch.nn.functional.conv2d(h(x), w, b, stride=stride, padding=padding)
    return orig_conv(sd_, name, x, stride, padding, reflect)
def prelu2(sd_, name, x):
    y = orig_prelu(sd_, name, x)
    if mode["on"] and name.startswith("amt_final_decoder.convblock."):
        stats["out:" + name] = max(stats.get("out:" + name, 0), float(y.abs().max()))
        return h(y)
    return y
O._conv, O._prelu = conv2, prelu2
mode["on"] = True
got = O.gimmvfi_r_forward(sd, xs, coord, tt)["imgt_pred"][0]
d = (got - ref).abs()
print("fp16-storage final decoder vs fp32: max %.3e mean %.3e" % (d.max(), d.mean()))
print("absmax of activations:", {k: round(v, 3) for k, v in list(stats.items())[:40]})
