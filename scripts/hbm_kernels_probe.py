"""GPU box: the HBM-bound kernels of the path at the 1080p bench shapes, one C-ABI call each (SURVEY 8(d) algorithmic bytes).

    python scripts/hbm_kernels_probe.py            # CUDA-event timing, L2 flushed between calls -> GB/s vs the measured HBM peak
    ncu --set full ... python scripts/hbm_kernels_probe.py --once    # one launch of each for the ncu capture

Shapes: softsplat 16 ch @1088x1920 (140 B/px); backwarp 64 ch @1088x1920 ((2C+2)*4 B/px); resize x4 128 ch 272x480 -> 1088x1920;
corr_lookup on a 4-level pyramid of N = 136x240 source pixels; instance norm 64 ch @2x544x960; convex upsample 136x240 -> 1088x1920."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import gpu_ops as K

once = "--once" in sys.argv
dev = "cuda"
torch.manual_seed(0)
H, W = 1088, 1920
peak = 6568.4
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p))["hbm_gbs"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def smooth_flow(n, h, w, amp=6.0):
    lo = torch.randn(n, 2, max(2, h // 64), max(2, w // 64), device=dev) * amp
    f = torch.nn.functional.interpolate(lo, size=(h, w), mode="bicubic", align_corners=False)
    return (f + torch.tensor([3.5, -2.25], device=dev).view(1, 2, 1, 1)).permute(0, 2, 3, 1).contiguous()


def timeit(name, fn, alg_bytes, reps=5):
    fn()
    torch.cuda.synchronize()
    if once:
        return
    ms = []
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    m = sorted(ms)[len(ms) // 2]
    gbs = alg_bytes / (m * 1e-3) / 1e9
    print("%-34s %8.3f ms   %8.1f MB algorithmic   %7.1f GB/s = %.3f of measured HBM peak (%.0f)" % (name, m, alg_bytes / 1e6, gbs, gbs / peak, peak), flush=True)


# ---- softsplat (memset + accumulate + normalise = the whole op the 140 B/px figure describes)
lat = torch.randn(1, H, W, 16, device=dev)
flow = smooth_flow(1, H, W)
metric = 0.5 + torch.rand(1, H, W, 1, device=dev)
t = torch.full((1,), 0.5, device=dev)
timeit("softsplat 16ch 1088x1920 (3 passes)", lambda: K.softsplat(lat, flow, metric, t, 0), H * W * 140.0)
amax = flow.abs().amax().reshape(1).contiguous()
timeit("softsplat 16ch 1088x1920 (one pass)", lambda: K.softsplat_fused(lat, flow, metric, t, 0, amax), H * W * 140.0)
# ---- backwarp 64 ch at full resolution
src = torch.randn(1, H, W, 64, device=dev)
timeit("backwarp 64ch 1088x1920", lambda: K.backwarp(src, flow), H * W * (2 * 64 + 2) * 4.0)
src3 = torch.randn(1, H, W, 4, device=dev)
# ---- resize x4 of 128 channels (ft_4 -> full resolution)
s4 = torch.randn(1, H // 4, W // 4, 128, device=dev)
timeit("resize x4 128ch 272x480->1088x1920", lambda: K.resize(s4, 4.0), (H // 4) * (W // 4) * 128 * 4.0 * (1 + 16))
# ---- instance norm (fnet stem size)
xin = torch.randn(2, H // 2, W // 2, 64, device=dev)
timeit("instnorm+relu 64ch 2x544x960", lambda: K.instnorm(xin, True), xin.numel() * 4.0 * 3)   # stats read + apply read + write
# ---- convex upsample
fl8 = torch.randn(2, H // 8, W // 8, 2, device=dev)
mk = torch.randn(2, H // 8, W // 8, 576, device=dev)
timeit("convex_upsample 2x136x240", lambda: K.convex_upsample(fl8, mk), (mk.numel() + fl8.numel() + 2 * 2 * H * W) * 4.0)
# ---- correlation lookup on a real-size pyramid (one direction: N = 32640 rows)
h, w = H // 8, W // 8
fa = torch.randn(1, h, w, 256, device=dev) * 0.5
fb = torch.randn(1, h, w, 256, device=dev) * 0.5
levels = K.corr_pyramid(fa, fb)
coords = (torch.stack(torch.meshgrid(torch.arange(w, device=dev), torch.arange(h, device=dev), indexing="xy"), -1).float()[None]
          + smooth_flow(1, h, w, 2.0)).contiguous()
timeit("corr_lookup 324ch 136x240", lambda: K.corr_lookup(levels, coords), h * w * (324 * 4.0 + 4 * 100 * 4.0))
# ---- the volume-free form of the same lookup (BidirCorrBlock): 100 dot products of 256 channels per (pixel, level); bytes = the half
# target pyramid once + the source features + the output (what a perfect cache would move)
hl = K.half_feature_pyramid(fb)
timeit("corr_lookup_direct 324ch 136x240", lambda: K.corr_lookup_direct(fa, fb, coords, levels=hl),
       h * w * (256 * 4.0 + 324 * 4.0) + sum(l.numel() * 2.0 for l in hl))
print("done", flush=True)
