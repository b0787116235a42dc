"""Clip throughput of the streaming API (gimmvfi_b200.video.VideoInterpolator, RAFT-encoder frame cache) vs independent
per-pair calls (interpolate_pair_u8): synthetic 1080p clip, uint8 frames resident on the GPU, N=2.  usage: video_bench.py [frames]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from gimmvfi_b200 import GIMMVFI_R
from gimmvfi_b200.synth import synth_batch
from gimmvfi_b200.video import VideoInterpolator, interpolate_pair_u8

nf = int(sys.argv[1]) if len(sys.argv) > 1 else 9
dev = "cuda"
m = GIMMVFI_R(seed=0).to(dev).eval()
b = synth_batch((nf + 1) // 2, 1080, 1920, seed=3)
frames = []
for i in range(b.shape[0]):
    for j in range(2):
        frames.append((b[i, :, j].permute(1, 2, 0) * 255).round().to(torch.uint8).to(dev))
frames = frames[:nf]


def run_pairs():
    return [interpolate_pair_u8(m, frames[i], frames[i + 1], N=2)[0] for i in range(nf - 1)]


def run_stream():
    vi = VideoInterpolator(m, N=2)
    out = []
    for f in frames:
        out += vi.push(f)
    return out


for name, fn in (("per-pair calls", run_pairs), ("VideoInterpolator (frame cache)", run_stream), ("per-pair calls", run_pairs),
                 ("VideoInterpolator (frame cache)", run_stream)):
    fn()  # warm
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    o = fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("%-34s %d pairs @1920x1080: %.1f ms/pair, %.2f interpolated frames/s" % (name, nf - 1, 1e3 * dt / (nf - 1), (nf - 1) / dt), flush=True)
