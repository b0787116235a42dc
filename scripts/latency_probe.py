"""GPU box: forward latency at BASELINE config 1's size (256x448, launch-bound: ~550 kernels of a few microseconds each) and at the bench
size, eager vs CUDA-graph replay: host enqueue time (wall clock until the call returns) and device time (CUDA events)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from gimmvfi_b200 import GIMMVFI_R
from gimmvfi_b200.synth import synth_batch

dev = "cuda"
m = GIMMVFI_R(seed=0).to(dev).eval()
torch.cuda.set_stream(torch.cuda.Stream())   # CUDA graphs cannot record the legacy default stream
for H, W in ((256, 448), (736, 1280), (1088, 1920)):
    xs = synth_batch(1, H, W, seed=6).to(dev)
    coords = m.sample_coord_input(1, (H, W), [0.5], device=dev).unsqueeze(0).contiguous()
    tt = torch.full((1, 1), 0.5, device=dev)
    eng = m.engine
    eng.set_tensor_cores(3)
    for graph in (False, True):
        eng.set_cuda_graph(graph)
        eng.static_outputs = graph
        for _ in range(4):
            eng.forward(xs, coords, tt, None, aux_outputs=False)
        torch.cuda.synchronize()
        host, devms = [], []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e0.record()
            eng.forward(xs, coords, tt, None, aux_outputs=False)
            e1.record()
            host.append((time.perf_counter() - t0) * 1e3)
            torch.cuda.synchronize()
            devms.append(e0.elapsed_time(e1))
        host.sort(); devms.sort()
        print("%4dx%-4d %-12s host enqueue %7.3f ms   device %8.3f ms   launches %d   graph replays %d"
              % (H, W, "CUDA graph" if graph else "eager", host[len(host) // 2], devms[len(devms) // 2], eng.last_launches, eng.graph_replays), flush=True)
    eng.set_cuda_graph(False)
    eng.static_outputs = False
