"""GPU box: which taps of the halo-reuse conv kernel are right?  Single-tap weights isolate the shifted-view descriptors."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F

import gpu_ops as K

torch.manual_seed(0)
dev = "cuda"
n, cin, cout, H, W = 1, 32, 32, 32, 24
x = torch.randn(n, cin, H, W, device=dev)
b = torch.zeros(cout, device=dev)
xn = K.nhwc(x)
for ky in range(3):
    for kx in range(3):
        w = torch.zeros(cout, cin, 3, 3, device=dev)
        w[:, :, ky, kx] = torch.randn(cout, cin, device=dev) / cin ** 0.5
        got = K.nchw(K.conv2d_halo(xn, w, b))
        ref = F.conv2d(K.tf32_trunc(x).double(), K.tf32_rn(w).double(), padding=1).float()
        d = (got - ref).abs()
        # also: does the result equal the reference of ANOTHER tap (i.e. a wrong shift)?
        best = None
        for qy in range(3):
            for qx in range(3):
                w2 = torch.zeros_like(w); w2[:, :, qy, qx] = w[:, :, ky, kx]
                r2 = F.conv2d(K.tf32_trunc(x).double(), K.tf32_rn(w2).double(), padding=1).float()
                e = (got - r2).abs().max().item()
                if best is None or e < best[0]:
                    best = (e, qy, qx)
        print("tap (ky=%d,kx=%d): max err %.3e (ref absmax %.2f); rows ok %d/%d; closest single-tap reference: (%d,%d) err %.3e"
              % (ky, kx, d.max().item(), ref.abs().max().item(), int((d.amax(dim=(0, 1, 3)) < 1e-2).sum()), H, best[1], best[2], best[0]), flush=True)
        if ky == 1 and kx == 1:
            bad = (d.amax(dim=(0, 1)) > 1e-2)
            print("   centre tap: bad pixel map (first 16 rows, '#' = wrong):")
            for yy in range(16):
                print("   ", "".join("#" if bad[yy, xx] else "." for xx in range(W)))
