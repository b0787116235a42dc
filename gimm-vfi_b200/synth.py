"""Deterministic synthetic frame pairs: a band-limited smooth
texture and a copy of it backward-warped by a smooth flow (global translation
(+3.5, -2.25) px plus a low-frequency component) with 1 % noise.  White noise
makes RAFT chaotic and precision comparisons meaningless, hence smooth inputs.
Pure CPU torch; used by tests, bench.py and the golden generator."""
import math

import torch
import torch.nn.functional as F


def synth_pair(H: int, W: int, seed: int = 0, max_disp: float = 8.0) -> torch.Tensor:
    """Returns img_xs (1, 3, 2, H, W) float32 in [0, 1]."""
    g = torch.Generator().manual_seed(seed)
    ys = torch.linspace(0, 1, H).view(H, 1)
    xs = torch.linspace(0, 1, W).view(1, W)
    img = torch.zeros(3, H, W)
    for c in range(3):
        for _ in range(16):
            fx, fy = (torch.rand(2, generator=g) * 12).tolist()
            ph = (torch.rand(1, generator=g) * 2 * math.pi).item()
            a = torch.rand(1, generator=g).item()
            img[c] += a * torch.sin(2 * math.pi * (fx * xs + fy * ys) + ph)
    lo = torch.rand(1, 3, max(H // 16, 2), max(W // 16, 2), generator=g)
    lo = F.interpolate(lo, size=(H, W), mode="bilinear", align_corners=False)[0]
    img = (img / img.abs().max() * 0.35 + 0.5 + 0.3 * (lo - 0.5)).clamp(0, 1)
    fl = torch.zeros(1, 2, H, W)
    fl[:, 0] = 3.5
    fl[:, 1] = -2.25
    lf = (torch.rand(1, 2, 4, 4, generator=g) - 0.5) * max_disp
    fl = fl + F.interpolate(lf, size=(H, W), mode="bicubic", align_corners=False)
    gx = torch.linspace(-1, 1, W).view(1, 1, W).expand(1, H, W)
    gy = torch.linspace(-1, 1, H).view(1, H, 1).expand(1, H, W)
    grid = torch.stack([gx + fl[:, 0] / ((W - 1) / 2), gy + fl[:, 1] / ((H - 1) / 2)], -1)
    img1 = F.grid_sample(img[None], grid, mode="bilinear", padding_mode="border", align_corners=True)[0]
    img1 = (img1 + 0.01 * torch.randn(3, H, W, generator=g)).clamp(0, 1)
    return torch.stack([img, img1], 1)[None].contiguous()


def synth_batch(B: int, H: int, W: int, seed: int = 0) -> torch.Tensor:
    return torch.cat([synth_pair(H, W, seed + i) for i in range(B)], 0)


def synth_flow_pair(B: int, H: int, W: int, seed: int = 0, max_disp: float = 6.0) -> torch.Tensor:
    """Seeded bidirectional flow fields for the GIMM-standalone path: (B,2,2,H,W), [:, :, 0] = f01 (smooth low-frequency
    field + a global translation), [:, :, 1] = f10 ~ -f01 plus a small inconsistency (so the forward/backward-consistency term
    of the splatting metric, gimm.py:106-121, is exercised)."""
    g = torch.Generator().manual_seed(seed)
    hs, ws = max(2, H // 16), max(2, W // 16)
    lo = torch.randn(B, 2, hs, ws, generator=g) * (max_disp / 2)
    f01 = F.interpolate(lo, size=(H, W), mode="bicubic", align_corners=False) + torch.tensor([2.5, -1.25]).view(1, 2, 1, 1)
    dev = torch.randn(B, 2, hs, ws, generator=g) * 0.35
    f10 = -f01 + F.interpolate(dev, size=(H, W), mode="bicubic", align_corners=False)
    return torch.stack([f01, f10], 2).contiguous()
