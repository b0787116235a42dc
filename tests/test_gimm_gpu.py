"""GIMM standalone on the GPU: gimmvfi_b200.GIMM vs the reference's golden outputs (gimm.py:129-214)."""
import os

import numpy as np
import pytest
import torch

import gimmvfi_r_oracle as O
from conftest import GOLDEN_DIR
from gimmvfi_b200.gimm import GIMM, GIMM_KEY_PREFIXES
from gimmvfi_b200.synth import synth_flow_pair

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("mode", [1, 0])
@pytest.mark.parametrize("name", ["gimm_64x96_t0.25_0.75", "gimm_b2_72x80_t0.5"])
def test_gimm_matches_reference_golden(name, mode, golden_manifest, weights0):
    meta = golden_manifest[name]
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    B, H, W, ts = meta["B"], meta["H"], meta["W"], meta["timesteps"]
    m = GIMM(seed=0).to(DEV).eval()
    m.load_state_dict({k: v for k, v in weights0.items() if k.startswith(GIMM_KEY_PREFIXES)}, strict=True)
    m.tensor_cores = mode
    ori = synth_flow_pair(B, H, W, seed=meta["flow_seed"])
    xs, _ = O.normalize_flow(ori)
    coord = [m.sample_coord_input(B, (H, W), [t], device=DEV) for t in ts]
    out = m(xs.to(DEV), coord, True, ori.to(DEV), [t * torch.ones(B, device=DEV) for t in ts])
    tol = 2e-5 if mode == 0 else 5e-3   # normalised flow in [0,1]; TF32 operands in the SIREN MLP
    for i in range(len(ts)):
        assert out[i].shape == (B, 2, 1, H, W)
        err = np.abs(out[i].cpu().numpy() - g["out_%d" % i]).max()
        print(name, "mode", mode, "t", ts[i], "max err %.3e" % err)
        assert err <= tol
    one = m(xs.to(DEV), coord[0], False, ori.to(DEV), ts[0] * torch.ones(B, device=DEV))   # tensor form, keep_xs_shape=False
    assert one.shape == (B, 1, H, W, 2)


def test_gimm_larger_resolution_matches_oracle(weights0):
    """448x256 flows (the reference's Vimeo-scale motion benchmark size), TF32 tensor-core path vs the CPU oracle"""
    B, H, W, ts = 1, 256, 448, [0.5]
    m = GIMM(seed=0).to(DEV).eval()
    m.load_state_dict({k: v for k, v in weights0.items() if k.startswith(GIMM_KEY_PREFIXES)}, strict=True)
    ori = synth_flow_pair(B, H, W, seed=9)
    xs, scale = O.normalize_flow(ori)
    coord = [m.sample_coord_input(B, (H, W), [0.5], device=DEV)]
    out = m(xs.to(DEV), coord, True, ori.to(DEV), [0.5 * torch.ones(B, device=DEV)])[0].cpu()
    with torch.no_grad():
        ref = O.gimm_forward(weights0, xs, [c.cpu() for c in coord], ori, [0.5 * torch.ones(B)])[0]
    d = (out - ref).abs()
    assert d.max().item() <= 5e-3 and d.mean().item() <= 5e-4
