"""Host-logic tests: the SAME engine sources compiled for the host (tests/hostsim) must
reproduce the oracle — validates orchestration, weight packing (BN folding, head
permutation, HypoNet normalisation), every thread-per-element kernel body and the
workspace planner without a GPU."""
import pytest
import torch

import gimmvfi_r_oracle as O
from gimmvfi_b200.synth import synth_batch


@pytest.fixture(scope="module")
def eng(weights0):
    import sys, os

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
    from harness import hostsim_engine

    e = hostsim_engine()
    e.load_state_dict(weights0)
    return e


CASES = [(1, 128, 160, [0.5], None, 3), (2, 128, 128, [0.25, 0.75], None, 4), (1, 256, 320, [0.5], 0.5, 5)]


@pytest.mark.parametrize("B,H,W,ts,ds,seed", CASES)
def test_hostsim_forward_matches_oracle(eng, weights0, B, H, W, ts, ds, seed):
    torch.set_grad_enabled(False)
    xs = synth_batch(B, H, W, seed=seed)
    r = ds or 1.0
    coords = torch.stack([O.sample_coord_input(B, (H, W), [t], r) for t in ts], 0).contiguous()
    tt = torch.stack([t * torch.ones(B) for t in ts], 0).contiguous()
    out = eng.forward(xs, coords, tt, ds)
    ref = O.gimmvfi_r_forward(weights0, xs, [(coords[i], None) for i in range(len(ts))], [tt[i] for i in range(len(ts))], ds_factor=ds)
    assert (out["raft_flow"] - ref["raft_flow"]).abs().max() <= 2e-4
    for i in range(len(ts)):
        assert (out["imgt_pred"][i] - ref["imgt_pred"][i]).abs().max() <= 2e-5
        assert (out["img_warp_4"][i] - ref["other_pred"][i][0]).abs().max() <= 2e-5
        assert (out["flowt"][i] - ref["flowt"][i].reshape(out["flowt"][i].shape)).abs().max() <= 2e-3
        assert (out["flowt0_1"][i] - ref["flowt0_pred"][i][0]).abs().max() <= 5e-4
        assert (out["flowt1_4"][i] - ref["flowt1_pred"][i][1]).abs().max() <= 5e-4
        assert (out["ninrflow"][i] - ref["ninrflow"][i]).abs().max() <= 1e-4


def test_planner_is_exact_and_rejects_bad_shapes(eng):
    n = eng.workspace_bytes(1, 128, 160, 1, None, 128, 160)
    assert n > 0 and eng.workspace_bytes(2, 128, 160, 1, None, 128, 160) > n
    from gimmvfi_b200._lib import GimmvfiError

    with pytest.raises(GimmvfiError):
        eng.workspace_bytes(1, 100, 160, 1, None, 100, 160)    # not a multiple of 8
    with pytest.raises(GimmvfiError):
        eng.workspace_bytes(1, 128, 160, 1, None, 64, 80)      # coord grid != network resolution


def test_frame_cache_reuses_second_frame_bit_exactly(eng):
    """pairs (A,B), (B,C) of a clip — the second call takes frame B's RAFT encoder products from the cache
    written by the first; every output must equal the uncached forward (same arithmetic, only skipped)."""
    torch.set_grad_enabled(False)
    H, W = 128, 160
    fr = synth_batch(2, H, W, seed=11)                       # (2,3,2,H,W): two pairs -> frames A,B and (reused) B,C
    A, Bf, Cf = fr[0, :, 0], fr[0, :, 1], fr[1, :, 1]
    pair = lambda a, b: torch.stack([a, b], 1).unsqueeze(0).contiguous()
    coords = O.sample_coord_input(1, (H, W), [0.5], 1.0).unsqueeze(0).contiguous()
    tt = 0.5 * torch.ones(1, 1)
    ref_ab = eng.forward(pair(A, Bf), coords, tt, None)
    ref_bc = eng.forward(pair(Bf, Cf), coords, tt, None)
    cache = torch.zeros(eng.frame_cache_bytes(1, H, W, 1, None, H, W), dtype=torch.uint8)
    got_ab = eng.forward(pair(A, Bf), coords, tt, None, frame_cache=(cache, False, True))
    got_bc = eng.forward(pair(Bf, Cf), coords, tt, None, frame_cache=(cache, True, True))
    for ref, got in ((ref_ab, got_ab), (ref_bc, got_bc)):
        assert torch.equal(ref["raft_flow"], got["raft_flow"])
        assert (ref["imgt_pred"] - got["imgt_pred"]).abs().max() <= 1e-6   # (splat atomics: summation order only)
    # a cache with other content must change the result (the mechanism is really in the path)
    cache.zero_()
    bad = eng.forward(pair(Bf, Cf), coords, tt, None, frame_cache=(cache, True, False))
    assert not torch.equal(ref_bc["raft_flow"], bad["raft_flow"])
    # a buffer no forward has stored into is refused
    from gimmvfi_b200._lib import GimmvfiError
    other = torch.zeros_like(cache)
    with pytest.raises(GimmvfiError):
        eng.forward(pair(Bf, Cf), coords, tt, None, frame_cache=(other, True, False))


@pytest.mark.parametrize("gimm_only", [False, True])
def test_hostsim_gimm_forward_matches_oracle(weights0, gimm_only):
    """GIMM standalone through the engine (run_gimm): full GIMM-VFI-R weights, and a GIMM-only checkpoint"""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
    from harness import hostsim_engine
    from gimmvfi_b200.gimm import GIMM_KEY_PREFIXES
    from gimmvfi_b200.synth import synth_flow_pair
    from gimmvfi_b200._lib import GimmvfiError

    torch.set_grad_enabled(False)
    e = hostsim_engine()
    sd = {k: v for k, v in weights0.items() if k.startswith(GIMM_KEY_PREFIXES)} if gimm_only else weights0
    e.load_state_dict(sd, gimm_only=gimm_only)
    B, H, W, ts = 2, 40, 56, [0.25, 0.75]
    ori = synth_flow_pair(B, H, W, seed=4)
    xs, _ = O.normalize_flow(ori)
    coords = torch.stack([O.sample_coord_input(B, (H, W), [t], 1.0) for t in ts], 0).contiguous()
    tt = torch.stack([t * torch.ones(B) for t in ts], 0).contiguous()
    got = e.gimm_forward(xs.contiguous(), ori, coords, tt)
    ref = O.gimm_forward(weights0, xs, [coords[i] for i in range(len(ts))], ori, [tt[i] for i in range(len(ts))])
    for i in range(len(ts)):
        assert (got[i] - ref[i]).abs().max() <= 2e-5
    if gimm_only:   # the interpolation path needs the whole GIMM-VFI-R state_dict
        c1 = O.sample_coord_input(1, (128, 160), [0.5], 1.0).unsqueeze(0).contiguous()
        with pytest.raises(GimmvfiError):
            e.forward(synth_batch(1, 128, 160, seed=1), c1, 0.5 * torch.ones(1, 1), None)


import os as _os


@pytest.mark.parametrize("mode", [4, 3, 2, 1] if _os.environ.get("GIMMVFI_HOSTSIM_ALL_MODES") else [4])   # (72 s per mode on 8 cores; 4 = the shipped default)
def test_hostsim_tensor_core_modes_match_oracle(eng, weights0, mode):
    """The engine's tensor-core ORCHESTRATION on the CPU: precision modes 1-4 with the tensor-core convolution's operand rounding
    emulated on the host (tests/hostsim/tc_hostsim.cu) — half-precision trunk tensors, merged GRU gates, stride-2 and pre-padded
    layers, tensor-core correlation — against the oracle within the product's tolerance (max|d imgt_pred| <= 1e-3)."""
    torch.set_grad_enabled(False)
    B, H, W, ts = 1, 128, 160, [0.5]
    xs = synth_batch(B, H, W, seed=3)
    coords = torch.stack([O.sample_coord_input(B, (H, W), [t], 1.0) for t in ts], 0).contiguous()
    tt = torch.stack([t * torch.ones(B) for t in ts], 0).contiguous()
    eng.set_tensor_cores(mode)
    try:
        out = eng.forward(xs, coords, tt, None)
    finally:
        eng.set_tensor_cores(0)
    ref = O.gimmvfi_r_forward(weights0, xs, [(coords[0], None)], [tt[0]])
    d_img = (out["imgt_pred"][0] - ref["imgt_pred"][0]).abs().max().item()
    d_raft = (out["raft_flow"] - ref["raft_flow"]).abs().max().item()
    print("hostsim mode", mode, "imgt_pred max %.3e raft_flow max %.3e" % (d_img, d_raft))
    assert d_img <= 1e-3
    assert d_raft <= (2e-3 if mode >= 2 else 5e-4)


def test_hostsim_f_synthesis_from_flow_matches_reference():
    """GIMM-VFI-F: the engine's forward_from_flow (everything downstream of the flow estimator) on the CPU build of the kernels,
    fed with the reference FlowFormer's outputs from the fixture, against the UNMODIFIED reference GIMMVFI_F's frame."""
    import json, os, sys
    import numpy as np
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
    import harness
    from conftest import GOLDEN_DIR
    from gimmvfi_b200._lib import GimmvfiError
    from gimmvfi_b200.weights import random_state_dict_f

    torch.set_grad_enabled(False)
    name = "f_128x160_t0.5"
    meta = json.load(open(os.path.join(GOLDEN_DIR, "manifest_f.json")))[name]
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    e = harness.hostsim_engine()
    e.load_state_dict(random_state_dict_f(0), synthesis_only=True)
    B, H, W = meta["B"], meta["H"], meta["W"]
    xs = synth_batch(B, H, W, seed=meta["input_seed"])
    coords = torch.stack([O.sample_coord_input(B, (H, W), [t], 1.0) for t in meta["timesteps"]], 0).contiguous()
    tt = torch.stack([t * torch.ones(B) for t in meta["timesteps"]], 0).contiguous()
    T = lambda k: torch.from_numpy(g[k])
    fi = dict(flows=T("flows"), feat4=[T("feat4_0"), T("feat4_1")], feat8=[T("feat8_0"), T("feat8_1")], fnet=[T("fnet_0"), T("fnet_1")])
    out = e.forward(xs, coords, tt, None, flow_inputs=fi)
    d = (out["imgt_pred"][0] - T("imgt_pred_0")).abs().max().item()
    print("hostsim F synthesis: imgt_pred max %.3e" % d)
    assert d <= 2e-5
    with pytest.raises(GimmvfiError):   # no flow estimator weights in this engine
        e.forward(xs, coords, tt, None)


def test_hostsim_volume_free_raft_lookup(weights0, monkeypatch):
    """RAFT's own lookups without the all-pairs volume (GIMMVFI_RAFT_CORR_DIRECT=1; chosen by size when the pyramid
    would not fit) against the oracle: the other frame's features enter as halves, so the flow moves by ~2e-4 px, the frame by < 1e-5."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
    import harness

    torch.set_grad_enabled(False)
    monkeypatch.setenv("GIMMVFI_RAFT_CORR_DIRECT", "1")
    e = harness.hostsim_engine()
    e.load_state_dict(weights0)
    B, H, W = 1, 128, 160
    xs = synth_batch(B, H, W, seed=3)
    coords = torch.stack([O.sample_coord_input(B, (H, W), [0.5], 1.0)], 0).contiguous()
    tt = 0.5 * torch.ones(1, B)
    out = e.forward(xs, coords, tt, None)
    ref = O.gimmvfi_r_forward(weights0, xs, [(coords[0], None)], [tt[0]])
    assert (out["raft_flow"] - ref["raft_flow"]).abs().max() <= 2e-3
    assert (out["imgt_pred"][0] - ref["imgt_pred"][0]).abs().max() <= 1e-4


@pytest.mark.parametrize("name,mode", [("ff_b2_128x128_t0.5", 0)])   # (the tensor-core modes: tests/test_f_gpu.py on the B200)
def test_hostsim_f_native_flowformer_matches_reference(name, mode):
    """GIMM-VFI-F end to end on the CPU build of the kernels: the NATIVE FlowFormer estimator (flowformer.cu: Twins x2, cost-perceiver
    memory encoder, 32-iteration GMA decoder; both directions batched) + the synthesis half, against the UNMODIFIED reference GIMMVFI_F
    on the same seeded weights (oracle/make_golden_ff.py).  B = 2 covers the reference's context.repeat batch order."""
    import json, os, sys
    import numpy as np
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
    import harness
    from conftest import GOLDEN_DIR
    from gimmvfi_b200.weights import random_state_dict_f

    torch.set_grad_enabled(False)
    meta = json.load(open(os.path.join(GOLDEN_DIR, "manifest_ff.json")))[name]
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    e = harness.hostsim_engine()
    e.load_state_dict(random_state_dict_f(0), full_f=True)
    e.set_tensor_cores(mode)
    B, H, W = meta["B"], meta["H"], meta["W"]
    xs = synth_batch(B, H, W, seed=meta["input_seed"])
    coords = torch.stack([O.sample_coord_input(B, (H, W), [t], 1.0) for t in meta["timesteps"]], 0).contiguous()
    tt = torch.stack([t * torch.ones(B) for t in meta["timesteps"]], 0).contiguous()
    out = e.forward(xs, coords, tt, None)
    T = lambda k: torch.from_numpy(g[k])
    d_flow = (out["raft_flow"] - T("flows")).abs().max().item()
    d_img = (out["imgt_pred"][0] - T("imgt_pred_0")).abs().max().item()
    print("hostsim native F %s mode %d: flows max %.3e px (|flow| <= %.1f), imgt_pred max %.3e" % (name, mode, d_flow, meta["flow_absmax"], d_img))
    assert d_flow <= 5e-4
    assert d_img <= (2e-4 if mode else 5e-5)
