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
