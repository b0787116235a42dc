"""GIMM-VFI-F (gimmvfi_f.py) against fixtures produced by the UNMODIFIED reference GIMMVFI_F:
  * ff_* (oracle/make_golden_ff.py): the whole model natively — FlowFormer estimator (csrc/flowformer.cu) + synthesis half — on seeded
    weights the tests rebuild (gimmvfi_b200.weights.random_state_dict_f);
  * f_* (oracle/make_golden_f.py): the synthesis half alone, fed with the reference FlowFormer's outputs carried in the fixture
    (`forward_from_flow` consumes them exactly where the reference's forward does)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from gimmvfi_b200.model_f import GIMMVFI_F
from gimmvfi_b200.synth import synth_batch
from gimmvfi_b200.weights import random_state_dict_f

pytestmark = pytest.mark.gpu
DEV = "cuda"

with open(os.path.join(GOLDEN_DIR, "manifest_f.json")) as _f:
    MANIFEST = json.load(_f)


def flow_inputs_of(g, dev):
    T = lambda k: torch.from_numpy(g[k]).to(dev)
    return dict(flows=T("flows"), feat4=[T("feat4_0"), T("feat4_1")], feat8=[T("feat8_0"), T("feat8_1")], fnet=[T("fnet_0"), T("fnet_1")])


@pytest.fixture(scope="module")
def model():
    m = GIMMVFI_F(seed=0).to(DEV).eval()
    m.load_state_dict(random_state_dict_f(0), strict=True)
    return m


@pytest.mark.parametrize("name", sorted(MANIFEST))
@pytest.mark.parametrize("mode", [4, 3, 0], ids=["default", "mode3", "fp32"])
def test_f_synthesis_matches_reference(name, mode, model):
    meta = MANIFEST[name]
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    B, H, W, ts, ds = meta["B"], meta["H"], meta["W"], meta["timesteps"], meta["ds_factor"]
    xs = synth_batch(B, H, W, seed=meta["input_seed"]).to(DEV)
    ratio = 1.0 if ds is None else ds
    coord = [(model.sample_coord_input(B, (H, W), [t], device=DEV, upsample_ratio=ratio), None) for t in ts]
    tt = [t * torch.ones(B, device=DEV) for t in ts]
    model.tensor_cores = mode
    out = model(xs, coord, t=tt, ds_factor=ds, flow_inputs=flow_inputs_of(g, DEV))
    torch.cuda.synchronize()
    tol = 1e-3 if mode else 2e-5
    assert torch.equal(out["raft_flow"].cpu(), torch.from_numpy(g["raft_flow"]))   # the estimator's flows are passed through
    for i in range(len(ts)):
        d = (out["imgt_pred"][i].cpu() - torch.from_numpy(g["imgt_pred_%d" % i])).abs()
        print(name, "mode", mode, "imgt_pred[%d] max %.3e rmse %.3e" % (i, d.max().item(), d.pow(2).mean().sqrt().item()))
        assert d.max().item() <= tol
        w4 = (out["other_pred"][i][0].cpu() - torch.from_numpy(g["img_warp_4_%d" % i])).abs().max().item()
        assert w4 <= tol
        nin = (out["ninrflow"][i].cpu() - torch.from_numpy(g["ninrflow_%d" % i])).abs().max().item()
        assert nin <= (5e-4 if mode else 5e-6), nin   # HypoNet itself is fp32-class; its latent input comes from TF32 convolutions in mode 3


with open(os.path.join(GOLDEN_DIR, "manifest_ff.json")) as _f:
    MANIFEST_FF = json.load(_f)


@pytest.mark.parametrize("name", sorted(MANIFEST_FF))
@pytest.mark.parametrize("mode", [4, 0], ids=["default", "fp32"])
def test_f_native_flowformer_matches_reference(name, mode, model):
    """forward() with no flow backend: the native FlowFormer (both directions batched) + synthesis.  Bars: the estimator's flows within
    2e-3 px of the reference's (|flow| up to 25 px; fp32 mode 2e-4), its feature maps within 1e-4, the interpolated frame within the
    1e-3 of BASELINE.json (fp32 mode: 5e-5)."""
    meta = MANIFEST_FF[name]
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    B, H, W, ts, ds = meta["B"], meta["H"], meta["W"], meta["timesteps"], meta["ds_factor"]
    xs = synth_batch(B, H, W, seed=meta["input_seed"]).to(DEV)
    ratio = 1.0 if ds is None else ds
    coord = [(model.sample_coord_input(B, (H, W), [t], device=DEV, upsample_ratio=ratio), None) for t in ts]
    tt = [t * torch.ones(B, device=DEV) for t in ts]
    model.tensor_cores = mode
    model.flow_backend = None
    model.engine.set_debug(True)
    try:
        out = model(xs, coord, t=tt, ds_factor=ds)
        torch.cuda.synchronize()
        f8 = model.engine.tap("ff.feat8").cpu()
        fp = model.engine.tap("ff.fproj").cpu()
    finally:
        model.engine.set_debug(False)
    d_flow = (out["raft_flow"].cpu() - torch.from_numpy(g["flows"])).abs().max().item()
    ref8 = torch.cat([torch.from_numpy(g["feat8_0"]), torch.from_numpy(g["feat8_1"])], 0).permute(0, 2, 3, 1)
    refp = torch.cat([torch.from_numpy(g["fnet_0"]), torch.from_numpy(g["fnet_1"])], 0).permute(0, 2, 3, 1)
    d8, dp = (f8 - ref8).abs().max().item(), (fp - refp).abs().max().item()
    print(name, "mode", mode, "flows max %.3e px, feat8 %.3e, fnet %.3e, launches %d" % (d_flow, d8, dp, model.engine.last_launches))
    assert d8 <= (1e-4 if mode else 2e-5) and dp <= (1e-4 if mode else 2e-5)
    assert d_flow <= (2e-3 if mode else 2e-4)
    for i in range(len(ts)):
        d = (out["imgt_pred"][i].cpu() - torch.from_numpy(g["imgt_pred_%d" % i])).abs()
        print("   imgt_pred[%d] max %.3e rmse %.3e" % (i, d.max().item(), d.pow(2).mean().sqrt().item()))
        assert d.max().item() <= (1e-3 if mode else 5e-5)


def test_f_boundary(model):
    """639-key state_dict (strict) and the optional backend hook with the reference's FlowFormer.forward signature."""
    sd = model.state_dict()
    with open(os.path.join(GOLDEN_DIR, "state_dict_spec_f.json")) as f:
        spec = json.load(f)
    assert list(sd.keys()) == [k for k, _, _ in spec] and all(list(sd[k].shape) == s for k, s, _ in spec)
    name = "f_128x160_t0.5"
    meta, g = MANIFEST[name], np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    xs = synth_batch(1, meta["H"], meta["W"], seed=meta["input_seed"]).to(DEV)
    coord = [(model.sample_coord_input(1, (meta["H"], meta["W"]), [0.5], device=DEV), None)]
    tt = [0.5 * torch.ones(1, device=DEV)]
    fi = flow_inputs_of(g, DEV)
    calls = []

    def backend(im0, im1):   # FlowFormer.forward(im0, im1, return_feat=True) -> (flow_predictions, cfeat, ffeat)
        j = len(calls)
        calls.append((float(im0.max()), float(im1.max())))
        return [fi["flows"][:, :, j]], [fi["feat4"][j], fi["feat8"][j]], fi["fnet"][j]

    model.flow_backend = backend
    model.tensor_cores = 3
    out = model(xs, coord, t=tt)
    model.flow_backend = None
    assert len(calls) == 2 and calls[0][0] > 1.5   # the backend sees 0..255 frames (gimmvfi_f.py:320-328)
    assert (out["imgt_pred"][0].cpu() - torch.from_numpy(g["imgt_pred_0"])).abs().max().item() <= 1e-3
    with pytest.raises(RuntimeError):               # the Twins sub-sampling needs a multiple of 32 at the network resolution
        c2 = [(model.sample_coord_input(1, (136, 160), [0.5], device=DEV), None)]
        model(synth_batch(1, 136, 160, seed=1).to(DEV), c2, t=tt)


def test_f_cuda_graph_replay_matches_eager(model):
    """The GIMM-VFI-F forward (native FlowFormer: ~1450 launches, device memsets, per-sample GEMM loops) recorded as a CUDA graph and
    replayed: the estimator's flows are atomics-free -> bit-identical to the eager sequence; the frame within the splat's jitter."""
    name = "ff_128x160_t0.5"
    meta = MANIFEST_FF[name]
    B, H, W = meta["B"], meta["H"], meta["W"]
    xs = synth_batch(B, H, W, seed=meta["input_seed"]).to(DEV)
    coord = [(model.sample_coord_input(B, (H, W), [0.5], device=DEV), None)]
    t = [0.5 * torch.ones(B, device=DEV)]
    model.tensor_cores = 4
    model.flow_backend = None
    eager = model(xs, coord, t=t)
    e_flow, e_img = eager["raft_flow"].clone(), eager["imgt_pred"][0].clone()
    eng = model.engine
    eng.set_cuda_graph(True)
    eng.static_outputs = True
    try:
        coords = coord[0][0].unsqueeze(0).contiguous()
        tt = t[0].reshape(1, B).contiguous()
        r0 = eng.graph_replays
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(4):
                out = eng.forward(xs, coords, tt, None)
        side.synchronize()
        torch.cuda.synchronize()
        assert eng.graph_replays - r0 >= 3
        assert torch.equal(out["raft_flow"], e_flow)
        assert (out["imgt_pred"][0] - e_img).abs().max().item() <= 6e-4
    finally:
        eng.set_cuda_graph(False)
        eng.static_outputs = False
