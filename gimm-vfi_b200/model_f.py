"""Drop-in ``GIMMVFI_F`` boundary (reference: src/models/generalizable_INR/gimmvfi_f.py:27-484): same constructor argument, the 639-key
``state_dict`` of the reference (FlowFormer included, so published GIMM-VFI-F / F-P checkpoints load with ``strict=True``), same
``forward(img_xs, coord, t, ds_factor)`` and returned dict.

Everything runs natively on the sm_100a engine (``gimmvfi_finalize_weights_f`` + ``gimmvfi_forward``):
  * the FlowFormer flow estimator (gimmvfi_f.py:114-138 -> LatentCostFormer/transformer.py:45-74): Twins-SVT-L encoders, cost-perceiver
    memory encoder, 32-iteration GMA memory decoder — csrc/flowformer.cu + csrc/ops_tokens.cu, both directions batched;
  * everything downstream of ``cal_bidirection_flow`` — the same kernels as GIMM-VFI-R (F has no feature projections,
    gimmvfi_f.py:37-60).
``model.flow_backend`` (optional) swaps the estimator for a caller-supplied callable with the signature of the reference's
``FlowFormer.forward(im0, im1, return_feat=True)``; ``forward(..., flow_inputs=...)`` takes precomputed estimator outputs
(``gimmvfi_forward_from_flow``).  There is no PyTorch re-implementation of FlowFormer on the product path."""
from typing import Callable, Optional

import torch
import torch.nn as nn

from .arch import param_spec_f
from .config import default_arch_config
from .engine import EngineHandle
from .model import _Node, sample_coords
from .weights import random_state_dict_f


class GIMMVFI_F(nn.Module):
    def __init__(self, config=None, seed: int = 0):
        super().__init__()
        self.config = config = (config.copy() if config is not None else default_arch_config())
        self.hyponet_config = config.hyponet
        self.raft_iter = getattr(config, "raft_iter", 20)
        self.num_flows = 3
        self.fwarp_type = getattr(config, "fwarp_type", "linear")
        if self.fwarp_type != "linear":
            raise NotImplementedError("only fwarp_type='linear' (the shipped configs' default, configs.py:44) is built")
        self.coord_range = list(config.coord_range)
        init = random_state_dict_f(seed)
        for key, shape, dt in param_spec_f():
            parts = key.split(".")
            mod = self
            for p in parts[:-1]:
                if p not in mod._modules:
                    mod.add_module(p, _Node())
                mod = mod._modules[p]
            leaf = parts[-1]
            if dt == "int64" or leaf in ("running_mean", "running_var"):
                mod.register_buffer(leaf, init[key].clone())
            else:
                mod.register_parameter(leaf, nn.Parameter(init[key].clone(), requires_grad=False))
        self._engine: Optional[EngineHandle] = None
        self._weights_dirty = True
        self.register_load_state_dict_post_hook(lambda m, k: setattr(m, "_weights_dirty", True))
        self.aux_outputs = True
        self.tensor_cores = 4
        self.flow_backend: Optional[Callable] = None   # see the module docstring

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._weights_dirty = True
        return r

    @property
    def engine(self) -> EngineHandle:
        dev = self.g_filter.device
        if dev.type != "cuda":
            raise RuntimeError("GIMMVFI_F (gimmvfi_b200) runs on CUDA devices only; there is no CPU path")
        if self._engine is None or self._engine.device != dev or self._weights_dirty:
            if self._engine is None or self._engine.device != dev:
                self._engine = EngineHandle(dev)
            self._engine.load_state_dict(self.state_dict(), full_f=True)
            self._engine.tensor_cores = None
            self._weights_dirty = False
        return self._engine

    def flow_estimator_state_dict(self) -> dict:
        """`flow_estimator.*` with the prefix stripped: what the reference FlowFormer module's load_state_dict takes."""
        return {k[len("flow_estimator."):]: v for k, v in self.state_dict().items() if k.startswith("flow_estimator.")}

    def sample_coord_input(self, batch_size, s_shape, t_ids, coord_range=None, upsample_ratio=1.0, device=None):
        """gimmvfi_f.py:391-407"""
        assert device is not None
        assert coord_range is None
        return sample_coords(batch_size, s_shape, t_ids, self.coord_range, upsample_ratio, device)

    def cal_bidirection_flow(self, im0, im1):
        """gimmvfi_f.py:114-138 through an EXTERNAL flow backend (model.flow_backend) -> the dict gimmvfi_forward_from_flow takes.
        Without a backend forward() runs the engine's native FlowFormer and never calls this."""
        assert self.flow_backend is not None
        f01, feats0, fnet0 = self.flow_backend(im0, im1)
        f10, feats1, fnet1 = self.flow_backend(im1, im0)
        f01 = f01[0] if isinstance(f01, (list, tuple)) else f01
        f10 = f10[0] if isinstance(f10, (list, tuple)) else f10
        return dict(flows=torch.stack([f01, f10], 2), feat4=[feats0[0], feats1[0]], feat8=[feats0[1], feats1[1]], fnet=[fnet0, fnet1])

    @torch.no_grad()
    def forward(self, img_xs, coord=None, t=None, ds_factor=None, flow_inputs=None):
        """gimmvfi_f.py:304-384 (inference form).  `flow_inputs` (optional) = precomputed outputs of cal_bidirection_flow."""
        assert isinstance(t, list)
        assert isinstance(coord, list)
        assert len(t) == len(coord)
        for c in coord:
            assert isinstance(c, tuple)
            if c[1] is not None:
                raise NotImplementedError("sub-sampled coordinates are the training path (gimmvfi_f.py:334-343); inference passes None")
        if img_xs.device.type != "cuda":
            raise RuntimeError("GIMMVFI_F (gimmvfi_b200): inputs must live on a CUDA device; there is no CPU path")
        eng = self.engine
        if getattr(eng, "tensor_cores", None) != int(self.tensor_cores):
            eng.set_tensor_cores(int(self.tensor_cores))
        B = img_xs.shape[0]
        xs = img_xs.to(torch.float32).contiguous()
        if flow_inputs is None and self.flow_backend is not None:
            x_net = xs
            if ds_factor is not None:   # gimmvfi_f.py:309-318: the estimator sees the down-scaled frames
                rs = lambda a: torch.nn.functional.interpolate(a, scale_factor=ds_factor, mode="bilinear", align_corners=False)
                x_net = torch.stack([rs(xs[:, :, 0]), rs(xs[:, :, 1])], 2)
            flow_inputs = self.cal_bidirection_flow(255 * x_net[:, :, 0], 255 * x_net[:, :, 1])
        coords = torch.stack([c[0].to(torch.float32) for c in coord], 0).contiguous()
        tt = torch.stack([x.reshape(-1).to(torch.float32).expand(B) for x in t], 0).contiguous()
        o = eng.forward(xs, coords, tt, ds_factor, aux_outputs=self.aux_outputs, flow_inputs=flow_inputs)
        T = len(t)
        out = {"imgt_pred": [o["imgt_pred"][i] for i in range(T)]}
        if self.aux_outputs:
            out.update({
                "other_pred": [[o["img_warp_4"][i]] for i in range(T)],
                "flowt0_pred": [[o["flowt0_1"][i], o["flowt0_4"][i]] for i in range(T)],
                "flowt1_pred": [[o["flowt1_1"][i], o["flowt1_4"][i]] for i in range(T)],
                "raft_flow": o["raft_flow"],
                "ninrflow": [o["ninrflow"][i] for i in range(T)],
                "nflow": o["nflow"],
                "flowt": [o["flowt"][i].squeeze() for i in range(T)],
            })
        return out

    def compute_psnr(self, preds, targets, reduction="mean"):
        """gimmvfi_f.py:389-..."""
        assert reduction in ["mean", "sum", "none"]
        mse = torch.reshape((preds - targets) ** 2, (preds.shape[0], -1)).mean(dim=-1)
        psnr = -10 * torch.log10(mse)
        return psnr.mean() if reduction == "mean" else (psnr.sum() if reduction == "sum" else psnr)
