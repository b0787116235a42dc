"""EngineHandle: owns one native engine (one per GPU / host thread), the packed
weights and a cached workspace; turns torch tensors into the raw pointers of the
C ABI (include/gimmvfi_b200.h).  PyTorch is used for device memory and streams only."""
import contextlib
import ctypes as C
from typing import Dict, Optional

import torch

from ._lib import IO, FlowInputs, Lib, Problem, View, default_lib


class EngineHandle:
    def __init__(self, device, lib: Optional[Lib] = None, allow_hostsim: bool = False):
        self.lib = lib if lib is not None else default_lib()
        self.device = torch.device(device)
        info = self.lib.build_info()
        self.hostsim = "HOSTSIM" in info
        if self.hostsim and not allow_hostsim:
            raise RuntimeError("gimmvfi_b200: refusing to run the test-only host simulation as a product path")
        if not self.hostsim and self.device.type != "cuda":
            raise RuntimeError("gimmvfi_b200 runs on CUDA devices only (got %s); there is no CPU fallback" % self.device)
        h = C.c_void_p()
        idx = self.device.index if self.device.type == "cuda" and self.device.index is not None else 0
        if self.device.type == "cuda":
            idx = torch.cuda.current_device() if self.device.index is None else self.device.index
        self.lib.check(self.lib.dll.gimmvfi_create(idx, C.byref(h)))
        self._h = h
        self._index = idx
        if self.device.type == "cuda":
            self.device = torch.device("cuda", idx)
        self._ws = None
        self._plans: Dict[tuple, int] = {}
        self.weights_loaded = False
        # static_outputs: forward() returns the same output tensors on every call of a problem shape (overwritten by the next call) -
        # with set_cuda_graph(True) and caller-owned input tensors this makes every pointer repeat, i.e. every forward a graph replay
        self.static_outputs = False
        self._static = None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self.lib.dll.gimmvfi_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd: Dict[str, torch.Tensor], gimm_only: bool = False, synthesis_only: bool = False, full_f: bool = False):
        """gimm_only: `sd` is a standalone GIMM checkpoint (gimm.py's module tree), only gimm_forward() is available.
        synthesis_only: `sd` is a GIMM-VFI-F state_dict; its `flow_estimator.*` (FlowFormer) entries are skipped and only
        forward_from_flow() is available.
        full_f: `sd` is a complete GIMM-VFI-F state_dict: forward() runs the native FlowFormer estimator + the synthesis half."""
        for k, v in sd.items():
            if not v.dtype.is_floating_point:
                continue  # num_batches_tracked
            if synthesis_only and k.startswith("flow_estimator."):
                continue
            t = v.detach().to("cpu", torch.float32).contiguous()
            shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
            self.lib.check(self.lib.dll.gimmvfi_load_weight(self._h, k.encode(), C.c_void_p(t.data_ptr()), shape, t.dim()), self._h)
        fin = self.lib.dll.gimmvfi_finalize_weights_gimm if gimm_only else (
            self.lib.dll.gimmvfi_finalize_weights_synthesis if synthesis_only else (
                self.lib.dll.gimmvfi_finalize_weights_f if full_f else self.lib.dll.gimmvfi_finalize_weights))
        self.synthesis_only = synthesis_only
        self.lib.check(fin(self._h), self._h)
        self.weights_loaded = True

    def set_debug(self, on: bool):
        self.lib.check(self.lib.dll.gimmvfi_set_debug(self._h, int(on)), self._h)
        self._plans.clear()

    def set_tensor_cores(self, mode: int):
        """0: fp32 CUDA cores; 1: post-RAFT convs on tcgen05 TF32; 2: + RAFT convs on tcgen05 3xTF32; 3: + the final decoder's
        residual trunk stored in fp16 on tcgen05 kind::f16 (conv_tc.cu); 4 (experimental): + the 32/64-channel full-resolution chains in
        fp16.  The workspace plan depends on the mode."""
        self.lib.check(self.lib.dll.gimmvfi_set_tensor_cores(self._h, int(mode)), self._h)
        self.tensor_cores = int(mode)
        self._plans.clear()

    def set_cuda_graph(self, on: bool):
        """Replay forward() from a recorded CUDA graph whenever problem, tensors (same addresses), stream and mode repeat; see
        include/gimmvfi_b200.h.  Callers that want replays keep their input / coordinate tensors alive and let outputs recycle."""
        self.lib.check(self.lib.dll.gimmvfi_set_cuda_graph(self._h, int(on)), self._h)

    @property
    def graph_replays(self) -> int:
        return int(self.lib.dll.gimmvfi_graph_replays(self._h))

    def set_profile(self, on: bool):
        self.lib.check(self.lib.dll.gimmvfi_set_profile(self._h, int(on)), self._h)

    def profile(self) -> dict:
        """Per-kernel CUDA-event totals of the last forward (set_profile(True) first)."""
        import json

        stream = torch.cuda.current_stream(self.device).cuda_stream if self.device.type == "cuda" else None
        return json.loads(self.lib.dll.gimmvfi_profile_json(self._h, C.c_void_p(stream)).decode())

    def set_raft_iters(self, iters: int):
        self.lib.check(self.lib.dll.gimmvfi_set_raft_iters(self._h, int(iters)), self._h)

    def set_flowformer_iters(self, iters: int):
        """decoder_depth of the native FlowFormer memory decoder (default 32)"""
        self.lib.check(self.lib.dll.gimmvfi_set_flowformer_iters(self._h, int(iters)), self._h)

    @property
    def last_launches(self) -> int:
        return int(self.lib.dll.gimmvfi_last_launches(self._h))

    @property
    def weights_version(self) -> int:
        return int(self.lib.dll.gimmvfi_weights_version(self._h))

    def _check_inputs(self, *tensors):
        """contiguous fp32 on THIS engine's device (a tensor of another GPU would hand the kernels a foreign pointer)"""
        for x in tensors:
            if x.dtype != torch.float32 or not x.is_contiguous():
                raise RuntimeError("gimmvfi_b200: inputs must be contiguous float32 tensors")
            if x.device.type != self.device.type or (self.device.type == "cuda" and x.device.index != self._index):
                raise RuntimeError("gimmvfi_b200: input on %s, engine on %s:%d" % (x.device, self.device.type, self._index))

    def _guard(self):
        """torch's current device = the engine's for the duration of a call (stream lookup, torch.empty)"""
        return torch.cuda.device(self._index) if self.device.type == "cuda" else contextlib.nullcontext()

    # ------------------------------------------------------------------ forward
    def _problem(self, B, Hf, Wf, T, ds, Hc, Wc) -> Problem:
        return Problem(B, Hf, Wf, T, float(ds) if ds else 0.0, Hc, Wc)

    def workspace_bytes(self, B, Hf, Wf, T, ds, Hc, Wc, from_flow: bool = False) -> int:
        key = (B, Hf, Wf, T, float(ds) if ds else 0.0, Hc, Wc)
        pkey = key + (bool(from_flow),)
        if pkey not in self._plans:
            p = self._problem(*key)
            n = C.c_size_t()
            plan = self.lib.dll.gimmvfi_plan_from_flow if from_flow else self.lib.dll.gimmvfi_plan
            self.lib.check(plan(self._h, C.byref(p), C.byref(n)), self._h)
            self._plans[pkey] = int(n.value)
        return self._plans[pkey]

    def frame_cache_bytes(self, B, Hf, Wf, T, ds, Hc, Wc) -> int:
        p = self._problem(B, Hf, Wf, T, ds, Hc, Wc)
        return int(self.lib.dll.gimmvfi_frame_cache_bytes(C.byref(p)))

    def forward(self, img_xs: torch.Tensor, coords: torch.Tensor, t: torch.Tensor, ds: Optional[float] = None,
                aux_outputs: bool = True, frame_cache=None, flow_inputs=None) -> Dict[str, torch.Tensor]:
        """img_xs (B,3,2,Hf,Wf), coords (T,B,1,Hc,Wc,3), t (T,B): contiguous fp32 on self.device.
        frame_cache = (uint8 device tensor of frame_cache_bytes(), load, store): see gimmvfi_set_frame_cache.
        flow_inputs = dict(flows (B,2,2,H,W), feat4 [2 x (B,128,H/4,W/4)], feat8 [2 x (B,256,H/8,W/8)], fnet [2 x (B,256,H/8,W/8)]):
        the outputs of an external flow estimator at the network resolution -> gimmvfi_forward_from_flow (GIMM-VFI-F)."""
        self._check_inputs(img_xs, coords, t)
        if flow_inputs is not None:
            flow_inputs = {k: (v.to(torch.float32).contiguous() if torch.is_tensor(v) else [u.to(torch.float32).contiguous() for u in v])
                           for k, v in flow_inputs.items()}
            self._check_inputs(flow_inputs["flows"], *flow_inputs["feat4"], *flow_inputs["feat8"], *flow_inputs["fnet"])
        with self._guard():
            return self._forward(img_xs, coords, t, ds, aux_outputs, frame_cache, flow_inputs)

    def _forward(self, img_xs, coords, t, ds, aux_outputs, frame_cache, flow_inputs=None):
        B, _, _, Hf, Wf = img_xs.shape
        T, _, _, Hc, Wc, _ = coords.shape
        H, W = (Hf, Wf) if not ds else (int(Hf * ds), int(Wf * ds))
        nbytes = self.workspace_bytes(B, Hf, Wf, T, ds, Hc, Wc, from_flow=flow_inputs is not None)
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        E = lambda *s: torch.empty(*s, dtype=torch.float32, device=self.device)
        okey = (B, Hf, Wf, T, float(ds) if ds else 0.0, Hc, Wc, bool(aux_outputs))
        if self.static_outputs and self._static is not None and self._static[0] == okey:
            out = self._static[1]   # the SAME tensors as the previous call (aliasing is the caller's choice: static_outputs = True)
            return self._launch(out, img_xs, coords, t, B, Hf, Wf, T, ds, Hc, Wc, frame_cache, flow_inputs)
        out = {"imgt_pred": E(T, B, 3, Hf, Wf)}
        if flow_inputs is not None:
            fi = flow_inputs
            assert tuple(fi["flows"].shape) == (B, 2, 2, H, W), (tuple(fi["flows"].shape), (B, 2, 2, H, W))
            for j in range(2):
                assert tuple(fi["feat4"][j].shape) == (B, 128, H // 4, W // 4) and tuple(fi["feat8"][j].shape) == (B, 256, H // 8, W // 8)
                assert tuple(fi["fnet"][j].shape) == (B, 256, H // 8, W // 8)
        if aux_outputs:
            out.update(
                img_warp_4=E(T, B, 3, H, W), flowt0_1=E(T, B, 3, 2, Hf, Wf), flowt1_1=E(T, B, 3, 2, Hf, Wf),
                flowt0_4=E(T, B, 2, H // 4, W // 4), flowt1_4=E(T, B, 2, H // 4, W // 4), raft_flow=E(B, 2, 2, H, W),
                nflow=E(B, 2, 2, H, W), ninrflow=E(T, B, 2, 1, Hc, Wc), flowt=E(T, B, 2, Hc, Wc))
        if self.static_outputs:
            self._static = (okey, out)
        return self._launch(out, img_xs, coords, t, B, Hf, Wf, T, ds, Hc, Wc, frame_cache, flow_inputs)

    def _launch(self, out, img_xs, coords, t, B, Hf, Wf, T, ds, Hc, Wc, frame_cache, flow_inputs):
        io = IO()
        io.img_xs, io.coords, io.t = img_xs.data_ptr(), coords.data_ptr(), t.data_ptr()
        for k, v in out.items():
            setattr(io, k, v.data_ptr())
        p = self._problem(B, Hf, Wf, T, ds, Hc, Wc)
        stream = torch.cuda.current_stream(self.device).cuda_stream if self.device.type == "cuda" else None
        if frame_cache is not None:
            buf, load, store = frame_cache
            assert buf.dtype == torch.uint8 and buf.is_contiguous() and buf.device == self.device
            self.lib.check(self.lib.dll.gimmvfi_set_frame_cache(self._h, C.c_void_p(buf.data_ptr()), buf.numel(), int(load), int(store)), self._h)
        try:
            if flow_inputs is not None:
                f = FlowInputs()
                f.flows = flow_inputs["flows"].data_ptr()
                for j in range(2):
                    f.feat4[j], f.feat8[j], f.fnet[j] = (flow_inputs[k][j].data_ptr() for k in ("feat4", "feat8", "fnet"))
                self.lib.check(self.lib.dll.gimmvfi_forward_from_flow(self._h, C.byref(p), C.byref(io), C.byref(f), C.c_void_p(self._ws.data_ptr()),
                                                                      self._ws.numel(), C.c_void_p(stream)), self._h)
            else:
                self.lib.check(self.lib.dll.gimmvfi_forward(self._h, C.byref(p), C.byref(io), C.c_void_p(self._ws.data_ptr()),
                                                            self._ws.numel(), C.c_void_p(stream)), self._h)
        finally:
            if frame_cache is not None:
                self.lib.check(self.lib.dll.gimmvfi_set_frame_cache(self._h, None, 0, 0, 0), self._h)
        return out

    def gimm_forward(self, xs: torch.Tensor, ori_flow: torch.Tensor, coords: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """GIMM.forward (gimm.py:129-214): xs, ori_flow (B,2,2,H,W), coords (T,B,1,H,W,3), t (T,B) -> (T,B,2,1,H,W)."""
        self._check_inputs(xs, ori_flow, coords, t)
        with self._guard():
            return self._gimm_forward(xs, ori_flow, coords, t)

    def _gimm_forward(self, xs, ori_flow, coords, t):
        B, _, _, H, W = xs.shape
        T = coords.shape[0]
        assert tuple(ori_flow.shape) == (B, 2, 2, H, W) and tuple(coords.shape) == (T, B, 1, H, W, 3) and tuple(t.shape) == (T, B)
        p = self._problem(B, H, W, T, None, H, W)
        key = ("gimm", B, H, W, T)
        if key not in self._plans:
            n = C.c_size_t()
            self.lib.check(self.lib.dll.gimmvfi_gimm_plan(self._h, C.byref(p), C.byref(n)), self._h)
            self._plans[key] = int(n.value)
        nbytes = self._plans[key]
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        out = torch.empty(T, B, 2, 1, H, W, dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream if self.device.type == "cuda" else None
        P = lambda x: C.c_void_p(x.data_ptr())
        self.lib.check(self.lib.dll.gimmvfi_gimm_forward(self._h, C.byref(p), P(xs), P(ori_flow), P(coords), P(t), P(out), P(self._ws),
                                                         self._ws.numel(), C.c_void_p(stream)), self._h)
        return out

    def tap(self, name: str) -> torch.Tensor:
        """Debug: copy of an intermediate NHWC tensor of the last forward (set_debug(True) first)."""
        v = View()
        self.lib.check(self.lib.dll.gimmvfi_get_tap(self._h, name.encode(), C.byref(v)), self._h)
        base = self._ws.data_ptr()
        off = (v.data - base) // 4
        flat = self._ws.view(torch.float32)
        t = torch.as_strided(flat, (v.n, v.h, v.w, v.c), (v.batch_stride, v.w * v.pixel_stride, v.pixel_stride, 1), off)
        return t.clone()
