"""GPU-side pre/post-processing of the reference's video driver (src/video_Nx.py:134-216) — SURVEY.md §8(f)
row 2.  uint8 frames go to the device once; /255, replicate-padding to a multiple of 32 (InputPadder,
src/utils/utils.py:156-185), the N-1 interpolations, un-padding and the uint8 (BGR) conversion all happen on
the GPU; only uint8 frames cross the PCIe bus."""
import ctypes as C
from typing import List, Optional

import torch

from ._lib import GimmvfiError, default_lib


class InputPadder:
    """src/utils/utils.py:156-185 — same pad arithmetic, GPU kernels instead of F.pad / slicing."""

    def __init__(self, dims, divisor: int = 32):
        self.ht, self.wd = dims[-2:]
        pad_ht = (((self.ht // divisor) + 1) * divisor - self.ht) % divisor
        pad_wd = (((self.wd // divisor) + 1) * divisor - self.wd) % divisor
        self._pad = [pad_wd // 2, pad_wd - pad_wd // 2, pad_ht // 2, pad_ht - pad_ht // 2]

    @property
    def padded_shape(self):
        return self.ht + self._pad[2] + self._pad[3], self.wd + self._pad[0] + self._pad[1]

    def pad_u8(self, frames_u8: torch.Tensor) -> torch.Tensor:
        """(n,h,w,3) uint8 RGB on the GPU -> (n,3,H,W) float32 in [0,1], replicate-padded."""
        assert frames_u8.is_cuda and frames_u8.dtype == torch.uint8 and frames_u8.dim() == 4 and frames_u8.shape[-1] == 3
        f = frames_u8.contiguous()
        n, h, w, _ = f.shape
        H, W = self.padded_shape
        out = torch.empty(n, 3, H, W, dtype=torch.float32, device=f.device)
        lib = default_lib()
        s = C.c_void_p(torch.cuda.current_stream(f.device).cuda_stream)
        lib.check(lib.dll.gimmvfi_op_frames_u8_to_padded_f32(C.c_void_p(f.data_ptr()), n, h, w, C.c_void_p(out.data_ptr()), H, W,
                                                             self._pad[2], self._pad[0], s))
        return out

    def unpad_u8(self, pred: torch.Tensor, bgr: bool = True) -> torch.Tensor:
        """(n,3,H,W) float32 -> (n,h,w,3) uint8, `(x*255).astype(uint8)`, BGR by default like video_Nx.py:190-196."""
        assert pred.is_cuda and pred.dtype == torch.float32 and pred.dim() == 4 and pred.shape[1] == 3
        p = pred.contiguous()
        n, _, H, W = p.shape
        out = torch.empty(n, self.ht, self.wd, 3, dtype=torch.uint8, device=p.device)
        lib = default_lib()
        s = C.c_void_p(torch.cuda.current_stream(p.device).cuda_stream)
        lib.check(lib.dll.gimmvfi_op_pred_to_u8(C.c_void_p(p.data_ptr()), n, H, W, C.c_void_p(out.data_ptr()), self.ht, self.wd,
                                                self._pad[2], self._pad[0], int(bgr), s))
        return out


@torch.no_grad()
def interpolate_pair_u8(model, frame0_u8: torch.Tensor, frame1_u8: torch.Tensor, N: int = 2, ds_factor: Optional[float] = None,
                        bgr: bool = True) -> List[torch.Tensor]:
    """One iteration of video_Nx.py's frame loop (:134-202): two (h,w,3) uint8 RGB frames -> the N-1 interpolated
    (h,w,3) uint8 frames at t = i/N."""
    dev = frame0_u8.device
    padder = InputPadder(frame0_u8.shape[:2], 32)
    x = padder.pad_u8(torch.stack([frame0_u8, frame1_u8], 0))          # (2,3,H,W)
    xs = torch.stack([x[0], x[1]], 1).unsqueeze(0).contiguous()          # (1,3,2,H,W)
    H, W = xs.shape[-2:]
    ratio = 1.0 if ds_factor is None else ds_factor
    coords = [(model.sample_coord_input(1, (H, W), [i / N], device=dev, upsample_ratio=ratio), None) for i in range(1, N)]
    ts = [i / N * torch.ones(1, device=dev) for i in range(1, N)]
    aux = model.aux_outputs
    model.aux_outputs = False
    try:
        out = model(xs, coords, t=ts, ds_factor=ds_factor)
    finally:
        model.aux_outputs = aux
    return [padder.unpad_u8(im, bgr)[0] for im in out["imgt_pred"]]


class VideoInterpolator:
    """video_Nx.py's frame loop (src/video_Nx.py:134-216) as a streaming object: push() the frames of a clip one by one and
    get the N-1 interpolated frames between the previous frame and the new one.  Consecutive pairs share a frame, so the
    RAFT encoder products of each pair's second frame (fnet map, cnet net/inp, projected context features) stay in a device
    cache and are not recomputed when that frame becomes the next pair's first frame; the outputs
    are bit-identical to independent per-pair calls."""

    def __init__(self, model, N: int = 2, ds_factor: Optional[float] = None, bgr: bool = True):
        self.model, self.N, self.ds, self.bgr = model, int(N), ds_factor, bgr
        self._prev = None      # padded float frame (3,H,W) of the last push
        self._padder = None
        self._cache = None     # uint8 device buffer holding the previous frame's encoder products
        self._cache_valid = False

    def reset(self):
        self._prev, self._cache_valid = None, False

    @torch.no_grad()
    def push(self, frame_u8: torch.Tensor) -> List[torch.Tensor]:
        dev = frame_u8.device
        if self._padder is None or (self._padder.ht, self._padder.wd) != tuple(frame_u8.shape[:2]):
            self._padder = InputPadder(frame_u8.shape[:2], 32)
            self.reset()
        cur = self._padder.pad_u8(frame_u8.unsqueeze(0))[0]              # (3,H,W)
        if self._prev is None:
            self._prev = cur
            return []
        m = self.model
        xs = torch.stack([self._prev, cur], 1).unsqueeze(0).contiguous()  # (1,3,2,H,W)
        H, W = xs.shape[-2:]
        ratio = 1.0 if self.ds is None else self.ds
        N = self.N
        coords = [(m.sample_coord_input(1, (H, W), [i / N], device=dev, upsample_ratio=ratio), None) for i in range(1, N)]
        ts = [i / N * torch.ones(1, device=dev) for i in range(1, N)]
        Hc, Wc = coords[0][0].shape[2:4]
        need = m.engine.frame_cache_bytes(1, H, W, N - 1, self.ds, Hc, Wc)
        if self._cache is None or self._cache.numel() < need or self._cache.device != dev:
            self._cache, self._cache_valid = torch.empty(need, dtype=torch.uint8, device=dev), False
        aux, m.aux_outputs = m.aux_outputs, False
        # The engine keeps its own record of what each cache buffer holds (problem size, precision mode; cleared when the weights
        # change): a refused load -> one full forward with load=False.  The signature adds what only the caller can know.
        sig = (id(m.engine), int(m.tensor_cores), m.engine.weights_version)
        load = self._cache_valid and getattr(self, "_cache_sig", None) == sig
        try:
            try:
                m._frame_cache = (self._cache, load, True)
                out = m(xs, coords, t=ts, ds_factor=self.ds)
            except GimmvfiError as ex:
                if not (load and "frame cache" in str(ex)):
                    raise
                m._frame_cache = (self._cache, False, True)
                out = m(xs, coords, t=ts, ds_factor=self.ds)
            self._cache_valid, self._cache_sig = True, sig
        except Exception:
            self._cache_valid = False
            raise
        finally:
            m.aux_outputs, m._frame_cache = aux, None
        self._prev = cur
        return [self._padder.unpad_u8(im, self.bgr)[0] for im in out["imgt_pred"]]
