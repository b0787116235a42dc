""".flo (Middlebury) optical-flow files — the format the reference's motion-modelling drivers exchange flows in
(src/utils/frame_utils.py:24-44 readFlow, :84-113 writeFlow; used by src/VTF.py / src/VSF.py).  Same bytes on disk:
float32 tag 202021.25, int32 width, int32 height, then height x width x (u, v) float32, little endian."""
import numpy as np

TAG = np.float32(202021.25)


def read_flo(path) -> np.ndarray:
    """-> (h, w, 2) float32; raises ValueError on a bad tag or a truncated file (the reference prints and returns None)"""
    with open(path, "rb") as f:
        head = f.read(12)
        if len(head) < 12 or np.frombuffer(head[:4], "<f4")[0] != TAG:
            raise ValueError("%s: not a .flo file (bad magic number)" % path)
        w, h = (int(v) for v in np.frombuffer(head[4:], "<i4"))
        if w <= 0 or h <= 0:
            raise ValueError("%s: invalid size %dx%d" % (path, w, h))
        data = np.frombuffer(f.read(8 * w * h), "<f4")
    if data.size != 2 * w * h:
        raise ValueError("%s: truncated (%d of %d values)" % (path, data.size, 2 * w * h))
    return data.reshape(h, w, 2).astype(np.float32, copy=True)


def write_flo(path, uv, v=None) -> None:
    """uv: (h, w, 2) array (or u with v given separately, as the reference allows)"""
    uv = np.asarray(uv)
    if v is not None:
        uv = np.stack([uv, np.asarray(v)], -1)
    if uv.ndim != 3 or uv.shape[2] != 2:
        raise ValueError("write_flo: expected (h, w, 2), got %s" % (uv.shape,))
    h, w = uv.shape[:2]
    with open(path, "wb") as f:
        f.write(np.asarray([TAG], "<f4").tobytes())
        f.write(np.asarray([w, h], "<i4").tobytes())
        f.write(np.ascontiguousarray(uv, dtype="<f4").tobytes())


def flo_to_tensor(path):
    """(1, 2, h, w) float32 torch tensor, channel 0 = u (x displacement) — the layout GIMM's ori_flow uses"""
    import torch

    return torch.from_numpy(read_flo(path)).permute(2, 0, 1).unsqueeze(0).contiguous()
