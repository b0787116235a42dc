"""Helpers that call the per-kernel C-ABI entry points with torch tensors (NHWC fp32)."""
import ctypes as C

import torch

from gimmvfi_b200._lib import View, default_lib, view_of


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream) if t.is_cuda else None


def nhwc(x):  # NCHW -> contiguous NHWC
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def pack_weight(w):  # (cout,cin,kh,kw) -> [kh*kw][cin][cout_ld]
    cout, cin, kh, kw = w.shape
    ld = (cout + 3) // 4 * 4
    p = torch.zeros(kh * kw, cin, ld, device=w.device)
    p[:, :, :cout] = w.permute(2, 3, 1, 0).reshape(kh * kw, cin, cout)
    return p.contiguous(), ld


def conv2d(x_nhwc, w, b, stride=1, pad=(0, 0), reflect=False, act=0, slope=None, residual=None, x1_nhwc=None, lib=None,
           in_view=None, in1_view=None):
    lib = lib or default_lib()
    cout, cin, kh, kw = w.shape
    pw, ld = pack_weight(w)
    bb = torch.zeros(ld, device=w.device)
    bb[:cout] = b
    n, h, wd, _ = x_nhwc.shape
    oh = (h + 2 * pad[0] - kh) // stride + 1
    ow = (wd + 2 * pad[1] - kw) // stride + 1
    out = torch.empty(n, oh, ow, cout, device=w.device)
    v0 = in_view if in_view is not None else view_of(x_nhwc)
    v1 = in1_view if in1_view is not None else (view_of(x1_nhwc) if x1_nhwc is not None else None)
    rv = view_of(residual) if residual is not None else None
    rc = lib.dll.gimmvfi_op_conv2d(C.byref(v0), C.byref(v1) if v1 is not None else None, C.c_void_p(pw.data_ptr()), C.c_void_p(bb.data_ptr()),
                                   cin, cout, ld, kh, kw, stride, pad[0], pad[1], int(reflect), act,
                                   C.c_void_p(slope.data_ptr()) if slope is not None else None,
                                   C.byref(rv) if rv is not None else None, C.byref(view_of(out)), _stream(out))
    lib.check(rc)
    return out


def softsplat(lat, flow, metric, t, t_mode, lib=None):
    lib = lib or default_lib()
    n, h, w, _ = lat.shape
    scratch = torch.empty(n, h, w, 20, device=lat.device)
    out = torch.empty(n, h, w, 16, device=lat.device)
    lib.check(lib.dll.gimmvfi_op_softsplat(C.byref(view_of(lat)), C.byref(view_of(flow)), C.byref(view_of(metric)), C.c_void_p(t.data_ptr()),
                                           t_mode, C.byref(view_of(scratch)), C.byref(view_of(out)), _stream(out)))
    return out


def softsplat_fused(lat, flow, metric, t, t_mode, absmax, lib=None):
    lib = lib or default_lib()
    n, h, w, _ = lat.shape
    out = torch.empty(n, h, w, 16, device=lat.device)
    lib.check(lib.dll.gimmvfi_op_softsplat_fused(C.byref(view_of(lat)), C.byref(view_of(flow)), C.byref(view_of(metric)), C.c_void_p(t.data_ptr()),
                                                 t_mode, C.c_void_p(absmax.data_ptr()), C.byref(view_of(out)), _stream(out)))
    return out


def backwarp(src, flow, lib=None):
    lib = lib or default_lib()
    out = torch.empty(flow.shape[0], flow.shape[1], flow.shape[2], src.shape[3], device=src.device)
    lib.check(lib.dll.gimmvfi_op_backwarp(C.byref(view_of(src)), C.byref(view_of(flow)), C.byref(view_of(out)), _stream(out)))
    return out


def resize(src, scale, mult=1.0, lib=None):
    lib = lib or default_lib()
    n, h, w, c = src.shape
    out = torch.empty(n, int(h * scale), int(w * scale), c, device=src.device)
    lib.check(lib.dll.gimmvfi_op_resize(C.byref(view_of(src)), C.byref(view_of(out)), float(scale), float(mult), _stream(out)))
    return out


def corr_pyramid(fa, fb, lib=None):
    """fa, fb (n,h,w,C) -> [level tensors (n*h*w, h_l, w_l)]"""
    lib = lib or default_lib()
    n, h, w, _ = fa.shape
    N = h * w
    lv = [torch.empty(n * N, h, w, device=fa.device)]
    lib.check(lib.dll.gimmvfi_op_corr_volume(C.byref(view_of(fa)), C.byref(view_of(fb)), C.c_void_p(lv[0].data_ptr()), _stream(fa)))
    hh, ww = h, w
    for _ in range(3):
        nxt = torch.empty(n * N, hh // 2, ww // 2, device=fa.device)
        lib.check(lib.dll.gimmvfi_op_corr_pool(C.c_void_p(lv[-1].data_ptr()), C.c_void_p(nxt.data_ptr()), n * N, hh, ww, _stream(fa)))
        lv.append(nxt)
        hh, ww = hh // 2, ww // 2
    return lv


def corr_lookup(levels, coords, lib=None):
    lib = lib or default_lib()
    n, h, w, _ = coords.shape
    out = torch.empty(n, h, w, 324, device=coords.device)
    ptrs = (C.c_void_p * 4)(*[l.data_ptr() for l in levels])
    hs = (C.c_int32 * 4)(*[l.shape[1] for l in levels])
    ws = (C.c_int32 * 4)(*[l.shape[2] for l in levels])
    lib.check(lib.dll.gimmvfi_op_corr_lookup(ptrs, hs, ws, C.byref(view_of(coords)), C.byref(view_of(out)), _stream(out)))
    return out


def half_feature_pyramid(tgt):
    """(n,h,w,c) fp32 features -> the 4 average-pooled levels in half, dense NHWC (what the engine prepares for the volume-free lookup)"""
    lv, cur = [], tgt.permute(0, 3, 1, 2)
    for l in range(4):
        lv.append(cur.permute(0, 2, 3, 1).contiguous().half())
        cur = torch.nn.functional.avg_pool2d(cur, 2, 2)
    return lv


def corr_lookup_direct(src, tgt, coords, lib=None, levels=None):
    """src, tgt: (n,h,w,256) fp32 features of the two frames; the half pooled target pyramid is prepared here with torch"""
    lib = lib or default_lib()
    n, h, w, c = src.shape
    lv = levels if levels is not None else half_feature_pyramid(tgt)
    out = torch.empty(n, h, w, 324, device=coords.device)
    ptrs = (C.c_void_p * 4)(*[l.data_ptr() for l in lv])
    hs = (C.c_int32 * 4)(*[l.shape[1] for l in lv])
    ws = (C.c_int32 * 4)(*[l.shape[2] for l in lv])
    lib.check(lib.dll.gimmvfi_op_corr_lookup_direct(C.byref(view_of(src)), ptrs, hs, ws, 1.0 / c ** 0.5, C.byref(view_of(coords)), C.byref(view_of(out)), _stream(out)))
    return out


def instnorm(x, relu, lib=None):
    lib = lib or default_lib()
    n, h, w, c = x.shape
    scratch = torch.empty(int(lib.dll.gimmvfi_instnorm_scratch_floats(n, c)) + 64, device=x.device)
    out = torch.empty_like(x)
    lib.check(lib.dll.gimmvfi_op_instnorm(C.byref(view_of(x)), int(relu), C.c_void_p(scratch.data_ptr()), C.byref(view_of(out)), _stream(x)))
    return out


def convex_upsample(flow, mask, lib=None):
    lib = lib or default_lib()
    n, h, w, _ = flow.shape
    out = torch.empty(n, 8 * h, 8 * w, 2, device=flow.device)
    lib.check(lib.dll.gimmvfi_op_convex_upsample(C.byref(view_of(flow)), C.byref(view_of(mask)), C.byref(view_of(out)), _stream(out)))
    return out


def pixel_shuffle(src, times, lib=None):
    lib = lib or default_lib()
    n, h, w, c = src.shape
    r = 2 ** times
    out = torch.empty(n, h * r, w * r, c // (r * r), device=src.device)
    lib.check(lib.dll.gimmvfi_op_pixel_shuffle(C.byref(view_of(src)), C.byref(view_of(out)), times, _stream(out)))
    return out


def tf32_rn(x):
    i = x.contiguous().view(torch.int32)
    i = (i + 0xFFF + ((i >> 13) & 1)) & ~0x1FFF
    return i.view(torch.float32)


def tf32_trunc(x):
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def tc_cout_pad(cout):
    c16 = (cout + 15) // 16 * 16
    tn = (c16 + 255) // 256
    bn = ((c16 + tn - 1) // tn + 15) // 16 * 16
    return bn * tn


def pack_weight_tc(w):  # (cout,cin,kh,kw) -> [2][kh*kw][cout_pad][cin_pad32]: TF32(w) and TF32(w - TF32(w))
    cout, cin, kh, kw = w.shape
    cp, kp = tc_cout_pad(cout), (cin + 31) // 32 * 32
    p = torch.zeros(2, kh * kw, cp, kp, device=w.device)
    hi = tf32_rn(w)
    lo = tf32_rn(w - hi)
    p[0, :, :cout, :cin] = hi.permute(2, 3, 0, 1).reshape(kh * kw, cout, cin)
    p[1, :, :cout, :cin] = lo.permute(2, 3, 0, 1).reshape(kh * kw, cout, cin)
    return p.contiguous()


def pack_weight_tc_split_f16(w):
    """(cout,cin,kh,kw) -> (half [2][kh*kw][cout_pad][cin_pad64], scale): fp16 hi / lo planes of w * 2^e, max|w| * 2^e in [2^13, 2^14)"""
    import math

    cout, cin, kh, kw = w.shape
    cp, kp = tc_cout_pad(cout), (cin + 63) // 64 * 64
    mx = w.abs().max().item()
    e = 0 if mx == 0 else max(-24, min(24, 14 - math.frexp(mx)[1]))
    ws = (w * (2.0 ** e)).permute(2, 3, 0, 1).reshape(kh * kw, cout, cin)
    hi = ws.half()
    lo = (ws - hi.float()).half()
    p = torch.zeros(2, kh * kw, cp, kp, device=w.device, dtype=torch.float16)
    p[0, :, :cout, :cin] = hi
    p[1, :, :cout, :cin] = lo
    return p.contiguous(), float(2.0 ** e)


def conv2d_tc(x_nhwc, w, b, act1=0, slope1=None, residual=None, act2=0, slope2=None, x1_nhwc=None, in_view=None, lib=None,
              mul=None, gru_z=None, gru_h=None, split=False, out=None, split_f16=False):
    lib = lib or default_lib()
    cout, cin, kh, kw = w.shape
    pw = pack_weight_tc(w)
    pws, wsc = pack_weight_tc_split_f16(w) if (split and split_f16) else (None, 1.0)
    bb = torch.zeros((tc_cout_pad(cout) + 31) // 32 * 32 + 128, device=w.device)
    bb[:cout] = b
    n, h, wd, _ = x_nhwc.shape
    if out is None:
        out = torch.empty(n, h, wd, cout, device=w.device)
    v0 = in_view if in_view is not None else view_of(x_nhwc)
    V = lambda t: C.byref(view_of(t)) if t is not None else None
    P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    rc = lib.dll.gimmvfi_op_conv2d_tc(C.byref(v0), V(x1_nhwc), P(pw), P(bb), cin, cout, kh, kw, act1, P(slope1), V(residual), act2, P(slope2),
                                      V(mul), V(gru_z), V(gru_h), int(split), C.byref(view_of(out)), P(pws), wsc, _stream(out))
    lib.check(rc)
    return out


def pack_weight_tc_f16(w):  # (cout,cin,kh,kw) -> [kh*kw][cout_pad][cin_pad64] half (RN)
    cout, cin, kh, kw = w.shape
    cp, kp = tc_cout_pad(cout), (cin + 63) // 64 * 64
    p = torch.zeros(kh * kw, cp, kp, device=w.device, dtype=torch.float16)
    p[:, :cout, :cin] = w.permute(2, 3, 0, 1).reshape(kh * kw, cout, cin).half()
    return p.contiguous()


def conv2d_tc_f16(x_nhwc, w, b, act1=0, slope1=None, residual=None, act2=0, slope2=None, x1_nhwc=None, out_half=True, lib=None):
    """x (and x1, residual) may be torch.float16 NHWC tensors: half operands run on kind::f16; fp32 x on TF32"""
    lib = lib or default_lib()
    cout, cin, kh, kw = w.shape
    in_half = x_nhwc.dtype == torch.float16
    pw_h = pack_weight_tc_f16(w)
    pw = pack_weight_tc(w)
    bb = torch.zeros((tc_cout_pad(cout) + 31) // 32 * 32 + 128, device=w.device)
    bb[:cout] = b
    n, h, wd, _ = x_nhwc.shape
    out = torch.empty(n, h, wd, cout, device=w.device, dtype=torch.float16 if out_half else torch.float32)
    mask = (1 if in_half else 0) | (2 if out_half else 0) | (4 if (residual is not None and residual.dtype == torch.float16) else 0)
    V = lambda t: C.byref(view_of(t)) if t is not None else None
    P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    lib.check(lib.dll.gimmvfi_op_conv2d_tc_f16(V(x_nhwc), V(x1_nhwc), P(pw_h), P(pw), P(bb), cin, cout, kh, kw, act1, P(slope1), V(residual), act2,
                                               P(slope2), mask, V(out), _stream(out)))
    return out


def conv2d_halo(x_nhwc, w, b, act1=0, slope1=None, residual=None, act2=0, slope2=None, out_half=False, prepadded=False, lib=None):
    """3x3 K-poor conv through csrc/conv_halo.cu; x (and residual) may be torch.float16 NHWC tensors"""
    lib = lib or default_lib()
    cout, cin, kh, kw = w.shape
    assert (kh, kw) in ((3, 3), (1, 1)) and not (prepadded and kh == 1)
    in_half = x_nhwc.dtype == torch.float16
    pw_h = pack_weight_tc_f16(w) if in_half else None
    pw = pack_weight_tc(w)
    bb = torch.zeros((tc_cout_pad(cout) + 31) // 32 * 32 + 128, device=w.device)
    bb[:cout] = b
    n, h, wd, _ = x_nhwc.shape
    oh, ow = (h - 2, wd - 2) if prepadded else (h, wd)
    out = torch.empty(n, oh, ow, cout, device=w.device, dtype=torch.float16 if out_half else torch.float32)
    mask = (1 if in_half else 0) | (2 if out_half else 0) | (4 if (residual is not None and residual.dtype == torch.float16) else 0)
    V = lambda t: C.byref(view_of(t)) if t is not None else None
    P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    lib.check(lib.dll.gimmvfi_op_conv2d_halo(V(x_nhwc), P(pw_h), P(pw), P(bb), cin, cout, act1, P(slope1), V(residual), act2, P(slope2), mask,
                                             2 if kh == 1 else int(prepadded), V(out), _stream(out)))
    return out
