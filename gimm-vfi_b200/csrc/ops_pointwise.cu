// HBM-bound thread-per-element kernels of the GIMM-VFI-R path (warp / resize /
// splat / normalise / blend ...).  Every kernel body is a functor executed by
// gv::parallel_for (CUDA grid-stride kernel, or an OpenMP loop under GV_HOSTSIM).
// Reference semantics are cited per kernel (paths relative to
// /root/reference/src/models/generalizable_INR/).
#include "common.h"

namespace gv {

// ------------------------------------------------------------------ helpers
// torch.linspace(-1, 1, n)[i]  (symmetric evaluation used by ATen's linspace)
GV_HD float linspace_m1_1(int i, int n) {
  if (n <= 1) return -1.f;
  float step = 2.f / (float)(n - 1);
  return (i < n / 2) ? (-1.f + step * (float)i) : (1.f - step * (float)(n - 1 - i));
}

struct BilinearTap {
  int x0, y0, x1, y1;
  float wx0, wx1, wy0, wy1;  // weights of x0/x1 and y0/y1
  bool vx0, vx1, vy0, vy1;
};

// grid_sample(bilinear, padding_mode="border", align_corners=True) at pixel
// (x + fx, y + fy) expressed the way modules/fi_utils.py:19-49 builds the grid:
// base = linspace(-1,1,W_flow), offset = flow / ((W_src-1)/2).
GV_HD BilinearTap border_tap(int x, int y, float fx, float fy, int wf, int hf, int ws, int hs) {
  float gx = linspace_m1_1(x, wf) + fx / (((float)ws - 1.0f) / 2.0f);
  float gy = linspace_m1_1(y, hf) + fy / (((float)hs - 1.0f) / 2.0f);
  float ix = ((gx + 1.f) / 2.f) * (float)(ws - 1);
  float iy = ((gy + 1.f) / 2.f) * (float)(hs - 1);
  ix = fminf(fmaxf(ix, 0.f), (float)(ws - 1));
  iy = fminf(fmaxf(iy, 0.f), (float)(hs - 1));
  BilinearTap t;
  float fx0 = floorf(ix), fy0 = floorf(iy);
  t.x0 = (int)fx0; t.y0 = (int)fy0; t.x1 = t.x0 + 1; t.y1 = t.y0 + 1;
  t.wx1 = ix - fx0; t.wx0 = (fx0 + 1.f) - ix;
  t.wy1 = iy - fy0; t.wy0 = (fy0 + 1.f) - iy;
  t.vx0 = t.x0 >= 0 && t.x0 < ws; t.vx1 = t.x1 >= 0 && t.x1 < ws;
  t.vy0 = t.y0 >= 0 && t.y0 < hs; t.vy1 = t.y1 >= 0 && t.y1 < hs;
  return t;
}

GV_HD float tap_fetch(const TV& s, int n, const BilinearTap& t, int ch) {
  const float* b = s.p + (int64_t)n * s.sn + ch;
  float v = 0.f;
  if (t.vy0 && t.vx0) v += b[((int64_t)t.y0 * s.w + t.x0) * s.ld] * (t.wx0 * t.wy0);
  if (t.vy0 && t.vx1) v += b[((int64_t)t.y0 * s.w + t.x1) * s.ld] * (t.wx1 * t.wy0);
  if (t.vy1 && t.vx0) v += b[((int64_t)t.y1 * s.w + t.x0) * s.ld] * (t.wx0 * t.wy1);
  if (t.vy1 && t.vx1) v += b[((int64_t)t.y1 * s.w + t.x1) * s.ld] * (t.wx1 * t.wy1);
  return v;
}

// decode a flat index over (n, y, x, c)
struct Idx4 { int n, y, x, c; };
GV_HD Idx4 decode4(int64_t i, int h, int w, int c) {
  Idx4 r;
  r.c = (int)(i % c); i /= c;
  r.x = (int)(i % w); i /= w;
  r.y = (int)(i % h); r.n = (int)(i / h);
  return r;
}

// ------------------------------------------------------------- layout moves
struct NchwToNhwcK {
  const float* src; int64_t sn, sc; TV dst; float scale, shift;
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, dst.h, dst.w, dst.c);
    float v = src[(int64_t)q.n * sn + (int64_t)q.c * sc + (int64_t)q.y * dst.w + q.x];
    dst.p[dst.off(q.n, q.y, q.x) + q.c] = v * scale + shift;
  }
};
void nchw_to_nhwc(Ctx& cx, const float* src, int64_t src_sn, int64_t src_sc, const TV& dst, float scale, float shift) {
  parallel_for(cx, dst.pixels() * dst.c, NchwToNhwcK{src, src_sn, src_sc, dst, scale, shift}, "nchw_to_nhwc");
}

struct NhwcToNchwK {
  TV src; float* dst; int64_t sn, sc; float scale, shift; int clamp01;
  GV_HD void operator()(int64_t i) const {
    // iterate in NCHW order so the writes are coalesced
    int x = (int)(i % src.w); int64_t r = i / src.w;
    int y = (int)(r % src.h); r /= src.h;
    int c = (int)(r % src.c); int n = (int)(r / src.c);
    float v = (src.p[src.off(n, y, x) + c] + shift) * scale;
    if (clamp01) v = fminf(fmaxf(v, 0.f), 1.f);
    dst[(int64_t)n * sn + (int64_t)c * sc + (int64_t)y * src.w + x] = v;
  }
};
void nhwc_to_nchw(Ctx& cx, const TV& src, float* dst, int64_t dst_sn, int64_t dst_sc, float scale, float shift, int clamp01) {
  parallel_for(cx, src.pixels() * src.c, NhwcToNchwK{src, dst, dst_sn, dst_sc, scale, shift, clamp01}, "nhwc_to_nchw");
}

struct CopyK {
  TV src, dst;
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, dst.h, dst.w, dst.c);
    dst.p[dst.off(q.n, q.y, q.x) + q.c] = src.p[src.off(q.n, q.y, q.x) + q.c];
  }
};
struct CopyK4 {
  TV src, dst;
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, dst.h, dst.w, dst.c / 4);
    st4(dst.p + dst.off(q.n, q.y, q.x) + q.c * 4, ld4(src.p + src.off(q.n, q.y, q.x) + q.c * 4));
  }
};
struct CopyAnyK {   // either side may be half precision (precision mode 4 concat buffers)
  TV src, dst;
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, dst.h, dst.w, dst.c);
    st1(dst, dst.off(q.n, q.y, q.x) + q.c, ld1(src, src.off(q.n, q.y, q.x) + q.c));
  }
};
void copy_channels(Ctx& cx, const TV& src, const TV& dst) {
  if (src.f16 || dst.f16) { parallel_for(cx, dst.pixels() * dst.c, CopyAnyK{src, dst}, "copy_channels"); return; }
  TV s4 = src; s4.c = dst.c;
  if (vec4_ok(s4) && vec4_ok(dst)) { parallel_for(cx, dst.pixels() * (dst.c / 4), CopyK4{src, dst}, "copy_channels"); return; }
  parallel_for(cx, dst.pixels() * dst.c, CopyK{src, dst}, "copy_channels");
}

// ---- driver-side pre/post-processing (src/video_Nx.py:40-50,152-153,182-202; src/utils/utils.py:156-185) ----
// uint8 HWC RGB frame -> float32 CHW in [0,1] (x / 255.0) replicate-padded to (H, W): InputPadder.pad
struct FramesU8ToF32K {
  const uint8_t* src; float* dst; int h, w, H, W, pt, pl;
  GV_HD void operator()(int64_t i) const {
    int X = (int)(i % W); int64_t r = i / W; int Y = (int)(r % H); r /= H; int c = (int)(r % 3); int n = (int)(r / 3);
    int y = Y - pt, x = X - pl;
    y = y < 0 ? 0 : (y >= h ? h - 1 : y); x = x < 0 ? 0 : (x >= w ? w - 1 : x);
    dst[i] = (float)src[(((int64_t)n * h + y) * w + x) * 3 + c] / 255.0f;
  }
};
void frames_u8_to_padded_f32(Ctx& cx, const uint8_t* src, int n, int h, int w, float* dst_nchw, int H, int W, int pad_top, int pad_left) {
  parallel_for(cx, (int64_t)n * 3 * H * W, FramesU8ToF32K{src, dst_nchw, h, w, H, W, pad_top, pad_left}, "frames_u8_to_f32");
}
// float32 CHW prediction -> unpadded uint8 HWC, (x * 255.0) truncated like numpy's astype(uint8); optional BGR
struct PredToU8K {
  const float* src; uint8_t* dst; int H, W, h, w, pt, pl, bgr;
  GV_HD void operator()(int64_t i) const {
    int c = (int)(i % 3); int64_t r = i / 3; int x = (int)(r % w); r /= w; int y = (int)(r % h); int n = (int)(r / h);
    int cs = bgr ? 2 - c : c;
    float v = src[(((int64_t)n * 3 + cs) * H + (y + pt)) * W + (x + pl)] * 255.0f;
    v = v < 0.f ? 0.f : (v > 255.f ? 255.f : v);
    dst[i] = (uint8_t)v;
  }
};
void pred_to_u8(Ctx& cx, const float* src_nchw, int n, int H, int W, uint8_t* dst, int h, int w, int pad_top, int pad_left, int bgr) {
  parallel_for(cx, (int64_t)n * h * w * 3, PredToU8K{src_nchw, dst, H, W, h, w, pad_top, pad_left, bgr}, "pred_to_u8");
}

// zero-padded copy of a 3-channel image stored with 4-float pixels (4th lane forced to 0)
struct PadImage4K {
  TV src, dst; int pad;
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, dst.h, dst.w, 4);
    int y = q.y - pad, x = q.x - pad;
    float v = 0.f;
    if (q.c < 3 && y >= 0 && y < src.h && x >= 0 && x < src.w) v = src.p[src.off(q.n, y, x) + q.c];
    dst.p[dst.off(q.n, q.y, q.x) + q.c] = v;
  }
};
void pad_image4(Ctx& cx, const TV& src, const TV& dst, int pad) {
  parallel_for(cx, dst.pixels() * 4, PadImage4K{src, dst, pad}, "pad_image4");
}

// reflect-padded copy (padding_mode="reflect"): lets the tensor-core conv, whose TMA can only zero-fill, run the
// reflect-padded layers (gimmvfi_r.py:94-96,106-108) as a "valid" conv on the padded buffer
struct PadReflectK {
  TV src, dst; int pad;
  GV_HD int refl(int v, int n) const { return v < 0 ? -v : (v >= n ? 2 * n - 2 - v : v); }
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, dst.h, dst.w, dst.c);
    int y = refl(q.y - pad, src.h), x = refl(q.x - pad, src.w);
    dst.p[dst.off(q.n, q.y, q.x) + q.c] = src.p[src.off(q.n, y, x) + q.c];
  }
};
struct PadReflectK4 {
  TV src, dst; int pad;
  GV_HD int refl(int v, int n) const { return v < 0 ? -v : (v >= n ? 2 * n - 2 - v : v); }
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, dst.h, dst.w, dst.c / 4);
    int y = refl(q.y - pad, src.h), x = refl(q.x - pad, src.w);
    st4(dst.p + dst.off(q.n, q.y, q.x) + q.c * 4, ld4(src.p + src.off(q.n, y, x) + q.c * 4));
  }
};
struct PadReflectK16 {   // half-precision tensors (2-byte elements)
  TV src, dst; int pad;
  GV_HD int refl(int v, int n) const { return v < 0 ? -v : (v >= n ? 2 * n - 2 - v : v); }
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, dst.h, dst.w, dst.c);
    int y = refl(q.y - pad, src.h), x = refl(q.x - pad, src.w);
    reinterpret_cast<uint16_t*>(dst.p)[dst.off(q.n, q.y, q.x) + q.c] = reinterpret_cast<const uint16_t*>(src.p)[src.off(q.n, y, x) + q.c];
  }
};
void pad_reflect(Ctx& cx, const TV& src, const TV& dst, int pad) {
  if (src.f16 != dst.f16) throw std::runtime_error("pad_reflect: source and destination must have the same storage type");
  if (src.f16) {
    // a pure copy: two half channels are one 32-bit element, so an even-strided half tensor is an fp32 tensor of half the width
    if (src.c % 8 == 0 && src.ld % 8 == 0 && dst.ld % 8 == 0 && src.sn % 8 == 0 && dst.sn % 8 == 0 &&
        (reinterpret_cast<uintptr_t>(src.p) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst.p) & 15) == 0) {
      TV s2 = src, d2 = dst;
      s2.f16 = d2.f16 = 0; s2.c /= 2; d2.c /= 2; s2.ld /= 2; d2.ld /= 2; s2.sn /= 2; d2.sn /= 2;
      parallel_for(cx, d2.pixels() * (d2.c / 4), PadReflectK4{s2, d2, pad}, "pad_reflect");
      return;
    }
    parallel_for(cx, dst.pixels() * dst.c, PadReflectK16{src, dst, pad}, "pad_reflect"); return;
  }
  if (vec4_ok(src) && vec4_ok(dst)) { parallel_for(cx, dst.pixels() * (dst.c / 4), PadReflectK4{src, dst, pad}, "pad_reflect"); return; }
  parallel_for(cx, dst.pixels() * dst.c, PadReflectK{src, dst, pad}, "pad_reflect");
}

// zero-padded copy: dst (n, h+2p, w+2p) gets src's channels in its interior, zeros elsewhere (all dst.ld lanes)
struct PadZeroK {
  TV src, dst; int pad;
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, dst.h, dst.w, dst.ld);
    int y = q.y - pad, x = q.x - pad;
    float v = 0.f;
    if (q.c < src.c && y >= 0 && y < src.h && x >= 0 && x < src.w) v = src.p[src.off(q.n, y, x) + q.c];
    dst.p[dst.off(q.n, q.y, q.x) + q.c] = v;
  }
};
struct PadZeroK4 {   // 4 lanes per thread (16-byte loads / stores); lanes >= src.c are written as zeros like PadZeroK does
  TV src, dst; int pad;
  GV_HD void operator()(int64_t i) const {
    const int g4 = dst.ld >> 2;
    const int g = (int)(i % g4); int64_t r = i / g4;
    const int px = (int)(r % dst.w); r /= dst.w; const int py = (int)(r % dst.h); const int n = (int)(r / dst.h);
    const int y = py - pad, x = px - pad, c0 = g * 4;
    F4 v = {0.f, 0.f, 0.f, 0.f};
    if (c0 < src.c && y >= 0 && y < src.h && x >= 0 && x < src.w) {
      v = ld4(src.p + src.off(n, y, x) + c0);   // (c0 + 3 < src.ld: both pixel strides are multiples of 4)
      if (c0 + 1 >= src.c) v.y = 0.f;
      if (c0 + 2 >= src.c) v.z = 0.f;
      if (c0 + 3 >= src.c) v.w = 0.f;
    }
    st4(dst.p + dst.off(n, py, px) + c0, v);
  }
};
void pad_zero(Ctx& cx, const TV& src, const TV& dst, int pad) {
  auto al16 = [](const TV& t) { return (reinterpret_cast<uintptr_t>(t.p) & 15) == 0 && t.ld % 4 == 0 && t.sn % 4 == 0 && !t.f16; };
  if (al16(src) && al16(dst) && ((src.c + 3) & ~3) <= src.ld) {
    parallel_for(cx, dst.pixels() * (dst.ld / 4), PadZeroK4{src, dst, pad}, "pad_zero");
    return;
  }
  parallel_for(cx, dst.pixels() * dst.ld, PadZeroK{src, dst, pad}, "pad_zero");
}

struct FillK {
  TV dst; float v;
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, dst.h, dst.w, dst.c);
    dst.p[dst.off(q.n, q.y, q.x) + q.c] = v;
  }
};
void fill(Ctx& cx, const TV& dst, float v) { parallel_for(cx, dst.pixels() * dst.c, FillK{dst, v}, "fill"); }

struct AxpbyK {
  TV a, b, out; float alpha, beta;
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, out.h, out.w, out.c);
    float v = alpha * a.p[a.off(q.n, q.y, q.x) + q.c];
    if (b.p) v += beta * b.p[b.off(q.n, q.y, q.x) + q.c];
    out.p[out.off(q.n, q.y, q.x) + q.c] = v;
  }
};
void axpby(Ctx& cx, const TV& a, float alpha, const TV& b, float beta, const TV& out) {
  parallel_for(cx, out.pixels() * out.c, AxpbyK{a, b, out, alpha, beta}, "axpby");
}

// ------------------------------------------------------------------ resize
// F.interpolate(mode="bilinear", align_corners=False) — modules/fi_utils.py:67-70.
// src coordinate = rscale * (dst + 0.5) - 0.5 clamped at 0 (ATen area_pixel_compute_source_index).
struct ResizeK {
  TV src, dst; float rsy, rsx, mult; int accumulate, act;
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, dst.h, dst.w, dst.c);
    float sy = rsy * ((float)q.y + 0.5f) - 0.5f; if (sy < 0.f) sy = 0.f;
    float sx = rsx * ((float)q.x + 0.5f) - 0.5f; if (sx < 0.f) sx = 0.f;
    int y0 = (int)sy, x0 = (int)sx;
    if (y0 > src.h - 1) y0 = src.h - 1;
    if (x0 > src.w - 1) x0 = src.w - 1;
    int y1 = y0 + (y0 < src.h - 1 ? 1 : 0), x1 = x0 + (x0 < src.w - 1 ? 1 : 0);
    float ly1 = sy - (float)y0, lx1 = sx - (float)x0;
    float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
    const float* b = src.p + (int64_t)q.n * src.sn + q.c;
    float v00 = b[((int64_t)y0 * src.w + x0) * src.ld], v01 = b[((int64_t)y0 * src.w + x1) * src.ld];
    float v10 = b[((int64_t)y1 * src.w + x0) * src.ld], v11 = b[((int64_t)y1 * src.w + x1) * src.ld];
    float v = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
    v *= mult;
    if (act == ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
    float* o = dst.p + dst.off(q.n, q.y, q.x) + q.c;
    *o = accumulate ? (*o + v) : v;
  }
};
struct ResizeK4 {   // 4 channels per thread, the per-component arithmetic of ResizeK unchanged
  TV src, dst; float rsy, rsx, mult; int accumulate, act;
  GV_HD float fin(float v00, float v01, float v10, float v11, float lx0, float lx1, float ly0, float ly1, float old) const {
    float v = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
    v *= mult;
    if (act == ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
    return accumulate ? (old + v) : v;
  }
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, dst.h, dst.w, dst.c / 4);
    float sy = rsy * ((float)q.y + 0.5f) - 0.5f; if (sy < 0.f) sy = 0.f;
    float sx = rsx * ((float)q.x + 0.5f) - 0.5f; if (sx < 0.f) sx = 0.f;
    int y0 = (int)sy, x0 = (int)sx;
    if (y0 > src.h - 1) y0 = src.h - 1;
    if (x0 > src.w - 1) x0 = src.w - 1;
    int y1 = y0 + (y0 < src.h - 1 ? 1 : 0), x1 = x0 + (x0 < src.w - 1 ? 1 : 0);
    float ly1 = sy - (float)y0, lx1 = sx - (float)x0;
    float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
    const float* b = src.p + (int64_t)q.n * src.sn + q.c * 4;
    const F4 a00 = ld4(b + ((int64_t)y0 * src.w + x0) * src.ld), a01 = ld4(b + ((int64_t)y0 * src.w + x1) * src.ld);
    const F4 a10 = ld4(b + ((int64_t)y1 * src.w + x0) * src.ld), a11 = ld4(b + ((int64_t)y1 * src.w + x1) * src.ld);
    float* o = dst.p + dst.off(q.n, q.y, q.x) + q.c * 4;
    F4 old = {0.f, 0.f, 0.f, 0.f};
    if (accumulate) old = ld4(o);
    F4 r;
    r.x = fin(a00.x, a01.x, a10.x, a11.x, lx0, lx1, ly0, ly1, old.x); r.y = fin(a00.y, a01.y, a10.y, a11.y, lx0, lx1, ly0, ly1, old.y);
    r.z = fin(a00.z, a01.z, a10.z, a11.z, lx0, lx1, ly0, ly1, old.z); r.w = fin(a00.w, a01.w, a10.w, a11.w, lx0, lx1, ly0, ly1, old.w);
    st4(o, r);
  }
};
struct ResizeAnyK4 {   // ResizeK4 with half-aware loads / stores (no accumulate mode)
  TV src, dst; float rsy, rsx, mult; int act;
  GV_HD float fin(float v00, float v01, float v10, float v11, float lx0, float lx1, float ly0, float ly1) const {
    float v = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
    v *= mult;
    if (act == ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
    return v;
  }
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, dst.h, dst.w, dst.c / 4);
    float sy = rsy * ((float)q.y + 0.5f) - 0.5f; if (sy < 0.f) sy = 0.f;
    float sx = rsx * ((float)q.x + 0.5f) - 0.5f; if (sx < 0.f) sx = 0.f;
    int y0 = (int)sy, x0 = (int)sx;
    if (y0 > src.h - 1) y0 = src.h - 1;
    if (x0 > src.w - 1) x0 = src.w - 1;
    int y1 = y0 + (y0 < src.h - 1 ? 1 : 0), x1 = x0 + (x0 < src.w - 1 ? 1 : 0);
    float ly1 = sy - (float)y0, lx1 = sx - (float)x0;
    float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
    const int64_t b = (int64_t)q.n * src.sn + q.c * 4;
    const F4 a00 = ld4v(src, b + ((int64_t)y0 * src.w + x0) * src.ld), a01 = ld4v(src, b + ((int64_t)y0 * src.w + x1) * src.ld);
    const F4 a10 = ld4v(src, b + ((int64_t)y1 * src.w + x0) * src.ld), a11 = ld4v(src, b + ((int64_t)y1 * src.w + x1) * src.ld);
    F4 r;
    r.x = fin(a00.x, a01.x, a10.x, a11.x, lx0, lx1, ly0, ly1); r.y = fin(a00.y, a01.y, a10.y, a11.y, lx0, lx1, ly0, ly1);
    r.z = fin(a00.z, a01.z, a10.z, a11.z, lx0, lx1, ly0, ly1); r.w = fin(a00.w, a01.w, a10.w, a11.w, lx0, lx1, ly0, ly1);
    st4v(dst, dst.off(q.n, q.y, q.x) + q.c * 4, r);
  }
};
void resize_bilinear(Ctx& cx, const TV& src, const TV& dst, float rscale_y, float rscale_x, float mult, int accumulate, int act) {
  if (src.f16 || dst.f16) {
    TV s4 = src; s4.c = dst.c;
    if (accumulate || !vec4_ok_any(s4) || !vec4_ok_any(dst)) throw std::runtime_error("resize_bilinear: half-precision tensors need 4-channel-aligned views and no accumulation");
    parallel_for(cx, dst.pixels() * (dst.c / 4), ResizeAnyK4{src, dst, rscale_y, rscale_x, mult, act}, "resize_bilinear");
    return;
  }
  {
    TV s4 = src; s4.c = dst.c;
    if (vec4_ok(s4) && vec4_ok(dst)) {
      parallel_for(cx, dst.pixels() * (dst.c / 4), ResizeK4{src, dst, rscale_y, rscale_x, mult, accumulate, act}, "resize_bilinear");
      return;
    }
  }
  parallel_for(cx, dst.pixels() * dst.c, ResizeK{src, dst, rscale_y, rscale_x, mult, accumulate, act}, "resize_bilinear");
}

// ---------------------------------------------------------------- backwarp
// modules/fi_utils.py:19-49.
struct BackwarpK {
  TV src, flow, dst;
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, dst.h, dst.w, dst.c);
    const float* f = flow.p + flow.off(q.n, q.y, q.x);
    BilinearTap t = border_tap(q.x, q.y, f[0], f[1], flow.w, flow.h, src.w, src.h);
    dst.p[dst.off(q.n, q.y, q.x) + q.c] = tap_fetch(src, q.n, t, q.c);
  }
};
struct BackwarpK4 {   // 4 channels per thread; same tap order / products as tap_fetch
  TV src, flow, dst;
  GV_HD void acc(F4& v, const float* p, float w) const { const F4 a = ld4(p); v.x += a.x * w; v.y += a.y * w; v.z += a.z * w; v.w += a.w * w; }
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, dst.h, dst.w, dst.c / 4);
    const float* f = flow.p + flow.off(q.n, q.y, q.x);
    const BilinearTap t = border_tap(q.x, q.y, f[0], f[1], flow.w, flow.h, src.w, src.h);
    const float* b = src.p + (int64_t)q.n * src.sn + q.c * 4;
    F4 v = {0.f, 0.f, 0.f, 0.f};
    if (t.vy0 && t.vx0) acc(v, b + ((int64_t)t.y0 * src.w + t.x0) * src.ld, t.wx0 * t.wy0);
    if (t.vy0 && t.vx1) acc(v, b + ((int64_t)t.y0 * src.w + t.x1) * src.ld, t.wx1 * t.wy0);
    if (t.vy1 && t.vx0) acc(v, b + ((int64_t)t.y1 * src.w + t.x0) * src.ld, t.wx0 * t.wy1);
    if (t.vy1 && t.vx1) acc(v, b + ((int64_t)t.y1 * src.w + t.x1) * src.ld, t.wx1 * t.wy1);
    st4(dst.p + dst.off(q.n, q.y, q.x) + q.c * 4, v);
  }
};
struct BackwarpAnyK4 {   // BackwarpK4 with half-aware loads / stores
  TV src, flow, dst;
  GV_HD void acc(F4& v, int64_t eoff, float w) const { const F4 a = ld4v(src, eoff); v.x += a.x * w; v.y += a.y * w; v.z += a.z * w; v.w += a.w * w; }
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, dst.h, dst.w, dst.c / 4);
    const float* f = flow.p + flow.off(q.n, q.y, q.x);
    const BilinearTap t = border_tap(q.x, q.y, f[0], f[1], flow.w, flow.h, src.w, src.h);
    const int64_t b = (int64_t)q.n * src.sn + q.c * 4;
    F4 v = {0.f, 0.f, 0.f, 0.f};
    if (t.vy0 && t.vx0) acc(v, b + ((int64_t)t.y0 * src.w + t.x0) * src.ld, t.wx0 * t.wy0);
    if (t.vy0 && t.vx1) acc(v, b + ((int64_t)t.y0 * src.w + t.x1) * src.ld, t.wx1 * t.wy0);
    if (t.vy1 && t.vx0) acc(v, b + ((int64_t)t.y1 * src.w + t.x0) * src.ld, t.wx0 * t.wy1);
    if (t.vy1 && t.vx1) acc(v, b + ((int64_t)t.y1 * src.w + t.x1) * src.ld, t.wx1 * t.wy1);
    st4v(dst, dst.off(q.n, q.y, q.x) + q.c * 4, v);
  }
};
struct BackwarpAnyK {   // scalar: 3-channel image warps into a half-precision concat buffer
  TV src, flow, dst;
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, dst.h, dst.w, dst.c);
    const float* f = flow.p + flow.off(q.n, q.y, q.x);
    const BilinearTap t = border_tap(q.x, q.y, f[0], f[1], flow.w, flow.h, src.w, src.h);
    const int64_t b = (int64_t)q.n * src.sn + q.c;
    float v = 0.f;
    if (t.vy0 && t.vx0) v += ld1(src, b + ((int64_t)t.y0 * src.w + t.x0) * src.ld) * (t.wx0 * t.wy0);
    if (t.vy0 && t.vx1) v += ld1(src, b + ((int64_t)t.y0 * src.w + t.x1) * src.ld) * (t.wx1 * t.wy0);
    if (t.vy1 && t.vx0) v += ld1(src, b + ((int64_t)t.y1 * src.w + t.x0) * src.ld) * (t.wx0 * t.wy1);
    if (t.vy1 && t.vx1) v += ld1(src, b + ((int64_t)t.y1 * src.w + t.x1) * src.ld) * (t.wx1 * t.wy1);
    st1(dst, dst.off(q.n, q.y, q.x) + q.c, v);
  }
};
void backwarp(Ctx& cx, const TV& src, const TV& flow, const TV& dst) {
  if (flow.f16) throw std::runtime_error("backwarp: the flow field must be fp32");
  if (src.f16 || dst.f16) {
    TV s4 = src; s4.c = dst.c;
    if (vec4_ok_any(s4) && vec4_ok_any(dst)) parallel_for(cx, dst.pixels() * (dst.c / 4), BackwarpAnyK4{src, flow, dst}, "backwarp");
    else parallel_for(cx, dst.pixels() * dst.c, BackwarpAnyK{src, flow, dst}, "backwarp");
    return;
  }
  {
    TV s4 = src; s4.c = dst.c;
    if (vec4_ok(s4) && vec4_ok(dst)) { parallel_for(cx, dst.pixels() * (dst.c / 4), BackwarpK4{src, flow, dst}, "backwarp"); return; }
  }
  parallel_for(cx, dst.pixels() * dst.c, BackwarpK{src, flow, dst}, "backwarp");
}

// ----------------------------------------------------------- pixel shuffle
// nn.PixelShuffle(2) applied `times` times (fi_components.py:235, :285-286).
struct PixelShuffleK {
  TV src, dst; int times;
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, dst.h, dst.w, dst.c);
    int y = q.y, x = q.x, c = q.c;
    for (int k = 0; k < times; ++k) {
      c = c * 4 + (y & 1) * 2 + (x & 1);
      y >>= 1; x >>= 1;
    }
    dst.p[dst.off(q.n, q.y, q.x) + q.c] = src.p[src.off(q.n, y, x) + c];
  }
};
void pixel_shuffle(Ctx& cx, const TV& src, const TV& dst, int times) {
  parallel_for(cx, dst.pixels() * dst.c, PixelShuffleK{src, dst, times}, "pixel_shuffle");
}

// ----------------------------------------------------------- instance norm
// nn.InstanceNorm2d(C): no affine, biased variance, eps 1e-5 (raft/extractor.py:30-34,133-134).
static const int IN_CHUNKS_MAX = 1024;
static int in_chunks(const TV& x) {  // ~64 pixels per thread, enough threads to fill 148 SMs
  int64_t hw = (int64_t)x.h * x.w; int64_t c = hw / 64;
  return (int)(c < 1 ? 1 : (c > IN_CHUNKS_MAX ? IN_CHUNKS_MAX : c));
}
int64_t instnorm_scratch_floats(const TV& x) { return (int64_t)x.n * (IN_CHUNKS_MAX + 32) * x.c * 4; }

struct InPartialK {
  TV x; double* part; int chunks;
  GV_HD void operator()(int64_t i) const {
    int c = (int)(i % x.c); int64_t r = i / x.c;
    int ch = (int)(r % chunks); int n = (int)(r / chunks);
    int64_t hw = (int64_t)x.h * x.w;
    int64_t per = (hw + chunks - 1) / chunks;
    int64_t a = (int64_t)ch * per, b = a + per; if (b > hw) b = hw;
    const float* p = x.p + (int64_t)n * x.sn + c;
    // four independent accumulator pairs: four loads in flight per thread (the single dependent chain was latency bound)
    double s = 0.0, s2 = 0.0, t = 0.0, t2 = 0.0, u = 0.0, u2 = 0.0, w = 0.0, w2 = 0.0;
    int64_t k = a;
    for (; k + 3 < b; k += 4) {
      const double v0 = (double)p[k * x.ld], v1 = (double)p[(k + 1) * x.ld], v2 = (double)p[(k + 2) * x.ld], v3 = (double)p[(k + 3) * x.ld];
      s += v0; s2 += v0 * v0; t += v1; t2 += v1 * v1; u += v2; u2 += v2 * v2; w += v3; w2 += v3 * v3;
    }
    for (; k < b; ++k) { const double v = (double)p[k * x.ld]; s += v; s2 += v * v; }
    part[i * 2] = (s + t) + (u + w); part[i * 2 + 1] = (s2 + t2) + (u2 + w2);
  }
};
// two-level tree over the chunk partials (a single thread per (n,c) looping over 1024 chunks was latency bound)
struct InFoldK {   // part[n][chunks][c] -> fold[n][32][c]
  const double* part; double* fold; int c, chunks;
  GV_HD void operator()(int64_t i) const {
    int ch = (int)(i % c); int64_t r = i / c; int g = (int)(r % 32); int n = (int)(r / 32);
    double s = 0.0, s2 = 0.0;
    for (int k = g; k < chunks; k += 32) { int64_t j = ((int64_t)n * chunks + k) * c + ch; s += part[j * 2]; s2 += part[j * 2 + 1]; }
    fold[i * 2] = s; fold[i * 2 + 1] = s2;
  }
};
struct InFinalK {
  const double* fold; float* mr; int c; double inv_hw;
  GV_HD void operator()(int64_t i) const {
    int ch = (int)(i % c); int n = (int)(i / c);
    double s = 0.0, s2 = 0.0;
    for (int g = 0; g < 32; ++g) { int64_t j = ((int64_t)n * 32 + g) * c + ch; s += fold[j * 2]; s2 += fold[j * 2 + 1]; }
    double m = s * inv_hw; double var = s2 * inv_hw - m * m; if (var < 0.0) var = 0.0;
    mr[i * 2] = (float)m; mr[i * 2 + 1] = (float)(1.0 / sqrt(var + 1e-5));
  }
};
void instnorm_stats(Ctx& cx, const TV& x, float* mean_rstd, float* scratch, int64_t) {
  double* part = reinterpret_cast<double*>(scratch);
  const int chunks = in_chunks(x);
  double* fold = part + (int64_t)x.n * IN_CHUNKS_MAX * x.c * 2;
  parallel_for(cx, (int64_t)x.n * chunks * x.c, InPartialK{x, part, chunks}, "instnorm_partial");
  parallel_for(cx, (int64_t)x.n * 32 * x.c, InFoldK{part, fold, x.c, chunks}, "instnorm_fold");
  parallel_for(cx, (int64_t)x.n * x.c, InFinalK{fold, mean_rstd, x.c, 1.0 / ((double)x.h * x.w)}, "instnorm_final");
}
struct InApplyK {
  TV x, res, out; const float* mr; int act1, act2;
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, out.h, out.w, out.c);
    const float* s = mr + ((int64_t)q.n * x.c + q.c) * 2;
    float v = (x.p[x.off(q.n, q.y, q.x) + q.c] - s[0]) * s[1];
    v = apply_act(v, act1, nullptr, 0);
    if (res.p) v += res.p[res.off(q.n, q.y, q.x) + q.c];
    v = apply_act(v, act2, nullptr, 0);
    out.p[out.off(q.n, q.y, q.x) + q.c] = v;
  }
};
struct InApplyK4 {
  TV x, res, out; const float* mr; int act1, act2;
  GV_HD float one(float xv, float m, float r, float rv) const {
    float v = (xv - m) * r;
    v = apply_act(v, act1, nullptr, 0);
    if (res.p) v += rv;
    return apply_act(v, act2, nullptr, 0);
  }
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, out.h, out.w, out.c / 4);
    const int c = q.c * 4;
    const float* s = mr + ((int64_t)q.n * x.c + c) * 2;   // (mean, rstd) pairs of 4 consecutive channels
    const F4 s01 = ld4(s), s23 = ld4(s + 4);
    const F4 a = ld4(x.p + x.off(q.n, q.y, q.x) + c);
    F4 rv = {0.f, 0.f, 0.f, 0.f};
    if (res.p) rv = ld4(res.p + res.off(q.n, q.y, q.x) + c);
    F4 o;
    o.x = one(a.x, s01.x, s01.y, rv.x); o.y = one(a.y, s01.z, s01.w, rv.y);
    o.z = one(a.z, s23.x, s23.y, rv.z); o.w = one(a.w, s23.z, s23.w, rv.w);
    st4(out.p + out.off(q.n, q.y, q.x) + c, o);
  }
};
void instnorm_apply(Ctx& cx, const TV& x, const float* mean_rstd, int act1, const TV& res, int act2, const TV& out) {
  if (vec4_ok(x) && vec4_ok(out) && (!res.p || vec4_ok(res)) && (reinterpret_cast<uintptr_t>(mean_rstd) & 15) == 0 && x.c == out.c) {
    parallel_for(cx, out.pixels() * (out.c / 4), InApplyK4{x, res, out, mean_rstd, act1, act2}, "instnorm_apply");
    return;
  }
  parallel_for(cx, out.pixels() * out.c, InApplyK{x, res, out, mean_rstd, act1, act2}, "instnorm_apply");
}

// ------------------------------------------------------- flow normalisation
// modules/fi_utils.py:52-60: scaler[n] = max |cat[f01, -f10]|.
static const int AM_CHUNKS = 4096;
int64_t absmax_scratch_floats(const TV& a) { return (int64_t)a.n * AM_CHUNKS; }
struct AbsmaxPartK {
  TV a, b; float* part; int chunks;
  GV_HD void operator()(int64_t i) const {
    int ch = (int)(i % chunks); int n = (int)(i / chunks);
    int64_t hw = (int64_t)a.h * a.w; int64_t per = (hw + chunks - 1) / chunks;
    int64_t s = (int64_t)ch * per, e = s + per; if (e > hw) e = hw;
    float m = 0.f;
    for (int64_t k = s; k < e; ++k) {
      const float* pa = a.p + (int64_t)n * a.sn + k * a.ld;
      const float* pb = b.p + (int64_t)n * b.sn + k * b.ld;
      for (int c = 0; c < a.c; ++c) { m = fmaxf(m, fabsf(pa[c])); m = fmaxf(m, fabsf(pb[c])); }
    }
    part[i] = m;
  }
};
struct AbsmaxFinK {
  const float* part; float* out; int chunks;
  GV_HD void operator()(int64_t n) const {
    float m = 0.f; for (int k = 0; k < chunks; ++k) m = fmaxf(m, part[n * chunks + k]);
    out[n] = m;
  }
};
void absmax_per_sample(Ctx& cx, const TV& a, const TV& b, float* out_n, float* scratch) {
  parallel_for(cx, (int64_t)a.n * AM_CHUNKS, AbsmaxPartK{a, b, scratch, AM_CHUNKS}, "absmax_partial");
  parallel_for(cx, (int64_t)a.n, AbsmaxFinK{scratch, out_n, AM_CHUNKS}, "absmax_final");
}

struct NormFlowK {
  TV f01, f10, n0, n1; const float* s;
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, f01.h, f01.w, 2);
    float sc = s[q.n];
    float a = f01.p[f01.off(q.n, q.y, q.x) + q.c];
    float b = -f10.p[f10.off(q.n, q.y, q.x) + q.c];
    n0.p[n0.off(q.n, q.y, q.x) + q.c] = (a / sc + 1.0f) / 2.0f;
    n1.p[n1.off(q.n, q.y, q.x) + q.c] = (b / sc + 1.0f) / 2.0f;
  }
};
void normalize_flow_pair(Ctx& cx, const TV& f01, const TV& f10, const float* scaler, const TV& n0, const TV& n1) {
  parallel_for(cx, f01.pixels() * 2, NormFlowK{f01, f10, n0, n1, scaler}, "normalize_flow");
}

struct UnnormK {
  TV a, out; const float* s;
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, out.h, out.w, out.c);
    out.p[out.off(q.n, q.y, q.x) + q.c] = (a.p[a.off(q.n, q.y, q.x) + q.c] * 2.0f - 1.0f) * s[q.n];
  }
};
void unnormalize_flow(Ctx& cx, const TV& ninr, const float* scaler, const TV& out) {
  parallel_for(cx, out.pixels() * out.c, UnnormK{ninr, out, scaler}, "unnormalize_flow");
}

// ------------------------------------------------------ RAFT coordinate ops
struct InitCoordsK {
  TV c;
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, c.h, c.w, 2);
    c.p[c.off(q.n, q.y, q.x) + q.c] = q.c == 0 ? (float)q.x : (float)q.y;
  }
};
void init_coords(Ctx& cx, const TV& coords) { parallel_for(cx, coords.pixels() * 2, InitCoordsK{coords}, "init_coords"); }

// flow = coords1 - coords0 (raft/raft.py:148), written to up to two consumers.
struct CoordsMinusGridK {
  TV c, a, b;
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, c.h, c.w, 2);
    float v = c.p[c.off(q.n, q.y, q.x) + q.c] - (q.c == 0 ? (float)q.x : (float)q.y);
    if (a.p) a.p[a.off(q.n, q.y, q.x) + q.c] = v;
    if (b.p) b.p[b.off(q.n, q.y, q.x) + q.c] = v;
  }
};
void coords_minus_grid(Ctx& cx, const TV& coords1, const TV& a, const TV& b) {
  parallel_for(cx, coords1.pixels() * 2, CoordsMinusGridK{coords1, a, b}, "coords_minus_grid");
}

// Convex x8 upsampling, raft/raft.py:86-97.  mask (n,h,w,576) already holds 0.25*conv
// (folded into the weights); channel = k*64 + i*8 + j, k = ky*3 + kx.
struct ConvexUpK {
  TV flow, mask, out;
  GV_HD void operator()(int64_t i) const {
    int X = (int)(i % out.w); int64_t r = i / out.w;
    int Y = (int)(r % out.h); int n = (int)(r / out.h);
    int x = X >> 3, j = X & 7, y = Y >> 3, ii = Y & 7;
    const float* m = mask.p + mask.off(n, y, x) + ii * 8 + j;
    float e[9]; float mx = -3.4e38f;
    for (int k = 0; k < 9; ++k) { e[k] = m[k * 64]; mx = fmaxf(mx, e[k]); }
    float sum = 0.f;
    for (int k = 0; k < 9; ++k) { e[k] = expf(e[k] - mx); sum += e[k]; }
    float ox = 0.f, oy = 0.f;
    for (int k = 0; k < 9; ++k) {
      int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
      float fx = 0.f, fy = 0.f;
      if (yy >= 0 && yy < flow.h && xx >= 0 && xx < flow.w) {
        const float* f = flow.p + flow.off(n, yy, xx);
        fx = 8.f * f[0]; fy = 8.f * f[1];
      }
      float wk = e[k] / sum;
      ox += wk * fx; oy += wk * fy;
    }
    float* o = out.p + out.off(n, Y, X);
    o[0] = ox; o[1] = oy;
  }
};
void convex_upsample(Ctx& cx, const TV& flow, const TV& mask, const TV& out) {
  parallel_for(cx, out.pixels(), ConvexUpK{flow, mask, out}, "convex_upsample");
}

// ----------------------------------------------------- GIMM splat weights
// gimmvfi_r.py:444-492 for one direction: var of a 3x3 gaussian (reflect pad) +
// forward/backward consistency  w = 1/(1+err*a_fe) + 1/(1+var*a_v).
struct SplatWeightsK {
  TV fs, fo, out; const float* g9; const float* afe; const float* av;
  GV_HD int refl(int v, int n) const { return v < 0 ? -v : (v >= n ? 2 * n - 2 - v : v); }
  GV_HD void operator()(int64_t i) const {
    int x = (int)(i % fs.w); int64_t r = i / fs.w;
    int y = (int)(r % fs.h); int n = (int)(r / fs.h);
    float var = 0.f;
    for (int c = 0; c < 2; ++c) {
      float sq = 0.f, m = 0.f;
      for (int k = 0; k < 9; ++k) {
        int yy = refl(y + k / 3 - 1, fs.h), xx = refl(x + k % 3 - 1, fs.w);
        float v = fs.p[fs.off(n, yy, xx) + c];
        sq += g9[k] * (v * v); m += g9[k] * v;
      }
      float d = sq - m * m; if (d < 1e-9f) d = 1e-9f;
      var += sqrtf(d);
    }
    var = var / 2.f;
    const float* f = fs.p + fs.off(n, y, x);
    BilinearTap t = border_tap(x, y, f[0], f[1], fs.w, fs.h, fo.w, fo.h);
    float e0 = fabsf(-tap_fetch(fo, n, t, 0) - f[0]);
    float e1 = fabsf(-tap_fetch(fo, n, t, 1) - f[1]);
    float err = (e0 + e1) / 2.f;
    out.p[out.off(n, y, x)] = 1.f / (1.f + err * afe[0]) + 1.f / (1.f + var * av[0]);
  }
};
void splat_weights(Ctx& cx, const TV& f_self, const TV& f_other, const float* g9, const float* alpha_fe, const float* alpha_v, const TV& out) {
  parallel_for(cx, out.pixels(), SplatWeightsK{f_self, f_other, out, g9, alpha_fe, alpha_v}, "splat_weights");
}

// ----------------------------------------------------------- forward splat
// "linear" softmax-splatting, modules/softsplat.py:307-308 + kernel :376-421.
// acc holds 17 channels per target pixel: 16 x sum(in*metric*w) and sum(metric*w).
// One thread per (source pixel, channel): NHWC makes the 17 atomics of a pixel
// land on consecutive addresses (the reference strides them by H*W).
struct SplatAccK {
  TV lat, flow, metric, acc; const float* t; int t_mode;
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, lat.h, lat.w, 17);
    float tt = t[q.n]; float sc = t_mode ? (1.f - tt) : tt;
    const float* f = flow.p + flow.off(q.n, q.y, q.x);
    float fx = (float)q.x + f[0] * sc;
    float fy = (float)q.y + f[1] * sc;
    if (!gv_isfinite(fx) || !gv_isfinite(fy)) return;
    float m = metric.p[metric.off(q.n, q.y, q.x)];
    float v = q.c < 16 ? lat.p[lat.off(q.n, q.y, q.x) + q.c] * m : m;
    float x0f = floorf(fx), y0f = floorf(fy);
    int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
    float wnw = ((float)x1 - fx) * ((float)y1 - fy);
    float wne = (fx - (float)x0) * ((float)y1 - fy);
    float wsw = ((float)x1 - fx) * (fy - (float)y0);
    float wse = (fx - (float)x0) * (fy - (float)y0);
    int W = acc.w, H = acc.h;
    float* base = acc.p + (int64_t)q.n * acc.sn + q.c;
    if (x0 >= 0 && x0 < W && y0 >= 0 && y0 < H) atomic_add_f(base + ((int64_t)y0 * W + x0) * acc.ld, v * wnw);
    if (x1 >= 0 && x1 < W && y0 >= 0 && y0 < H) atomic_add_f(base + ((int64_t)y0 * W + x1) * acc.ld, v * wne);
    if (x0 >= 0 && x0 < W && y1 >= 0 && y1 < H) atomic_add_f(base + ((int64_t)y1 * W + x0) * acc.ld, v * wsw);
    if (x1 >= 0 && x1 < W && y1 >= 0 && y1 < H) atomic_add_f(base + ((int64_t)y1 * W + x1) * acc.ld, v * wse);
  }
};
#ifndef GV_HOSTSIM
// The B200 form of the scatter (the functor above is the reference kernel's thread-per-(pixel, channel) decomposition, kept for the
// host simulation and for ragged layouts).  ONE thread per source pixel:
//   * flow, metric, target cell and the four bilinear weights are computed once (the reference does it 17 times per pixel);
//   * the 16 latent channels arrive as four 16-byte loads; with the metric they form five float4 groups = the 80-byte accumulator
//     pixel [16 x sum(in*m*w) | sum(m*w) | 3 pad lanes];
//   * lanes of a warp hold x-consecutive source pixels.  Under a smooth flow lane l's EAST column of target cells is lane l+1's
//     WEST column: the west contributions travel one lane down by warp shuffle and are summed in registers, so an interior
//     lane issues 2 instead of 4 corner updates;
//   * every update is one red.global.add.v4.f32 (16 bytes per atomic): 10-20 vector reductions per pixel instead of 68 scalar ones.
// Summation order differs from the reference's (whose atomics are unordered anyway): fp32 re-association, ~1e-7 relative.
__device__ __forceinline__ void red_add_v4(float* addr, const F4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__global__ void __launch_bounds__(256) softsplat_acc_v4_kernel(TV lat, TV flow, TV metric, TV acc, const float* __restrict__ t, int t_mode,
                                                              int64_t npix) {
  const int lane = threadIdx.x & 31;
  const int W = acc.w, H = acc.h;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  const int64_t npad = (npix + 31) & ~(int64_t)31;   // whole warps take part in the shuffles
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npad; i += step) {
    bool act = i < npix;
    int x = 0, y = 0, n = 0, x0 = 0, y0 = 0;
    float wnw = 0.f, wne = 0.f, wsw = 0.f, wse = 0.f, m = 0.f;
    F4 v[5];
#pragma unroll
    for (int g = 0; g < 5; ++g) v[g] = F4{0.f, 0.f, 0.f, 0.f};
    if (act) {
      x = (int)(i % W); const int64_t r = i / W; y = (int)(r % H); n = (int)(r / H);
      const float tt = __ldg(t + n), sc = t_mode ? (1.f - tt) : tt;
      const float2 f = *reinterpret_cast<const float2*>(flow.p + flow.off(n, y, x));
      const float fx = (float)x + f.x * sc, fy = (float)y + f.y * sc;
      act = gv_isfinite(fx) && gv_isfinite(fy);
      if (act) {
        const float x0f = floorf(fx), y0f = floorf(fy);
        x0 = (int)x0f; y0 = (int)y0f;
        const float x1f = (float)(x0 + 1), y1f = (float)(y0 + 1);
        wnw = (x1f - fx) * (y1f - fy); wne = (fx - (float)x0) * (y1f - fy);
        wsw = (x1f - fx) * (fy - (float)y0); wse = (fx - (float)x0) * (fy - (float)y0);
        m = metric.p[metric.off(n, y, x)];
        const float* lp = lat.p + lat.off(n, y, x);
#pragma unroll
        for (int g = 0; g < 4; ++g) { const F4 a = ld4(lp + 4 * g); v[g] = F4{a.x * m, a.y * m, a.z * m, a.w * m}; }
        v[4] = F4{m, 0.f, 0.f, 0.f};
      }
    }
    // x-neighbour merge: lane l+1 is pixel x+1 of the same row (rows never straddle a merge: x + 1 < W is required)
    const int px0 = __shfl_down_sync(0xffffffffu, x0, 1), py0 = __shfl_down_sync(0xffffffffu, y0, 1);
    const int pact = __shfl_down_sync(0xffffffffu, (int)act, 1);
    const float pwnw = __shfl_down_sync(0xffffffffu, wnw, 1), pwsw = __shfl_down_sync(0xffffffffu, wsw, 1);
    const bool take = lane < 31 && act && pact && x + 1 < W && px0 == x0 + 1 && py0 == y0;   // my east column == its west column
    const bool taken = __shfl_up_sync(0xffffffffu, (int)take, 1) != 0 && lane > 0;          // lane l-1 carries my west column
    const int x1 = x0 + 1, y1 = y0 + 1;
    const bool in_y0 = y0 >= 0 && y0 < H, in_y1 = y1 >= 0 && y1 < H, in_x0 = x0 >= 0 && x0 < W, in_x1 = x1 >= 0 && x1 < W;
    float* base = acc.p + (int64_t)n * acc.sn;
    float* a_nw = base + ((int64_t)y0 * W + x0) * acc.ld;
    float* a_sw = a_nw + (int64_t)W * acc.ld;
#pragma unroll
    for (int g = 0; g < 5; ++g) {
      // the neighbour's west-column VALUES (its own v * its own west weights), fetched by every lane (shuffles are warp collective)
      F4 pv;
      pv.x = __shfl_down_sync(0xffffffffu, v[g].x, 1); pv.y = __shfl_down_sync(0xffffffffu, v[g].y, 1);
      pv.z = __shfl_down_sync(0xffffffffu, v[g].z, 1); pv.w = __shfl_down_sync(0xffffffffu, v[g].w, 1);
      if (!act) continue;
      F4 ne = F4{v[g].x * wne, v[g].y * wne, v[g].z * wne, v[g].w * wne};
      F4 se = F4{v[g].x * wse, v[g].y * wse, v[g].z * wse, v[g].w * wse};
      if (take) {
        ne.x += pv.x * pwnw; ne.y += pv.y * pwnw; ne.z += pv.z * pwnw; ne.w += pv.w * pwnw;
        se.x += pv.x * pwsw; se.y += pv.y * pwsw; se.z += pv.z * pwsw; se.w += pv.w * pwsw;
      }
      if (!taken) {
        if (in_x0 && in_y0) red_add_v4(a_nw + 4 * g, F4{v[g].x * wnw, v[g].y * wnw, v[g].z * wnw, v[g].w * wnw});
        if (in_x0 && in_y1) red_add_v4(a_sw + 4 * g, F4{v[g].x * wsw, v[g].y * wsw, v[g].z * wsw, v[g].w * wsw});
      }
      if (in_x1 && in_y0) red_add_v4(a_nw + acc.ld + 4 * g, ne);
      if (in_x1 && in_y1) red_add_v4(a_sw + acc.ld + 4 * g, se);
    }
  }
}
// "zeroeps" normalisation, 4 channels per thread
struct SplatNormK4 {
  TV acc, out;
  GV_HD void operator()(int64_t i) const {
    const int g = (int)(i & 3); int64_t r = i >> 2;
    const int x = (int)(r % out.w); r /= out.w; const int y = (int)(r % out.h); const int n = (int)(r / out.h);
    const float* a = acc.p + acc.off(n, y, x);
    float d = a[16]; if (d == 0.f) d = 1.f;
    const F4 v = ld4(a + 4 * g);
    st4(out.p + out.off(n, y, x) + 4 * g, F4{v.x / d, v.y / d, v.z / d, v.w / d});
  }
};
#endif

#ifndef GV_HOSTSIM
// The whole "linear-zeroeps" splat (modules/softsplat.py:286-352) in ONE pass over HBM: a CTA owns a 32 x 32 tile of TARGET pixels whose
// 17-channel accumulators live in shared memory (68 KB), scans the source pixels that can reach it - the tile grown by
// R = ceil(max|flow| * s) + 1, the per-sample flow bound the engine computes anyway for normalize_flow (fi_utils.py:52-60) - adds the
// landing ones with shared-memory atomics, normalises in place and writes the 16 output channels.  DRAM traffic: flows (8 B, re-read
// by neighbouring tiles mostly from L2) + latents and metric of the landing pixels (68 B) + the output (64 B) = the op's algorithmic
// 140 B/px; the three-pass form (memset, global vector reductions, normalise) moves ~430 B/px through a 167 MB accumulator that
// does not fit L2.  Summation order differs from the reference's (unordered atomics there too): fp32 re-association only.
constexpr int SPT = 32;   // tile edge
__global__ void __launch_bounds__(256) softsplat_tile_kernel(TV lat, TV flow, TV metric, TV out, const float* __restrict__ t, int t_mode,
                                                            const float* __restrict__ absmax, int tiles_x, int tiles_y, int n_img) {
  extern __shared__ float sacc[];   // [SPT * SPT][17]
  const int W = out.w, H = out.h, tid = threadIdx.x;
  const int num_tiles = n_img * tiles_x * tiles_y;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    int r = tile; const int tx = r % tiles_x; r /= tiles_x; const int ty = r % tiles_y; const int n = r / tiles_y;
    const int X0 = tx * SPT, Y0 = ty * SPT;
    for (int i = tid; i < SPT * SPT * 17; i += 256) sacc[i] = 0.f;
    const float tt = __ldg(t + n), sc = t_mode ? (1.f - tt) : tt;
    const float bound = __ldg(absmax + n) * fabsf(sc);
    const int R = (bound < 1e6f ? (int)ceilf(bound) : 1000000) + 1;   // (a non-finite / absurd bound degrades to scanning the frame)
    const int xa = max(0, X0 - R), xb = min(W, X0 + SPT + R), ya = max(0, Y0 - R), yb = min(H, Y0 + SPT + R);
    const int rw = xb - xa;
    __syncthreads();
    for (int i = tid; i < rw * (yb - ya); i += 256) {
      const int y = ya + i / rw, x = xa + i % rw;
      const float2 f = *reinterpret_cast<const float2*>(flow.p + flow.off(n, y, x));
      const float fx = (float)x + f.x * sc, fy = (float)y + f.y * sc;
      if (!gv_isfinite(fx) || !gv_isfinite(fy)) continue;
      const float x0f = floorf(fx), y0f = floorf(fy);
      const int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
      if (x1 < X0 || x0 >= X0 + SPT || y1 < Y0 || y0 >= Y0 + SPT) continue;   // no corner in this tile
      const float m = metric.p[metric.off(n, y, x)];
      float v[17];
      const float* lp = lat.p + lat.off(n, y, x);
#pragma unroll
      for (int g = 0; g < 4; ++g) { const F4 a = ld4(lp + 4 * g); v[4 * g] = a.x * m; v[4 * g + 1] = a.y * m; v[4 * g + 2] = a.z * m; v[4 * g + 3] = a.w * m; }
      v[16] = m;
      const float x1f = (float)x1, y1f = (float)y1;
      const float wgt[4] = {(x1f - fx) * (y1f - fy), (fx - (float)x0) * (y1f - fy), (x1f - fx) * (fy - (float)y0), (fx - (float)x0) * (fy - (float)y0)};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int cx = (k & 1) ? x1 : x0, cy = (k & 2) ? y1 : y0;
        if (cx < X0 || cx >= X0 + SPT || cy < Y0 || cy >= Y0 + SPT || cx >= W || cy >= H || cx < 0 || cy < 0) continue;
        float* a = sacc + ((cy - Y0) * SPT + (cx - X0)) * 17;
#pragma unroll
        for (int c = 0; c < 17; ++c) atomicAdd(a + c, v[c] * wgt[k]);
      }
    }
    __syncthreads();
    for (int i = tid; i < SPT * SPT * 4; i += 256) {   // "zeroeps" normalisation + store, 4 channels per thread
      const int g = i & 3, p = i >> 2, lx = p % SPT, ly = p / SPT;
      const int x = X0 + lx, y = Y0 + ly;
      if (x >= W || y >= H) continue;
      const float* a = sacc + p * 17;
      float d = a[16]; if (d == 0.f) d = 1.f;
      st4(out.p + out.off(n, y, x) + 4 * g, F4{a[4 * g] / d, a[4 * g + 1] / d, a[4 * g + 2] / d, a[4 * g + 3] / d});
    }
    __syncthreads();
  }
}
#endif
// fused splat: returns false when the layout does not qualify (caller falls back to memset + accumulate + normalise)
bool softsplat_fused(Ctx& cx, const TV& lat, const TV& flow, const TV& metric, const float* t_per_sample, int t_mode, const float* flow_absmax,
                     const TV& out, bool force) {
#ifndef GV_HOSTSIM
  static int on = -1;
  // off by default: DRAM traffic is 0.85x the op's algorithmic bytes (ncu), but fp32 (and 64-bit integer) atomic adds on SHARED memory are
  // compare-and-swap loops on sm_100a (ATOMS.CAST.SPIN): 115 M warp instructions, 0.318 ms vs 0.249 ms for the three-pass form whose
  // red.global.add.v4.f32 run in the L2 atomic units (profiles/r02_hbm_kernels_probe.log, r02_ncu_softsplat_tile.jsonl)
  if (on < 0) { const char* s = getenv("GIMMVFI_SPLAT_TILE"); on = s ? atoi(s) : 0; }
  if ((!on && !force) || !flow_absmax || lat.c != 16 || out.c != 16 || !vec4_ok(lat) || !vec4_ok(out) || (reinterpret_cast<uintptr_t>(flow.p) & 7) || flow.ld % 2 || flow.sn % 2 ||
      flow.f16 || metric.f16)
    return false;
  if (cx.dry) return true;
  cx.launches++;
  const int tiles_x = (out.w + SPT - 1) / SPT, tiles_y = (out.h + SPT - 1) / SPT;
  const int smem = SPT * SPT * 17 * 4;
  static volatile unsigned char attr[64];
  gv_set_max_smem(softsplat_tile_kernel, smem, attr);
  if (cx.prof) cx.prof->begin(cx.stream, "softsplat_fused", (double)lat.pixels() * 35.0);   // 140 B per pixel
  const int tiles = out.n * tiles_x * tiles_y;
  const int grid = tiles < cx.sm_count * 3 ? tiles : cx.sm_count * 3;
  softsplat_tile_kernel<<<grid, 256, smem, cx.stream>>>(lat, flow, metric, out, t_per_sample, t_mode, flow_absmax, tiles_x, tiles_y, out.n);
  gv_check_launch("softsplat_fused");
  if (cx.prof) cx.prof->end(cx.stream);
  return true;
#else
  (void)cx; (void)lat; (void)flow; (void)metric; (void)t_per_sample; (void)t_mode; (void)flow_absmax; (void)out;
  return false;
#endif
}

void softsplat_accumulate(Ctx& cx, const TV& lat, const TV& flow, const TV& metric, const float* t_per_sample, int t_mode, const TV& acc) {
#ifndef GV_HOSTSIM
  // vector form: 16-byte aligned latent / accumulator pixels (acc.ld % 4: the pad lanes receive +0)
  const bool v4 = lat.c == 16 && vec4_ok(lat) && (reinterpret_cast<uintptr_t>(acc.p) & 15) == 0 && acc.ld % 4 == 0 && acc.ld >= 20 && acc.sn % 4 == 0 &&
                  (reinterpret_cast<uintptr_t>(flow.p) & 7) == 0 && flow.ld % 2 == 0 && flow.sn % 2 == 0 && !acc.f16 && !flow.f16 && !metric.f16;
  if (v4) {
    if (cx.dry) return;
    cx.launches++;
    const int64_t npix = lat.pixels();
    if (cx.prof) cx.prof->begin(cx.stream, "softsplat_accumulate", (double)npix * 35.0);   // 140 B per pixel: 16+1+2 read, 16 written
    int64_t blocks = (npix + 255) / 256, cap = (int64_t)cx.sm_count * 8;
    if (blocks > cap) blocks = cap;
    softsplat_acc_v4_kernel<<<(unsigned)blocks, 256, 0, cx.stream>>>(lat, flow, metric, acc, t_per_sample, t_mode, npix);
    gv_check_launch("softsplat_accumulate");
    if (cx.prof) cx.prof->end(cx.stream);
    return;
  }
#endif
  parallel_for(cx, lat.pixels() * 17, SplatAccK{lat, flow, metric, acc, t_per_sample, t_mode}, "softsplat_accumulate");
}
// "zeroeps" normalisation, modules/softsplat.py:330-344.
struct SplatNormK {
  TV acc, out;
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, out.h, out.w, 16);
    const float* a = acc.p + acc.off(q.n, q.y, q.x);
    float d = a[16]; if (d == 0.f) d = 1.f;
    out.p[out.off(q.n, q.y, q.x) + q.c] = a[q.c] / d;
  }
};
void softsplat_normalize(Ctx& cx, const TV& acc, const TV& out) {
#ifndef GV_HOSTSIM
  if (out.c == 16 && vec4_ok(out) && (reinterpret_cast<uintptr_t>(acc.p) & 15) == 0 && acc.ld % 4 == 0 && acc.sn % 4 == 0) {
    parallel_for(cx, out.pixels() * 4, SplatNormK4{acc, out}, "softsplat_normalize");
    return;
  }
#endif
  parallel_for(cx, out.pixels() * 16, SplatNormK{acc, out}, "softsplat_normalize");
}

// ------------------------------------------------------------ synthesis glue
// gimmvfi_r.py:239-240: flow_t->0 = flow_t * (-t), flow_t->1 = flow_t * (1 - t)
struct ScaleFlowTK {
  TV ft, f0, f1; const float* t;
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, ft.h, ft.w, 2);
    float v = ft.p[ft.off(q.n, q.y, q.x) + q.c]; float tt = t[q.n];
    f0.p[f0.off(q.n, q.y, q.x) + q.c] = v * (-tt);
    f1.p[f1.off(q.n, q.y, q.x) + q.c] = v * (1.0f - tt);
  }
};
void scale_flow_t(Ctx& cx, const TV& flow_t, const float* t_per_sample, const TV& f0, const TV& f1) {
  parallel_for(cx, flow_t.pixels() * 2, ScaleFlowTK{flow_t, f0, f1, t_per_sample}, "scale_flow_t");
}

struct HypoPackK {
  const float* coord; TV dst;
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, dst.h, dst.w, 3);
    dst.p[dst.off(q.n, q.y, q.x) + q.c] = coord[i];
  }
};
void hypo_pack_input(Ctx& cx, const float* coord, const TV& dst) {
  parallel_for(cx, dst.pixels() * 3, HypoPackK{coord, dst}, "hypo_pack_input");
}

// gimmvfi_r.py:494-504: coord + flow * (1/(1-t))  (mode 0)  |  coord + flow * (1/t)  (mode 1)
struct LookupCoordsK {
  TV flow, out; const float* t; int mode;
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, out.h, out.w, 2);
    float tt = t[q.n];
    float sc = mode == 0 ? 1.0f / (1.0f - tt) : 1.0f / tt;
    float g = q.c == 0 ? (float)q.x : (float)q.y;
    out.p[out.off(q.n, q.y, q.x) + q.c] = g + flow.p[flow.off(q.n, q.y, q.x) + q.c] * sc;
  }
};
void lookup_coords(Ctx& cx, const TV& flow, const float* t_per_sample, int mode, const TV& out) {
  parallel_for(cx, out.pixels() * 2, LookupCoordsK{flow, out, t_per_sample, mode}, "lookup_coords");
}

// fi_components.py:272-275 with the init-decoder head re-ordered at pack time to
// [ft(128) | dflow0(2) | dflow1(2) | mask(1)].
struct FlowMaskSplitK {
  TV o, f0i, f1i, f0, f1, mask;
  GV_HD void operator()(int64_t i) const {
    int x = (int)(i % o.w); int64_t r = i / o.w; int y = (int)(r % o.h); int n = (int)(r / o.h);
    const float* s = o.p + o.off(n, y, x) + 128;
    const float* a = f0i.p + f0i.off(n, y, x); const float* b = f1i.p + f1i.off(n, y, x);
    float* A = f0.p + f0.off(n, y, x); float* B = f1.p + f1.off(n, y, x);
    A[0] = a[0] + s[0]; A[1] = a[1] + s[1]; B[0] = b[0] + s[2]; B[1] = b[1] + s[3];
    mask.p[mask.off(n, y, x)] = s[4];
  }
};
void flow_mask_split(Ctx& cx, const TV& out133, const TV& f0_in, const TV& f1_in, const TV& f0, const TV& f1, const TV& mask) {
  parallel_for(cx, out133.pixels(), FlowMaskSplitK{out133, f0_in, f1_in, f0, f1, mask}, "flow_mask_split");
}

struct AddInplaceK {
  TV dst, src;
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, dst.h, dst.w, dst.c);
    dst.p[dst.off(q.n, q.y, q.x) + q.c] += src.p[src.off(q.n, q.y, q.x) + q.c];
  }
};
void add_inplace_slices(Ctx& cx, const TV& dst, const TV& src) {
  parallel_for(cx, dst.pixels() * dst.c, AddInplaceK{dst, src}, "add_inplace");
}

// fi_components.py:331-340: split 24 -> 6/6/3/9, add the repeated base flows / mask, sigmoid.
struct FinalHeadsK {
  TV o, fl0, fl1, m, of0, of1, om, ores;
  GV_HD void operator()(int64_t i) const {
    int x = (int)(i % o.w); int64_t r = i / o.w; int y = (int)(r % o.h); int n = (int)(r / o.h);
    const float* s = o.p + o.off(n, y, x);
    const float* a = fl0.p + fl0.off(n, y, x); const float* b = fl1.p + fl1.off(n, y, x);
    float mm = m.p[m.off(n, y, x)];
    float* A = of0.p + of0.off(n, y, x); float* B = of1.p + of1.off(n, y, x);
    float* M = om.p + om.off(n, y, x); float* R = ores.p + ores.off(n, y, x);
    for (int k = 0; k < 6; ++k) { A[k] = s[k] + a[k & 1]; B[k] = s[6 + k] + b[k & 1]; }
    for (int k = 0; k < 3; ++k) M[k] = 1.f / (1.f + expf(-(s[12 + k] + mm)));
    for (int k = 0; k < 9; ++k) R[k] = s[15 + k];
  }
};
void final_heads(Ctx& cx, const TV& out24, const TV& flow0, const TV& flow1, const TV& mask, const TV& oflow0, const TV& oflow1,
                 const TV& omask, const TV& ores) {
  parallel_for(cx, out24.pixels(), FinalHeadsK{out24, flow0, flow1, mask, oflow0, oflow1, omask, ores}, "final_heads");
}

// warp_w_mask, gimmvfi_r.py:213-220 (flows / sigmoid(mask) already at full res).
struct WarpBlendK {
  TV i0, i1, f0, f1, m, out;
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, out.h, out.w, 3);
    const float* a = f0.p + f0.off(q.n, q.y, q.x); const float* b = f1.p + f1.off(q.n, q.y, q.x);
    BilinearTap t0 = border_tap(q.x, q.y, a[0], a[1], f0.w, f0.h, i0.w, i0.h);
    BilinearTap t1 = border_tap(q.x, q.y, b[0], b[1], f1.w, f1.h, i1.w, i1.h);
    float mm = m.p[m.off(q.n, q.y, q.x)];
    out.p[out.off(q.n, q.y, q.x) + q.c] = mm * tap_fetch(i0, q.n, t0, q.c) + (1.f - mm) * tap_fetch(i1, q.n, t1, q.c);
  }
};
void warp_blend(Ctx& cx, const TV& img0, const TV& img1, const TV& f0, const TV& f1, const TV& mask, const TV& out) {
  parallel_for(cx, out.pixels() * 3, WarpBlendK{img0, img1, f0, f1, mask, out}, "warp_blend");
}

// multi_flow_combine, fi_components.py:57-88: 3 x (2 warps, blend, + residual) and their mean.
struct MultiFlowBlendK {
  TV i0, i1, f0, f1, m, res, w9, mean3;
  GV_HD void operator()(int64_t i) const {
    Idx4 q = decode4(i, w9.h, w9.w, 3);
    float acc = 0.f;
    for (int k = 0; k < 3; ++k) {
      const float* a = f0.p + f0.off(q.n, q.y, q.x) + 2 * k; const float* b = f1.p + f1.off(q.n, q.y, q.x) + 2 * k;
      BilinearTap t0 = border_tap(q.x, q.y, a[0], a[1], f0.w, f0.h, i0.w, i0.h);
      BilinearTap t1 = border_tap(q.x, q.y, b[0], b[1], f1.w, f1.h, i1.w, i1.h);
      float mm = m.p[m.off(q.n, q.y, q.x) + k];
      float v = mm * tap_fetch(i0, q.n, t0, q.c) + (1.f - mm) * tap_fetch(i1, q.n, t1, q.c);
      v += res.p[res.off(q.n, q.y, q.x) + 3 * k + q.c];
      w9.p[w9.off(q.n, q.y, q.x) + 3 * k + q.c] = v;
      acc += v;
    }
    mean3.p[mean3.off(q.n, q.y, q.x) + q.c] = acc / 3.f;
  }
};
// One thread per PIXEL (the form above runs one per (pixel, colour): every thread re-derives the three bilinear taps of both frames and
// re-reads both flow triples): the taps are computed once per flow pair, the four corners of each frame are fetched as 16-byte pixels
// (3 colours + padding lane), the nine blended values leave as three 16-byte stores.  Same arithmetic and summation order per value.
struct MultiFlowBlendPxK {
  TV i0, i1, f0, f1, m, res, w9, mean3;
  GV_HD static void fetch3(const TV& s, int n, const BilinearTap& t, float* v) {
    const float* b = s.p + (int64_t)n * s.sn;
    v[0] = v[1] = v[2] = 0.f;
    if (t.vy0 && t.vx0) { const F4 p = ld4(b + ((int64_t)t.y0 * s.w + t.x0) * s.ld); const float w = t.wx0 * t.wy0; v[0] += p.x * w; v[1] += p.y * w; v[2] += p.z * w; }
    if (t.vy0 && t.vx1) { const F4 p = ld4(b + ((int64_t)t.y0 * s.w + t.x1) * s.ld); const float w = t.wx1 * t.wy0; v[0] += p.x * w; v[1] += p.y * w; v[2] += p.z * w; }
    if (t.vy1 && t.vx0) { const F4 p = ld4(b + ((int64_t)t.y1 * s.w + t.x0) * s.ld); const float w = t.wx0 * t.wy1; v[0] += p.x * w; v[1] += p.y * w; v[2] += p.z * w; }
    if (t.vy1 && t.vx1) { const F4 p = ld4(b + ((int64_t)t.y1 * s.w + t.x1) * s.ld); const float w = t.wx1 * t.wy1; v[0] += p.x * w; v[1] += p.y * w; v[2] += p.z * w; }
  }
  GV_HD void operator()(int64_t i) const {
    const int x = (int)(i % w9.w); int64_t r = i / w9.w; const int y = (int)(r % w9.h); const int n = (int)(r / w9.h);
    const float* a = f0.p + f0.off(n, y, x); const float* b = f1.p + f1.off(n, y, x);
    const float* mp = m.p + m.off(n, y, x); const float* rp = res.p + res.off(n, y, x);
    float out[12]; float acc[3] = {0.f, 0.f, 0.f};
    out[9] = out[10] = out[11] = 0.f;
    for (int k = 0; k < 3; ++k) {
      const BilinearTap t0 = border_tap(x, y, a[2 * k], a[2 * k + 1], f0.w, f0.h, i0.w, i0.h);
      const BilinearTap t1 = border_tap(x, y, b[2 * k], b[2 * k + 1], f1.w, f1.h, i1.w, i1.h);
      float v0[3], v1[3];
      fetch3(i0, n, t0, v0); fetch3(i1, n, t1, v1);
      const float mm = mp[k];
      for (int c = 0; c < 3; ++c) {
        float v = mm * v0[c] + (1.f - mm) * v1[c];
        v += rp[3 * k + c];
        out[3 * k + c] = v;
        acc[c] += v;
      }
    }
    float* o = w9.p + w9.off(n, y, x);
    F4 s0 = {out[0], out[1], out[2], out[3]}, s1 = {out[4], out[5], out[6], out[7]}, s2 = {out[8], out[9], out[10], out[11]};
    st4(o, s0); st4(o + 4, s1); st4(o + 8, s2);
    F4 mn = {acc[0] / 3.f, acc[1] / 3.f, acc[2] / 3.f, 0.f};
    st4(mean3.p + mean3.off(n, y, x), mn);
  }
};
void multi_flow_blend(Ctx& cx, const TV& img0, const TV& img1, const TV& f0, const TV& f1, const TV& mask, const TV& res,
                      const TV& warps9, const TV& mean3) {
  auto al16 = [](const TV& t) { return (reinterpret_cast<uintptr_t>(t.p) & 15) == 0 && t.ld % 4 == 0 && t.sn % 4 == 0 && !t.f16; };
  if (al16(img0) && al16(img1) && img0.ld >= 4 && img1.ld >= 4 && al16(warps9) && warps9.ld >= 12 && al16(mean3) && mean3.ld >= 4 && !f0.f16 && !f1.f16 &&
      !mask.f16 && !res.f16) {
    parallel_for(cx, warps9.pixels(), MultiFlowBlendPxK{img0, img1, f0, f1, mask, res, warps9, mean3}, "multi_flow_blend");
    return;
  }
  parallel_for(cx, warps9.pixels() * 3, MultiFlowBlendK{img0, img1, f0, f1, mask, res, warps9, mean3}, "multi_flow_blend");
}

// fi_components.py:89-92 + gimmvfi_r.py:308: clamp((mean + comb + 1)/2, 0, 1) -> NCHW
struct CombineOutK {
  TV mean3, conv3; float* dst;
  GV_HD void operator()(int64_t i) const {
    int x = (int)(i % mean3.w); int64_t r = i / mean3.w;
    int y = (int)(r % mean3.h); r /= mean3.h; int c = (int)(r % 3); int n = (int)(r / 3);
    float v = mean3.p[mean3.off(n, y, x) + c] + conv3.p[conv3.off(n, y, x) + c];
    v = (v + 1.0f) / 2.f;
    dst[i] = fminf(fmaxf(v, 0.f), 1.f);
  }
};
void combine_output(Ctx& cx, const TV& mean3, const TV& conv3, float* dst_nchw) {
  parallel_for(cx, mean3.pixels() * 3, CombineOutK{mean3, conv3, dst_nchw}, "combine_output");
}

}  // namespace gv
