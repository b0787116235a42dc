// All-pairs correlation volume, its 2x2 average-pooled pyramid and the 4-level
// 9x9 bilinear lookup (reference: raft/corr.py:127-175, :23-93; sampler
// raft/utils/utils.py:66-80).  fp32 throughout (the RAFT recurrence amplifies
// rounding: see DESIGN.md "precision plan").
#include "common.h"

namespace gv {

// ----------------------------------------------------------------- volume
#ifndef GV_HOSTSIM
// C[n][i][j] = scale * sum_k A[n,i,k] * B[n,j,k];  128x128 tile, BK=16, 8x8 per thread.
__global__ void __launch_bounds__(256) corr_gemm_nt_kernel(TV fa, TV fb, float* __restrict__ vol, float scale) {
  const int M = fa.h * fa.w, K = fa.c;
  const int n = blockIdx.z;
  const int m0 = blockIdx.y * 128, n0 = blockIdx.x * 128;
  __shared__ float As[16][128 + 4];
  __shared__ float Bs[16][128 + 4];
  const int tid = threadIdx.x;
  const int tm = (tid / 16) * 8, tn = (tid % 16) * 8;
  const float* A = fa.p + (int64_t)n * fa.sn;
  const float* B = fb.p + (int64_t)n * fb.sn;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      int f4 = tid + it * 256;       // 512 float4 per operand tile
      int row = f4 >> 2, q = f4 & 3;
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
      if (m0 + row < M) va = *reinterpret_cast<const float4*>(A + (int64_t)(m0 + row) * fa.ld + k0 + q * 4);
      if (n0 + row < M) vb = *reinterpret_cast<const float4*>(B + (int64_t)(n0 + row) * fb.ld + k0 + q * 4);
      As[q * 4 + 0][row] = va.x; As[q * 4 + 1][row] = va.y; As[q * 4 + 2][row] = va.z; As[q * 4 + 3][row] = va.w;
      Bs[q * 4 + 0][row] = vb.x; Bs[q * 4 + 1][row] = vb.y; Bs[q * 4 + 2][row] = vb.z; Bs[q * 4 + 3][row] = vb.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i] = As[k][tm + i]; b[i] = Bs[k][tn + i]; }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* C = vol + (int64_t)n * M * M;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int r = m0 + tm + i;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int c = n0 + tn + j;
      if (c < M) C[(int64_t)r * M + c] = acc[i][j] * scale;
    }
  }
}
#endif

void corr_volume(Ctx& cx, const TV& fa, const TV& fb, float* vol, float scale) {
  if (cx.dry) return;
  cx.launches++;
  const int M = fa.h * fa.w;
#ifdef GV_HOSTSIM
  const int K = fa.c;
  for (int n = 0; n < fa.n; ++n) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < M; ++i) {
      const float* a = fa.p + (int64_t)n * fa.sn + (int64_t)i * fa.ld;
      for (int j = 0; j < M; ++j) {
        const float* b = fb.p + (int64_t)n * fb.sn + (int64_t)j * fb.ld;
        float s = 0.f;
        for (int k = 0; k < K; ++k) s += a[k] * b[k];
        vol[((int64_t)n * M + i) * M + j] = s * scale;
      }
    }
  }
#else
  if (fa.c % 16 != 0 || fa.ld % 4 != 0 || fb.ld % 4 != 0) throw std::runtime_error("corr_volume: feature dim must be a multiple of 16");
  dim3 grid((M + 127) / 128, (M + 127) / 128, fa.n);
  if (cx.prof) cx.prof->begin(cx.stream, "corr_gemm_nt", 2.0 * fa.n * (double)M * M * fa.c);
  corr_gemm_nt_kernel<<<grid, 256, 0, cx.stream>>>(fa, fb, vol, scale);
  gv_check_launch("corr_volume");
  if (cx.prof) cx.prof->end(cx.stream);
#endif
}

// [2][pixels][c] planes for the 3xTF32 correlation GEMM: rn_tf32(x) and rn_tf32(x - rn_tf32(x))
struct SplitPlanesK {
  TV src; float* planes; int64_t plane;
  GV_HD float rn(float x) const {  // round-to-nearest-even to TF32
    uint32_t u; memcpy(&u, &x, 4);
    u += 0xfffu + ((u >> 13) & 1u); u &= 0xffffe000u;
    float y; memcpy(&y, &u, 4); return y;
  }
  GV_HD void operator()(int64_t i) const {
    int c = (int)(i % src.c); int64_t px = i / src.c;
    int x = (int)(px % src.w); int64_t r = px / src.w; int y = (int)(r % src.h); int n = (int)(r / src.h);
    float v = src.p[src.off(n, y, x) + c];
    float hi = rn(v);
    planes[i] = hi; planes[plane + i] = rn(v - hi);
  }
};
void split_planes(Ctx& cx, const TV& src, float* planes) {
  int64_t n = src.pixels() * src.c;
  parallel_for(cx, n, SplitPlanesK{src, planes, n}, "split_planes");
}
// half [2][pixels][c] planes for the 3xF16 correlation GEMM: hi = rn_f16(x), lo = rn_f16(x - hi)  (feature maps are O(1): no scaling)
struct SplitPlanesF16K {
  TV src; uint16_t* planes; int64_t plane;
  GV_HD void operator()(int64_t i) const {
    int c = (int)(i % src.c); int64_t px = i / src.c;
    int x = (int)(px % src.w); int64_t r = px / src.w; int y = (int)(r % src.h); int n = (int)(r / src.h);
    const float v = src.p[src.off(n, y, x) + c];
    const uint16_t hi = gv_f2h(v);
    planes[i] = hi; planes[plane + i] = gv_f2h(v - gv_h2f(hi));
  }
};
void split_planes_f16(Ctx& cx, const TV& src, void* planes) {
  int64_t n = src.pixels() * src.c;
  parallel_for(cx, n, SplitPlanesF16K{src, static_cast<uint16_t*>(planes), n}, "split_planes");
}

// 2x2 average of an NHWC feature map (floor semantics): the targets of pyramid level l are the level-(l-1) features pooled
struct PoolFeatK {
  TV src, dst;
  GV_HD void operator()(int64_t i) const {
    const int c = (int)(i % dst.c); int64_t r = i / dst.c;
    const int x = (int)(r % dst.w); r /= dst.w; const int y = (int)(r % dst.h); const int n = (int)(r / dst.h);
    const float* s = src.p + src.off(n, 2 * y, 2 * x) + c;
    const int64_t dx = src.ld, dy = (int64_t)src.w * src.ld;
    dst.p[dst.off(n, y, x) + c] = (s[0] + s[dx] + s[dy] + s[dy + dx]) * 0.25f;
  }
};
void avgpool2_features(Ctx& cx, const TV& src, const TV& dst) {
  parallel_for(cx, dst.pixels() * dst.c, PoolFeatK{src, dst}, "avgpool2_features");
}

// ------------------------------------------------------------------- pool
// F.avg_pool2d(corr, 2, stride=2) over the trailing (h, w) image of every row (raft/corr.py:139-142).
struct CorrPoolK {
  const float* src; float* dst; int h, w, ho, wo;
  GV_HD void operator()(int64_t i) const {
    int x = (int)(i % wo); int64_t r = i / wo; int y = (int)(r % ho); int64_t row = r / ho;
    const float* s = src + row * ((int64_t)h * w) + (int64_t)(2 * y) * w + 2 * x;
    dst[i] = (s[0] + s[1] + s[w] + s[w + 1]) * 0.25f;
  }
};
void corr_pool(Ctx& cx, const float* src, float* dst, int64_t rows, int h, int w) {
  int ho = h / 2, wo = w / 2;
  parallel_for(cx, rows * ho * wo, CorrPoolK{src, dst, h, w, ho, wo}, "corr_pool");
}

#ifndef GV_HOSTSIM
// All three pooled levels of one correlation row in one pass: level 0 is read from HBM once, levels 1 and 2 stay in
// shared memory for their successors (the per-level kernels re-read 1.3x the bytes).  One CTA per row (n, pixel).
__global__ void __launch_bounds__(256) corr_pyramid_kernel(const float* __restrict__ l0, float* __restrict__ l1, float* __restrict__ l2,
                                                            float* __restrict__ l3, int64_t rows, int h, int w) {
  extern __shared__ float sm[];
  const int h1 = h / 2, w1 = w / 2, h2 = h1 / 2, w2 = w1 / 2, h3 = h2 / 2, w3 = w2 / 2;
  float* s1 = sm; float* s2 = sm + h1 * w1;
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const float* src = l0 + row * ((int64_t)h * w);
    float* d1 = l1 + row * ((int64_t)h1 * w1);
    if ((w & 3) == 0) {   // two pooled values per thread from one float4 per source row (16-byte aligned: w % 4 == 0)
      const int wp = w1 >> 1;
#pragma unroll 4
      for (int i = threadIdx.x; i < h1 * wp; i += blockDim.x) {
        const int y = i / wp, xp = i - y * wp;
        const float* s = src + (int64_t)(2 * y) * w + 4 * xp;
        const float4 a = __ldcs(reinterpret_cast<const float4*>(s)), b = __ldcs(reinterpret_cast<const float4*>(s + w));
        const float2 v = make_float2((a.x + a.y + b.x + b.y) * 0.25f, (a.z + a.w + b.z + b.w) * 0.25f);
        *reinterpret_cast<float2*>(s1 + y * w1 + 2 * xp) = v;
        *reinterpret_cast<float2*>(d1 + y * w1 + 2 * xp) = v;
      }
    } else
    for (int i = threadIdx.x; i < h1 * w1; i += blockDim.x) {
      const int y = i / w1, x = i - y * w1;
      const float* s = src + (int64_t)(2 * y) * w + 2 * x;
      float2 a, b;
      if ((w & 1) == 0) { a = __ldcs(reinterpret_cast<const float2*>(s)); b = __ldcs(reinterpret_cast<const float2*>(s + w)); }
      else { a = make_float2(s[0], s[1]); b = make_float2(s[w], s[w + 1]); }
      const float v = (a.x + a.y + b.x + b.y) * 0.25f;
      s1[i] = v; d1[i] = v;
    }
    __syncthreads();
    float* d2 = l2 + row * ((int64_t)h2 * w2);
    for (int i = threadIdx.x; i < h2 * w2; i += blockDim.x) {
      const int y = i / w2, x = i - y * w2;
      const float* s = s1 + (2 * y) * w1 + 2 * x;
      const float v = (s[0] + s[1] + s[w1] + s[w1 + 1]) * 0.25f;
      s2[i] = v; d2[i] = v;
    }
    __syncthreads();
    float* d3 = l3 + row * ((int64_t)h3 * w3);
    for (int i = threadIdx.x; i < h3 * w3; i += blockDim.x) {
      const int y = i / w3, x = i - y * w3;
      const float* s = s2 + (2 * y) * w2 + 2 * x;
      d3[i] = (s[0] + s[1] + s[w2] + s[w2 + 1]) * 0.25f;
    }
    __syncthreads();   // s1/s2 are rewritten by the next row
  }
}
#endif

// levels 1..3 of the pyramid from level 0 (raft/corr.py:139-142: three successive avg_pool2d(2, 2))
void corr_pool_pyramid(Ctx& cx, const float* l0, float* l1, float* l2, float* l3, int64_t rows, int h, int w) {
#ifndef GV_HOSTSIM
  const int h1 = h / 2, w1 = w / 2, h2 = h1 / 2, w2 = w1 / 2;
  const size_t smem = ((size_t)h1 * w1 + (size_t)h2 * w2) * sizeof(float);
  if (smem <= 96 * 1024 && h2 / 2 > 0 && w2 / 2 > 0) {
    if (cx.dry) return;
    cx.launches++;
    static volatile unsigned char attr[64];
    gv_set_max_smem(corr_pyramid_kernel, 96 * 1024, attr);
    if (cx.prof) cx.prof->begin(cx.stream, "corr_pool_pyramid", (double)rows * h * w);
    const int64_t grid = rows < (int64_t)cx.sm_count * 8 ? rows : (int64_t)cx.sm_count * 8;
    corr_pyramid_kernel<<<(unsigned)grid, 256, smem, cx.stream>>>(l0, l1, l2, l3, rows, h, w);
    gv_check_launch("corr_pool_pyramid");
    if (cx.prof) cx.prof->end(cx.stream);
    return;
  }
#endif
  corr_pool(cx, l0, l1, rows, h, w);
  corr_pool(cx, l1, l2, rows, h / 2, w / 2);
  corr_pool(cx, l2, l3, rows, h / 4, w / 4);
}

// ----------------------------------------------------------------- lookup
// raft/corr.py:144-165.  out channel = lvl*81 + a*9 + b samples level `lvl` of row
// (n, pixel) at (x/2^lvl + (a-4), y/2^lvl + (b-4)) — the "transposed window" of the
// reference's meshgrid(dy, dx) — bilinear, zero padding, align_corners=True, going
// through the same normalise / un-normalise float round trip as bilinear_sampler +
// grid_sample.
struct CorrLookupK {
  CorrPyr pyr; TV coords, out;
  GV_HD void operator()(int64_t i) const {
    const int nch = pyr.nl * 81;
    int ch = (int)(i % nch); int64_t r = i / nch;
    int x = (int)(r % coords.w); r /= coords.w; int y = (int)(r % coords.h); int n = (int)(r / coords.h);
    int lvl = ch / 81, k = ch % 81, a = k / 9, b = k % 9;
    const float* c = coords.p + coords.off(n, y, x);
    float inv = 1.0f / (float)(1 << lvl);
    int H = pyr.h[lvl], W = pyr.w[lvl];
    float px = c[0] * inv + (float)(a - 4);
    float py = c[1] * inv + (float)(b - 4);
    float xg = 2.f * px / (float)(W - 1) - 1.f;
    float yg = 2.f * py / (float)(H - 1) - 1.f;
    float ix = ((xg + 1.f) / 2.f) * (float)(W - 1);
    float iy = ((yg + 1.f) / 2.f) * (float)(H - 1);
    float x0f = floorf(ix), y0f = floorf(iy);
    int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
    float wx1 = ix - x0f, wx0 = (x0f + 1.f) - ix, wy1 = iy - y0f, wy0 = (y0f + 1.f) - iy;
    int64_t row = (int64_t)n * pyr.rows_per_sample + (int64_t)y * coords.w + x;
    const float* img = pyr.lvl[lvl] + row * ((int64_t)H * W);
    float v = 0.f;
    bool vx0 = x0 >= 0 && x0 < W, vx1 = x1 >= 0 && x1 < W, vy0 = y0 >= 0 && y0 < H, vy1 = y1 >= 0 && y1 < H;
    if (vy0 && vx0) v += img[(int64_t)y0 * W + x0] * (wx0 * wy0);
    if (vy0 && vx1) v += img[(int64_t)y0 * W + x1] * (wx1 * wy0);
    if (vy1 && vx0) v += img[(int64_t)y1 * W + x0] * (wx0 * wy1);
    if (vy1 && vx1) v += img[(int64_t)y1 * W + x1] * (wx1 * wy1);
    out.p[out.off(n, y, x) + ch] = v;
  }
};
#ifndef GV_HOSTSIM
// One warp per (source pixel, level): the 9 x-offsets and 9 y-offsets of the window go through the grid_sample round trip ONCE
// (lanes 0..8 / 9..17: 18 coordinate evaluations instead of 2 x 81), the 81 outputs take their (x0, wx) / (y0, wy) by shuffle and issue
// their four taps back to back; a warp's stores are 32 consecutive channels.  Arithmetic and summation order are those of CorrLookupK
// (bit-identical results); the thread-per-output form spent ~100 instructions per output on the coordinate transform.
__global__ void __launch_bounds__(256) corr_lookup_warp_kernel(CorrPyr pyr, TV coords, TV out, int64_t n_items) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t it = warp0; it < n_items; it += nwarps) {
    const int lvl = (int)(it % pyr.nl); int64_t r = it / pyr.nl;
    const int x = (int)(r % coords.w); r /= coords.w; const int y = (int)(r % coords.h); const int n = (int)(r / coords.h);
    const float* c = coords.p + coords.off(n, y, x);
    const float inv = 1.0f / (float)(1 << lvl);
    const int H = pyr.h[lvl], W = pyr.w[lvl];
    // lane l < 9: x offset a = l; lanes 9..17: y offset b = l - 9
    const bool isx = lane < 9;
    const int d = isx ? lane : lane - 9;
    const int S = isx ? W : H;
    const float pc = (isx ? c[0] : c[1]) * inv + (float)(d - 4);
    const float g = 2.f * pc / (float)(S - 1) - 1.f;
    const float ic = ((g + 1.f) / 2.f) * (float)(S - 1);
    const float c0f = floorf(ic);
    const int my0 = (int)c0f;
    const float mw1 = ic - c0f, mw0 = (c0f + 1.f) - ic;
    const int64_t row = (int64_t)n * pyr.rows_per_sample + (int64_t)y * coords.w + x;
    const float* img = pyr.lvl[lvl] + row * ((int64_t)H * W);
    float* o = out.p + out.off(n, y, x) + lvl * 81;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
      const int k = pass * 32 + lane, kk = k < 81 ? k : 80;
      const int a = kk / 9, b = kk - a * 9;
      const int x0 = __shfl_sync(0xffffffffu, my0, a), y0 = __shfl_sync(0xffffffffu, my0, 9 + b);
      const float wx0 = __shfl_sync(0xffffffffu, mw0, a), wx1 = __shfl_sync(0xffffffffu, mw1, a);
      const float wy0 = __shfl_sync(0xffffffffu, mw0, 9 + b), wy1 = __shfl_sync(0xffffffffu, mw1, 9 + b);
      if (k < 81) {
        const int x1 = x0 + 1, y1 = y0 + 1;
        const bool vx0 = x0 >= 0 && x0 < W, vx1 = x1 >= 0 && x1 < W, vy0 = y0 >= 0 && y0 < H, vy1 = y1 >= 0 && y1 < H;
        const float t00 = (vy0 && vx0) ? __ldg(img + (int64_t)y0 * W + x0) : 0.f, t01 = (vy0 && vx1) ? __ldg(img + (int64_t)y0 * W + x1) : 0.f;
        const float t10 = (vy1 && vx0) ? __ldg(img + (int64_t)y1 * W + x0) : 0.f, t11 = (vy1 && vx1) ? __ldg(img + (int64_t)y1 * W + x1) : 0.f;
        float v = 0.f;
        if (vy0 && vx0) v += t00 * (wx0 * wy0);
        if (vy0 && vx1) v += t01 * (wx1 * wy0);
        if (vy1 && vx0) v += t10 * (wx0 * wy1);
        if (vy1 && vx1) v += t11 * (wx1 * wy1);
        o[k] = v;
      }
    }
  }
}
#endif
void corr_lookup(Ctx& cx, const CorrPyr& pyr, const TV& coords, const TV& out) {
#ifndef GV_HOSTSIM
  if (!out.f16 && !coords.f16) {
    if (cx.dry) return;
    cx.launches++;
    const int64_t items = coords.pixels() * pyr.nl;
    if (cx.prof) cx.prof->begin(cx.stream, "corr_lookup", (double)coords.pixels() * pyr.nl * 81);
    int64_t blocks = (items + 7) / 8, cap = (int64_t)cx.sm_count * 16;
    if (blocks > cap) blocks = cap;
    corr_lookup_warp_kernel<<<(unsigned)blocks, 256, 0, cx.stream>>>(pyr, coords, out, items);
    gv_check_launch("corr_lookup");
    if (cx.prof) cx.prof->end(cx.stream);
    return;
  }
#endif
  parallel_for(cx, coords.pixels() * pyr.nl * 81, CorrLookupK{pyr, coords, out}, "corr_lookup");
}

// -------------------------------------------------------- volume-free lookup
// BidirCorrBlock (raft/corr.py:23-93) is looked up ONCE per interpolated frame: building the all-pairs volume pyramid for it writes
// N^2 x 4/3 x 2 values (11 GB at 1088x1920) to read back 2 x N x 324.  Bilinear interpolation of the volume is linear in the volume,
// and a level-l volume entry is the dot product with the 2^l x 2^l average-pooled target feature, so the 81 window samples of a level
// only need the dot products with the (at most) 10 x 10 integer neighbours of the window: 100 dots of 256 channels per (pixel, level)
// = 8.6 GFLOP per pair instead of the 1.45 TFLOP GEMM.  Target features are IEEE half (the same 11-bit significand the plain-TF32
// volume GEMM gave them), the source feature and the accumulation are fp32.
struct ToHalfK {
  TV src; uint16_t* dst;
  GV_HD void operator()(int64_t i) const {
    const int c = (int)(i % src.c); int64_t r = i / src.c;
    const int x = (int)(r % src.w); r /= src.w; const int y = (int)(r % src.h); const int n = (int)(r / src.h);
    dst[i] = gv_f2h(src.p[src.off(n, y, x) + c]);
  }
};
void features_to_half(Ctx& cx, const TV& src, void* dst) {
  parallel_for(cx, src.pixels() * src.c, ToHalfK{src, static_cast<uint16_t*>(dst)}, "features_to_half");
}

// grid_sample(align_corners=True) round trip of one window coordinate (raft/utils/utils.py:66-80), as CorrLookupK evaluates it
GV_HD void corr_axis(float c, float inv, int d, int S, int& i0, float& w0, float& w1) {
  const float pc = c * inv + (float)(d - 4);
  const float g = 2.f * pc / (float)(S - 1) - 1.f;
  const float ic = ((g + 1.f) / 2.f) * (float)(S - 1);
  const float c0f = floorf(ic);
  i0 = (int)c0f; w1 = ic - c0f; w0 = (c0f + 1.f) - ic;
}

// thread-per-output form (host build of the engine; the reference for the warp kernel below): four taps, each a C-channel dot
struct CorrLookupDirectK {
  TV src; CorrFeat tgt; TV coords, out;
  GV_HD float dot(const float* a, int n, int lvl, int ty, int tx) const {
    const uint16_t* b = tgt.lvl[lvl] + ((int64_t)n * tgt.h[lvl] * tgt.w[lvl] + (int64_t)ty * tgt.w[lvl] + tx) * tgt.c;
    float acc = 0.f;
    for (int k = 0; k < tgt.c; ++k) acc = fmaf(a[k], gv_h2f(b[k]), acc);
    return acc * tgt.scale;
  }
  GV_HD void operator()(int64_t i) const {
    const int ch = (int)(i % 324); int64_t r = i / 324;
    const int x = (int)(r % coords.w); r /= coords.w; const int y = (int)(r % coords.h); const int n = (int)(r / coords.h);
    const int lvl = ch / 81, k = ch % 81, a = k / 9, b = k % 9;
    const float* c = coords.p + coords.off(n, y, x);
    const float inv = 1.0f / (float)(1 << lvl);
    const int H = tgt.h[lvl], W = tgt.w[lvl];
    int x0, y0; float wx0, wx1, wy0, wy1;
    corr_axis(c[0], inv, a, W, x0, wx0, wx1);
    corr_axis(c[1], inv, b, H, y0, wy0, wy1);
    const int x1 = x0 + 1, y1 = y0 + 1;
    const float* f = src.p + src.off(n, y, x);
    const bool vx0 = x0 >= 0 && x0 < W, vx1 = x1 >= 0 && x1 < W, vy0 = y0 >= 0 && y0 < H, vy1 = y1 >= 0 && y1 < H;
    float v = 0.f;
    if (vy0 && vx0) v += dot(f, n, lvl, y0, x0) * (wx0 * wy0);
    if (vy0 && vx1) v += dot(f, n, lvl, y0, x1) * (wx1 * wy0);
    if (vy1 && vx0) v += dot(f, n, lvl, y1, x0) * (wx0 * wy1);
    if (vy1 && vx1) v += dot(f, n, lvl, y1, x1) * (wx1 * wy1);
    out.p[out.off(n, y, x) + ch] = v;
  }
};

#ifndef GV_HOSTSIM
// One warp per (source pixel, level); the 8 warps of a CTA take 8 x-adjacent pixels of one level, whose windows overlap by ~90 % and
// stay in L1.  A lane owns 8 of the 256 channels: the source feature sits in registers, every neighbour costs one 16-byte load per
// lane (512 contiguous bytes per warp), and 32 neighbours at a time are reduced across the warp with a transposing butterfly
// (31 shuffles for 32 sums).  The 100 dots land in shared memory; the 81 outputs are then the four-tap blends of CorrLookupK.
constexpr int CLD_WARPS = 8;
__global__ void __launch_bounds__(256) corr_lookup_direct_kernel(TV src, CorrFeat tgt, TV coords, TV out, int xblocks, int64_t n_groups) {
  __shared__ float dots_s[CLD_WARPS][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* dots = dots_s[warp];
  for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    int64_t r = grp;
    const int xb = (int)(r % xblocks); r /= xblocks;
    const int y = (int)(r % coords.h); r /= coords.h;
    const int n = (int)(r % coords.n); const int lvl = (int)(r / coords.n);
    const int x = xb * CLD_WARPS + warp;
    if (x < coords.w) {   // (warp-uniform)
      const float* c = coords.p + coords.off(n, y, x);
      const float inv = 1.0f / (float)(1 << lvl);
      const int H = tgt.h[lvl], W = tgt.w[lvl];
      const bool isx = lane < 9;
      int my0; float mw0, mw1;
      corr_axis(isx ? c[0] : c[1], inv, isx ? lane : (lane < 18 ? lane - 9 : 0), isx ? W : H, my0, mw0, mw1);
      const int bx = __shfl_sync(0xffffffffu, my0, 0), by = __shfl_sync(0xffffffffu, my0, 9);
      float f[8];
      {
        const float4* fp = reinterpret_cast<const float4*>(src.p + src.off(n, y, x) + lane * 8);
        const float4 f0 = __ldg(fp), f1 = __ldg(fp + 1);
        f[0] = f0.x; f[1] = f0.y; f[2] = f0.z; f[3] = f0.w; f[4] = f1.x; f[5] = f1.y; f[6] = f1.z; f[7] = f1.w;
      }
      const uint16_t* tb = tgt.lvl[lvl] + (int64_t)n * H * W * 256 + lane * 8;
#pragma unroll 1
      for (int g = 0; g < 4; ++g) {
        float p[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int nb = g * 32 + j;           // neighbour (gy, gx) of the 10 x 10 grid based at (by, bx)
          const int gy = nb / 10, gx = nb - gy * 10;
          const int ty = by + gy, tx = bx + gx;
          float acc = 0.f;
          if (nb < 100 && ty >= 0 && ty < H && tx >= 0 && tx < W) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(tb + ((int64_t)ty * W + tx) * 256));
            const float2 a0 = __half22float2(*reinterpret_cast<const __half2*>(&q.x)), a1 = __half22float2(*reinterpret_cast<const __half2*>(&q.y));
            const float2 a2 = __half22float2(*reinterpret_cast<const __half2*>(&q.z)), a3 = __half22float2(*reinterpret_cast<const __half2*>(&q.w));
            acc = fmaf(f[0], a0.x, acc); acc = fmaf(f[1], a0.y, acc); acc = fmaf(f[2], a1.x, acc); acc = fmaf(f[3], a1.y, acc);
            acc = fmaf(f[4], a2.x, acc); acc = fmaf(f[5], a2.y, acc); acc = fmaf(f[6], a3.x, acc); acc = fmaf(f[7], a3.y, acc);
          }
          p[j] = acc;
        }
        // transposing reduction: after the step with stride s, a lane keeps the half of the neighbour indices whose bit s equals its own
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) {
          const bool up = (lane & s) != 0;
#pragma unroll
          for (int i = 0; i < s; ++i) {
            const float send = up ? p[i] : p[i + s], keep = up ? p[i + s] : p[i];
            p[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
          }
        }
        dots[g * 32 + lane] = p[0] * tgt.scale;
      }
      __syncwarp();
      float* o = out.p + out.off(n, y, x) + lvl * 81;
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {
        const int k = pass * 32 + lane, kk = k < 81 ? k : 80;
        const int a = kk / 9, b = kk - a * 9;
        int x0 = __shfl_sync(0xffffffffu, my0, a), y0 = __shfl_sync(0xffffffffu, my0, 9 + b);
        float wx0 = __shfl_sync(0xffffffffu, mw0, a), wx1 = __shfl_sync(0xffffffffu, mw1, a);
        float wy0 = __shfl_sync(0xffffffffu, mw0, 9 + b), wy1 = __shfl_sync(0xffffffffu, mw1, 9 + b);
        if (k < 81) {
          // window column a normally starts at bx + a; a floor() that lands one off through rounding moves the (then ~zero-weight)
          // far tap outside the 10-wide grid: fold it onto the grid edge
          int ix = x0 - bx, iy = y0 - by;
          if (ix < 0) { ix = 0; x0 = bx; wx0 = wx1; wx1 = 0.f; } else if (ix > 8) { ix = 8; x0 = bx + 8; wx1 = wx0; wx0 = 0.f; }
          if (iy < 0) { iy = 0; y0 = by; wy0 = wy1; wy1 = 0.f; } else if (iy > 8) { iy = 8; y0 = by + 8; wy1 = wy0; wy0 = 0.f; }
          const int x1 = x0 + 1, y1 = y0 + 1;
          const bool vx0 = x0 >= 0 && x0 < W, vx1 = x1 >= 0 && x1 < W, vy0 = y0 >= 0 && y0 < H, vy1 = y1 >= 0 && y1 < H;
          const float* d = dots + iy * 10 + ix;
          float v = 0.f;
          if (vy0 && vx0) v += d[0] * (wx0 * wy0);
          if (vy0 && vx1) v += d[1] * (wx1 * wy0);
          if (vy1 && vx0) v += d[10] * (wx0 * wy1);
          if (vy1 && vx1) v += d[11] * (wx1 * wy1);
          o[k] = v;
        }
      }
      __syncwarp();
    }
  }
}
#endif

bool corr_lookup_direct_supported(const TV& src, const CorrFeat& tgt, const TV& coords, const TV& out) {
  return !src.f16 && !coords.f16 && !out.f16 && src.c == 256 && tgt.c == 256 && src.ld % 4 == 0 && src.sn % 4 == 0 &&
         (reinterpret_cast<uintptr_t>(src.p) & 15) == 0 && src.h == coords.h && src.w == coords.w && src.n == coords.n && out.c >= 324;
}

void corr_lookup_direct(Ctx& cx, const TV& src, const CorrFeat& tgt, const TV& coords, const TV& out) {
  if (!corr_lookup_direct_supported(src, tgt, coords, out)) throw std::runtime_error("corr_lookup_direct: unsupported tensor layout");
#ifndef GV_HOSTSIM
  if (cx.dry) return;
  cx.launches++;
  for (int l = 0; l < 4; ++l)
    if ((reinterpret_cast<uintptr_t>(tgt.lvl[l]) & 15) != 0) throw std::runtime_error("corr_lookup_direct: target features must be 16-byte aligned");
  const int xblocks = (coords.w + CLD_WARPS - 1) / CLD_WARPS;
  const int64_t groups = (int64_t)4 * coords.n * coords.h * xblocks;
  if (cx.prof) cx.prof->begin(cx.stream, "corr_lookup_direct", (double)coords.pixels() * 4 * 100 * 256 * 2);
  int64_t blocks = groups, cap = (int64_t)cx.sm_count * 8;
  if (blocks > cap) blocks = cap;
  corr_lookup_direct_kernel<<<(unsigned)blocks, 256, 0, cx.stream>>>(src, tgt, coords, out, xblocks, groups);
  gv_check_launch("corr_lookup_direct");
  if (cx.prof) cx.prof->end(cx.stream);
#else
  parallel_for(cx, coords.pixels() * 324, CorrLookupDirectK{src, tgt, coords, out}, "corr_lookup_direct");
#endif
}


}  // namespace gv
