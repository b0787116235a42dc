// Tensor-core implicit-GEMM convolution for sm_100a: TMA tile loads with halo
// coordinates -> 128B-swizzled shared memory -> tcgen05.mma (kind::tf32, fp32
// accumulate in TMEM) -> fused, coalesced epilogue from TMEM.
//
//   GEMM view   M = 128 output pixels (an 8 x 16 spatial tile of one image)
//               N = BN <= 256 output channels per tile (cout tiled if larger)
//               K = taps x cin, walked as (tap, 32-channel block): each K step is ONE
//                   TMA box {32 ch, 16 x, 8 y, 1 n} of the NHWC activation at the
//                   tap-shifted coordinate (out-of-bounds -> zero fill = zero padding,
//                   no im2col, no halo staging code) plus one {32 k, BN} weight box.
//   roles       warp 0: TMA producer | warp 1: MMA issuer (+TMEM alloc) | 4 or 8 epilogue warps
//               | 4 operand-splitter warps (SPLIT only)
//   pipelines   smem full/empty ring, 2 TMEM accumulators (full/empty) so the epilogue of
//               tile i overlaps the MMAs of tile i+1; persistent over tiles.
//   CTA pairs   PAIR: the two CTAs of a cluster issue ONE cta_group::2 MMA of M = 256 (two pixel tiles), each holding
//               half of the weight tile; TMA loads of both signal the leader's barrier, commits are multicast.
//               (Without PAIR, CL = 2 only multicasts the weight tile.)
//   stride 2    TMA element strides {1, 2, 2, 1}: the box spans 32 x 16 pixels, every second one is delivered.
//
// Three operand formats:
//   TF32         (SPLIT=false) the tensor core ignores the low 13 mantissa bits of A; weights are rounded RN at
//                pack time; outputs are stored TF32-rounded so the next layer's truncation is exact.  Post-RAFT layers.
//   FP16 storage (SPLIT=false, Params::f16_in) activations and weights in IEEE half, kind::f16 MMAs with K = 64 per
//                128-byte smem row, fp32 accumulation; half outputs / residuals as flagged per tensor (TV::f16).  The final
//                decoder's 256-channel residual trunk in precision mode 3.
//   3xTF32       (SPLIT=true) D += Ahi*Blo + Alo*Bhi + Ahi*Bhi with Ahi = rn_tf32(a), Alo = rn_tf32(a - Ahi).  The splitter
//                warps read the fp32 A tile from shared memory once and write Ahi / Alo into TENSOR MEMORY (tcgen05.st,
//                64 columns per stage); the MMAs take A from there (tcgen05.mma [d], [a_tmem], b_desc), Bhi / Blo are
//                pre-split at pack time.  The K loop is cut into segments of Params::seg K steps whose accumulators are
//                drained into fp32 registers (the tensor core's accumulate add truncates): ~2^-21 relative error, i.e.
//                fp32-class accuracy at 3 MMAs per K step.  Used for the RAFT recurrence, which amplifies operand rounding.
//                (GIMMVFI_TC_ATMEM=0 keeps the first version, which rewrote Ahi / Alo in shared memory.)
#include "common.h"

#ifndef GV_HOSTSIM
#include <cuda.h>
#include <cuda_fp16.h>

namespace gv {

namespace tc {

constexpr int TILE_H = 8, TILE_W = 16, BM = 128, BK = 32, MAX_STAGES = 8;
constexpr int A_BYTES = BM * BK * 4;                // 16 KB
constexpr int STG_PITCH = 32;                       // floats per staged row; the 16-byte chunk c of row r sits at chunk c ^ (r & 7): conflict-free
                                                    // 128-bit row writes (phase 1) and row-segment reads (phase 2) without padding - the 4.5 KB the
                                                    // padded form cost is what lets the 3xF16 kernel keep 4 pipeline stages at N = 128
constexpr int STG_WARP_BYTES = 32 * STG_PITCH * 4;  // staging area of one epilogue warp (32 rows)
constexpr int BAR_BYTES = 512;

struct Params {
  int taps, kw, ph, pw;
  int stride;                      // 1 or 2: output pixel (y, x) reads input (y * stride + ky - ph, ...) through TMA element strides
  int kblocks, c0_blocks;          // 32-channel K blocks in total / from segment 0
  int tiles_x, tiles_y, n_img, tiles_n;
  int H, W;
  int BN;                          // MMA N (multiple of 16)
  int stages;                      // smem ring depth (<= MAX_STAGES), sized on the host
  int cout;
  int round_out;                   // store TF32-rounded (RN) values
  int spin_limit;                  // mbarrier try_wait attempts before trapping (0 = wait forever)
  int dbg;                         // timing experiments only (GIMMVFI_TC_DEBUG): 1 = splitter idles, 2 = segment drain skips its TMEM loads
  unsigned long long* stall;       // stall profiling (GIMMVFI_TC_STALL_BUF): 16 counters per CTA, else nullptr
  int f16_in;                      // A / B operands are IEEE half (kind::f16, 64-element K blocks); else fp32 read as TF32
  int bk;                          // K elements per 128-byte smem row: 32 (tf32) or 64 (f16)
  int atmem;                       // SPLIT: A_hi / A_lo live in tensor memory (written by the splitter warps), not in smem
  int seg;                         // SPLIT: K steps accumulated in TMEM before promotion to fp32 registers
  int split_f16;                   // SPLIT: the three MMAs run on kind::f16 with fp16 hi / lo operand pairs (K step = 64 elements = two
                                   // 32-channel fp32 A boxes; weights pre-split and pre-scaled at pack time): twice the MMA rate of 3xTF32
  float out_scale;                 // accumulator scale applied before the bias (all-pairs correlation: 1/sqrt(C))
  const float* bias;               // padded to tiles_n * BN
  int act1; const float* slope1;
  int act2; const float* slope2;
  TV res, mul, gru_z, gru_h, out;
  TV out2; int split_c;            // channels >= split_c go to out2 (and only they see `mul`): merged z | r gate convolution
};

#include "tc_ptx.cuh"

#include "tc_epilogue.cuh"

// Residual values of one chunk, fetched one chunk ahead of their use (8 rows x 4 channels per lane, raw bits: 16 B per row
// for fp32, 8 B for half).  A lane whose 4 channels are ragged or misaligned reports false and loads in place later.
__device__ __forceinline__ bool res_prefetch(const Params& p, int cbase, int quarter, int lane, int tx, int ty, int n, uint4* rp) {
  const int q8 = lane & 7, rsub = lane >> 3;
  const int c = cbase + q8 * 4;
  if (!p.res.p || c + 3 >= p.cout || n >= p.n_img || (p.res.ld & 3)) return false;
  const int y0 = ty * TILE_H + quarter * 2, x0 = tx * TILE_W + rsub;
  const int64_t base = p.res.off(n, y0, x0) + c;
  const char* q = reinterpret_cast<const char*>(p.res.p) + base * (p.res.f16 ? 2 : 4);
  if (reinterpret_cast<uintptr_t>(q) & (p.res.f16 ? 7 : 15)) return false;
  const int64_t rowb = (int64_t)p.res.w * p.res.ld * (p.res.f16 ? 2 : 4), xb = (int64_t)4 * p.res.ld * (p.res.f16 ? 2 : 4);
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int yy = it >> 2, xx = (it & 3) * 4;
    rp[it] = make_uint4(0u, 0u, 0u, 0u);
    if (y0 + yy < p.H && x0 + xx < p.W) {
      const char* a = q + yy * rowb + (it & 3) * xb;
      if (p.res.f16) { const uint2 t = *reinterpret_cast<const uint2*>(a); rp[it].x = t.x; rp[it].y = t.y; }
      else rp[it] = *reinterpret_cast<const uint4*>(a);
    }
  }
  return true;
}
__device__ __forceinline__ void res_unpack(const Params& p, const uint4& r, float* t) {
  if (p.res.f16) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    t[0] = a.x; t[1] = a.y; t[2] = b.x; t[3] = b.y;
  } else {
    t[0] = __uint_as_float(r.x); t[1] = __uint_as_float(r.y); t[2] = __uint_as_float(r.z); t[3] = __uint_as_float(r.w);
  }
}
// The SepConvGRU gate epilogue as its OWN kernel instantiation (template parameter EPI = 1 of conv2d_tc_kernel), chosen on the host
// when a 3xF16 layer on CTA pairs has side tensors and the layout below holds for the whole layer:
//   act1 = none; fp32 side tensors / outputs with 16-byte aligned 4-channel groups (cout % 4 == 0, strides % 4 == 0); act2 in {none,
//   sigmoid, tanh}; then  v = acc + bias (+ residual) -> act2 -> (x gate multiplier | GRU blend with z, h).
// Why a separate instantiation: in the generic kernel this path is 2.6 K executed instructions per 32 x 32 chunk inside a 15 K
// instruction kernel (L1.5 I-cache 32 KB) - ncu: 20 K clk per chunk, 44 % of the stall samples "no instruction" - and the split
// kernel cannot hide it (its drain warps ARE the epilogue warps).  Moving it out of line or batching its loads inline cost the hot
// loops registers (measured); here the lean / general / prefetched-residual paths do not exist, so the straight-line
// form - one base pointer and two strides per tensor, every load of a 4-row group issued before the first use - fits the 128
// registers a thread of the 14-warp CTA gets.
__device__ __forceinline__ void epi_chunk_gate(const Params& p, const uint32_t* v, uint32_t stg_s, int cbase, int quarter, int lane, int tx, int ty, int n) {
  const int q8 = lane & 7, rsub = lane >> 3;
  {   // phase 1 (thread = pixel row): accumulator x scale + bias, staged (swizzled) for the row-major phase 2
    float o[32];
    const float4* b4p = reinterpret_cast<const float4*>(p.bias + cbase);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b4 = __ldg(b4p + j);
      o[4 * j] = fmaf(__uint_as_float(v[4 * j]), p.out_scale, b4.x); o[4 * j + 1] = fmaf(__uint_as_float(v[4 * j + 1]), p.out_scale, b4.y);
      o[4 * j + 2] = fmaf(__uint_as_float(v[4 * j + 2]), p.out_scale, b4.z); o[4 * j + 3] = fmaf(__uint_as_float(v[4 * j + 3]), p.out_scale, b4.w);
    }
    const uint32_t srow = stg_s + (uint32_t)(lane * STG_PITCH * 4);
#pragma unroll
    for (int j = 0; j < 8; ++j) sts128(srow + (uint32_t)((j ^ (lane & 7)) << 4), o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
  }
  __syncwarp();
  const int c = cbase + q8 * 4;
  if (c < p.cout && n < p.n_img) {
    const int y0 = ty * TILE_H + quarter * 2, x0 = tx * TILE_W + rsub;
    const uint32_t sb_even = stg_s + (uint32_t)(rsub * STG_PITCH * 4 + ((q8 ^ rsub) << 4));
    const uint32_t sb_odd = stg_s + (uint32_t)(rsub * STG_PITCH * 4 + ((q8 ^ (rsub + 4)) << 4));
    const bool second = p.split_c > 0 && cbase >= p.split_c;          // merged z | r convolution: this chunk belongs to out2
    const int c_loc = c - (second ? p.split_c : 0);
    const bool use_mul = p.mul.p && (p.split_c == 0 || second);
    float* oq = (second ? p.out2.p : p.out.p) + (second ? p.out2.off(n, y0, x0) : p.out.off(n, y0, x0)) + c_loc;
    const int64_t o_x = (int64_t)4 * (second ? p.out2.ld : p.out.ld), o_row = (int64_t)p.W * (second ? p.out2.ld : p.out.ld);
    const float* rq = p.res.p ? p.res.p + p.res.off(n, y0, x0) + c : nullptr;
    const float* bq = use_mul ? p.mul.p + p.mul.off(n, y0, x0) + c_loc : (p.gru_z.p ? p.gru_z.p + p.gru_z.off(n, y0, x0) + c : nullptr);
    const float* hq = p.gru_z.p ? p.gru_h.p + p.gru_h.off(n, y0, x0) + c : nullptr;
    const int64_t r_x = (int64_t)4 * p.res.ld, r_row = (int64_t)p.res.w * p.res.ld;
    const int64_t b_x = (int64_t)4 * (use_mul ? p.mul.ld : p.gru_z.ld), b_row = use_mul ? (int64_t)p.mul.w * p.mul.ld : (int64_t)p.gru_z.w * p.gru_z.ld;
    const int64_t h_x = (int64_t)4 * p.gru_h.ld, h_row = (int64_t)p.gru_h.w * p.gru_h.ld;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      if (y0 + g < p.H) {
        bool ok[4];
        float4 r4[4], b4[4], h4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ok[j] = x0 + j * 4 < p.W;
          r4[j] = (rq && ok[j]) ? *reinterpret_cast<const float4*>(rq + g * r_row + j * r_x) : zero4;
          b4[j] = (bq && ok[j]) ? *reinterpret_cast<const float4*>(bq + g * b_row + j * b_x) : zero4;
          h4[j] = (hq && ok[j]) ? *reinterpret_cast<const float4*>(hq + g * h_row + j * h_x) : zero4;
        }
        float o[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 sv = lds128(((j & 1) ? sb_odd : sb_even) + (uint32_t)((g * 4 + j) * 4 * STG_PITCH * 4));
          o[4 * j] = sv.x + r4[j].x; o[4 * j + 1] = sv.y + r4[j].y; o[4 * j + 2] = sv.z + r4[j].z; o[4 * j + 3] = sv.w + r4[j].w;
        }
        if (p.act2 == ACT_SIGMOID) {
#pragma unroll
          for (int u = 0; u < 16; ++u) o[u] = __fdividef(1.f, 1.f + exp2f(-1.4426950408889634f * o[u]));
        } else if (p.act2 == ACT_TANH) {
#pragma unroll
          for (int u = 0; u < 16; ++u) o[u] = 1.f - __fdividef(2.f, 1.f + exp2f(2.8853900817779268f * o[u]));
        }
        if (hq) {          // h = (1 - z) * h + z * q   (raft/update.py:58,66)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            o[4 * j] = (1.f - b4[j].x) * h4[j].x + b4[j].x * o[4 * j]; o[4 * j + 1] = (1.f - b4[j].y) * h4[j].y + b4[j].y * o[4 * j + 1];
            o[4 * j + 2] = (1.f - b4[j].z) * h4[j].z + b4[j].z * o[4 * j + 2]; o[4 * j + 3] = (1.f - b4[j].w) * h4[j].w + b4[j].w * o[4 * j + 3];
          }
        } else if (bq) {   // gate multiply (r * h)
#pragma unroll
          for (int j = 0; j < 4; ++j) { o[4 * j] *= b4[j].x; o[4 * j + 1] *= b4[j].y; o[4 * j + 2] *= b4[j].z; o[4 * j + 3] *= b4[j].w; }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (ok[j]) *reinterpret_cast<float4*>(oq + g * o_row + j * o_x) = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
      }
    }
  }
  __syncwarp();
}

// EPI = 2: the plain epilogue of the 3xF16 layers (bias, none / ReLU, fp32 output, no side tensors), same two-phase shape as epi_chunk
// (staged transpose, 128-byte row stores) but nothing else compiled in: a 5 K instead of 15 K instruction kernel for the layers whose
// epilogue code is otherwise fetched cold (ncu: 40 % of the plain epilogue's stall samples are instruction fetch).
// (A direct-store variant - thread = pixel, 128 contiguous bytes per thread, no staging - was measured and is slower here: 96.97 vs
// 93.52 ms per frame on one box.)
__device__ __forceinline__ void epi_chunk_plain(const Params& p, const uint32_t* v, uint32_t stg_s, int cbase, int quarter, int lane, int tx, int ty, int n) {
  const int q8 = lane & 7, rsub = lane >> 3;
  {
    float o[32];
    const float4* b4p = reinterpret_cast<const float4*>(p.bias + cbase);
    const bool relu = p.act1 == ACT_RELU;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b4 = __ldg(b4p + j);
      o[4 * j] = fmaf(__uint_as_float(v[4 * j]), p.out_scale, b4.x); o[4 * j + 1] = fmaf(__uint_as_float(v[4 * j + 1]), p.out_scale, b4.y);
      o[4 * j + 2] = fmaf(__uint_as_float(v[4 * j + 2]), p.out_scale, b4.z); o[4 * j + 3] = fmaf(__uint_as_float(v[4 * j + 3]), p.out_scale, b4.w);
    }
    if (relu) {
#pragma unroll
      for (int u = 0; u < 32; ++u) o[u] = fmaxf(o[u], 0.f);
    }
    const uint32_t srow = stg_s + (uint32_t)(lane * STG_PITCH * 4);
#pragma unroll
    for (int j = 0; j < 8; ++j) sts128(srow + (uint32_t)((j ^ (lane & 7)) << 4), o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
  }
  __syncwarp();
  const int c = cbase + q8 * 4;
  if (c < p.cout && n < p.n_img) {
    const int y0 = ty * TILE_H + quarter * 2, x0 = tx * TILE_W + rsub;
    const uint32_t sb_even = stg_s + (uint32_t)(rsub * STG_PITCH * 4 + ((q8 ^ rsub) << 4));
    const uint32_t sb_odd = stg_s + (uint32_t)(rsub * STG_PITCH * 4 + ((q8 ^ (rsub + 4)) << 4));
    float* oq = p.out.p + p.out.off(n, y0, x0) + c;
    const int64_t o_x = (int64_t)4 * p.out.ld, o_row = (int64_t)p.W * p.out.ld;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int yy = it >> 2, xx = (it & 3) * 4;
      const float4 sv = lds128(((it & 1) ? sb_odd : sb_even) + (uint32_t)(it * 4 * STG_PITCH * 4));
      if (y0 + yy < p.H && x0 + xx < p.W) *reinterpret_cast<float4*>(oq + yy * o_row + (it & 3) * o_x) = sv;
    }
  }
  __syncwarp();
}

// One 32-row x 32-column chunk of the output tile: v[j] = accumulator of (this thread's pixel row, column c0 + j).
// stg_s = shared-window address of this warp's 32 x STG_PITCH staging area; cbase = first output channel of the chunk.
__device__ __forceinline__ void epi_chunk(const Params& p, const uint32_t* v, uint32_t stg_s, int cbase, int quarter, int lane, int tx, int ty,
                                          int n, long long* tprof /* stall profiling: [0] phase 1, [1] phase 2, [2] chunks */,
                                          const uint4* rp = nullptr /* res_prefetch() of this chunk */, bool rp_valid = false) {
  const int q8 = lane & 7, rsub = lane >> 3;
  const long long tp0 = tprof ? clock64() : 0;
  // phase 1 (thread = pixel row): bias + act1, stage to smem.  bias is padded to tiles_n*BN (+slack) on the host.
  {
    float o[32];
    const float4* b4p = reinterpret_cast<const float4*>(p.bias + cbase);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b4 = __ldg(b4p + j);
      o[4 * j] = fmaf(__uint_as_float(v[4 * j]), p.out_scale, b4.x); o[4 * j + 1] = fmaf(__uint_as_float(v[4 * j + 1]), p.out_scale, b4.y);
      o[4 * j + 2] = fmaf(__uint_as_float(v[4 * j + 2]), p.out_scale, b4.z); o[4 * j + 3] = fmaf(__uint_as_float(v[4 * j + 3]), p.out_scale, b4.w);
    }
    act_n<32>(o, p.act1, p.slope1, cbase, p.cout);
    const uint32_t srow = stg_s + (uint32_t)(lane * STG_PITCH * 4);
#pragma unroll
    for (int j = 0; j < 8; ++j) sts128(srow + (uint32_t)((j ^ (lane & 7)) << 4), o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
  }
  __syncwarp();
  const long long tp1 = tprof ? clock64() : 0;
  // phase 2 (8 lanes = one 128-byte row, 4 rows per instruction): + residual, act2, gate multiply,
  // ConvGRU blend, optional TF32 rounding, coalesced store.  Row `it` of this lane is tile pixel
  // (yy, xx) = (quarter * 2 + it / 4, (it % 4) * 4 + rsub).
  const int c = cbase + q8 * 4;
  if (c < p.cout && n < p.n_img) {
    const bool full4 = c + 3 < p.cout;
    const int y0 = ty * TILE_H + quarter * 2, x0 = tx * TILE_W + rsub;
    const bool interior = y0 + 1 < p.H && tx * TILE_W + TILE_W <= p.W;
    // staged row it * 4 + rsub, logical chunk q8: its swizzle phase (row & 7) is rsub for even `it`, rsub + 4 for odd
    const uint32_t sb_even = stg_s + (uint32_t)(rsub * STG_PITCH * 4 + ((q8 ^ rsub) << 4));
    const uint32_t sb_odd = stg_s + (uint32_t)(rsub * STG_PITCH * 4 + ((q8 ^ (rsub + 4)) << 4));
    const bool second = p.split_c > 0 && cbase >= p.split_c;          // merged z | r convolution: this chunk belongs to out2
    const TV& O = second ? p.out2 : p.out;
    const int c_loc = c - (second ? p.split_c : 0);
    const bool use_mul = p.mul.p && (p.split_c == 0 || second);
    const bool lean = !p.res.p && !use_mul && !p.gru_z.p;
    const int64_t opix0 = (int64_t)y0 * p.W + x0;
    const int64_t obase = (int64_t)n * O.sn + opix0 * O.ld + c_loc;    // element offset of (row 0 of this lane, its first channel)
    const int64_t o_row = (int64_t)p.W * O.ld, o_x = (int64_t)4 * O.ld;
    const uintptr_t oaddr = reinterpret_cast<uintptr_t>(O.p) + (uintptr_t)obase * (O.f16 ? 2 : 4);
    const bool ovec = full4 && ((oaddr & (O.f16 ? 7 : 15)) == 0) && (O.ld % 4 == 0);
    if (lean) {
      float o[32];
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const float4 sv = lds128(((it & 1) ? sb_odd : sb_even) + (uint32_t)(it * 4 * STG_PITCH * 4));
        o[4 * it] = sv.x; o[4 * it + 1] = sv.y; o[4 * it + 2] = sv.z; o[4 * it + 3] = sv.w;
      }
      if (p.act2 != ACT_NONE) act_rows<8>(o, p.act2, p.slope2, c, p.cout);
      if (p.round_out) {
        // store TF32-representable values (round-to-nearest-even): the next TF32 layer then truncates
        // nothing, i.e. its operands are RN- instead of toward-zero-rounded (unbiased)
#pragma unroll
        for (int u = 0; u < 32; ++u) o[u] = rn_tf32(o[u]);
      }
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int yy = it >> 2, xx = (it & 3) * 4;
        if (!interior && (y0 + yy >= p.H || x0 + xx >= p.W)) continue;
        store4(O, obase + yy * o_row + (it & 3) * o_x, o + 4 * it, c, p.cout, ovec);
      }
    } else {
      // side tensors (residual / gate / GRU state): two groups of 4 rows, every global load of a group issued before its
      // first use (one warp's epilogue is latency bound: the 2-row loop this replaces took 2x the main loop at N = 256)
      const int64_t rbase = p.res.p ? p.res.off(n, y0, x0) + c : 0, mbase = use_mul ? p.mul.off(n, y0, x0) + c_loc : 0;
      const int64_t zbase = p.gru_z.p ? p.gru_z.off(n, y0, x0) + c : 0, hbase = p.gru_z.p ? p.gru_h.off(n, y0, x0) + c : 0;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        float o[16];
        bool ok[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ok[j] = interior || (y0 + g < p.H && x0 + j * 4 < p.W);
          const float4 sv = lds128(((j & 1) ? sb_odd : sb_even) + (uint32_t)((g * 4 + j) * 4 * STG_PITCH * 4));
          o[4 * j] = sv.x; o[4 * j + 1] = sv.y; o[4 * j + 2] = sv.z; o[4 * j + 3] = sv.w;
        }
        if (p.res.p) {
          float t[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (rp_valid) res_unpack(p, rp[g * 4 + j], t + 4 * j);
            else if (ok[j]) load4_any(p.res, rbase + (int64_t)g * p.res.w * p.res.ld + (int64_t)j * 4 * p.res.ld, c, p.cout, t + 4 * j);
            else { t[4 * j] = t[4 * j + 1] = t[4 * j + 2] = t[4 * j + 3] = 0.f; }
          }
#pragma unroll
          for (int u = 0; u < 16; ++u) o[u] += t[u];
        }
        if (p.act2 != ACT_NONE) act_rows<4>(o, p.act2, p.slope2, c, p.cout);
        if (use_mul) {
          float t[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (ok[j]) load4(p.mul.p + mbase + (int64_t)g * p.mul.w * p.mul.ld + (int64_t)j * 4 * p.mul.ld, c, p.cout, t + 4 * j);
            else { t[4 * j] = t[4 * j + 1] = t[4 * j + 2] = t[4 * j + 3] = 0.f; }
          }
#pragma unroll
          for (int u = 0; u < 16; ++u) o[u] *= t[u];
        }
        if (p.gru_z.p) {  // h = (1 - z) * h + z * q   (raft/update.py:58,66)
          float z[16], hh[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (ok[j]) {
              load4(p.gru_z.p + zbase + (int64_t)g * p.gru_z.w * p.gru_z.ld + (int64_t)j * 4 * p.gru_z.ld, c, p.cout, z + 4 * j);
              load4(p.gru_h.p + hbase + (int64_t)g * p.gru_h.w * p.gru_h.ld + (int64_t)j * 4 * p.gru_h.ld, c, p.cout, hh + 4 * j);
            } else {
#pragma unroll
              for (int u = 0; u < 4; ++u) { z[4 * j + u] = 0.f; hh[4 * j + u] = 0.f; }
            }
          }
#pragma unroll
          for (int u = 0; u < 16; ++u) o[u] = (1.f - z[u]) * hh[u] + z[u] * o[u];
        }
        if (p.round_out) {
#pragma unroll
          for (int u = 0; u < 16; ++u) o[u] = rn_tf32(o[u]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (ok[j]) store4(O, obase + g * o_row + j * o_x, o + 4 * j, c, p.cout, ovec);
      }
    }
  }
  __syncwarp();
  if (tprof) { const long long tp2 = clock64(); tprof[0] += tp1 - tp0; tprof[1] += tp2 - tp1; tprof[2] += 1; }
}

// CL = thread-block-cluster size (1 or 2).  CL == 2: the two CTAs of a cluster work on neighbouring pixel tiles
// with the same weights; each loads HALF of every weight box and multicasts it to both, halving the weight
// traffic out of L2 (the limiter at N = 256).  A stage may be refilled only when BOTH CTAs' MMAs retired it,
// so the MMA commit is multicast to both CTAs' empty barriers (count 2).
// EW = epilogue warps (4 or 8).  With K-poor layers (1x1 convs, correlation, 32/64 channels) the epilogue, not the
// MMA, is the critical path and one warp per SM sub-partition is instruction-latency bound: EW == 8 puts two warps
// on every TMEM lane quarter, alternating over the 32-column chunks.
// PAIR (CL == 2, !SPLIT): the two CTAs of the cluster issue ONE tcgen05.mma.cta_group::2 of M = 256 (rows 0-127 = the
// leader's pixel tile, 128-255 = the peer's); each CTA holds its own A tile and HALF of the weight tile, so every SM's
// shared memory carries 2/3 of the bytes per MMA of the single-CTA form (whose limiter it is: 128 B/clk).  Only the
// leader (cluster rank 0) issues MMAs; both CTAs' TMA loads signal the leader's full barrier, the MMA commits are
// multicast to both CTAs' empty / accumulator-full barriers, and the peer's epilogue releases accumulators remotely.
template <bool SPLIT, int CL, int EW, bool PAIR = false, int EPI = 0>
__global__ void __launch_bounds__(64 + 32 * EW + (SPLIT ? 128 : 0), 1)
conv2d_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: per stage [A 16 KB | (A_lo 16 KB) | B BN*128 B | (B_lo)], then the epilogue staging area, then barriers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_bytes = (PAIR ? p.BN / 2 : p.BN) * BK * 4;   // PAIR: this CTA's half of the weight rows
  const bool ATM = SPLIT && p.atmem;   // TMEM columns: accumulators at 0 / 128, A ring (64 columns per stage) from 256
  const bool SF16 = SPLIT && p.split_f16;   // (implies ATM)
  const int a_all = (SPLIT && (!ATM || SF16)) ? 2 * A_BYTES : A_BYTES;   // SF16: two 32-channel fp32 sub-tiles per 64-element K step
  const uint32_t acc_stride = ATM ? 128u : 256u;
  const int stage_bytes = a_all + (SPLIT ? 2 * b_bytes : b_bytes);
  const int STAGES = p.stages;
  float* stg_base = reinterpret_cast<float*>(smem + STAGES * stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes + EW * STG_WARP_BYTES);
  uint64_t* full_bar = bars;                          // [MAX_STAGES]  TMA bytes landed
  uint64_t* empty_bar = bars + MAX_STAGES;            // [MAX_STAGES]  MMAs reading the stage retired
  uint64_t* xf_bar = bars + 2 * MAX_STAGES;           // [MAX_STAGES]  (SPLIT) A_lo written
  uint64_t* tfull_bar = bars + 3 * MAX_STAGES;        // [2]
  uint64_t* tempty_bar = bars + 3 * MAX_STAGES + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * MAX_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile index -> (pixel tile, N tile): consecutive indices 2k, 2k+1 (one cluster) share the N tile (same weights)
  // and take neighbouring pixel tiles; the pixel-tile count is padded to CL so both CTAs run the same K-step count.
  const int pix_tiles = (p.n_img * p.tiles_y * p.tiles_x + CL - 1) / CL * CL;
  const int num_tiles = pix_tiles * p.tiles_n;
  const int ksteps = p.taps * p.kblocks;
  const int SPIN = p.spin_limit;
  long long st_a = 0, st_b = 0;           // stall-profile accumulators of this thread's role
  long long st_c[3] = {0, 0, 0};
  long long* const ST_A = p.stall ? &st_a : nullptr;
  long long* const ST_B = p.stall ? &st_b : nullptr;
  const long long t_start = p.stall ? clock64() : 0;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA0)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA1)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], (PAIR && !SPLIT) ? 2 : 1); mbar_init(&empty_bar[s], PAIR ? 1 : CL); mbar_init(&xf_bar[s], (PAIR && SPLIT) ? 10 : 4); }
      for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], PAIR ? 2 * EW : EW); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    // 512 columns: two 256-column fp32 accumulators (1 CTA / SM, so the whole TMEM is ours)
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CL > 1) cluster_sync_all(); else __syncthreads();   // barriers initialised cluster-wide before any remote arrive
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  uint32_t cta_rank = 0;
  if (CL > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));

  if (warp == 0) {
    // ===================================================== TMA producer
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int nt = (tile / CL) % p.tiles_n; int r = (tile / CL) / p.tiles_n * CL + tile % CL;
        const int tx = r % p.tiles_x; r /= p.tiles_x; const int ty = r % p.tiles_y; const int n = r / p.tiles_y;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int ky = tap / p.kw, kx = tap % p.kw;
          const int x0 = tx * TILE_W * p.stride + kx - p.pw, y0 = ty * TILE_H * p.stride + ky - p.ph;
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait_t(&empty_bar[stage], phase ^ 1, SPIN, ST_A);
            uint8_t* a_dst = smem + stage * stage_bytes;
            uint8_t* b_dst = a_dst + a_all;
            if (PAIR && !SPLIT) {   // both CTAs' bytes are counted by the leader's barrier (2 arrivals + 2 x (A + B/2) bytes)
              const uint32_t lead_full = mapa_shared(smem_u32(&full_bar[stage]), 0u);
              mbar_expect_tx_cluster(lead_full, (uint32_t)(A_BYTES + b_bytes));
              if (kb < p.c0_blocks) tma_load_4d_2sm(a_dst, &tmA0, lead_full, kb * p.bk, x0, y0, n);
              else tma_load_4d_2sm(a_dst, &tmA1, lead_full, (kb - p.c0_blocks) * p.bk, x0, y0, n);
              tma_load_3d_2sm(b_dst, &tmB, lead_full, kb * p.bk, nt * p.BN + (int)cta_rank * (p.BN / 2), tap);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
              continue;
            }
            if (PAIR && SPLIT) {
              // A feeds this CTA's own splitter warps (local barrier); the two weight-plane halves are operands of the leader's
              // MMAs: their bytes are counted by the LEADER's xf barrier, next to the 8 splitter-warp arrivals of both CTAs
              mbar_expect_tx(&full_bar[stage], (uint32_t)(SF16 ? 2 * A_BYTES : A_BYTES));
              if (kb < p.c0_blocks) tma_load_4d(a_dst, &tmA0, &full_bar[stage], kb * p.bk, x0, y0, n);
              else tma_load_4d(a_dst, &tmA1, &full_bar[stage], (kb - p.c0_blocks) * p.bk, x0, y0, n);
              if (SF16) {   // channels [32, 64) of the K block (beyond the tensor: zero fill)
                if (kb < p.c0_blocks) tma_load_4d(a_dst + A_BYTES, &tmA0, &full_bar[stage], kb * p.bk + 32, x0, y0, n);
                else tma_load_4d(a_dst + A_BYTES, &tmA1, &full_bar[stage], (kb - p.c0_blocks) * p.bk + 32, x0, y0, n);
              }
              const uint32_t lead_xf = mapa_shared(smem_u32(&xf_bar[stage]), 0u);
              mbar_expect_tx_cluster(lead_xf, (uint32_t)(2 * b_bytes));
              const int row0 = nt * p.BN + (int)cta_rank * (p.BN / 2);
              tma_load_3d_2sm(b_dst, &tmB, lead_xf, kb * p.bk, row0, tap);
              tma_load_3d_2sm(b_dst + b_bytes, &tmB, lead_xf, kb * p.bk, row0, tap + p.taps);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
              continue;
            }
            mbar_expect_tx(&full_bar[stage], (uint32_t)((SF16 ? 2 * A_BYTES : A_BYTES) + (SPLIT ? 2 * b_bytes : b_bytes)));
            if (kb < p.c0_blocks) tma_load_4d(a_dst, &tmA0, &full_bar[stage], kb * p.bk, x0, y0, n);
            else tma_load_4d(a_dst, &tmA1, &full_bar[stage], (kb - p.c0_blocks) * p.bk, x0, y0, n);
            if (SF16) {
              if (kb < p.c0_blocks) tma_load_4d(a_dst + A_BYTES, &tmA0, &full_bar[stage], kb * p.bk + 32, x0, y0, n);
              else tma_load_4d(a_dst + A_BYTES, &tmA1, &full_bar[stage], (kb - p.c0_blocks) * p.bk + 32, x0, y0, n);
            }
            if (CL == 1) {
              tma_load_3d(b_dst, &tmB, &full_bar[stage], kb * p.bk, nt * p.BN, tap);
              if (SPLIT) tma_load_3d(b_dst + b_bytes, &tmB, &full_bar[stage], kb * p.bk, nt * p.BN, tap + p.taps);
            } else {  // tmB's box is BN/CL rows: my slice of the weight tile, delivered to every CTA of the cluster
              const int rows = p.BN / CL, roff = (int)cta_rank * rows;
              tma_load_3d_mc(b_dst + roff * 128, &tmB, &full_bar[stage], kb * p.bk, nt * p.BN + roff, tap, (uint16_t)((1u << CL) - 1));
              if (SPLIT)
                tma_load_3d_mc(b_dst + b_bytes + roff * 128, &tmB, &full_bar[stage], kb * p.bk, nt * p.BN + roff, tap + p.taps, (uint16_t)((1u << CL) - 1));
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
      if (p.stall) p.stall[(size_t)blockIdx.x * 16] = (unsigned long long)st_a;
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    // instruction descriptor: D=f32, A=B=tf32, both K-major, N>>3, M>>4
    // (kind::f16: a/b format 0 = F16, K = 16 per instruction = the same 32 bytes of every smem row)
    const uint32_t fmt = (p.f16_in || (SPLIT && p.split_f16)) ? 0u : 2u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.BN >> 3) << 17) |
                           ((uint32_t)((PAIR ? 2 * BM : BM) >> 4) << 24);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < ((PAIR && cta_rank != 0) ? 0 : num_tiles); tile += gridDim.x) {
      // plain: one TMEM accumulator per tile.  SPLIT: the K loop is cut into segments; every segment
      // (p.seg K steps) starts a fresh accumulator (alternating buffers) that the epilogue warps drain into fp32 registers.
      // The tensor core's accumulator add truncates; short chains + a true fp32 sum across segments keep the
      // result at CUDA-core fp32 accuracy (measured: whole-K chains were ~10x worse).
      for (int ks = 0; ks < ksteps; ++ks) {
        const bool seg_start = SPLIT ? (ks % p.seg == 0) : (ks == 0);
        const bool seg_end = SPLIT ? (ks % p.seg == p.seg - 1 || ks == ksteps - 1) : (ks == ksteps - 1);
        if (seg_start) {
          mbar_wait_t(&tempty_bar[acc], acc_phase ^ 1, SPIN, ST_B);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_stride;
        mbar_wait_t(SPLIT ? &xf_bar[stage] : &full_bar[stage], phase, SPIN, ST_A);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
          const uint64_t adesc = make_smem_desc(a_addr);
          const uint64_t bdesc = make_smem_desc(a_addr + a_all);
          const uint64_t alo = make_smem_desc(a_addr + A_BYTES);
          const uint64_t blo = make_smem_desc(a_addr + a_all + b_bytes);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {  // UMMA_K = 8 tf32 = 32 B -> +2 in the (addr >> 4) field
            const uint64_t ko = (uint64_t)(k * 2);
            if (ATM && PAIR) {
              const uint32_t ahi_t = tmem_base + 256u + (uint32_t)(stage * 64 + k * 8), alo_t = ahi_t + 32u;
              if (SF16) {   // 8 TMEM columns = 16 packed halves = one K = 16 instruction
                mma_f16_ts_2sm(d_tmem, ahi_t, blo + ko, idesc, (!seg_start || k > 0) ? 1u : 0u);
                mma_f16_ts_2sm(d_tmem, alo_t, bdesc + ko, idesc, 1u);
                mma_f16_ts_2sm(d_tmem, ahi_t, bdesc + ko, idesc, 1u);
              } else {
                mma_tf32_ts_2sm(d_tmem, ahi_t, blo + ko, idesc, (!seg_start || k > 0) ? 1u : 0u);
                mma_tf32_ts_2sm(d_tmem, alo_t, bdesc + ko, idesc, 1u);
                mma_tf32_ts_2sm(d_tmem, ahi_t, bdesc + ko, idesc, 1u);
              }
            } else if (ATM) {
              const uint32_t ahi_t = tmem_base + 256u + (uint32_t)(stage * 64 + k * 8), alo_t = ahi_t + 32u;
              if (SF16) {
                mma_f16_ts(d_tmem, ahi_t, blo + ko, idesc, (!seg_start || k > 0) ? 1u : 0u);
                mma_f16_ts(d_tmem, alo_t, bdesc + ko, idesc, 1u);
                mma_f16_ts(d_tmem, ahi_t, bdesc + ko, idesc, 1u);
              } else {
                mma_tf32_ts(d_tmem, ahi_t, blo + ko, idesc, (!seg_start || k > 0) ? 1u : 0u);     // A_hi * B_lo
                mma_tf32_ts(d_tmem, alo_t, bdesc + ko, idesc, 1u);                                 // A_lo * B_hi
                mma_tf32_ts(d_tmem, ahi_t, bdesc + ko, idesc, 1u);                                 // A_hi * B_hi
              }
            } else if (SPLIT) {   // small terms first
              mma_tf32(d_tmem, adesc + ko, blo + ko, idesc, (!seg_start || k > 0) ? 1u : 0u);   // A_hi * B_lo
              mma_tf32(d_tmem, alo + ko, bdesc + ko, idesc, 1u);                                 // A_lo * B_hi
              mma_tf32(d_tmem, adesc + ko, bdesc + ko, idesc, 1u);                               // A_hi * B_hi
            } else if (PAIR) {
              mma_2sm(d_tmem, adesc + ko, bdesc + ko, idesc, (!seg_start || k > 0) ? 1u : 0u, p.f16_in != 0);
            } else if (p.f16_in) {
              mma_f16(d_tmem, adesc + ko, bdesc + ko, idesc, (!seg_start || k > 0) ? 1u : 0u);
            } else {
              mma_tf32(d_tmem, adesc + ko, bdesc + ko, idesc, (!seg_start || k > 0) ? 1u : 0u);
            }
          }
          if (PAIR) {
            mma_commit_2sm(&empty_bar[stage], (uint16_t)3);                      // both CTAs' producers
            if (seg_end) mma_commit_2sm(&tfull_bar[acc], (uint16_t)3);           // both CTAs' epilogues
          } else {
            if (CL == 1) mma_commit(&empty_bar[stage]);    // frees the smem slot when these MMAs retire
            else mma_commit_mc(&empty_bar[stage], (uint16_t)((1u << CL) - 1));   // ... in every CTA that shares the weight tile
            if (seg_end) mma_commit(&tfull_bar[acc]);      // accumulator (segment) complete
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
        if (seg_end) { if (++acc == 2) { acc = 0; acc_phase ^= 1; } }
      }
    }
  } else if (warp < 2 + EW) {
    // ===================================================== epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1)
    // TMEM gives each thread one pixel (row) x 32 consecutive channels; the 32x32 chunk is transposed
    // through shared memory so global reads/writes are full 128-byte rows (8 lanes x float4).
    const int quarter = warp & 3;
    const uint32_t stg_s = smem_u32(stg_base + (warp - 2) * 32 * STG_PITCH);
    const int chunk0 = (warp - 2) / 4, chunk_step = EW / 4;   // EW == 8: the two warps of a quarter interleave chunks
    const int q8 = lane & 7, rsub = lane >> 3;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int nt = (tile / CL) % p.tiles_n; int r = (tile / CL) / p.tiles_n * CL + tile % CL;
      const int tx = r % p.tiles_x; r /= p.tiles_x; const int ty = r % p.tiles_y; const int n = r / p.tiles_y;
      constexpr int NLC = SPLIT ? 4 / (EW / 4) : 1;   // 32-column chunks owned by this warp in SPLIT mode (4, or 2 with EW == 8)
      float racc[SPLIT ? 32 * NLC : 1];   // SPLIT: fp32 register accumulators (BN <= 128), summed across K segments
      uint4 rp_cur[SPLIT ? 1 : 8], rp_nxt[SPLIT ? 1 : 8];   // plain: residual of the current / next chunk (res_prefetch)
      bool rp_ok = false, rp_ok_nxt = false;
      if (SPLIT) {
#pragma unroll
        for (int i = 0; i < (SPLIT ? 32 * NLC : 1); ++i) racc[i] = 0.f;
        const int nseg = (ksteps + p.seg - 1) / p.seg;
        for (int sg = 0; sg < nseg; ++sg) {
          mbar_wait_t(&tfull_bar[acc], acc_phase, SPIN, ST_A);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t ta = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * acc_stride;
#pragma unroll
          for (int lc = 0; lc < NLC; ++lc) {
            const int ch = chunk0 + lc * chunk_step;
            if (ch * 32 < p.BN && !(p.dbg & 2)) {
              uint32_t t[32];
              tmem_ld32(ta + (uint32_t)(ch * 32), t);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
              for (int j = 0; j < 32; ++j) racc[(SPLIT ? lc * 32 : 0) + (SPLIT ? j : 0)] += __uint_as_float(t[j]);
            }
          }
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            if (PAIR) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[acc]), 0u));
            else mbar_arrive(&tempty_bar[acc]);
          }
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      } else {
        // the residual of this warp's first chunk is requested before the accumulator is even complete
        if (p.res.p) rp_ok = res_prefetch(p, nt * p.BN + chunk0 * 32, quarter, lane, tx, ty, n, rp_cur);
        mbar_wait_t(&tfull_bar[acc], acc_phase, SPIN, ST_A);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * acc_stride;
      const long long t_out0 = p.stall ? clock64() : 0;
#pragma unroll (SPLIT ? NLC : 1)
      for (int lcq = 0; lcq < (SPLIT ? NLC : 8); ++lcq) {
        const int chq = chunk0 + lcq * chunk_step;   // SPLIT: static register index lcq; plain: TMEM column chunk
        const int c0 = chq * 32;
        if (c0 >= p.BN) break;
        const int cbase = nt * p.BN + c0;
        if (cbase >= p.cout) break;   // warp-uniform
        uint32_t v[32];
        if (SPLIT) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(racc[(SPLIT ? lcq * 32 : 0) + (SPLIT ? j : 0)]);
        } else {
          tmem_ld32(taddr + (uint32_t)c0, v);
          if (p.res.p) {   // next chunk's residual: in flight during this chunk's phases
            const int c0n = (chq + chunk_step) * 32;
            rp_ok_nxt = (c0n < p.BN && nt * p.BN + c0n < p.cout) ? res_prefetch(p, nt * p.BN + c0n, quarter, lane, tx, ty, n, rp_nxt) : false;
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        if (EPI == 1) epi_chunk_gate(p, v, stg_s, cbase, quarter, lane, tx, ty, n);
        else if (EPI == 2) epi_chunk_plain(p, v, stg_s, cbase, quarter, lane, tx, ty, n);
        else epi_chunk(p, v, stg_s, cbase, quarter, lane, tx, ty, n, p.stall ? st_c : nullptr, SPLIT ? nullptr : rp_cur, SPLIT ? false : rp_ok);
        if (!SPLIT && p.res.p) {
#pragma unroll
          for (int i = 0; i < (SPLIT ? 1 : 8); ++i) rp_cur[i] = rp_nxt[i];
          rp_ok = rp_ok_nxt;
        }
      }
      if (p.stall) st_b += clock64() - t_out0;
      if (!SPLIT) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[acc]), 0u));   // the leader's MMA warp waits for both epilogues
          else mbar_arrive(&tempty_bar[acc]);
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (SPLIT) {
    // ===================================================== operand splitter (warps 6..9)
    // A is rewritten in place as A_hi = rn_tf32(a) and A_lo = rn_tf32(a - A_hi) goes to the second buffer at the
    // same (swizzled) offsets, so one descriptor shape serves both.  Round-to-nearest on both terms keeps the
    // split unbiased (truncation left a coherent ~2^-20 relative error per product, i.e. ~1e-6*sqrt(K)).
    const int t = threadIdx.x - 32 * (2 + EW);  // 0..127
    int stage = 0; uint32_t phase = 0;
    if (ATM) {
      // tensor-memory variant: thread = pixel row of its warp's TMEM lane quarter; hi / lo go to TMEM columns
      // [256 + 64 * stage, +32) / (+32, +64) and the MMAs take A from there, so the shared-memory port only carries
      // the TMA fill, one read of A and the weight reads (it is the limiter: 128 B/clk vs 3 MMAs per K step).
      const int quarter = warp & 3, r = quarter * 32 + lane;
      const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + 256u;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait_t(&full_bar[stage], phase, SPIN, ST_A);
          const uint8_t* arow = smem + stage * stage_bytes + r * 128;
          if (SF16 && !(p.dbg & 1)) {
            // fp16 hi / lo pairs: a = hi + lo with hi = rn_f16(a), lo = rn_f16(a - hi) (both exact-product operands of kind::f16
            // MMAs: 11 x 11 significand bits fit the fp32 accumulator); two K elements per 32-bit TMEM column.
            // K elements [32 s, 32 s + 32) of the block come from sub-tile s -> hi columns [16 s, 16 s + 16), lo columns 32 + the same.
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
              uint32_t hi[16], lo[16];
#pragma unroll
              for (int c = 0; c < 8; ++c) {   // 16-byte chunk c of row r sits at chunk ^ (r & 7)  (SWIZZLE_128B)
                const float4 v = *reinterpret_cast<const float4*>(arow + sub * A_BYTES + ((c ^ (r & 7)) << 4));
                const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
                const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
                const __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
                hi[2 * c] = *reinterpret_cast<const uint32_t*>(&h0); hi[2 * c + 1] = *reinterpret_cast<const uint32_t*>(&h1);
                lo[2 * c] = *reinterpret_cast<const uint32_t*>(&l0); lo[2 * c + 1] = *reinterpret_cast<const uint32_t*>(&l1);
              }
              tmem_st16(trow + (uint32_t)(stage * 64 + sub * 16), hi);
              tmem_st16(trow + (uint32_t)(stage * 64 + 32 + sub * 16), lo);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          } else if (!(p.dbg & 1)) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              uint32_t hi[16], lo[16];
#pragma unroll
              for (int c = 0; c < 4; ++c) {   // 16-byte chunk (half * 4 + c) of row r sits at chunk ^ (r & 7)  (SWIZZLE_128B)
                const float4 v = *reinterpret_cast<const float4*>(arow + (((half * 4 + c) ^ (r & 7)) << 4));
                const float h0 = rn_tf32(v.x), h1 = rn_tf32(v.y), h2 = rn_tf32(v.z), h3 = rn_tf32(v.w);
                hi[c * 4 + 0] = __float_as_uint(h0); lo[c * 4 + 0] = __float_as_uint(rn_tf32(v.x - h0));
                hi[c * 4 + 1] = __float_as_uint(h1); lo[c * 4 + 1] = __float_as_uint(rn_tf32(v.y - h1));
                hi[c * 4 + 2] = __float_as_uint(h2); lo[c * 4 + 2] = __float_as_uint(rn_tf32(v.z - h2));
                hi[c * 4 + 3] = __float_as_uint(h3); lo[c * 4 + 3] = __float_as_uint(rn_tf32(v.w - h3));
              }
              tmem_st16(trow + (uint32_t)(stage * 64 + half * 16), hi);
              tmem_st16(trow + (uint32_t)(stage * 64 + 32 + half * 16), lo);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          }
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            if (PAIR) mbar_arrive_cluster(mapa_shared(smem_u32(&xf_bar[stage]), 0u));   // the leader's MMA consumes both CTAs' A rows
            else mbar_arrive(&xf_bar[stage]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(&full_bar[stage], phase, SPIN);
        float4* a = reinterpret_cast<float4*>(smem + stage * stage_bytes);
        float4* lo = reinterpret_cast<float4*>(smem + stage * stage_bytes + A_BYTES);
        if (!(p.dbg & 1))
#pragma unroll
        for (int i = 0; i < A_BYTES / 16 / 128; ++i) {
          const float4 v = a[t + i * 128];
          float4 hh, ll;
          hh.x = rn_tf32(v.x); ll.x = rn_tf32(v.x - hh.x);
          hh.y = rn_tf32(v.y); ll.y = rn_tf32(v.y - hh.y);
          hh.z = rn_tf32(v.z); ll.z = rn_tf32(v.z - hh.z);
          hh.w = rn_tf32(v.w); ll.w = rn_tf32(v.w - hh.w);
          a[t + i * 128] = hh;
          lo[t + i * 128] = ll;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&xf_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  }
  if (p.stall && lane == 0) {
    // [0] producer: empty waits  [1] MMA: operand waits  [2] MMA: accumulator waits  [3] epilogue warp 2: tfull waits
    // [4] epilogue warp 2: output phase  [5] splitter warp: full waits  [6] CTA cycles (MMA warp)
    unsigned long long* o = p.stall + (size_t)blockIdx.x * 16;
    if (warp == 1) { o[1] = (unsigned long long)st_a; o[2] = (unsigned long long)st_b; o[6] = (unsigned long long)(clock64() - t_start); }
    if (warp == 2) { o[3] = (unsigned long long)st_a; o[4] = (unsigned long long)st_b; o[8] = (unsigned long long)st_c[0]; o[9] = (unsigned long long)st_c[1]; o[10] = (unsigned long long)st_c[2]; }
    if (warp == 2 + EW) o[5] = (unsigned long long)st_a;
  }
  if (CL > 1) cluster_sync_all(); else __syncthreads();   // no CTA leaves while its peer may still write to it
  if (warp == 1) {
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || !p) throw std::runtime_error("conv_tc: cuTensorMapEncodeTiled is unavailable");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

void encode(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box,
            bool f16, int pixel_stride) {
  cuuint32_t estr[5] = {1, (cuuint32_t)pixel_stride, (cuuint32_t)pixel_stride, 1, 1};   // (activation maps: dims 1, 2 = x, y)
  CUresult r = encode_fn()(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("conv_tc: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
}

// stride 2: the box spans 2x the tile in x and y and TMA keeps every second element (elementStrides), so shared memory
// receives the same dense 16 x 8 pixel tile as at stride 1
static void encode_act(CUtensorMap* m, const TV& t, int stride = 1) {
  cuuint64_t dims[4] = {(cuuint64_t)t.c, (cuuint64_t)t.w, (cuuint64_t)t.h, (cuuint64_t)t.n};
  const cuuint64_t es = t.f16 ? 2 : 4;
  cuuint64_t str[3] = {(cuuint64_t)t.ld * es, (cuuint64_t)t.w * t.ld * es, (cuuint64_t)t.sn * es};
  cuuint32_t box[4] = {(cuuint32_t)(t.f16 ? 2 * BK : BK), (cuuint32_t)(TILE_W * stride), (cuuint32_t)(TILE_H * stride), 1};   // 128-byte rows either way
  encode(m, t.p, 4, dims, str, box, t.f16 != 0, stride);
}

}  // namespace tc

// launch with an optional (2,1,1) thread-block cluster
template <bool SPLIT, int CL, int EW, bool PAIR = false, int EPI = 0>
static void launch_tc(int grid, int threads, int smem, gvStream_t stream, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                      const tc::Params& p) {
  static volatile unsigned char attr_set[64];   // per (instantiation, device)
  gv_set_max_smem(tc::conv2d_tc_kernel<SPLIT, CL, EW, PAIR, EPI>, 227 * 1024, attr_set);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t er = cudaLaunchKernelEx(&cfg, tc::conv2d_tc_kernel<SPLIT, CL, EW, PAIR, EPI>, a0, a1, b, p);
  if (er != cudaSuccess) throw std::runtime_error(std::string("conv_tc: launch failed: ") + cudaGetErrorString(er));
}

static int tc_debug() {   // GIMMVFI_TC_DEBUG: timing experiments that break the numerics (never set in production)
  static int v = -1;
  if (v < 0) { const char* s = getenv("GIMMVFI_TC_DEBUG"); v = s ? atoi(s) : 0; }
  return v;
}
static unsigned long long* tc_stall_buf() {   // GIMMVFI_TC_STALL_BUF=<device address of >= 148*16 u64>: scripts/tc_split_probe.py
  const char* s = getenv("GIMMVFI_TC_STALL_BUF");
  return s ? reinterpret_cast<unsigned long long*>(strtoull(s, nullptr, 0)) : nullptr;
}
static int tc_atmem() {   // GIMMVFI_TC_ATMEM=0: keep the split A operand in shared memory (the first implementation)
  static int v = -1;
  if (v < 0) { const char* s = getenv("GIMMVFI_TC_ATMEM"); v = s ? atoi(s) : 1; }
  return v;
}
static int tc_pair() {   // GIMMVFI_TC_PAIR: 0 = single-CTA MMAs (+ weight multicast) everywhere; 1 = cta_group::2 CTA pairs in the plain / f16 kernel;
                         // 2 (default) = also in the 3xTF32 kernel
  static int v = -1;
  if (v < 0) { const char* s = getenv("GIMMVFI_TC_PAIR"); v = s ? atoi(s) : 2; }
  return v;
}
static int tc_cluster() {   // GIMMVFI_TC_CLUSTER=1 disables the 2-CTA weight multicast
  static int c = -1;
  if (c < 0) { const char* s = getenv("GIMMVFI_TC_CLUSTER"); c = s ? atoi(s) : 2; if (c != 1 && c != 2) c = 2; }
  return c;
}

static bool tc_epi8() {   // GIMMVFI_TC_EPI8=0 keeps 4 epilogue warps everywhere
  static int v = -1;
  if (v < 0) { const char* s = getenv("GIMMVFI_TC_EPI8"); v = s ? atoi(s) : 1; }
  return v != 0;
}

static int tc_split_epi8() {   // GIMMVFI_TC_SPLIT_EPI8: 0 = 4 drain warps in the 3xTF32 kernel, 1 = 4 for K-rich full-width tiles, 2 (default) = always 8
  static int v = -1;
  if (v < 0) { const char* s = getenv("GIMMVFI_TC_SPLIT_EPI8"); v = s ? atoi(s) : 2; }
  return v;
}

static int tc_split_f16() {   // GIMMVFI_TC_SPLIT_F16=0: keep the 3xTF32 form of the split kernel
  static int v = -1;
  if (v < 0) { const char* s = getenv("GIMMVFI_TC_SPLIT_F16"); v = s ? atoi(s) : 1; }
  return v;
}
// conditions of the gate-epilogue instantiation (epi_chunk_gate)
static bool gate_epilogue_ok(const tc::Params& p, const ConvW& w, const ConvEpi& e, const TV& out) {
  static int on = -1;
  if (on < 0) { const char* s = getenv("GIMMVFI_TC_GATE_EPI"); on = s ? atoi(s) : 1; }
  auto vec = [](const TV& t) { return !t.p || (!t.f16 && t.ld % 4 == 0 && t.sn % 4 == 0 && (reinterpret_cast<uintptr_t>(t.p) & 15) == 0); };
  if (!on || !(e.res.p || e.mul.p || e.gru_z.p) || (e.mul.p && e.gru_z.p) || e.act1 != ACT_NONE) return false;
  if (e.act2 != ACT_NONE && e.act2 != ACT_SIGMOID && e.act2 != ACT_TANH) return false;
  if (w.cout % 4 || p.round_out || (e.split_c && e.split_c % 4)) return false;
  return vec(e.res) && vec(e.mul) && vec(e.gru_z) && vec(e.gru_h) && vec(out) && out.p && vec(e.out2);
}
// conditions of the plain-epilogue instantiation (epi_chunk_plain)
static bool plain_epilogue_ok(const tc::Params& p, const ConvW& w, const ConvEpi& e, const TV& out) {
  static int on = -1;
  if (on < 0) { const char* s = getenv("GIMMVFI_TC_PLAIN_EPI"); on = s ? atoi(s) : 1; }
  if (!on || e.res.p || e.mul.p || e.gru_z.p || e.out2.p || e.split_c || e.act2 != ACT_NONE) return false;
  if (e.act1 != ACT_NONE && e.act1 != ACT_RELU) return false;
  if (w.cout % 4 || p.round_out || out.f16 || out.ld % 4 || out.sn % 4 || (reinterpret_cast<uintptr_t>(out.p) & 15)) return false;
  return true;
}
static int tc_seg_f16() {   // K steps (of 64 elements) per TMEM accumulation segment of the 3xF16 form.  2 (default): draining the
                            // 64 KB accumulator (TMEM reads: 64 B/clk/SM) every step costs more than the step's MMAs; full-frame parity at
                            // 1088x1920: 0 of 6.27 M values off by > 1e-3 with 1 and with 2, one outlier appears with 3 (profiles/r02_fullframe_parity.log)
  static int seg = -1;
  if (seg < 0) { const char* s = getenv("GIMMVFI_TC_SEG_F16"); seg = s ? atoi(s) : 2; if (seg < 1) seg = 1; }
  return seg;
}
static int tc_seg() {
  static int seg = -1;
  if (seg < 0) { const char* s = getenv("GIMMVFI_TC_SEG"); seg = s ? atoi(s) : 2; if (seg < 1) seg = 1; }
  return seg;
}

void conv2d_tc(Ctx& cx, const TV& in0, const TV& in1, const ConvW& w, const ConvGeom& g, const ConvEpi& e, const TV& out, bool split) {
  using namespace tc;
  CUtensorMap mA0, mA1, mB;
  encode_act(&mA0, in0, g.stride);
  if (in1.p) encode_act(&mA1, in1, g.stride); else mA1 = mA0;
  int BN, tiles_n;
  const int pix_tiles_host = out.n * ((out.h + TILE_H - 1) / TILE_H) * ((out.w + TILE_W - 1) / TILE_W);
  if (split) {  // register-promoted accumulation holds BN fp32 accumulators per thread -> BN <= 128
    const int c16 = (w.cout + 15) & ~15;
    tiles_n = (c16 + 127) / 128;
    BN = (((c16 + tiles_n - 1) / tiles_n) + 15) & ~15;   // weight rows beyond cout_pad are TMA zero-fill
    // several N tiles: a tile's last 32-column epilogue chunk must not reach into the next tile's channels (it would store columns
    // the MMA never wrote) -> tile width = whole chunks
    if (tiles_n > 1) BN = (BN + 31) & ~31;
  } else {
    tc_tile_n(w.cout, &BN, &tiles_n);
    if (BN * tiles_n != w.cout_pad) throw std::runtime_error("conv_tc: weight padding does not match the N tiling");
  }
  // 2-CTA clusters pay off when there are at least two pixel tiles per SM; BN/2 must keep the 8-row swizzle atom
  const int CL = (tc_cluster() == 2 && pix_tiles_host >= 2 * cx.sm_count && BN % 16 == 0) ? 2 : 1;
  const int taps = w.kh * w.kw;
  const bool f16 = in0.f16 != 0;
  // 3xF16: the split kernel on kind::f16 with fp16 hi / lo operand pairs (two 32-channel activation boxes per 64-element K step);
  // needs the TMEM-resident A operand, pre-split weights, and a first segment of whole K blocks
  const bool sf16 = split && tc_split_f16() && tc_atmem() && w.w_tc_s != nullptr && (!in1.p || in0.c % 64 == 0);
  const int bk = (f16 || sf16) ? 2 * BK : BK, cin_pad = f16 ? w.cin_pad_h : (sf16 ? w.cin_pad_s : w.cin_pad);
  if (sf16) {
    cuuint64_t dims[3] = {(cuuint64_t)cin_pad, (cuuint64_t)w.cout_pad, (cuuint64_t)(taps * 2)};
    cuuint64_t str[2] = {(cuuint64_t)cin_pad * 2, (cuuint64_t)cin_pad * w.cout_pad * 2};
    cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)(BN / CL), 1};
    encode(&mB, w.w_tc_s, 3, dims, str, box, true);
  } else if (f16) {
    cuuint64_t dims[3] = {(cuuint64_t)cin_pad, (cuuint64_t)w.cout_pad, (cuuint64_t)taps};
    cuuint64_t str[2] = {(cuuint64_t)cin_pad * 2, (cuuint64_t)cin_pad * w.cout_pad * 2};
    cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)(BN / CL), 1};
    encode(&mB, w.w_tc_h, 3, dims, str, box, true);
  } else {
    cuuint64_t dims[3] = {(cuuint64_t)w.cin_pad, (cuuint64_t)w.cout_pad, (cuuint64_t)(taps * (w.has_lo ? 2 : 1))};
    cuuint64_t str[2] = {(cuuint64_t)w.cin_pad * 4, (cuuint64_t)w.cin_pad * w.cout_pad * 4};
    cuuint32_t box[3] = {BK, (cuuint32_t)(BN / CL), 1};   // CL == 2: each CTA fetches half of the weight rows
    encode(&mB, w.w_tc, 3, dims, str, box);
  }
  Params p = Params();
  p.taps = taps; p.kw = w.kw; p.ph = g.ph; p.pw = g.pw; p.stride = g.stride;
  p.f16_in = f16 ? 1 : 0; p.bk = bk; p.split_f16 = sf16 ? 1 : 0;
  p.c0_blocks = in1.p ? in0.c / bk : (in0.c + bk - 1) / bk;
  p.kblocks = cin_pad / bk;
  p.tiles_x = (out.w + TILE_W - 1) / TILE_W; p.tiles_y = (out.h + TILE_H - 1) / TILE_H; p.n_img = out.n; p.tiles_n = tiles_n;
  p.H = out.h; p.W = out.w; p.BN = BN; p.cout = w.cout;
  p.round_out = (split || out.f16) ? 0 : 1;   // (a half store already rounds to 10 mantissa bits)
  p.out_scale = sf16 ? 1.f / w.w_scale : 1.f;
  p.seg = sf16 ? tc_seg_f16() : tc_seg();
  static int spin = -1;
  if (spin < 0) { const char* s = getenv("GIMMVFI_TC_SPIN_LIMIT"); spin = s ? atoi(s) : 400; }
  p.spin_limit = spin; p.dbg = tc_debug(); p.atmem = split ? tc_atmem() : 0; p.stall = tc_stall_buf();
  p.bias = w.b; p.act1 = e.act1; p.slope1 = e.slope1; p.act2 = e.act2; p.slope2 = e.slope2;
  p.res = e.res; p.mul = e.mul; p.gru_z = e.gru_z; p.gru_h = e.gru_h; p.out = out; p.out2 = e.out2; p.split_c = e.split_c;
  // 3xTF32: 8 drain/epilogue warps help K-poor layers (+20-30 %) but steal issue slots from the splitter warps on
  // K-rich full-width tiles (SepConvGRU gates: -15 %): measured in profiles/r01_tc_microbench_split_epi8.log
  const bool sew8 = split && tc_epi8() && tc_split_epi8() && (tc_split_epi8() == 2 || !(BN == 128 && taps * (w.cin_pad / 32) >= 48));
  // CTA pairs: N/2 rows per CTA must keep the 8-row swizzle atom (and N % 16); the 3xTF32 kernel pairs only in its TMEM-A, 8-drain-warp form
  const bool pair = CL == 2 && BN % 32 == 0 && (split ? (tc_pair() >= 2 && p.atmem && sew8) : (tc_pair() != 0));
  const int stage_bytes = split ? (((p.atmem && !sf16) ? 1 : 2) * A_BYTES + 2 * (pair ? BN / 2 : BN) * BK * 4) : (A_BYTES + (pair ? BN / 2 : BN) * BK * 4);
  static int ew8_wide = -1;   // GIMMVFI_TC_EPI8_WIDE=0: 4 epilogue warps for N > 128 tiles of CTA pairs
  if (ew8_wide < 0) { const char* q = getenv("GIMMVFI_TC_EPI8_WIDE"); ew8_wide = q ? atoi(q) : 1; }
  // K-poor plain layers are epilogue bound: 8 epilogue warps.  CTA pairs halve the per-stage smem footprint, which leaves room
  // for the 8-warp staging area next to >= 5 stages even at N = 256 (where the f16 trunk's epilogue is as long as its main loop)
  const bool ew8 = !split && (BN <= 128 || (pair && ew8_wide)) && tc_epi8();
  const int stg_bytes = ((ew8 || sew8) ? 8 : 4) * STG_WARP_BYTES;
  const int budget = 227 * 1024 - 1024 /*align*/ - stg_bytes - BAR_BYTES;
  p.stages = budget / stage_bytes;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  if (p.atmem && p.stages > 4) p.stages = 4;   // the TMEM A ring has 4 slots of 64 columns
  if (p.stages < 2) throw std::runtime_error("conv_tc: not enough shared memory for 2 pipeline stages");
  const int smem = p.stages * stage_bytes + stg_bytes + BAR_BYTES + 1024;
  const int padded_tiles = (pix_tiles_host + CL - 1) / CL * CL * tiles_n;
  int grid = padded_tiles < cx.sm_count ? padded_tiles : cx.sm_count;
  grid -= grid % CL;
  cx.launches++;
  if (cx.prof) {
    char nm[128];
    snprintf(nm, sizeof nm, "conv2d_tc_%s k%dx%d c%d>%d @%dx%dx%d", split ? (sf16 ? "3xf16" : "3xtf32") : (f16 ? "f16" : "tf32"), w.kh, w.kw, w.cin, w.cout, out.n, out.h, out.w);
    cx.prof->begin(cx.stream, prof_intern(nm), 2.0 * (double)out.n * out.h * out.w * w.cout * (double)w.cin * w.kh * w.kw);
  }
  if (split && sew8) {
    if (pair && gate_epilogue_ok(p, w, e, out)) launch_tc<true, 2, 8, true, 1>(grid, 448, smem, cx.stream, mA0, mA1, mB, p);
    else if (pair && plain_epilogue_ok(p, w, e, out)) launch_tc<true, 2, 8, true, 2>(grid, 448, smem, cx.stream, mA0, mA1, mB, p);
    else if (pair) launch_tc<true, 2, 8, true>(grid, 448, smem, cx.stream, mA0, mA1, mB, p);
    else if (CL == 2) launch_tc<true, 2, 8>(grid, 448, smem, cx.stream, mA0, mA1, mB, p);
    else launch_tc<true, 1, 8>(grid, 448, smem, cx.stream, mA0, mA1, mB, p);
  }
  else if (split) { if (CL == 2) launch_tc<true, 2, 4>(grid, 320, smem, cx.stream, mA0, mA1, mB, p); else launch_tc<true, 1, 4>(grid, 320, smem, cx.stream, mA0, mA1, mB, p); }
  else if (ew8) {
    if (pair) launch_tc<false, 2, 8, true>(grid, 320, smem, cx.stream, mA0, mA1, mB, p);
    else if (CL == 2) launch_tc<false, 2, 8>(grid, 320, smem, cx.stream, mA0, mA1, mB, p);
    else launch_tc<false, 1, 8>(grid, 320, smem, cx.stream, mA0, mA1, mB, p);
  } else {
    if (pair) launch_tc<false, 2, 4, true>(grid, 192, smem, cx.stream, mA0, mA1, mB, p);
    else if (CL == 2) launch_tc<false, 2, 4>(grid, 192, smem, cx.stream, mA0, mA1, mB, p);
    else launch_tc<false, 1, 4>(grid, 192, smem, cx.stream, mA0, mA1, mB, p);
  }
  gv_check_launch("conv2d_tc");
  if (cx.prof) cx.prof->end(cx.stream);
}


// All-pairs correlation vol[i][j] = scale * <fa[i,:], fb[j,:]> (raft/corr.py:167-175) as a 3xTF32 GEMM on the
// same kernel: A = fa pixels (M), "weights" = the other frame's features, K-major as they lie in NHWC memory.
// fb_planes = [2][N][C]: plane 0 = rn_tf32(x), plane 1 = rn_tf32(x - plane 0)  (split_planes in corr.cu).
void corr_volume_tc(Ctx& cx, const TV& fa, const float* fb_planes, const float* zero_bias, float* vol, float scale, bool split, int n_targets) {
  using namespace tc;
  if (fa.n != 1) throw std::runtime_error("corr_volume_tc: one sample per launch");
  // M = the source pixels of fa; N = n_targets rows of fb (0: as many as sources).  Pooled pyramid levels are GEMMs against the
  // avg-pooled target features (pooling and the dot product commute), so the level-0 volume is never re-read.
  const int Ms = fa.h * fa.w, N = n_targets > 0 ? n_targets : Ms, C = fa.c;
  CUtensorMap mA, mB;
  encode_act(&mA, fa);
  // split: fb_planes = [2][N][C] (rn / residual planes).  plain TF32: fb_planes = the raw K-major features [N][C] of the
  // other frame (the tensor core truncates them).  BN = 128 in both (register accumulators / 8 epilogue warps).
  const int BN = 128, tiles_n = (N + BN - 1) / BN;
  const int pix_tiles_host = ((fa.h + TILE_H - 1) / TILE_H) * ((fa.w + TILE_W - 1) / TILE_W);
  const bool e8 = tc_epi8() && (!split || tc_split_epi8());
  // CTA pairs as in conv2d_tc: two neighbouring source-pixel tiles share the target-pixel ("weight") tile, half of it per CTA
  const bool atm = split && tc_atmem();
  const bool sf16 = split && corr_volume_tc_wants_f16_planes() && C % 64 == 0;
  const bool pair = e8 && pix_tiles_host >= 2 && (split ? (tc_pair() >= 2 && atm) : (tc_pair() != 0)) && pix_tiles_host * tiles_n >= 2 * cx.sm_count;
  const int CL = pair ? 2 : 1;
  if (sf16) {   // half planes, 64-element K blocks (128-byte rows)
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)N, 2};
    cuuint64_t str[2] = {(cuuint64_t)C * 2, (cuuint64_t)N * C * 2};
    cuuint32_t box[3] = {2 * BK, (cuuint32_t)(BN / CL), 1};
    encode(&mB, fb_planes, 3, dims, str, box, true);
  } else {
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)N, (cuuint64_t)(split ? 2 : 1)};
    cuuint64_t str[2] = {(cuuint64_t)C * 4, (cuuint64_t)N * C * 4};
    cuuint32_t box[3] = {BK, (cuuint32_t)(BN / CL), 1};
    encode(&mB, fb_planes, 3, dims, str, box);
  }
  Params p = Params();
  p.taps = 1; p.kw = 1; p.ph = 0; p.pw = 0; p.stride = 1;
  p.kblocks = sf16 ? C / 64 : C / 32; p.c0_blocks = p.kblocks; p.f16_in = 0; p.bk = sf16 ? 2 * BK : BK; p.split_f16 = sf16 ? 1 : 0;
  p.tiles_x = (fa.w + TILE_W - 1) / TILE_W; p.tiles_y = (fa.h + TILE_H - 1) / TILE_H; p.n_img = 1; p.tiles_n = tiles_n;
  p.H = fa.h; p.W = fa.w; p.BN = BN; p.cout = N; p.round_out = 0; p.out_scale = scale; p.seg = sf16 ? tc_seg_f16() : tc_seg();
  static int spin = -1;
  if (spin < 0) { const char* s = getenv("GIMMVFI_TC_SPIN_LIMIT"); spin = s ? atoi(s) : 400; }
  p.spin_limit = spin; p.dbg = tc_debug(); p.atmem = atm ? 1 : 0; p.stall = tc_stall_buf();
  p.bias = zero_bias; p.act1 = ACT_NONE; p.slope1 = nullptr; p.act2 = ACT_NONE; p.slope2 = nullptr;
  p.out = make_tv(vol, 1, fa.h, fa.w, N, N); p.out2 = TV(); p.split_c = 0;
  const int bn_cta = pair ? BN / 2 : BN;
  const int stage_bytes = split ? (((p.atmem && !sf16) ? 1 : 2) * A_BYTES + 2 * bn_cta * BK * 4) : (A_BYTES + bn_cta * BK * 4);
  const int stg_bytes = (e8 ? 8 : 4) * STG_WARP_BYTES;
  p.stages = (227 * 1024 - 1024 - stg_bytes - BAR_BYTES) / stage_bytes;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  if (p.atmem && p.stages > 4) p.stages = 4;
  const int smem = p.stages * stage_bytes + stg_bytes + BAR_BYTES + 1024;
  const int num_tiles = (pix_tiles_host + CL - 1) / CL * CL * tiles_n;
  int grid = num_tiles < cx.sm_count ? num_tiles : cx.sm_count;
  grid -= grid % CL;
  cx.launches++;
  if (cx.prof) cx.prof->begin(cx.stream, split ? (sf16 ? "corr_gemm_tc_3xf16" : "corr_gemm_tc_3xtf32") : "corr_gemm_tc_tf32", 2.0 * (double)Ms * N * C);
  // plain fp32 rows of the volume: the compact plain-epilogue instantiation when every row is a whole number of 16-byte groups
  static int plain_on = -1;
  if (plain_on < 0) { const char* s = getenv("GIMMVFI_TC_PLAIN_EPI"); plain_on = s ? atoi(s) : 1; }
  if (split && pair && plain_on && N % 4 == 0 && (reinterpret_cast<uintptr_t>(vol) & 15) == 0) launch_tc<true, 2, 8, true, 2>(grid, 448, smem, cx.stream, mA, mA, mB, p);
  else if (split && pair) launch_tc<true, 2, 8, true>(grid, 448, smem, cx.stream, mA, mA, mB, p);
  else if (split && e8) launch_tc<true, 1, 8>(grid, 448, smem, cx.stream, mA, mA, mB, p);
  else if (split) launch_tc<true, 1, 4>(grid, 320, smem, cx.stream, mA, mA, mB, p);
  else if (pair) launch_tc<false, 2, 8, true>(grid, 320, smem, cx.stream, mA, mA, mB, p);
  else if (e8) launch_tc<false, 1, 8>(grid, 320, smem, cx.stream, mA, mA, mB, p);
  else launch_tc<false, 1, 4>(grid, 192, smem, cx.stream, mA, mA, mB, p);
  gv_check_launch("corr_volume_tc");
  if (cx.prof) cx.prof->end(cx.stream);
}

bool corr_volume_tc_wants_f16_planes() { return tc_split_f16() != 0 && tc_atmem() != 0; }

}  // namespace gv
#endif  // GV_HOSTSIM
