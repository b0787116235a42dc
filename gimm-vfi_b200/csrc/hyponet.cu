// Fused HypoNet (modules/hyponet.py:71-146) for sm_100a: the whole 5-layer per-pixel INR-MLP
//     x = [latent32 | t, y, x] -> 128 -> 128 -> 128 -> 128 -> 2     (sin between layers, fan-in-normalised columns, +output_bias)
// in ONE kernel.  Per 128-pixel tile: the latent rows arrive by TMA as a 128B-swizzled K-major tile, every layer is a tcgen05.mma
// with its fp32 accumulator in tensor memory, the epilogue warps read the accumulator (tcgen05.ld), add bias, apply sin and write
// the activations back to SHARED memory as the next layer's K-major fp16 A operand — the 128-channel activations never touch HBM
// (the five 1x1 convolutions this replaces made three 2.1 GB round trips at 1088x1920).  All five weight matrices stay resident in
// shared memory (pre-swizzled at pack time: 116 KB).
//   layer 0   A = latent32 fp32 (TF32 MMA, K = 32); the (t, y, x) coordinates and the bias enter as an exact fp32 affine term in the
//             epilogue (3 FMAs per output): no packed [latent | coords] tensor, full-precision coordinates.
//   layers 1-3 A = sin() activations in IEEE half (|x| <= 1; same 10-bit mantissa as TF32), kind::f16 MMAs, K = 128.
//   layer 4   N = 16 (2 used), + bias (output_bias folded in) -> normalised flow, fp32.
// Three independent 128-thread groups per CTA (one pixel tile each, own accumulator / activation buffer / barriers) share the weights:
// while one group's epilogue occupies the MUFU / FMA pipes, another group's MMAs occupy the tensor pipe.
#include "common.h"

#ifndef GV_HOSTSIM
#include <cuda.h>
#include <cuda_fp16.h>

namespace gv {
namespace tc {
#include "tc_ptx.cuh"

// constants live in SHARED memory: explicit ld.shared (a generic-pointer load goes through the global / local queue and throttles)
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory"); }

struct HypoParams {
  const uint8_t* blob; const float* coords; float* out;
  long long P; int num_tiles; int out_ld; int spin_limit;
};

// one epilogue pass over the 128 accumulator columns of this thread's pixel row: z = acc + bias (+ affine) -> sin -> half -> smem
template <bool FIRST>
__device__ __forceinline__ void hypo_epilogue(uint32_t my_t, uint8_t* abuf, int row, uint32_t bias_s, uint32_t aff_s, float ct, float cy, float cx) {
#pragma unroll 1
  for (int ch = 0; ch < 4; ++ch) {
    uint32_t v[32];
    tmem_ld32(my_t + (uint32_t)(ch * 32), v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t h[16];
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const int c = ch * 32 + j;
      const float4 b = lds_f4(bias_s + 4u * c);
      float z0 = __uint_as_float(v[j]) + b.x, z1 = __uint_as_float(v[j + 1]) + b.y, z2 = __uint_as_float(v[j + 2]) + b.z, z3 = __uint_as_float(v[j + 3]) + b.w;
      if (FIRST) {   // (t, y, x) . W0[32..34, :] in fp32
        const float4 wt = lds_f4(aff_s + 4u * c), wy = lds_f4(aff_s + 4u * (128 + c)), wx = lds_f4(aff_s + 4u * (256 + c));
        z0 = fmaf(ct, wt.x, fmaf(cy, wy.x, fmaf(cx, wx.x, z0))); z1 = fmaf(ct, wt.y, fmaf(cy, wy.y, fmaf(cx, wx.y, z1)));
        z2 = fmaf(ct, wt.z, fmaf(cy, wy.z, fmaf(cx, wx.z, z2))); z3 = fmaf(ct, wt.w, fmaf(cy, wy.w, fmaf(cx, wx.w, z3)));
      }
      const __half2 p0 = __floats2half2_rn(__sinf(z0), __sinf(z1)), p1 = __floats2half2_rn(__sinf(z2), __sinf(z3));
      h[j / 2] = *reinterpret_cast<const uint32_t*>(&p0); h[j / 2 + 1] = *reinterpret_cast<const uint32_t*>(&p1);
    }
    // K index = ch * 32 + j -> K block (64 halves = one 128-byte row) ch >> 1, 16-byte chunk (ch & 1) * 4 + q, swizzled with the row
    uint8_t* rowp = abuf + (ch >> 1) * 16384 + row * 128;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int chunk = ((ch & 1) * 4 + q) ^ (row & 7);
      *reinterpret_cast<uint4*>(rowp + (chunk << 4)) = make_uint4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
    }
  }
}

constexpr int HN_W0 = hypo::W0, HN_W1 = hypo::W1, HN_W4 = hypo::W4, HN_AFF = hypo::AFF, HN_B1 = hypo::B1, HN_B4 = hypo::B4, HN_BLOB = hypo::BLOB,
              HN_BLOB_SMEM = (HN_BLOB + 1023) / 1024 * 1024, HN_GROUPS = 3;

__global__ void __launch_bounds__(128 * HN_GROUPS, 1) hyponet_fused_kernel(const __grid_constant__ CUtensorMap tmLat, const __grid_constant__ HypoParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wsm = smem;
  uint8_t* act = smem + HN_BLOB_SMEM;
  uint64_t* bars = reinterpret_cast<uint64_t*>(act + HN_GROUPS * 32768);   // [g] input landed, [HN_GROUPS + g] MMAs retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * HN_GROUPS);
  const int tid = threadIdx.x, warp = tid >> 5, g = warp >> 2, wq = warp & 3, row = tid - g * 128;
  for (int i = tid; i < HN_BLOB / 16; i += 128 * HN_GROUPS) reinterpret_cast<uint4*>(wsm)[i] = __ldg(reinterpret_cast<const uint4*>(p.blob) + i);
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmLat)) : "memory");
    for (int i = 0; i < 2 * HN_GROUPS; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the weights were written through the generic proxy; MMAs read them through the async proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_t = tmem_base + (uint32_t)(g * 128);
  const uint32_t my_t = acc_t + ((uint32_t)(wq * 32) << 16);
  uint8_t* abuf = act + g * 32768;
  uint64_t* bar_in = &bars[g];
  uint64_t* bar_mma = &bars[HN_GROUPS + g];
  uint32_t ph_in = 0, ph_mma = 0;
  const bool leader = row == 0;
  const int SPIN = p.spin_limit;
  // instruction descriptors: D = f32; A/B K-major; tf32 (format 2) for layer 0, f16 (format 0) after; N >> 3 at bit 17, M >> 4 at bit 24
  const uint32_t id_tf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t id_f16 = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t id_f16_n16 = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t a_s = smem_u32(abuf), w_s = smem_u32(wsm);
  const uint32_t aff_s = w_s + HN_AFF;

  for (int tile = blockIdx.x * HN_GROUPS + g; tile < p.num_tiles; tile += gridDim.x * HN_GROUPS) {
    const long long pix = (long long)tile * 128 + row;
    const bool valid = pix < p.P;
    float ct = 0.f, cy = 0.f, cx = 0.f;
    if (valid) { const float* c = p.coords + pix * 3; ct = __ldg(c); cy = __ldg(c + 1); cx = __ldg(c + 2); }
    if (leader) {
      mbar_expect_tx(bar_in, 16384u);
      tma_load_2d(abuf, &tmLat, bar_in, 0, tile * 128);   // rows beyond P are zero-filled
      mbar_wait(bar_in, ph_in, SPIN);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t ad = make_smem_desc(a_s), bd = make_smem_desc(w_s + HN_W0);
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_tf32(acc_t, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), id_tf32, k > 0 ? 1u : 0u);
      mma_commit(bar_mma);
    }
    ph_in ^= 1;
    mbar_wait(bar_mma, ph_mma, SPIN); ph_mma ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    hypo_epilogue<true>(my_t, abuf, row, aff_s + 4u * 384, aff_s, ct, cy, cx);
#pragma unroll 1
    for (int l = 1; l <= 4; ++l) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // activations (generic proxy) -> visible to the MMAs
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      group_sync(g);
      if (leader) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t wl = w_s + (l < 4 ? HN_W1 + (l - 1) * 32768 : HN_W4);
        const uint32_t kb_stride = l < 4 ? 16384u : 2048u;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t ad = make_smem_desc(a_s + kb * 16384), bd = make_smem_desc(wl + kb * kb_stride);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_f16(acc_t, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), l < 4 ? id_f16 : id_f16_n16, (kb | k) ? 1u : 0u);
        }
        mma_commit(bar_mma);
      }
      mbar_wait(bar_mma, ph_mma, SPIN); ph_mma ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (l < 4) {
        hypo_epilogue<false>(my_t, abuf, row, w_s + HN_B1 + (l - 1) * 512, aff_s, 0.f, 0.f, 0.f);
      } else {
        uint32_t v[16];
        tmem_ld16(my_t, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const float4 b4 = lds_f4(w_s + HN_B4);
        if (valid) {
          float* o = p.out + pix * p.out_ld;
          o[0] = __uint_as_float(v[0]) + b4.x; o[1] = __uint_as_float(v[1]) + b4.y;
        }
      }
    }
    // every thread of the group has drained the accumulator (and the last MMAs have read the activation buffer) before the next tile's TMA / MMAs
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    group_sync(g);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}


// ------------------------------------------------------------------------------------------------------------------------------
// fp32-class variant.  The flow a pixel gets is (2 o - 1) * max|flow| with o the MLP output: at 40 px motion an error of 1e-5 in o is
// 1e-3 px, so TF32 / half operands (2^-11) are not enough once the frames have real edges (the reference's demo frames: max|d imgt_pred|
// 2e-3 with the kernel above, 4e-4 with this one).  Same structure, but
//   * layers 1-3: D += A_hi B_lo + A_lo B_hi + A_hi B_hi on kind::f16 with fp16 hi / lo pairs (11 + 11 significand bits; every product
//     is exact in the fp32 accumulator).  The weights are pre-scaled by 2^6 (lo parts stay normal) and pre-split at pack time and stay
//     resident in shared memory (192 KB); the activations are split by the epilogue threads and written to TENSOR memory
//     (tcgen05.st, two halves per 32-bit column), from where the MMAs take their A operand - they never touch shared memory;
//   * layer 0 (K = 32 latents + (t, y, x) + bias) and layer 4 (2 outputs) run on the CUDA cores in plain fp32 from registers.
// Two 128-thread groups per CTA, each with its own accumulator (128 columns) and A_hi / A_lo regions (64 + 64 columns) in TMEM.
constexpr int H3_GROUPS = 2;

__device__ __forceinline__ void split_store(uint32_t t_hi, uint32_t t_lo, const float* s /*32 values, K order*/) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const __half2 h = __floats2half2_rn(s[2 * j], s[2 * j + 1]);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(s[2 * j] - f.x, s[2 * j + 1] - f.y);
    hi[j] = *reinterpret_cast<const uint32_t*>(&h); lo[j] = *reinterpret_cast<const uint32_t*>(&l);
  }
  tmem_st16(t_hi, hi);
  tmem_st16(t_lo, lo);
}

__global__ void __launch_bounds__(128 * H3_GROUPS, 1) hyponet_fused3_kernel(const __grid_constant__ HypoParams p, const float* __restrict__ lat, int lat_ld) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (hypo3::BLOB + 15) / 16 * 16);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + H3_GROUPS);
  const int tid = threadIdx.x, warp = tid >> 5, g = warp >> 2, wq = warp & 3, row = tid - g * 128;
  for (int i = tid; i < hypo3::BLOB / 16; i += 128 * H3_GROUPS) reinterpret_cast<uint4*>(smem)[i] = __ldg(reinterpret_cast<const uint4*>(p.blob) + i);
  if (tid == 0) {
    for (int i = 0; i < H3_GROUPS; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_t = tmem_base + (uint32_t)(g * 256);          // accumulator columns [0, 128) of this group's half
  const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
  const uint32_t ahi_t = acc_t + 128u, alo_t = acc_t + 192u;       // A_hi / A_lo: 64 columns each = 128 packed halves per row
  uint64_t* bar_mma = &bars[g];
  uint32_t ph = 0;
  const bool leader = row == 0;
  const int SPIN = p.spin_limit;
  const uint32_t id_f16 = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t w_s = smem_u32(smem);
  const uint32_t W0s = w_s + hypo3::W0A, B13s = w_s + hypo3::B13, W4s = w_s + hypo3::W4;
  const float inv_scale = 1.0f / hypo3::W_SCALE;

  for (int tile = blockIdx.x * H3_GROUPS + g; tile < p.num_tiles; tile += gridDim.x * H3_GROUPS) {
    const long long pix = (long long)tile * 128 + row;
    const bool valid = pix < p.P;
    // ---- layer 0 on the CUDA cores: z = [latent32 | t, y, x | 1] . W0A   (hyponet.py:101-117 with the ones column as bias)
    float x[35];
#pragma unroll
    for (int k = 0; k < 35; ++k) x[k] = 0.f;
    if (valid) {
      const float4* lp = reinterpret_cast<const float4*>(lat + pix * lat_ld);
#pragma unroll
      for (int q = 0; q < 8; ++q) { const float4 v = __ldg(lp + q); x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w; }
      const float* c = p.coords + pix * 3;
      x[32] = __ldg(c); x[33] = __ldg(c + 1); x[34] = __ldg(c + 2);
    }
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {
      float z[32];
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = lds_f4(W0s + 4u * (35 * 128 + ch * 32 + j));
        z[j] = b.x; z[j + 1] = b.y; z[j + 2] = b.z; z[j + 3] = b.w;
      }
#pragma unroll
      for (int k = 0; k < 35; ++k) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 w = lds_f4(W0s + 4u * (k * 128 + ch * 32 + j));
          z[j] = fmaf(x[k], w.x, z[j]); z[j + 1] = fmaf(x[k], w.y, z[j + 1]); z[j + 2] = fmaf(x[k], w.z, z[j + 2]); z[j + 3] = fmaf(x[k], w.w, z[j + 3]);
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) z[j] = __sinf(z[j]);
      split_store(ahi_t + lane_off + (uint32_t)(ch * 16), alo_t + lane_off + (uint32_t)(ch * 16), z);
    }
    float o0 = 0.f, o1 = 0.f;
#pragma unroll 1
    for (int l = 0; l < 3; ++l) {
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      group_sync(g);
      if (leader) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t bhi = make_smem_desc(w_s + hypo3::W13 + ((l * 2 + 0) * 2 + kb) * 16384), blo = make_smem_desc(w_s + hypo3::W13 + ((l * 2 + 1) * 2 + kb) * 16384);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t ah = ahi_t + (uint32_t)(kb * 32 + k * 8), al = alo_t + (uint32_t)(kb * 32 + k * 8);
            mma_f16_ts(acc_t, ah, blo + (uint64_t)(2 * k), id_f16, (kb | k) ? 1u : 0u);   // A_hi * B_lo
            mma_f16_ts(acc_t, al, bhi + (uint64_t)(2 * k), id_f16, 1u);                   // A_lo * B_hi
            mma_f16_ts(acc_t, ah, bhi + (uint64_t)(2 * k), id_f16, 1u);                   // A_hi * B_hi
          }
        }
        mma_commit(bar_mma);
      }
      mbar_wait(bar_mma, ph, SPIN); ph ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t v[32];
        tmem_ld32(acc_t + lane_off + (uint32_t)(ch * 32), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float z[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b = lds_f4(B13s + 4u * (l * 128 + ch * 32 + j));
          z[j] = __sinf(fmaf(__uint_as_float(v[j]), inv_scale, b.x)); z[j + 1] = __sinf(fmaf(__uint_as_float(v[j + 1]), inv_scale, b.y));
          z[j + 2] = __sinf(fmaf(__uint_as_float(v[j + 2]), inv_scale, b.z)); z[j + 3] = __sinf(fmaf(__uint_as_float(v[j + 3]), inv_scale, b.w));
        }
        if (l < 2) {
          split_store(ahi_t + lane_off + (uint32_t)(ch * 16), alo_t + lane_off + (uint32_t)(ch * 16), z);
        } else {   // layer 4 on the fly: o += h3 . W4
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float4 w = lds_f4(W4s + 4u * ((ch * 32 + j) * 2));
            o0 = fmaf(z[j], w.x, o0); o1 = fmaf(z[j], w.y, o1); o0 = fmaf(z[j + 1], w.z, o0); o1 = fmaf(z[j + 1], w.w, o1);
          }
        }
      }
    }
    if (valid) {
      const float4 b4 = lds_f4(w_s + hypo3::B4);
      float* o = p.out + pix * p.out_ld;
      o[0] = o0 + b4.x; o[1] = o1 + b4.y;
    }
    // (the next tile's layer-0 stores into A_hi / A_lo are ordered after this tile's last MMAs: every thread waited on their commit)
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

}  // namespace tc

bool hyponet_fused_supported(const TV& lat, const TV& out) {
  return lat.c == 32 && !lat.f16 && !out.f16 && lat.ld % 4 == 0 && (reinterpret_cast<uintptr_t>(lat.p) & 15) == 0 && lat.sn == (int64_t)lat.h * lat.w * lat.ld &&
         out.c == 2 && out.sn == (int64_t)out.h * out.w * out.ld && out.n == lat.n && out.h == lat.h && out.w == lat.w;
}

void hyponet_fused(Ctx& cx, const TV& lat, const float* coords, const void* blob, const TV& out) {
  using namespace tc;
  if (!hyponet_fused_supported(lat, out)) throw std::runtime_error("hyponet_fused: unsupported tensor layout");
  if (cx.dry) return;
  const long long P = lat.pixels();
  CUtensorMap m;
  cuuint64_t dims[2] = {32, (cuuint64_t)P};
  cuuint64_t str[1] = {(cuuint64_t)lat.ld * 4};
  cuuint32_t box[2] = {32, 128};
  encode(&m, lat.p, 2, dims, str, box);
  HypoParams p;
  p.blob = static_cast<const uint8_t*>(blob); p.coords = coords; p.out = out.p; p.P = P; p.num_tiles = (int)((P + 127) / 128); p.out_ld = out.ld;
  static int spin = -1;
  if (spin < 0) { const char* s = getenv("GIMMVFI_TC_SPIN_LIMIT"); spin = s ? atoi(s) : 400; }
  p.spin_limit = spin;
  const int smem = HN_BLOB_SMEM + HN_GROUPS * 32768 + 128 + 1024;
  static volatile unsigned char attr[64];
  gv_set_max_smem(hyponet_fused_kernel, smem, attr);
  int grid = (p.num_tiles + HN_GROUPS - 1) / HN_GROUPS;
  if (grid > cx.sm_count) grid = cx.sm_count;
  cx.launches++;
  if (cx.prof) cx.prof->begin(cx.stream, "hyponet_fused", 2.0 * (double)P * (35.0 * 128 + 3.0 * 128 * 128 + 128.0 * 2));
  hyponet_fused_kernel<<<grid, 128 * HN_GROUPS, smem, cx.stream>>>(m, p);
  gv_check_launch("hyponet_fused");
  if (cx.prof) cx.prof->end(cx.stream);
}

void hyponet_fused3(Ctx& cx, const TV& lat, const float* coords, const void* blob3, const TV& out) {
  using namespace tc;
  if (!hyponet_fused_supported(lat, out)) throw std::runtime_error("hyponet_fused3: unsupported tensor layout");
  if (cx.dry) return;
  const long long P = lat.pixels();
  HypoParams p;
  p.blob = static_cast<const uint8_t*>(blob3); p.coords = coords; p.out = out.p; p.P = P; p.num_tiles = (int)((P + 127) / 128); p.out_ld = out.ld;
  static int spin = -1;
  if (spin < 0) { const char* s = getenv("GIMMVFI_TC_SPIN_LIMIT"); spin = s ? atoi(s) : 400; }
  p.spin_limit = spin;
  const int smem = (hypo3::BLOB + 15) / 16 * 16 + 64 + 1024;
  static volatile unsigned char attr[64];
  gv_set_max_smem(hyponet_fused3_kernel, smem, attr);
  int grid = (p.num_tiles + H3_GROUPS - 1) / H3_GROUPS;
  if (grid > cx.sm_count) grid = cx.sm_count;
  cx.launches++;
  if (cx.prof) cx.prof->begin(cx.stream, "hyponet_fused3", 2.0 * (double)P * (35.0 * 128 + 3.0 * 128 * 128 + 128.0 * 2));
  hyponet_fused3_kernel<<<grid, 128 * H3_GROUPS, smem, cx.stream>>>(p, lat.p, lat.ld);
  gv_check_launch("hyponet_fused3");
  if (cx.prof) cx.prof->end(cx.stream);
}

}  // namespace gv
#endif  // GV_HOSTSIM
