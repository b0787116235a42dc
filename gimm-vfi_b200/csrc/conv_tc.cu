// Tensor-core implicit-GEMM convolution for sm_100a: TMA tile loads with halo
// coordinates -> 128B-swizzled shared memory -> tcgen05.mma (kind::tf32, fp32
// accumulate in TMEM) -> fused epilogue from TMEM.
//
//   GEMM view   M = 128 output pixels (an 8 x 16 spatial tile of one image)
//               N = cout (padded to 16, <= 256, one tile)
//               K = taps x cin, walked as (tap, 32-channel block): each K step is ONE
//                   TMA box {32 ch, 16 x, 8 y, 1 n} of the NHWC activation at the
//                   tap-shifted coordinate (out-of-bounds -> zero fill = zero padding,
//                   no im2col, no halo staging code) plus one {32 k, N} weight box.
//   roles       warp 0: TMA producer | warp 1: MMA issuer (+TMEM alloc) | warps 2-5: epilogue
//   pipelines   smem full/empty ring (4 stages), 2 TMEM accumulators (full/empty) so the
//               epilogue of tile i overlaps the MMAs of tile i+1; persistent over tiles.
//
// Used for the post-RAFT networks (DESIGN.md "precision plan": TF32 operands there
// move imgt_pred by < 3e-4; the RAFT recurrence stays on the fp32 path in conv.cu).
#include "common.h"

#ifndef GV_HOSTSIM
#include <cuda.h>

namespace gv {

namespace tc {

constexpr int TILE_H = 8, TILE_W = 16, BM = 128, BK = 32, MAX_STAGES = 8;
constexpr int STG_PITCH = 36;                       // floats per staged row (32 + 4: keeps float4 alignment)
constexpr int STG_BYTES = 4 * 32 * STG_PITCH * 4;   // 4 epilogue warps x 32 rows
constexpr int A_BYTES = BM * BK * 4;  // 16 KB
constexpr int NUM_THREADS = 192;

struct Params {
  int taps, kw, ph, pw;
  int kblocks, c0_blocks;          // 32-channel K blocks in total / from segment 0
  int tiles_x, tiles_y, n_img;
  int H, W;
  int BN;                          // MMA N (multiple of 16)
  int stages;                      // smem ring depth (<= MAX_STAGES), sized from BN on the host
  int cout;
  const float* bias;
  int act1; const float* slope1;
  int act2; const float* slope2;
  TV res, out;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  // try_wait suspends the thread up to ~10 ms per attempt; a pipeline that makes no progress for
  // ~4 s is a bug -> trap (turns a would-be hang into a launch failure the host reports).
  for (int spin = 0; spin < 400; ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(ok)
        : "r"(addr), "r"(parity), "r"(0x989680u)
        : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
// K-major, SWIZZLE_128B operand tile: rows of 128 B, 8-row atoms of 1024 B (SBO), version 1 (sm_100).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address  [0,14)
  d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major) [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset: 8 rows x 128 B [32,46)
  d |= (uint64_t)1 << 46;                            // descriptor version [46,48)
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B [61,64)
  return d;
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// Rare activations (sigmoid / tanh / sin) go through ONE out-of-line copy: inlining the accurate sinf/tanhf/expf
// paths at every element site made the epilogue ~25k SASS instructions (instruction-cache bound, ~10 us per
// 32-column chunk measured).
__device__ __noinline__ float act_slow(float v, int act) {
  switch (act) {
    case ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case ACT_TANH: return tanhf(v);
    case ACT_SIN: return sinf(v);
    default: return v;
  }
}
// 4 consecutive channels starting at c (c % 4 == 0; slope padded like the bias is not guaranteed -> bounds via cout)
__device__ __forceinline__ void act4(float* o, int act, const float* slope, int c, int cout) {
  if (act == ACT_NONE) return;
  if (act == ACT_RELU) {
#pragma unroll
    for (int u = 0; u < 4; ++u) o[u] = fmaxf(o[u], 0.f);
  } else if (act == ACT_LRELU) {
#pragma unroll
    for (int u = 0; u < 4; ++u) o[u] = o[u] > 0.f ? o[u] : 0.1f * o[u];
  } else if (act == ACT_PRELU) {
#pragma unroll
    for (int u = 0; u < 4; ++u) { const float sl = (c + u < cout) ? slope[c + u] : 0.f; o[u] = o[u] > 0.f ? o[u] : sl * o[u]; }
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u) o[u] = act_slow(o[u], act);
  }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv2d_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB,
                 const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stage A (16 KB) | stage B (BN*128 B)] x STAGES, then barriers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_bytes = p.BN * BK * 4;
  const int stage_bytes = A_BYTES + b_bytes;
  const int STAGES = p.stages;
  float* stg_base = reinterpret_cast<float*>(smem + STAGES * stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes + STG_BYTES);
  uint64_t* full_bar = bars;                          // [MAX_STAGES]
  uint64_t* empty_bar = bars + MAX_STAGES;            // [MAX_STAGES]
  uint64_t* tfull_bar = bars + 2 * MAX_STAGES;        // [2]
  uint64_t* tempty_bar = bars + 2 * MAX_STAGES + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.n_img * p.tiles_y * p.tiles_x;
  const int ksteps = p.taps * p.kblocks;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA0)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA1)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    // 512 columns: two 256-column fp32 accumulators (1 CTA / SM, so the whole TMEM is ours)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int tx = tile % p.tiles_x; const int r = tile / p.tiles_x; const int ty = r % p.tiles_y; const int n = r / p.tiles_y;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int ky = tap / p.kw, kx = tap % p.kw;
          const int x0 = tx * TILE_W + kx - p.pw, y0 = ty * TILE_H + ky - p.ph;
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* a_dst = smem + stage * stage_bytes;
            uint8_t* b_dst = a_dst + A_BYTES;
            mbar_expect_tx(&full_bar[stage], (uint32_t)(A_BYTES + b_bytes));
            if (kb < p.c0_blocks) tma_load_4d(a_dst, &tmA0, &full_bar[stage], kb * BK, x0, y0, n);
            else tma_load_4d(a_dst, &tmA1, &full_bar[stage], (kb - p.c0_blocks) * BK, x0, y0, n);
            tma_load_3d(b_dst, &tmB, &full_bar[stage], kb * BK, 0, tap);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    // instruction descriptor: D=f32, A=B=tf32, both K-major, N>>3, M>>4
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(&full_bar[stage], phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
          const uint64_t adesc = make_smem_desc(a_addr);
          const uint64_t bdesc = make_smem_desc(a_addr + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k)  // UMMA_K = 8 tf32 = 32 B -> +2 in the (addr >> 4) field
            mma_tf32(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (ks > 0 || k > 0) ? 1u : 0u);
          mma_commit(&empty_bar[stage]);                       // frees the smem slot when these MMAs retire
          if (ks == ksteps - 1) mma_commit(&tfull_bar[acc]);   // accumulator complete
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================================================== epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1)
    // TMEM gives each thread one pixel (row) x 32 consecutive channels.  Writing that straight to NHWC
    // scatters 16-byte pieces over 32 different lines per store (measured: ~230 GB/s chip-wide), so the
    // 32x32 chunk is transposed through shared memory and written as full 128-byte rows.
    const int quarter = warp & 3;
    float* stg = stg_base + (warp - 2) * 32 * STG_PITCH;
    const int q8 = lane & 7, rsub = lane >> 3;      // write phase: 8 lanes x float4 = one 128-byte row, 4 rows / instruction
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int tx = tile % p.tiles_x; const int r = tile / p.tiles_x; const int ty = r % p.tiles_y; const int n = r / p.tiles_y;
      // rows this lane writes in the coalesced phase: rr = quarter*32 + it*4 + rsub
      int64_t ooff[8], roff[8]; uint32_t vmask = 0;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int rr = quarter * 32 + it * 4 + rsub;
        const int y = ty * TILE_H + rr / TILE_W, x = tx * TILE_W + rr % TILE_W;
        const bool ok = y < p.H && x < p.W;
        vmask |= (ok ? 1u : 0u) << it;
        ooff[it] = ok ? p.out.off(n, y, x) : 0;
        roff[it] = (ok && p.res.p) ? p.res.off(n, y, x) : 0;
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 256);
      for (int c0 = 0; c0 < p.BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)c0, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        // phase 1 (thread = pixel row): bias + act1, stage to smem.  bias is padded to BN on the host.
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(p.bias + c0 + j);
          float o[4] = {__uint_as_float(v[j]) + b4.x, __uint_as_float(v[j + 1]) + b4.y, __uint_as_float(v[j + 2]) + b4.z,
                        __uint_as_float(v[j + 3]) + b4.w};
          act4(o, p.act1, p.slope1, c0 + j, p.cout);
          *reinterpret_cast<float4*>(stg + lane * STG_PITCH + j) = make_float4(o[0], o[1], o[2], o[3]);
        }
        __syncwarp();
        // phase 2 (8 lanes = one 128-byte row): + residual, act2, TF32 round-to-nearest, coalesced store
        const int c = c0 + q8 * 4;
        if (c < p.cout) {
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            if (!((vmask >> it) & 1u)) continue;
            const float4 sv = *reinterpret_cast<const float4*>(stg + (it * 4 + rsub) * STG_PITCH + q8 * 4);
            float o[4] = {sv.x, sv.y, sv.z, sv.w};
            float* optr = p.out.p + ooff[it] + c;
            const float* rptr = p.res.p ? p.res.p + roff[it] + c : nullptr;
            const bool full4 = c + 3 < p.cout;
            if (rptr) {
              if (full4 && ((reinterpret_cast<uintptr_t>(rptr) & 15) == 0)) {
                const float4 rv = *reinterpret_cast<const float4*>(rptr);
                o[0] += rv.x; o[1] += rv.y; o[2] += rv.z; o[3] += rv.w;
              } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) if (c + u < p.cout) o[u] += rptr[u];
              }
            }
            act4(o, p.act2, p.slope2, c, p.cout);
            // store TF32-representable values (round-to-nearest-even): the next tensor-core layer then
            // truncates nothing, i.e. its operands are RN- instead of toward-zero-rounded (unbiased)
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              uint32_t bits = __float_as_uint(o[u]);
              bits += 0xfffu + ((bits >> 13) & 1u);
              o[u] = __uint_as_float(bits & 0xffffe000u);
            }
            if (full4 && ((reinterpret_cast<uintptr_t>(optr) & 15) == 0)) {
              *reinterpret_cast<float4*>(optr) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
              for (int u = 0; u < 4; ++u) if (c + u < p.cout) optr[u] = o[u];
            }
          }
        }
        __syncwarp();
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
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

static void encode(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("conv_tc: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
}

static void encode_act(CUtensorMap* m, const TV& t) {
  cuuint64_t dims[4] = {(cuuint64_t)t.c, (cuuint64_t)t.w, (cuuint64_t)t.h, (cuuint64_t)t.n};
  cuuint64_t str[3] = {(cuuint64_t)t.ld * 4, (cuuint64_t)t.w * t.ld * 4, (cuuint64_t)t.sn * 4};
  cuuint32_t box[4] = {BK, TILE_W, TILE_H, 1};
  encode(m, t.p, 4, dims, str, box);
}

}  // namespace tc

bool conv2d_tc_supported(const TV& in0, const TV& in1, const ConvW& w, const ConvGeom& g, const ConvEpi& e, const TV& out) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!w.w_tc || g.stride != 1 || g.reflect) return false;
  if (e.mul.p || e.gru_z.p) return false;
  if (w.cout_pad > 256) return false;
  if (g.ph != w.kh / 2 || g.pw != w.kw / 2) return false;
  if (!al16(in0.p) || in0.ld % 4 || in0.sn % 4) return false;
  if (in1.p && (!al16(in1.p) || in1.ld % 4 || in1.sn % 4 || in0.c % 32)) return false;
  if (in0.h != out.h || in0.w != out.w) return false;
  return true;
}

void conv2d_tc(Ctx& cx, const TV& in0, const TV& in1, const ConvW& w, const ConvGeom& g, const ConvEpi& e, const TV& out) {
  using namespace tc;
  CUtensorMap mA0, mA1, mB;
  encode_act(&mA0, in0);
  if (in1.p) encode_act(&mA1, in1); else mA1 = mA0;
  {
    cuuint64_t dims[3] = {(cuuint64_t)w.cin_pad, (cuuint64_t)w.cout_pad, (cuuint64_t)(w.kh * w.kw)};
    cuuint64_t str[2] = {(cuuint64_t)w.cin_pad * 4, (cuuint64_t)w.cin_pad * w.cout_pad * 4};
    cuuint32_t box[3] = {BK, (cuuint32_t)w.cout_pad, 1};
    encode(&mB, w.w_tc, 3, dims, str, box);
  }
  Params p;
  p.taps = w.kh * w.kw; p.kw = w.kw; p.ph = g.ph; p.pw = g.pw;
  p.c0_blocks = in1.p ? in0.c / 32 : (in0.c + 31) / 32;
  p.kblocks = w.cin_pad / 32;
  p.tiles_x = (out.w + TILE_W - 1) / TILE_W; p.tiles_y = (out.h + TILE_H - 1) / TILE_H; p.n_img = out.n;
  p.H = out.h; p.W = out.w; p.BN = w.cout_pad; p.cout = w.cout;
  p.bias = w.b; p.act1 = e.act1; p.slope1 = e.slope1; p.act2 = e.act2; p.slope2 = e.slope2; p.res = e.res; p.out = out;
  const int num_tiles = p.n_img * p.tiles_y * p.tiles_x;
  const int stage_bytes = A_BYTES + p.BN * BK * 4;
  const int budget = 227 * 1024 - 1024 /*align*/ - STG_BYTES - 512 /*barriers*/;
  p.stages = budget / stage_bytes;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  if (p.stages < 2) throw std::runtime_error("conv_tc: not enough shared memory for 2 pipeline stages");
  const int smem = p.stages * stage_bytes + STG_BYTES + 512 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t er = cudaFuncSetAttribute(conv2d_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (er != cudaSuccess) throw std::runtime_error(std::string("conv_tc: cudaFuncSetAttribute: ") + cudaGetErrorString(er));
    attr_set = true;
  }
  const int grid = num_tiles < cx.sm_count ? num_tiles : cx.sm_count;
  cx.launches++;
  if (cx.prof) {
    char nm[128];
    snprintf(nm, sizeof nm, "conv2d_tc_tf32 k%dx%d c%d>%d @%dx%dx%d", w.kh, w.kw, w.cin, w.cout, out.n, out.h, out.w);
    cx.prof->begin(cx.stream, prof_intern(nm), 2.0 * (double)out.n * out.h * out.w * w.cout * (double)w.cin * w.kh * w.kw);
  }
  conv2d_tc_kernel<<<grid, NUM_THREADS, smem, cx.stream>>>(mA0, mA1, mB, p);
  gv_check_launch("conv2d_tc");
  if (cx.prof) cx.prof->end(cx.stream);
}

}  // namespace gv
#endif  // GV_HOSTSIM
