// 3x3 convolution for K-POOR layers (cin <= 32 fp32 / <= 64 half, cout <= 64) on tcgen05: the activation tile is loaded ONCE with
// its halo and every tap is an MMA on a SHIFTED VIEW of that tile.
//
// Why: conv_tc.cu walks K as (tap, 32-channel block) and fetches one TMA box per K step - 9 boxes of 16 KB for a 3x3 layer whose
// distinct input is 23 KB.  For the 32 -> 32 full-resolution layers (9 per frame at 1088x1920) that is 4.8 GB of L2 -> SM traffic
// per launch and the measured limiter (0.80 ms = 6 TB/s; 96 TFLOP/s; HBM floor 0.17 ms).
//
// How: output tile = 8 (x) x 16 (y) pixels, so the 8-row groups of the M = 128 operand are image rows.  ONE TMA box
// {32 ch, 16 x, 18 y} lands the tile plus halo 128B-swizzled in shared memory with a row pitch of P = 16 pixels (2048 bytes: the
// stride between 8-row groups must be a multiple of the 1024-byte swizzle pattern - a 10-pixel pitch measured wrong); tap (ky, kx)
// is the UMMA descriptor whose start address is advanced by (ky * P + kx) pixel rows, stride between 8-row groups P * 128 bytes.  The 128B swizzle is a function of the ABSOLUTE shared-memory address bits on both
// the TMA write and the MMA read (measured: profiles/r02_halo_diag.log - every tap exact with the descriptor's "matrix base offset"
// field left 0, garbage when it is set to the start row's phase), so a row-shifted start needs no further treatment.
// Out-of-bounds box elements are zero-filled = the convolution's zero padding.  The 9 weight tiles (<= 72 KB) stay resident in shared memory for the whole
// persistent CTA.  Per tile: 36 KB of TMA traffic instead of 144 KB (+ weights), 36 MMAs, no other shared-memory writes.
//
// Roles: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM alloc), 8 epilogue warps (two per TMEM lane quarter, each BN/2 columns;
// a thread stores its pixel's 16 / 32 channels straight from registers).
#include "common.h"

#ifndef GV_HOSTSIM
#include <cuda.h>
#include <cuda_fp16.h>

namespace gv {
namespace tc {
#include "tc_ptx.cuh"
#include "tc_epilogue.cuh"

constexpr int HL_TW = 8, HL_TH = 16;                  // output tile (x, y)
constexpr int HL_BW = 16, HL_BH = HL_TH + 2;          // halo box: 16 x 18 pixels (x extent padded to 16: 2048-byte row pitch)
constexpr int HL_HALO_BYTES = HL_BW * HL_BH * 128;    // 36 KB
constexpr int HL_MAX_STAGES = 4;

struct HaloParams {
  int tiles_x, tiles_y, n_img, H, W, cout, BN, stages, origin, f16_in, round_out, spin_limit, dbg;
  int copies3;      // 1: three x-shifted copies {K, 8 x, 18 y} of the tile (aligned descriptors only); 0: one {K, P x, 18 y} halo tile + shifted views
  int pitch;        // P: pixels per halo row (16 or 10)
  int halo_bytes;   // bytes of one pipeline stage (multiple of 1024)
  int kinstr;       // MMAs per tap: ceil(cin bytes / 32) <= 4
  int ntaps;        // 9 (3x3) or 1 (1x1: the "halo" box is the 8 x 16 tile itself, pitch 8)
  const float* bias; int act1; const float* slope1; int act2; const float* slope2;
  TV res, out;
};

__device__ __forceinline__ uint64_t make_desc_sbo(uint32_t saddr, uint32_t sbo_bytes, int dbg) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  if (dbg & 1) d |= (uint64_t)((saddr >> 7) & 7) << 49;   // experiment only: "matrix base offset" = start row phase -> WRONG results on sm_100a
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(320, 1) conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                                                             const __grid_constant__ HaloParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int w_bytes = p.ntaps * p.BN * 128;
  const int w_region = (w_bytes + 1023) & ~1023;
  uint8_t* wsm = smem;
  uint8_t* halo = smem + w_region;
  uint64_t* bars = reinterpret_cast<uint64_t*>(halo + p.stages * p.halo_bytes);
  uint64_t* w_bar = bars;                         // weights landed
  uint64_t* full_bar = bars + 1;                  // [HL_MAX_STAGES]
  uint64_t* empty_bar = bars + 1 + HL_MAX_STAGES; // [HL_MAX_STAGES]
  uint64_t* tfull_bar = bars + 1 + 2 * HL_MAX_STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                 // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.n_img * p.tiles_y * p.tiles_x;
  const int SPIN = p.spin_limit;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(w_bar, 1);
      for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 8); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(w_bar, (uint32_t)w_bytes);
      tma_load_3d(wsm, &tmW, w_bar, 0, 0, 0);   // {K block, BN rows, 9 taps}
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int r = tile; const int tx = r % p.tiles_x; r /= p.tiles_x; const int ty = r % p.tiles_y; const int n = r / p.tiles_y;
        mbar_wait(&empty_bar[stage], phase ^ 1, SPIN);
        mbar_expect_tx(&full_bar[stage], (uint32_t)p.halo_bytes);
        if (p.copies3) {   // copy kx = the 8-pixel-wide column of the tile shifted by kx: every tap's operand starts 1024-byte aligned
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
            tma_load_4d(halo + stage * p.halo_bytes + kx * (HL_TW * HL_BH * 128), &tmA, &full_bar[stage], 0, tx * HL_TW + p.origin + kx, ty * HL_TH + p.origin, n);
        } else {
          tma_load_4d(halo + stage * p.halo_bytes, &tmA, &full_bar[stage], 0, tx * HL_TW + p.origin, ty * HL_TH + p.origin, n);
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t fmt = p.f16_in ? 0u : 2u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    mbar_wait(w_bar, 0, SPIN);
    int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1, SPIN);
      mbar_wait(&full_bar[stage], phase, SPIN);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 64);
        const uint32_t h_s = smem_u32(halo + stage * p.halo_bytes), w_s = smem_u32(wsm);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          if (tap >= p.ntaps) break;
          const int ky = tap / 3, kx = tap % 3;
          const uint64_t ad = p.copies3 ? make_smem_desc(h_s + (uint32_t)(kx * (HL_TW * HL_BH * 128) + ky * HL_TW * 128))          // copy kx, image row ky
                                        : make_desc_sbo(h_s + (uint32_t)((ky * p.pitch + kx) * 128), (uint32_t)(p.pitch * 128), p.dbg);   // shifted view of the halo tile
          const uint64_t bd = make_smem_desc(w_s + (uint32_t)(tap * p.BN * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) {   // 32 bytes of K per instruction (8 tf32 / 16 half); channels beyond cin are zero fill: not issued
            if (k >= p.kinstr) break;
            if (p.f16_in) mma_f16(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (tap | k) ? 1u : 0u);
            else mma_tf32(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (tap | k) ? 1u : 0u);
          }
        }
        mma_commit(&empty_bar[stage]);
        mma_commit(&tfull_bar[acc]);
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ---- epilogue: warp -> TMEM lane quarter (warp % 4) = image rows [4 q, 4 q + 4) of the tile; the two warps of a quarter split the
    // columns.  A thread owns ONE output pixel (TMEM lane = tile row m = y * 8 + x) and 16 consecutive channels = 64 contiguous bytes of
    // the NHWC output: bias / activation / residual / rounding and four 16-byte stores straight from registers.  (The first version
    // transposed through shared memory like conv_tc.cu's epilogue: 555 instructions per warp and tile, the kernel's limiter per ncu -
    // tensor pipe 20 % active; a pixel's other 16-channel halves are written by the partner warp, so every 32-byte sector is still
    // written in full.)
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const int cpw = p.BN / 2;                                  // columns per warp: 16 or 32
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int r = tile; const int tx = r % p.tiles_x; r /= p.tiles_x; const int ty = r % p.tiles_y; const int n = r / p.tiles_y;
      const int y = ty * HL_TH + quarter * 4 + (lane >> 3), x = tx * HL_TW + (lane & 7);
      const bool inside = y < p.H && x < p.W;
      mbar_wait(&tfull_bar[acc], acc_phase, SPIN);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 64 + half * cpw);
      for (int u = 0; u < cpw / 16; ++u) {
        const int cbase = half * cpw + u * 16;
        uint32_t v[16];
        tmem_ld16(taddr + (uint32_t)(u * 16), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (cbase < p.cout && inside) {
          float o[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + cbase) + j);
            o[4 * j] = __uint_as_float(v[4 * j]) + b4.x; o[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b4.y;
            o[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b4.z; o[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b4.w;
          }
          act_n<16>(o, p.act1, p.slope1, cbase, p.cout);
          if (p.res.p) {
            const int64_t rb = p.res.off(n, y, x) + cbase;
#pragma unroll
            for (int j = 0; j < 4; ++j) { float t[4]; load4_any(p.res, rb + 4 * j, cbase + 4 * j, p.cout, t); o[4 * j] += t[0]; o[4 * j + 1] += t[1]; o[4 * j + 2] += t[2]; o[4 * j + 3] += t[3]; }
          }
          if (p.act2 != ACT_NONE) act_n<16>(o, p.act2, p.slope2, cbase, p.cout);
          if (p.round_out) {
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = rn_tf32(o[j]);
          }
          const int64_t eoff = p.out.off(n, y, x) + cbase;
          const uintptr_t oaddr = reinterpret_cast<uintptr_t>(p.out.p) + (uintptr_t)eoff * (p.out.f16 ? 2 : 4);
          const bool al = (oaddr & (p.out.f16 ? 7 : 15)) == 0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = cbase + 4 * j;
            if (c < p.cout) store4(p.out, eoff + 4 * j, o + 4 * j, c, p.cout, al && c + 3 < p.cout);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
}

}  // namespace tc

static int halo_mode() {   // GIMMVFI_HALO_MODE: 1 (default) = one 16-wide halo tile (36 KB) + row-shifted descriptor starts;
                           // 3 = three x-shifted 8-wide copies (54 KB, 1024-aligned starts only).  (A 10-wide halo tile with a 1280-byte
                           // group stride does NOT work: measured wrong - the stride between 8-row groups has to be a multiple of 1024.)
  static int v = -1;
  if (v < 0) { const char* s = getenv("GIMMVFI_HALO_MODE"); v = s ? atoi(s) : 1; }
  return v;
}
static int tc_halo() {   // GIMMVFI_TC_HALO=0: K-poor 3x3 layers stay on the per-tap TMA path of conv_tc.cu
  static int v = -1;
  if (v < 0) { const char* s = getenv("GIMMVFI_TC_HALO"); v = s ? atoi(s) : 1; }
  return v;
}

bool conv2d_halo_supported(const TV& in0, const TV& in1, const ConvW& w, const ConvGeom& g, const ConvEpi& e, const TV& out) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!tc_halo() || in1.p || !((w.kh == 3 && w.kw == 3) || (w.kh == 1 && w.kw == 1)) || g.stride != 1 || g.reflect || w.cout > 64) return false;
  if (e.mul.p || e.gru_z.p || e.out2.p || e.split_c) return false;
  const bool h = in0.f16 != 0;
  if (h ? (!w.w_tc_h || w.cin_pad_h != 64 || in0.c > 64) : (!w.w_tc || w.cin_pad != 32 || in0.c > 32)) return false;
  if (!al16(in0.p) || in0.ld % (h ? 8 : 4) || in0.sn % (h ? 8 : 4)) return false;
  const int pad = w.kh / 2;
  if (g.loose_w ? (g.ph != 0 || g.pw != 0 || in0.h < out.h + 2 * pad || in0.w < out.w + 2 * pad || pad == 0) : (g.ph != pad || g.pw != pad || in0.h != out.h || in0.w != out.w)) return false;
  if (out.ld % 4 || (e.res.p && e.res.ld % 4)) return false;
  return in0.n == out.n;
}

void conv2d_halo(Ctx& cx, const TV& in0, const ConvW& w, const ConvGeom& g, const ConvEpi& e, const TV& out) {
  using namespace tc;
  const bool f16 = in0.f16 != 0;
  const bool one = w.kh == 1;                          // 1x1: no halo, one tap
  const int ntaps = one ? 1 : 9;
  const int c16 = (w.cout + 15) & ~15;
  const int BN = c16 <= 32 ? 32 : 64;                 // two epilogue warps per lane quarter split BN into 16- or 32-column halves
  if (w.cout_pad < c16) throw std::runtime_error("conv2d_halo: weight padding");
  CUtensorMap mA, mW;
  {
    const cuuint64_t es = f16 ? 2 : 4;
    cuuint64_t dims[4] = {(cuuint64_t)in0.c, (cuuint64_t)in0.w, (cuuint64_t)in0.h, (cuuint64_t)in0.n};
    cuuint64_t str[3] = {(cuuint64_t)in0.ld * es, (cuuint64_t)in0.w * in0.ld * es, (cuuint64_t)in0.sn * es};
    cuuint32_t box[4] = {(cuuint32_t)(f16 ? 64 : 32), (cuuint32_t)(one ? HL_TW : (halo_mode() == 3 ? HL_TW : HL_BW)), (cuuint32_t)(one ? HL_TH : HL_BH), 1};
    encode(&mA, in0.p, 4, dims, str, box, f16);
  }
  {   // weights [tap][cout_pad][K block]: rows beyond cout_pad are zero-filled by TMA
    const int kpad = f16 ? w.cin_pad_h : w.cin_pad;
    const cuuint64_t es = f16 ? 2 : 4;
    cuuint64_t dims[3] = {(cuuint64_t)kpad, (cuuint64_t)w.cout_pad, (cuuint64_t)ntaps};
    cuuint64_t str[2] = {(cuuint64_t)kpad * es, (cuuint64_t)kpad * w.cout_pad * es};
    cuuint32_t box[3] = {(cuuint32_t)(f16 ? 64 : 32), (cuuint32_t)BN, (cuuint32_t)ntaps};
    encode(&mW, f16 ? w.w_tc_h : static_cast<const void*>(w.w_tc), 3, dims, str, box, f16);
  }
  HaloParams p;
  p.tiles_x = (out.w + HL_TW - 1) / HL_TW; p.tiles_y = (out.h + HL_TH - 1) / HL_TH; p.n_img = out.n; p.H = out.h; p.W = out.w; p.cout = w.cout; p.BN = BN;
  p.origin = (g.loose_w || one) ? 0 : -1; p.f16_in = f16 ? 1 : 0; p.round_out = out.f16 ? 0 : 1;
  static int spin = -1;
  if (spin < 0) { const char* s = getenv("GIMMVFI_TC_SPIN_LIMIT"); spin = s ? atoi(s) : 400; }
  p.spin_limit = spin;
  { const char* d = getenv("GIMMVFI_HALO_DBG"); p.dbg = d ? atoi(d) : 0; }
  p.bias = w.b; p.act1 = e.act1; p.slope1 = e.slope1; p.act2 = e.act2; p.slope2 = e.slope2; p.res = e.res; p.out = out;
  p.copies3 = (!one && halo_mode() == 3) ? 1 : 0;
  p.ntaps = ntaps;
  p.kinstr = (in0.c * (f16 ? 2 : 4) + 31) / 32;
  p.pitch = one ? HL_TW : HL_BW;
  p.halo_bytes = one ? HL_TW * HL_TH * 128 : (p.copies3 ? 3 * HL_TW * HL_BH * 128 : ((p.pitch * HL_BH * 128 + 1023) & ~1023));
  const int w_region = (ntaps * BN * 128 + 1023) & ~1023;
  const int fixed = w_region + 256 + 1024;
  p.stages = (227 * 1024 - fixed) / p.halo_bytes;
  if (p.stages > HL_MAX_STAGES) p.stages = HL_MAX_STAGES;
  if (p.stages < 2) throw std::runtime_error("conv2d_halo: not enough shared memory");
  const int smem = fixed + p.stages * p.halo_bytes;
  static volatile unsigned char attr[64];
  gv_set_max_smem(conv3x3_halo_kernel, 227 * 1024, attr);   // set once per device: the layer-dependent size below must always fit
  const int num_tiles = p.n_img * p.tiles_y * p.tiles_x;
  const int grid = num_tiles < cx.sm_count ? num_tiles : cx.sm_count;
  cx.launches++;
  if (cx.prof) {
    char nm[128];
    snprintf(nm, sizeof nm, "conv2d_halo_%s k%dx%d c%d>%d @%dx%dx%d", f16 ? "f16" : "tf32", w.kh, w.kw, w.cin, w.cout, out.n, out.h, out.w);
    cx.prof->begin(cx.stream, prof_intern(nm), 2.0 * (double)out.n * out.h * out.w * w.cout * (double)w.cin * ntaps);
  }
  conv3x3_halo_kernel<<<grid, 320, smem, cx.stream>>>(mA, mW, p);
  gv_check_launch("conv2d_halo");
  if (cx.prof) cx.prof->end(cx.stream);
}

}  // namespace gv
#endif  // GV_HOSTSIM
