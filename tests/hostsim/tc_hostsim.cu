// TEST-ONLY (tests/hostsim, -DGV_HOSTSIM): host emulation of the ARITHMETIC of the tensor-core convolution (conv_tc.cu) so that the
// engine's tensor-core orchestration — precision modes 1-3, half-precision tensors, two-segment and merged two-output layers,
// stride 2, pre-padded (reflect / x-packed) inputs, the tensor-core correlation — runs end to end on the CPU against the oracle.
// What is emulated: TF32 truncation of the activations and RN-rounded weights (plain mode), IEEE-half operands (f16 mode),
// hi + lo weight planes with exact fp32 activations (3xTF32 mode), fp32 accumulation, and the fused epilogue with its store
// rounding (RN to TF32 / half).  What is not: the accumulation ORDER and the tensor core's truncating accumulate — the device
// kernel is checked against PyTorch on the GPU (tests/test_conv_tc_gpu.py).  Never part of the product build.
#include "common.h"

#ifdef GV_HOSTSIM
namespace gv {

static inline float tf32_trunc(float x) { uint32_t u; std::memcpy(&u, &x, 4); u &= 0xffffe000u; float y; std::memcpy(&y, &u, 4); return y; }
static inline float tf32_rn_host(float x) {
  uint32_t u; std::memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return x;
  u += 0xfffu + ((u >> 13) & 1u); u &= 0xffffe000u;
  float y; std::memcpy(&y, &u, 4); return y;
}
static inline float ld_any(const TV& t, int64_t eoff) {
  return t.f16 ? gv_f16_to_f32(reinterpret_cast<const uint16_t*>(t.p)[eoff]) : t.p[eoff];
}
static inline void st_any(const TV& t, int64_t eoff, float v) {
  if (t.f16) reinterpret_cast<uint16_t*>(t.p)[eoff] = gv_f32_to_f16(v); else t.p[eoff] = v;
}

void conv2d_tc(Ctx& cx, const TV& in0, const TV& in1, const ConvW& w, const ConvGeom& g, const ConvEpi& e, const TV& out, bool split) {
  if (cx.dry) return;
  cx.launches++;
  const bool f16 = in0.f16 != 0;
  const int taps = w.kh * w.kw, cin = w.cin, cout = w.cout;
  const size_t plane = (size_t)taps * w.cout_pad * w.cin_pad;
  const uint16_t* wh = reinterpret_cast<const uint16_t*>(w.w_tc_h);
  const int64_t M = (int64_t)out.n * out.h * out.w;
#pragma omp parallel for schedule(static)
  for (int64_t m = 0; m < M; ++m) {
    const int ox = (int)(m % out.w); int64_t r = m / out.w; const int oy = (int)(r % out.h); const int n = (int)(r / out.h);
    std::vector<float> acc(cout, 0.f);
    for (int ky = 0; ky < w.kh; ++ky)
      for (int kx = 0; kx < w.kw; ++kx) {
        const int iy = oy * g.stride + ky - g.ph, ix = ox * g.stride + kx - g.pw;   // TMA box coordinate; outside the map = zero fill
        if (iy < 0 || iy >= in0.h || ix < 0 || ix >= in0.w) continue;
        const int tap = ky * w.kw + kx;
        for (int ci = 0; ci < cin; ++ci) {
          const bool seg0 = !in1.p || ci < in0.c;
          const TV& src = seg0 ? in0 : in1;
          float a = ld_any(src, src.off(n, iy, ix) + (seg0 ? ci : ci - in0.c));
          if (!f16 && !split) a = tf32_trunc(a);
          for (int co = 0; co < cout; ++co) {
            float wv;
            if (f16) wv = gv_f16_to_f32(wh[((size_t)tap * w.cout_pad + co) * w.cin_pad_h + ci]);
            else {
              const size_t o = ((size_t)tap * w.cout_pad + co) * w.cin_pad + ci;
              wv = split ? (w.w_tc[o] + w.w_tc[plane + o]) : w.w_tc[o];
            }
            acc[co] += a * wv;
          }
        }
      }
    for (int co = 0; co < cout; ++co) {
      const bool second = e.split_c > 0 && co >= e.split_c;
      const TV& O = second ? e.out2 : out;
      const int cl = co - (second ? e.split_c : 0);
      float v = acc[co] + w.b[co];
      v = apply_act(v, e.act1, e.slope1, co);
      if (e.res.p) v += ld_any(e.res, e.res.off(n, oy, ox) + co);
      v = apply_act(v, e.act2, e.slope2, co);
      if (e.mul.p && (e.split_c == 0 || second)) v *= e.mul.p[e.mul.off(n, oy, ox) + cl];
      if (e.gru_z.p) { const float z = e.gru_z.p[e.gru_z.off(n, oy, ox) + co], hh = e.gru_h.p[e.gru_h.off(n, oy, ox) + co]; v = (1.f - z) * hh + z * v; }
      if (!split && !O.f16) v = tf32_rn_host(v);          // Params::round_out
      st_any(O, O.off(n, oy, ox) + cl, v);
    }
  }
}

bool corr_volume_tc_wants_f16_planes() { return false; }   // the emulation takes the TF32 hi / lo planes (same 22-bit values)
void corr_volume_tc(Ctx& cx, const TV& fa, const float* fb_planes, const float* zero_bias, float* vol, float scale, bool split, int n_targets) {
  (void)zero_bias;
  if (cx.dry) return;
  cx.launches++;
  const int64_t Ms = (int64_t)fa.h * fa.w, N = n_targets > 0 ? n_targets : Ms; const int C = fa.c;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < Ms; ++i) {
    const float* a = fa.p + i * fa.ld;
    for (int64_t j = 0; j < N; ++j) {
      float s = 0.f;
      if (split) { const float* hi = fb_planes + j * C; const float* lo = fb_planes + N * C + j * C; for (int k = 0; k < C; ++k) s += a[k] * (hi[k] + lo[k]); }
      else { const float* b = fb_planes + j * C; for (int k = 0; k < C; ++k) s += tf32_trunc(a[k]) * tf32_trunc(b[k]); }
      vol[i * N + j] = s * scale;
    }
  }
}


// hyponet.cu's arithmetic: TF32 layer 0 on the latent + exact fp32 affine term for (t, y, x); half-precision activations and
// weights in the hidden layers; fp32 accumulation; fp32 output.  (The device kernel uses MUFU sin: |error| ~1e-6, not emulated.)
bool hyponet_fused_supported(const TV& lat, const TV& out) {
  return lat.c == 32 && !lat.f16 && !out.f16 && out.c == 2 && out.n == lat.n && out.h == lat.h && out.w == lat.w;
}
void hyponet_fused(Ctx& cx, const TV& lat, const float* coords, const void* blob_v, const TV& out) {
  if (cx.dry) return;
  cx.launches++;
  const uint8_t* blob = static_cast<const uint8_t*>(blob_v);
  const float* aff = reinterpret_cast<const float*>(blob + hypo::AFF);
  const int64_t P = lat.pixels();
#pragma omp parallel for schedule(static)
  for (int64_t pix = 0; pix < P; ++pix) {
    const int x = (int)(pix % lat.w); int64_t r = pix / lat.w; const int y = (int)(r % lat.h); const int n = (int)(r / lat.h);
    const float* lp = lat.p + lat.off(n, y, x);
    const float* c = coords + pix * 3;
    float a[128], b[128];
    for (int o = 0; o < 128; ++o) {
      float acc = 0.f;
      for (int k = 0; k < 32; ++k) { float w; std::memcpy(&w, blob + hypo::W0 + hypo::swz(o, k * 4), 4); acc += tf32_trunc(lp[k]) * w; }
      const float z = std::fmaf(c[0], aff[o], std::fmaf(c[1], aff[128 + o], std::fmaf(c[2], aff[256 + o], acc + aff[384 + o])));
      a[o] = gv_f16_to_f32(gv_f32_to_f16(sinf(z)));
    }
    for (int l = 1; l <= 3; ++l) {
      const float* bias = reinterpret_cast<const float*>(blob + hypo::B1 + (l - 1) * 512);
      for (int o = 0; o < 128; ++o) {
        float acc = 0.f;
        for (int k = 0; k < 128; ++k) {
          uint16_t h; std::memcpy(&h, blob + hypo::W1 + (l - 1) * 32768 + (k / 64) * 16384 + hypo::swz(o, (k % 64) * 2), 2);
          acc += a[k] * gv_f16_to_f32(h);
        }
        b[o] = gv_f16_to_f32(gv_f32_to_f16(sinf(acc + bias[o])));
      }
      std::memcpy(a, b, sizeof a);
    }
    const float* b4 = reinterpret_cast<const float*>(blob + hypo::B4);
    for (int o = 0; o < 2; ++o) {
      float acc = 0.f;
      for (int k = 0; k < 128; ++k) {
        uint16_t h; std::memcpy(&h, blob + hypo::W4 + (k / 64) * 2048 + hypo::swz(o, (k % 64) * 2), 2);
        acc += a[k] * gv_f16_to_f32(h);
      }
      out.p[out.off(n, y, x) + o] = acc + b4[o];
    }
  }
}


// hyponet_fused3: fp32 layers 0 and 4; layers 1-3 with the weights' and activations' fp16 hi + lo pairs (= their 22-bit values)
void hyponet_fused3(Ctx& cx, const TV& lat, const float* coords, const void* blob_v, const TV& out) {
  if (cx.dry) return;
  cx.launches++;
  const uint8_t* blob = static_cast<const uint8_t*>(blob_v);
  const float* W0 = reinterpret_cast<const float*>(blob + hypo3::W0A);
  const float* W4 = reinterpret_cast<const float*>(blob + hypo3::W4);
  const float* b4 = reinterpret_cast<const float*>(blob + hypo3::B4);
  auto pair = [](float v) { const float hi = gv_f16_to_f32(gv_f32_to_f16(v)); return hi + gv_f16_to_f32(gv_f32_to_f16(v - hi)); };
  const int64_t P = lat.pixels();
#pragma omp parallel for schedule(static)
  for (int64_t pix = 0; pix < P; ++pix) {
    const int x = (int)(pix % lat.w); int64_t r = pix / lat.w; const int y = (int)(r % lat.h); const int n = (int)(r / lat.h);
    const float* lp = lat.p + lat.off(n, y, x);
    const float* c = coords + pix * 3;
    float a[128], b[128];
    for (int o = 0; o < 128; ++o) {
      float z = W0[35 * 128 + o];
      for (int k = 0; k < 35; ++k) z = std::fmaf(k < 32 ? lp[k] : c[k - 32], W0[k * 128 + o], z);
      a[o] = pair(sinf(z));
    }
    for (int l = 0; l < 3; ++l) {
      const float* bias = reinterpret_cast<const float*>(blob + hypo3::B13 + l * 512);
      for (int o = 0; o < 128; ++o) {
        float acc = 0.f;
        for (int k = 0; k < 128; ++k) {
          uint16_t hi, lo;
          std::memcpy(&hi, blob + hypo3::W13 + ((l * 2 + 0) * 2 + k / 64) * 16384 + hypo::swz(o, (k % 64) * 2), 2);
          std::memcpy(&lo, blob + hypo3::W13 + ((l * 2 + 1) * 2 + k / 64) * 16384 + hypo::swz(o, (k % 64) * 2), 2);
          acc += a[k] * (gv_f16_to_f32(hi) + gv_f16_to_f32(lo));
        }
        const float s = sinf(acc / hypo3::W_SCALE + bias[o]);
        b[o] = l < 2 ? pair(s) : s;
      }
      std::memcpy(a, b, sizeof a);
    }
    for (int o = 0; o < 2; ++o) {
      float acc = 0.f;
      for (int k = 0; k < 128; ++k) acc = std::fmaf(a[k], W4[k * 2 + o], acc);
      out.p[out.off(n, y, x) + o] = acc + b4[o];
    }
  }
}

}  // namespace gv
#endif  // GV_HOSTSIM
