// Host orchestration of the GIMM-VFI-R per-pair interpolation path.
// One stream, one pre-planned workspace, zero allocations / host syncs in
// forward().  Reference call stack: gimmvfi_r.py:324-407.
#include "engine.h"
#include "net.h"

#include <algorithm>

namespace gv {

// ---------------------------------------------------------------------------
// device memory helpers
// ---------------------------------------------------------------------------
#ifdef GV_HOSTSIM
void gv_check_launch(const char*) {}
void* dev_alloc(size_t bytes) { return std::malloc(bytes ? bytes : 1); }
void dev_free(void* p) { std::free(p); }
void dev_upload(void* dst, const void* src, size_t bytes) { std::memcpy(dst, src, bytes); }
void dev_download(void* dst, const void* src, size_t bytes, gvStream_t) { std::memcpy(dst, src, bytes); }
void dev_memset(void* dst, int v, size_t bytes, gvStream_t) { std::memset(dst, v, bytes); }
void dev_copy(void* dst, const void* src, size_t bytes, gvStream_t) { std::memcpy(dst, src, bytes); }
void dev_sync(gvStream_t) {}
#else
static void cuda_ok(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e));
}
void gv_check_launch(const char* what) { cuda_ok(cudaGetLastError(), what); }
void* dev_alloc(size_t bytes) { void* p = nullptr; cuda_ok(cudaMalloc(&p, bytes ? bytes : 1), "cudaMalloc"); return p; }
void dev_free(void* p) { cudaFree(p); }
void dev_upload(void* dst, const void* src, size_t bytes) { cuda_ok(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice), "cudaMemcpy H2D"); }
void dev_download(void* dst, const void* src, size_t bytes, gvStream_t s) {
  cuda_ok(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s), "cudaMemcpy D2H");
  cuda_ok(cudaStreamSynchronize(s), "sync");
}
void dev_memset(void* dst, int v, size_t bytes, gvStream_t s) { cuda_ok(cudaMemsetAsync(dst, v, bytes, s), "cudaMemsetAsync"); }
void dev_copy(void* dst, const void* src, size_t bytes, gvStream_t s) { cuda_ok(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s), "cudaMemcpy D2D"); }
void dev_sync(gvStream_t s) { cuda_ok(cudaStreamSynchronize(s), "cudaStreamSynchronize"); }
#endif

// RAII: make the engine's device current for the duration of a call and restore the caller's (an engine on cuda:1 must
// not change torch's current device, and its uploads / launches must not land on whatever device happens to be current)
struct DeviceGuard {
#ifndef GV_HOSTSIM
  int prev = -1; bool switched = false;
  explicit DeviceGuard(int dev) {
    cuda_ok(cudaGetDevice(&prev), "cudaGetDevice");
    if (prev != dev) { cuda_ok(cudaSetDevice(dev), "cudaSetDevice"); switched = true; }
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
#else
  explicit DeviceGuard(int) {}
#endif
  DeviceGuard(const DeviceGuard&) = delete; DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// ---------------------------------------------------------------------------
// Profiler (CUDA events around every launch; off by default)
// ---------------------------------------------------------------------------
const char* prof_intern(const std::string& s) {
  static std::map<std::string, std::string> pool;
  auto it = pool.find(s);
  if (it == pool.end()) it = pool.emplace(s, s).first;
  return it->second.c_str();
}
#ifdef GV_HOSTSIM
void* Profiler::get_event() { return nullptr; }
void Profiler::begin(gvStream_t, const char*, double) {}
void Profiler::end(gvStream_t) {}
Profiler::~Profiler() {}
#else
void* Profiler::get_event() {
  if (used == pool.size()) { cudaEvent_t e; cuda_ok(cudaEventCreate(&e), "cudaEventCreate"); pool.push_back(e); }
  return pool[used++];
}
void Profiler::begin(gvStream_t s, const char* name, double work) {
  ProfRec r; r.name = name; r.work = work; r.ev0 = get_event(); r.ev1 = get_event();
  cuda_ok(cudaEventRecord((cudaEvent_t)r.ev0, s), "cudaEventRecord");
  recs.push_back(r);
}
void Profiler::end(gvStream_t s) { cuda_ok(cudaEventRecord((cudaEvent_t)recs.back().ev1, s), "cudaEventRecord"); }
Profiler::~Profiler() { for (void* e : pool) cudaEventDestroy((cudaEvent_t)e); }
#endif

std::string Engine::profile_json(gvStream_t stream) {
  std::string out = "{";
  DeviceGuard dg(device_);
#ifndef GV_HOSTSIM
  cuda_ok(cudaStreamSynchronize(stream), "sync");
  struct Acc { double ms = 0, work = 0; int64_t n = 0; };
  std::map<std::string, Acc> acc;
  for (const ProfRec& r : prof_.recs) {
    float ms = 0.f;
    cuda_ok(cudaEventElapsedTime(&ms, (cudaEvent_t)r.ev0, (cudaEvent_t)r.ev1), "cudaEventElapsedTime");
    Acc& a = acc[r.name]; a.ms += ms; a.work += r.work; a.n++;
  }
  bool first = true;
  for (auto& kv : acc) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s\"%s\": {\"ms\": %.6f, \"work\": %.6e, \"launches\": %lld}", first ? "" : ", ", kv.first.c_str(), kv.second.ms,
             kv.second.work, (long long)kv.second.n);
    out += buf; first = false;
  }
#else
  (void)stream;
#endif
  return out + "}";
}

// ---------------------------------------------------------------------------
// Engine: weights
// ---------------------------------------------------------------------------
Engine::Engine(int device) : device_(device) {
  if (const char* kn = getenv("GIMMVFI_PRECISE")) precise_ = atoi(kn);   // builder experiments: override the 3xTF32 stage mask
  if (const char* kn = getenv("GIMMVFI_HYPO_FAST")) hypo_fast_ = atoi(kn) != 0;
  if (const char* kn = getenv("GIMMVFI_CORR_DIRECT")) corr_direct_max_t_ = atoi(kn);
  if (const char* kn = getenv("GIMMVFI_RAFT_CORR_DIRECT")) raft_direct_ = atoi(kn);            // 1: always volume-free, 0: never, unset: by size
  if (const char* kn = getenv("GIMMVFI_RAFT_CORR_DIRECT_GB")) raft_direct_auto_bytes_ = atof(kn) * 1e9;
  if (const char* kn = getenv("GIMMVFI_GRU_HOIST")) gru_hoist_ = atoi(kn) != 0;
#ifndef GV_HOSTSIM
  int count = 0;
  cuda_ok(cudaGetDeviceCount(&count), "cudaGetDeviceCount");
  if (device < 0 || device >= count) throw std::runtime_error("gimmvfi: no CUDA device " + std::to_string(device));
  int sms = 148;
  cuda_ok(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device), "cudaDeviceGetAttribute");   // (no cudaSetDevice: the caller's current device stays)
  sm_count_ = sms;
#endif
}

Engine::~Engine() {
  clear_graphs();
  DeviceGuard dg(device_);
  for (void* p : dev_allocs_) dev_free(p);
}

void Engine::load_weight(const std::string& key, const float* host, const int64_t* shape, int ndim) {
  HostTensor t;
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) { t.shape.push_back(shape[i]); n *= shape[i]; }
  t.data.assign(host, host + n);
  raw_[key] = std::move(t);
  finalized_ = false;
}

const HostTensor& Engine::raw(const std::string& key) const {
  auto it = raw_.find(key);
  if (it == raw_.end()) throw std::runtime_error("gimmvfi: missing weight '" + key + "'");
  return it->second;
}

const float* Engine::upload(const std::vector<float>& v) {
  void* d = dev_alloc(v.size() * sizeof(float));
  dev_allocs_.push_back(d);
  dev_upload(d, v.data(), v.size() * sizeof(float));
  return static_cast<const float*>(d);
}

const float* Engine::vec(const std::string& key) {
  auto it = vec_.find(key);
  if (it != vec_.end()) return it->second;
  const float* d = upload(raw(key).data);
  vec_[key] = d;
  return d;
}

// Conv2d weight (cout,cin,kh,kw) -> [tap][cin][cout_ld]; optional eval-mode BatchNorm
// folding (bn = name of the BatchNorm2d), output scale and output-channel permutation.
ConvW Engine::pack_conv(const std::string& name, const std::string& bn, float out_scale, const std::vector<int>* perm) {
  const HostTensor& W = raw(name + ".weight");
  if (W.shape.size() != 4 && W.shape.size() != 2) throw std::runtime_error("gimmvfi: '" + name + ".weight' is neither a Conv2d (4-D) nor a Linear (2-D) weight");
  const bool lin = W.shape.size() == 2;   // nn.Linear (cout, cin) = a 1x1 convolution over tokens laid out as an NHWC map
  const int cout = (int)W.shape[0], cin = (int)W.shape[1], kh = lin ? 1 : (int)W.shape[2], kw = lin ? 1 : (int)W.shape[3];
  HostTensor zero_bias;
  if (!raw_.count(name + ".bias")) { zero_bias.shape = {cout}; zero_bias.data.assign(cout, 0.f); }   // bias=False layers
  const HostTensor& Bv = raw_.count(name + ".bias") ? raw(name + ".bias") : zero_bias;
  std::vector<float> s(cout, out_scale), sh(cout, 0.f);
  if (!bn.empty()) {
    const auto& g = raw(bn + ".weight").data; const auto& be = raw(bn + ".bias").data;
    const auto& mu = raw(bn + ".running_mean").data; const auto& var = raw(bn + ".running_var").data;
    for (int co = 0; co < cout; ++co) {
      float k = g[co] / std::sqrt(var[co] + 1e-5f);
      s[co] = k * out_scale; sh[co] = (be[co] - mu[co] * k) * out_scale;
    }
  }
  const int cout_ld = (cout + 3) & ~3;
  std::vector<float> pw((size_t)kh * kw * cin * cout_ld, 0.f), pb(((tc_cout_pad(cout) + 31) & ~31) + 128, 0.f);  // bias padded past any tensor-core N tiling (float4 reads)
  for (int co = 0; co < cout; ++co) {
    const int src = perm ? (*perm)[co] : co;
    for (int ci = 0; ci < cin; ++ci)
      for (int t = 0; t < kh * kw; ++t)
        pw[((size_t)t * cin + ci) * cout_ld + co] = W.data[((size_t)src * cin + ci) * kh * kw + t] * s[src];
    pb[co] = Bv.data[src] * s[src] + sh[src];
  }
  ConvW c;
  c.w = upload(pw); c.b = upload(pb); c.cin = cin; c.cout = cout; c.kh = kh; c.kw = kw; c.cout_ld = cout_ld;
  pack_tc(c, pw);
  // layers of the final decoder's residual trunk read half-precision activations in precision mode 3 (run())
  if (name.rfind("amt_final_decoder.convblock.", 0) == 0 && name.rfind("amt_final_decoder.convblock.0", 0) != 0) pack_tc_f16(c, pw);
  // precision mode 4 (experimental): the 32/64-channel full-resolution chains are stored in half as well -> their consumers
  static const char* const kHalfFed[] = {"cnn_encoder.3.", "cnn_encoder.4.", "cnn_encoder.5.", "cnn_encoder.7", "res_conv.1", "res_conv.3.", "res_conv.5",
                                         "amt_final_decoder.upsample.3.", "amt_final_decoder.upsample.4.", "amt_final_decoder.upsample.5.",
                                         "amt_final_decoder.upsample.6.", "amt_final_decoder.upsample.7", "amt_init_decoder.convblock.1.",
                                         "amt_init_decoder.convblock.2.", "amt_init_decoder.convblock.3.", "amt_init_decoder.convblock.4",
                                         "amt_final_decoder.convblock.0."};
  for (const char* pre : kHalfFed)
    if (name.rfind(pre, 0) == 0) { pack_tc_f16(c, pw); break; }
  conv_[name] = c;
  return c;
}

// round-to-nearest-even to TF32 (10-bit mantissa): the tensor core ignores the low 13 bits
static float tf32_rn(float x) {
  uint32_t u; std::memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return x;
  u += 0xfffu + ((u >> 13) & 1u);
  u &= 0xffffe000u;
  float y; std::memcpy(&y, &u, 4);
  return y;
}

// [tap][cin][cout_ld] fp32  ->  [2][tap][cout_pad][cin_pad]: plane 0 = TF32(w) (K-major rows for the UMMA B
// operand), plane 1 = TF32(w - plane0), the low term of the 3xTF32 split
void Engine::pack_tc(ConvW& c, const std::vector<float>& pw) {
  const int cout_pad = tc_cout_pad(c.cout), cin_pad = (c.cin + 31) & ~31, taps = c.kh * c.kw;
  const size_t plane = (size_t)taps * cout_pad * cin_pad;
  std::vector<float> t(2 * plane, 0.f);
  for (int tp = 0; tp < taps; ++tp)
    for (int ci = 0; ci < c.cin; ++ci)
      for (int co = 0; co < c.cout; ++co) {
        const float w = pw[((size_t)tp * c.cin + ci) * c.cout_ld + co];
        const float hi = tf32_rn(w);
        const size_t o = ((size_t)tp * cout_pad + co) * cin_pad + ci;
        t[o] = hi; t[plane + o] = tf32_rn(w - hi);
      }
  c.w_tc = upload(t); c.cout_pad = cout_pad; c.cin_pad = cin_pad; c.has_lo = true;
  // fp16 hi / lo planes of w * 2^e for the 3xF16 split: 64-element K blocks (two 32-channel fp32 activation boxes per K step)
  {
    float mx = 0.f;
    for (float v : pw) mx = std::max(mx, std::fabs(v));
    int e = 0;
    if (mx > 0.f) { int ex; std::frexp(mx, &ex); e = 14 - ex; }   // mx * 2^e in [2^13, 2^14)
    if (e > 24) e = 24;
    if (e < -24) e = -24;
    const float sc = std::ldexp(1.f, e);
    const int cps = (c.cin + 63) & ~63;
    const size_t pl = (size_t)taps * cout_pad * cps;
    std::vector<float> hs((2 * pl + 1) / 2, 0.f);
    uint16_t* h = reinterpret_cast<uint16_t*>(hs.data());
    for (int tp = 0; tp < taps; ++tp)
      for (int ci = 0; ci < c.cin; ++ci)
        for (int co = 0; co < c.cout; ++co) {
          const float w = pw[((size_t)tp * c.cin + ci) * c.cout_ld + co] * sc;
          const uint16_t hi = gv_f32_to_f16(w), lo = gv_f32_to_f16(w - gv_f16_to_f32(hi));
          const size_t o = ((size_t)tp * cout_pad + co) * cps + ci;
          h[o] = hi; h[pl + o] = lo;
        }
    c.w_tc_s = upload(hs); c.cin_pad_s = cps; c.w_scale = sc;
  }
}

// [tap][cin][cout_ld] fp32 -> [tap][cout_pad][cin_pad_h] fp16 (K-major rows of 64-element / 128-byte K blocks): the
// B operand of the kind::f16 tensor-core path for layers whose activations are stored in half precision
void Engine::pack_tc_f16(ConvW& c, const std::vector<float>& pw) {
  const int cout_pad = tc_cout_pad(c.cout), cin_pad = (c.cin + 63) & ~63, taps = c.kh * c.kw;
  std::vector<float> t(((size_t)taps * cout_pad * cin_pad + 1) / 2, 0.f);
  uint16_t* h = reinterpret_cast<uint16_t*>(t.data());
  for (int tp = 0; tp < taps; ++tp)
    for (int ci = 0; ci < c.cin; ++ci)
      for (int co = 0; co < c.cout; ++co)
        h[((size_t)tp * cout_pad + co) * cin_pad + ci] = gv_f32_to_f16(pw[((size_t)tp * c.cin + ci) * c.cout_ld + co]);
  c.w_tc_h = upload(t); c.cin_pad_h = cin_pad;
}

// RAFT stem (7x7, stride 2, cin 3) with the x-taps packed into the channel axis: on a zero-padded image with
// 4-float pixels, the 7 x-taps of one kernel row are 28 contiguous floats, so the conv becomes a (7 x 1) kernel
// over 28 "channels" (K = 7 x 28 instead of 49 taps x 3-of-16 used lanes: 3.5x fewer K chunks).
void Engine::pack_stem(const std::string& name, const std::string& bn) {
  const HostTensor& W = raw(name + ".weight");
  const HostTensor& Bv = raw(name + ".bias");
  const int cout = (int)W.shape[0], cin = (int)W.shape[1], kh = (int)W.shape[2], kw = (int)W.shape[3];
  if (cin != 3 || kh != 7 || kw != 7) throw std::runtime_error("pack_stem: expected a 7x7 conv on 3 channels");
  std::vector<float> s(cout, 1.f), sh(cout, 0.f);
  if (!bn.empty()) {
    const auto& g = raw(bn + ".weight").data; const auto& be = raw(bn + ".bias").data;
    const auto& mu = raw(bn + ".running_mean").data; const auto& var = raw(bn + ".running_var").data;
    for (int co = 0; co < cout; ++co) { float k = g[co] / std::sqrt(var[co] + 1e-5f); s[co] = k; sh[co] = be[co] - mu[co] * k; }
  }
  const int cout_ld = (cout + 3) & ~3, K = 28;
  std::vector<float> pw((size_t)kh * K * cout_ld, 0.f), pb(((tc_cout_pad(cout) + 31) & ~31) + 128, 0.f);
  for (int co = 0; co < cout; ++co) {
    for (int ky = 0; ky < kh; ++ky)
      for (int kx = 0; kx < kw; ++kx)
        for (int c = 0; c < 3; ++c)
          pw[((size_t)ky * K + kx * 4 + c) * cout_ld + co] = W.data[(((size_t)co * cin + c) * kh + ky) * kw + kx] * s[co];
    pb[co] = Bv.data[co] * s[co] + sh[co];
  }
  ConvW c;
  c.w = upload(pw); c.b = upload(pb); c.cin = K; c.cout = cout; c.kh = kh; c.kw = 1; c.cout_ld = cout_ld;
  pack_tc(c, pw);   // stride 2 runs on the tensor-core path too (TMA element strides)
  conv_[name + "#xpacked"] = c;
}

// k x k convolutions (k = 7, 5) on few channels with the x-taps packed into the channel axis (see pack_stem): the input is a
// zero-padded copy with `ldp` floats per pixel, the kernel becomes (k x 1) over K = k*ldp "channels" - k K-step groups instead of k*k.
void Engine::pack_xpacked(const std::string& name, int ldp) {
  const HostTensor& W = raw(name + ".weight");
  const HostTensor& Bv = raw(name + ".bias");
  const int cout = (int)W.shape[0], cin = (int)W.shape[1], kh = (int)W.shape[2], kw = (int)W.shape[3];
  if (kh != kw || cin > ldp) throw std::runtime_error("pack_xpacked: expected a k x k conv with cin <= ldp");
  const int cout_ld = (cout + 3) & ~3, K = kw * ldp;
  std::vector<float> pw((size_t)kh * K * cout_ld, 0.f), pb(((tc_cout_pad(cout) + 31) & ~31) + 128, 0.f);
  for (int co = 0; co < cout; ++co) {
    for (int ky = 0; ky < kh; ++ky)
      for (int kx = 0; kx < kw; ++kx)
        for (int c = 0; c < cin; ++c)
          pw[((size_t)ky * K + kx * ldp + c) * cout_ld + co] = W.data[(((size_t)co * cin + c) * kh + ky) * kw + kx];
    pb[co] = Bv.data[co];
  }
  ConvW c;
  c.w = upload(pw); c.b = upload(pb); c.cin = K; c.cout = cout; c.kh = kh; c.kw = 1; c.cout_ld = cout_ld;
  pack_tc(c, pw);
  conv_[name + "#xp"] = c;
}

void Engine::finalize_weights() {
  DeviceGuard dg(device_);
  for (void* p : dev_allocs_) dev_free(p);
  dev_allocs_.clear(); conv_.clear(); vec_.clear();
  fc_valid_.clear(); ++weights_version_; clear_graphs();   // a frame cache stored under the old weights must not be loaded again
  // --- RAFT encoders (raft/extractor.py:122-171); cnet's BatchNorm is folded
  for (int e = 0; e < 2; ++e) {
    const std::string p = e == 0 ? "flow_estimator.fnet" : "flow_estimator.cnet";
    const bool bn = e == 1;
    pack_conv(p + ".conv1", bn ? p + ".norm1" : "");
    pack_stem(p + ".conv1", bn ? p + ".norm1" : "");
    for (int li = 1; li <= 3; ++li)
      for (int bi = 0; bi < 2; ++bi) {
        const std::string q = p + ".layer" + std::to_string(li) + "." + std::to_string(bi);
        pack_conv(q + ".conv1", bn ? q + ".norm1" : "");
        pack_conv(q + ".conv2", bn ? q + ".norm2" : "");
        if (bi == 0 && li > 1) pack_conv(q + ".downsample.0", bn ? q + ".downsample.1" : "");
      }
    pack_conv(p + ".conv2");
  }
  // --- RAFT update block (raft/update.py:94-154)
  const std::string u = "flow_estimator.update_block";
  for (const char* n : {".encoder.convc1", ".encoder.convc2", ".encoder.convf1", ".encoder.convf2", ".encoder.conv",
                        ".gru.convz1", ".gru.convr1", ".gru.convq1", ".gru.convz2", ".gru.convr2", ".gru.convq2",
                        ".flow_head.conv1", ".flow_head.conv2", ".mask.0"})
    pack_conv(u + n);
  for (const char* sfx : {"1", "2"}) {   // z | r gates read the same input: one 384 -> 256 convolution (raft/update.py:47-49,55-57)
    const HostTensor& wz = raw(u + ".gru.convz" + sfx + ".weight"); const HostTensor& wr = raw(u + ".gru.convr" + sfx + ".weight");
    const HostTensor& bz = raw(u + ".gru.convz" + sfx + ".bias"); const HostTensor& br = raw(u + ".gru.convr" + sfx + ".bias");
    HostTensor w = wz, b = bz;
    w.shape[0] = wz.shape[0] + wr.shape[0]; w.data.insert(w.data.end(), wr.data.begin(), wr.data.end());
    b.shape[0] = bz.shape[0] + br.shape[0]; b.data.insert(b.data.end(), br.data.begin(), br.data.end());
    raw_[u + ".gru.convzr" + sfx + ".weight"] = w; raw_[u + ".gru.convzr" + sfx + ".bias"] = b;
    pack_conv(u + ".gru.convzr" + sfx);
    // The GRU input is [h | inp | motion features] (raft/update.py:122-124) and `inp` (the context network's output) does not change
    // over the iterations: its share of every gate convolution is computed once per pair ("_inp", carries the bias) and enters the
    // per-iteration convolutions over [h | motion] ("_hm", K = 256 instead of 384) as a pre-activation term.
    for (const char* gate : {"zr", "q"}) {
      const std::string g = u + ".gru.conv" + gate + sfx;
      const HostTensor& wf = raw(g + ".weight"); const HostTensor& bf = raw(g + ".bias");
      const int64_t co = wf.shape[0], ci = wf.shape[1], kk = wf.shape[2] * wf.shape[3];
      if (ci != 384) throw std::runtime_error("gimmvfi: SepConvGRU input width");
      HostTensor w_hm, w_in, b0 = bf;
      w_hm.shape = {co, 256, wf.shape[2], wf.shape[3]}; w_in.shape = {co, 128, wf.shape[2], wf.shape[3]};
      for (int64_t o = 0; o < co; ++o)
        for (int64_t c = 0; c < ci; ++c) {
          std::vector<float>& dst = (c >= 128 && c < 256) ? w_in.data : w_hm.data;
          dst.insert(dst.end(), wf.data.begin() + (o * ci + c) * kk, wf.data.begin() + (o * ci + c + 1) * kk);
        }
      std::fill(b0.data.begin(), b0.data.end(), 0.f);
      raw_[g + "_hm.weight"] = w_hm; raw_[g + "_hm.bias"] = b0;
      raw_[g + "_inp.weight"] = w_in; raw_[g + "_inp.bias"] = bf;
      pack_conv(g + "_hm"); pack_conv(g + "_inp");
    }
  }
  pack_conv(u + ".mask.2", "", 0.25f);  // "scale mask to balance gradients" raft/update.py:153
  pack_xpacked(u + ".encoder.convf1", 4);
  // --- feature projections (gimmvfi_r.py:51-53)
  pack_conv("amt_last_cproj"); pack_conv("amt_second_last_cproj"); pack_conv("amt_fproj");
  finalize_decoders();
  finalize_gimm_part();
  finalized_ = true; gimm_only_ = false; synth_only_ = false; ff_ = false;
}

// AMT decoders / update blocks / combine block (fi_components.py:229-305): shared by GIMM-VFI-R and -F
void Engine::finalize_decoders() {
  // --- decoders (fi_components.py:229-305)
  {
    const std::string p = "amt_init_decoder";
    for (int i = 1; i <= 5; ++i) { pack_conv(p + ".upsample." + std::to_string(i) + ".0"); vec(p + ".upsample." + std::to_string(i) + ".1.weight"); }
    pack_conv(p + ".upsample.6", p + ".upsample.7");
    pack_conv(p + ".convblock.0.0"); vec(p + ".convblock.0.1.weight");
    // head re-ordered to [ft(128) | dflow0(2) | dflow1(2) | mask(1)] so ft_ is an aligned slice
    std::vector<int> perm(133);
    for (int i = 0; i < 128; ++i) perm[i] = 5 + i;
    for (int i = 0; i < 5; ++i) perm[128 + i] = i;
    pack_conv(p + ".convblock.4", "", 1.f, &perm);
  }
  {
    const std::string p = "amt_final_decoder";
    for (int i = 2; i <= 6; ++i) { pack_conv(p + ".upsample." + std::to_string(i) + ".0"); vec(p + ".upsample." + std::to_string(i) + ".1.weight"); }
    pack_conv(p + ".upsample.7", p + ".upsample.8");
    pack_conv(p + ".convblock.0.0"); vec(p + ".convblock.0.1.weight");
    pack_conv(p + ".convblock.4");
  }
  for (const char* d : {"amt_init_decoder", "amt_final_decoder"})
    for (int k = 1; k <= 3; ++k) {
      const std::string q = std::string(d) + ".convblock." + std::to_string(k);
      for (int c = 1; c <= 4; ++c) { pack_conv(q + ".conv" + std::to_string(c) + ".0"); vec(q + ".conv" + std::to_string(c) + ".1.weight"); }
      pack_conv(q + ".conv5"); vec(q + ".prelu.weight");
    }
  for (const char* d : {"amt_update4_low", "amt_update4_high"})
    for (const char* n : {".convc1", ".convc2", ".convf1", ".convf2", ".conv", ".gru.0", ".gru.2", ".feat_head.0", ".feat_head.2", ".flow_head.0", ".flow_head.2"})
      pack_conv(std::string(d) + n);
  pack_conv("amt_comb_block.0"); vec("amt_comb_block.1.weight"); pack_conv("amt_comb_block.2");
  pack_xpacked("amt_comb_block.0", 12); pack_xpacked("amt_comb_block.2", 20);
  pack_xpacked("amt_update4_low.convf1", 4); pack_xpacked("amt_update4_high.convf1", 4);
  pack_xpacked("amt_final_decoder.upsample.2.0", 8);   // 5x5 on 8 channels at full resolution: 5 x 40 lanes instead of 25 taps x 8-of-32 lanes
}

// GIMM-VFI-F's parameter tree minus flow_estimator.* (gimmvfi_f.py:37-111): the synthesis half runs natively on the outputs of an
// external flow estimator (forward_from_flow); there are no feature projections in F (twins features are used as they are)
void Engine::finalize_weights_synthesis() {
  DeviceGuard dg(device_);
  for (void* p : dev_allocs_) dev_free(p);
  dev_allocs_.clear(); conv_.clear(); vec_.clear();
  fc_valid_.clear(); ++weights_version_; clear_graphs();
  finalize_decoders();
  finalize_gimm_part();
  finalized_ = true; gimm_only_ = false; synth_only_ = true; ff_ = false;
}

// GIMM-VFI-F complete (gimmvfi_f.py:27-111): the synthesis half + the native FlowFormer estimator (flowformer.cu)
void Engine::finalize_weights_f() {
  finalize_weights_synthesis();
  DeviceGuard dg(device_);
  finalize_flowformer();
  synth_only_ = false; ff_ = true;
}

void Engine::tap_copy(Ctx& cx, const std::string& name, const TV& tv) {
  if (!debug_) return;
  TV c = cx.arena.tensor(tv.n, tv.h, tv.w, tv.c);
  copy_channels(cx, tv, c);
  taps_[name] = c;
}

// GIMM's own parameters (gimm.py:36-80 == gimmvfi_r.py:86-111): motion encoder, latent refiner, HypoNet, splat-metric scalars
void Engine::finalize_gimm_part() {
  // --- GIMM encoders (gimmvfi_r.py:86-109)
  for (const char* n : {"cnn_encoder.0", "cnn_encoder.1", "cnn_encoder.3.layers.0", "cnn_encoder.3.layers.2", "cnn_encoder.4.layers.0",
                        "cnn_encoder.4.layers.2", "cnn_encoder.5.layers.0", "cnn_encoder.5.layers.2", "cnn_encoder.7", "res_conv.0",
                        "res_conv.1", "res_conv.3.layers.0", "res_conv.3.layers.2", "res_conv.5"})
    pack_conv(n);
  // --- HypoNet (hyponet.py:99-143): fan-in-normalised columns, bias row, +output_bias on the last layer
  for (int l = 0; l < 5; ++l) {
    const std::string key = "hyponet.params_dict.linear_wb" + std::to_string(l);
    const HostTensor& wb = raw(key);
    const int fin = (int)wb.shape[0] - 1, fout = (int)wb.shape[1];
    const int ld = (fout + 3) & ~3;
    std::vector<float> pw((size_t)fin * ld, 0.f), pb(((tc_cout_pad(fout) + 31) & ~31) + 128, 0.f);
    for (int co = 0; co < fout; ++co) {
      double nrm = 0.0;
      for (int ci = 0; ci < fin; ++ci) { double v = wb.data[(size_t)ci * fout + co]; nrm += v * v; }
      float d = std::max((float)std::sqrt(nrm), 1e-12f);
      for (int ci = 0; ci < fin; ++ci) pw[(size_t)ci * ld + co] = wb.data[(size_t)ci * fout + co] / d;
      pb[co] = wb.data[(size_t)fin * fout + co] + (l == 4 ? 0.5f : 0.f);
    }
    ConvW c; c.w = upload(pw); c.b = upload(pb); c.cin = fin; c.cout = fout; c.kh = c.kw = 1; c.cout_ld = ld;
    pack_tc(c, pw);
    if (l >= 1) pack_tc_f16(c, pw);   // precision mode 4: the sin() activations between the layers (|x| <= 1) are stored in half
    conv_[key] = c;
  }
  // --- the same five matrices packed for the fused kernel (hyponet.cu): pre-swizzled K-major tiles, fp32 affine rows for (t, y, x)
  {
    std::vector<float> blobf((hypo::BLOB + 3) / 4, 0.f);
    uint8_t* blob = reinterpret_cast<uint8_t*>(blobf.data());
    auto norm_col = [&](const HostTensor& wb, int fin, int fout, int co) {
      double nrm = 0.0;
      for (int ci = 0; ci < fin; ++ci) { double v = wb.data[(size_t)ci * fout + co]; nrm += v * v; }
      return std::max((float)std::sqrt(nrm), 1e-12f);
    };
    const HostTensor& w0 = raw("hyponet.params_dict.linear_wb0");
    if (w0.shape[0] != 36 || w0.shape[1] != 128) throw std::runtime_error("hyponet: expected linear_wb0 of shape (36, 128)");
    float* aff = reinterpret_cast<float*>(blob + hypo::AFF);
    for (int n = 0; n < 128; ++n) {
      const float d = norm_col(w0, 35, 128, n);
      for (int k = 0; k < 32; ++k) { const float v = tf32_rn(w0.data[(size_t)k * 128 + n] / d); std::memcpy(blob + hypo::W0 + hypo::swz(n, k * 4), &v, 4); }
      for (int j = 0; j < 3; ++j) aff[j * 128 + n] = w0.data[(size_t)(32 + j) * 128 + n] / d;
      aff[3 * 128 + n] = w0.data[(size_t)35 * 128 + n];
    }
    for (int l = 1; l <= 4; ++l) {
      const HostTensor& wb = raw("hyponet.params_dict.linear_wb" + std::to_string(l));
      const int fout = (int)wb.shape[1];
      if (wb.shape[0] != 129 || fout != (l < 4 ? 128 : 2)) throw std::runtime_error("hyponet: unexpected hidden / output layer shape");
      const int rows = l < 4 ? 128 : 16, base = l < 4 ? hypo::W1 + (l - 1) * 32768 : hypo::W4, kb_bytes = rows * 128;
      float* bias = reinterpret_cast<float*>(blob + (l < 4 ? hypo::B1 + (l - 1) * 512 : hypo::B4));
      for (int n = 0; n < fout; ++n) {
        const float d = norm_col(wb, 128, fout, n);
        for (int k = 0; k < 128; ++k) {
          const uint16_t h = gv_f32_to_f16(wb.data[(size_t)k * fout + n] / d);
          std::memcpy(blob + base + (k / 64) * kb_bytes + hypo::swz(n, (k % 64) * 2), &h, 2);
        }
        bias[n] = wb.data[(size_t)128 * fout + n] + (l == 4 ? 0.5f : 0.f);
      }
    }
    hypo_blob_ = upload(blobf);
    // fp32-class variant (hyponet_fused3): scaled fp16 hi / lo planes of layers 1-3, fp32 rows of layers 0 and 4
    std::vector<float> b3f((hypo3::BLOB + 3) / 4, 0.f);
    uint8_t* b3 = reinterpret_cast<uint8_t*>(b3f.data());
    float* w0a = reinterpret_cast<float*>(b3 + hypo3::W0A);
    for (int n = 0; n < 128; ++n) {
      const float d = norm_col(w0, 35, 128, n);
      for (int k = 0; k < 35; ++k) w0a[k * 128 + n] = w0.data[(size_t)k * 128 + n] / d;
      w0a[35 * 128 + n] = w0.data[(size_t)35 * 128 + n];
    }
    for (int l = 1; l <= 3; ++l) {
      const HostTensor& wb = raw("hyponet.params_dict.linear_wb" + std::to_string(l));
      float* bias = reinterpret_cast<float*>(b3 + hypo3::B13 + (l - 1) * 512);
      for (int n = 0; n < 128; ++n) {
        const float d = norm_col(wb, 128, 128, n);
        for (int k = 0; k < 128; ++k) {
          const float v = wb.data[(size_t)k * 128 + n] / d * hypo3::W_SCALE;
          const uint16_t hi = gv_f32_to_f16(v), lo = gv_f32_to_f16(v - gv_f16_to_f32(hi));
          std::memcpy(b3 + hypo3::W13 + (((l - 1) * 2 + 0) * 2 + k / 64) * 16384 + hypo::swz(n, (k % 64) * 2), &hi, 2);
          std::memcpy(b3 + hypo3::W13 + (((l - 1) * 2 + 1) * 2 + k / 64) * 16384 + hypo::swz(n, (k % 64) * 2), &lo, 2);
        }
        bias[n] = wb.data[(size_t)128 * 128 + n];
      }
    }
    {
      const HostTensor& wb = raw("hyponet.params_dict.linear_wb4");
      float* w4 = reinterpret_cast<float*>(b3 + hypo3::W4);
      float* b4 = reinterpret_cast<float*>(b3 + hypo3::B4);
      for (int n = 0; n < 2; ++n) {
        const float d = norm_col(wb, 128, 2, n);
        for (int k = 0; k < 128; ++k) w4[k * 2 + n] = wb.data[(size_t)k * 2 + n] / d;
        b4[n] = wb.data[(size_t)128 * 2 + n] + 0.5f;
      }
    }
    hypo_blob3_ = upload(b3f);
  }
  g9_ = vec("g_filter"); alpha_fe_ = vec("alpha_fe"); alpha_v_ = vec("alpha_v");
}

// standalone GIMM checkpoint: only the keys of gimm.py's module tree are present
void Engine::finalize_weights_gimm() {
  DeviceGuard dg(device_);
  for (void* p : dev_allocs_) dev_free(p);
  dev_allocs_.clear(); conv_.clear(); vec_.clear();
  fc_valid_.clear(); ++weights_version_; clear_graphs();
  finalize_gimm_part();
  finalized_ = true; gimm_only_ = true; ff_ = false;
}


// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
static ConvW slice_cout(const ConvW& w, int co0, int cnt) {
  ConvW s = w; s.w = w.w + co0; s.b = w.b + co0; s.cout = cnt; s.w_tc = nullptr;  // TC layout is not sliceable this way
  return s;
}

// raft/extractor.py:173-220.  x: (n,H,W,3) -> fmap (n,H/8,W/8,256) [+ layer2/layer3 outputs]
static void raft_encoder(Net& N, const std::string& p, bool instance, const TV& x, const TV& xpad /*zero-padded by 3, 4-float pixels*/,
                         const TV& out256, TV* feat4, TV* feat8,
                         const ConvW* split_tanh_relu_out /*cnet: tanh/relu halves*/, const TV& net_out, const TV& inp_out) {
  Ctx& cx = N.cx; Arena& A = cx.arena;
  const int n = x.n, H2 = x.h / 2, W2 = x.w / 2;
  float* mr = nullptr; float* scratch = nullptr;
  if (instance) {
    mr = A.alloc_f((size_t)n * 128 * 2);
    TV probe = make_tv(nullptr, n, H2, W2, 128);
    scratch = A.alloc_f((size_t)instnorm_scratch_floats(probe));
  }
  auto norm_act = [&](const TV& raw, int act1, const TV& res, int act2, const TV& dst) {
    instnorm_stats(cx, raw, mr, scratch, 0);
    instnorm_apply(cx, raw, mr, act1, res, act2, dst);
  };
  // stem
  TV a = A.tensor(n, H2, W2, 64);
  auto stem = [&](const TV& dst, int act) {   // 7x7 stride-2 stem through the x-packed weights (pack_stem)
    TV in = xpad; in.c = 28;
    ConvGeom g; g.stride = 2; g.ph = 0; g.pw = 0; g.loose_w = 1;
    ConvEpi e; e.act1 = act;
    conv2d(cx, in, TV(), N.W(p + ".conv1#xpacked"), g, e, dst);   // (tensor-core path: TMA element strides take every 2nd pixel)
  };
  if (instance) {
    TV r = A.tensor(n, H2, W2, 64);
    stem(r, ACT_NONE);
    norm_act(r, ACT_RELU, TV(), ACT_NONE, a);
  } else {
    stem(a, ACT_RELU);
  }
  TV cur = a;
  int cin = 64;
  const int dims[3] = {64, 96, 128};
  for (int li = 1; li <= 3; ++li) {
    const int dim = dims[li - 1];
    for (int bi = 0; bi < 2; ++bi) {
      const std::string q = p + ".layer" + std::to_string(li) + "." + std::to_string(bi);
      const int stride = (bi == 0 && li > 1) ? 2 : 1;
      const int oh = cur.h / stride, ow = cur.w / stride;
      TV y1 = A.tensor(n, oh, ow, dim), y2 = A.tensor(n, oh, ow, dim), outb = A.tensor(n, oh, ow, dim);
      TV skip = cur;
      if (instance) {
        TV r = A.tensor(n, oh, ow, dim);
        N.conv(q + ".conv1", cur, r, ACT_NONE, nullptr, stride);
        norm_act(r, ACT_RELU, TV(), ACT_NONE, y1);
        N.conv(q + ".conv2", y1, r, ACT_NONE);
        if (stride != 1) {
          TV d = A.tensor(n, oh, ow, dim);
          N.conv(q + ".downsample.0", cur, d, ACT_NONE, nullptr, stride);
          norm_act(d, ACT_NONE, TV(), ACT_NONE, y2);  // y2 reused as the normalised skip
          skip = y2;
        }
        norm_act(r, ACT_RELU, skip, ACT_RELU, outb);  // relu(x + relu(norm2(conv2)))
      } else {
        N.conv(q + ".conv1", cur, y1, ACT_RELU, nullptr, stride);
        if (stride != 1) { N.conv(q + ".downsample.0", cur, y2, ACT_NONE, nullptr, stride); skip = y2; }
        ConvEpi e; e.act1 = ACT_RELU; e.res = skip; e.act2 = ACT_RELU;
        N.conv_e(q + ".conv2", y1, TV(), outb, e);
      }
      cur = outb; cin = dim;
    }
    if (li == 2 && feat4) *feat4 = cur;
    if (li == 3 && feat8) *feat8 = cur;
  }
  (void)cin;
  if (split_tanh_relu_out) {
    // cnet: net = tanh(first 128), inp = relu(last 128)  (raft/raft.py:133-136)
    const ConvW& w = *split_tanh_relu_out;
    ConvEpi e1; e1.act1 = ACT_TANH;
    conv2d(cx, cur, TV(), slice_cout(w, 0, 128), ConvGeom(), e1, net_out);
    ConvEpi e2; e2.act1 = ACT_RELU;
    conv2d(cx, cur, TV(), slice_cout(w, 128, 128), ConvGeom(), e2, inp_out);
  } else {
    N.conv(p + ".conv2", cur, out256);
  }
}

struct Pyramid {
  float* lvl[4]; int h[4], w[4]; int64_t N;  // rows per sample
  CorrPyr view(int sample0) const {
    CorrPyr c;
    for (int l = 0; l < 4; ++l) { c.lvl[l] = lvl[l] + (int64_t)sample0 * N * h[l] * w[l]; c.h[l] = h[l]; c.w[l] = w[l]; }
    c.rows_per_sample = N;
    return c;
  }
};

// Volumes for both directions stacked on the batch axis: samples [0,B) hold
// <F[s], F[s+B]>, samples [B,2B) hold <F[s], F[s-B]> (= the transposed volume,
// raft/corr.py:32).  Level l is (2B*N) x (h_l*w_l).
// tc: 0 = fp32 CUDA-core GEMM, 1 = tcgen05 plain TF32, 2 = tcgen05 3xTF32 (promoted accumulation)
static Pyramid build_pyramid(Ctx& cx, const TV& F /*2B,h,w,256*/, int B, int tensor_cores) {
  Arena& A = cx.arena;
  Pyramid P; P.N = (int64_t)F.h * F.w;
  int h = F.h, w = F.w;
  for (int l = 0; l < 4; ++l) { P.h[l] = h; P.w[l] = w; P.lvl[l] = A.alloc_f((size_t)2 * B * P.N * h * w); h /= 2; w /= 2; }
  const float scale = 1.0f / std::sqrt((float)F.c);
  if (tensor_cores && F.c % 32 == 0 && F.ld % 4 == 0 && F.ld == F.c) {
    // tcgen05 GEMMs; the other frame's features act as the K-major "weights".  Level l > 0 is the GEMM against the 2^l x 2^l
    // AVERAGE-POOLED target features: avg_pool2d of the volume over the target axes (raft/corr.py:139-142) commutes with the dot
    // product, so the pooled levels (1/4 + 1/16 + 1/64 of level 0) cost a third more GEMM work and the 8.5 GB level-0 volume
    // (1088x1920) is never re-read (the pooling pass over it took as long as the GEMM that wrote it).
    const bool split = tensor_cores >= 2;
    const size_t mk = A.mark();
    const int64_t plane = (int64_t)P.N * F.c;
    float* planes = A.alloc_f((size_t)2 * plane);
    const int64_t nz = ((P.N + 255) / 256) * 256 + 512;
    float* zeros = A.alloc_f((size_t)nz);
    if (!cx.dry) dev_memset(zeros, 0, (size_t)nz * sizeof(float), cx.stream);
    TV Fl[4]; Fl[0] = F;
    for (int l = 1; l < 4; ++l) { Fl[l] = A.tensor(2 * B, P.h[l], P.w[l], F.c); avgpool2_features(cx, Fl[l - 1], Fl[l]); }
    for (int s = 0; s < 2 * B; ++s) {
      const int other = s < B ? s + B : s - B;
      for (int l = 0; l < 4; ++l) {
        const int nt = P.h[l] * P.w[l];
        if (nt <= 0) continue;
        if (split && corr_volume_tc_wants_f16_planes()) split_planes_f16(cx, Fl[l].batch(other, 1), planes);
        else if (split) split_planes(cx, Fl[l].batch(other, 1), planes);
        if (!cx.dry) corr_volume_tc(cx, F.batch(s, 1), split ? planes : Fl[l].batch(other, 1).p, zeros, P.lvl[l] + (int64_t)s * P.N * nt, scale, split, nt);
      }
    }
    A.release(mk);
    return P;
  } else
  {
    corr_volume(cx, F.batch(0, B), F.batch(B, B), P.lvl[0], scale);
    corr_volume(cx, F.batch(B, B), F.batch(0, B), P.lvl[0] + (int64_t)B * P.N * P.N, scale);
  }
  corr_pool_pyramid(cx, P.lvl[0], P.lvl[1], P.lvl[2], P.lvl[3], (int64_t)2 * B * P.N, P.h[0], P.w[0]);
  return P;
}

// AMT update block (fi_components.py:157-222).  `ft` (128 ch view) and `fl4` (4 ch) are updated in place.
static void amt_update(Net& N, const std::string& p, bool low, const TV& ft, const TV& fl4, const TV& flow_in /*4ch at block res*/,
                       const TV& corr648) {
  Ctx& cx = N.cx; Arena& A = cx.arena;
  const size_t mk = A.mark();
  const int n = corr648.n, h = corr648.h, w = corr648.w;
  TV inp = A.tensor(n, h, w, 320);
  if (low) resize_bilinear(cx, ft, inp.slice(192, 128), 2.f, 2.f, 1.f, 0, ACT_NONE);  // resize(net, 1/2)
  else copy_channels(cx, ft, inp.slice(192, 128));
  copy_channels(cx, flow_in, inp.slice(188, 4));
  TV cor = A.tensor(n, h, w, 256), cf = A.tensor(n, h, w, 256), flo = A.tensor(n, h, w, 128);
  N.conv(p + ".convc1", corr648, cor, ACT_LRELU);
  N.conv(p + ".convc2", cor, cf.slice(0, 192), ACT_LRELU);
  N.conv7x(p + ".convf1", flow_in, flo, ACT_LRELU);
  N.conv(p + ".convf2", flo, cf.slice(192, 64), ACT_LRELU);
  N.conv(p + ".conv", cf, inp.slice(0, 188), ACT_LRELU);
  TV g1 = A.tensor(n, h, w, 192), g2 = A.tensor(n, h, w, 192), hd = A.tensor(n, h, w, 192);
  N.conv(p + ".gru.0", inp, g1, ACT_LRELU);
  N.conv(p + ".gru.2", g1, g2);
  TV dnet = A.tensor(n, h, w, 128), dflow = A.tensor(n, h, w, 4);
  N.conv(p + ".feat_head.0", g2, hd, ACT_LRELU);
  N.conv(p + ".feat_head.2", hd, dnet);
  N.conv(p + ".flow_head.0", g2, hd, ACT_LRELU);
  N.conv(p + ".flow_head.2", hd, dflow);
  if (low) {
    resize_bilinear(cx, dnet, ft, 0.5f, 0.5f, 1.f, 1, ACT_NONE);
    resize_bilinear(cx, dflow, fl4, 0.5f, 0.5f, 2.f, 1, ACT_NONE);
  } else {
    add_inplace_slices(cx, ft, dnet);
    add_inplace_slices(cx, fl4, dflow);
  }
  A.release(mk);
}

// ResBlock (fi_components.py:97-154), in place on x.
static void resblock(Net& N, const std::string& q, const TV& x, int C, int S) {
  Ctx& cx = N.cx; Arena& A = cx.arena;
  const size_t mk = A.mark();
  TV a = A.tensor_like(x, C), b = A.tensor_like(x, C), s1 = A.tensor_like(x, S), s2 = A.tensor_like(x, S);
  N.convrelu(q + ".conv1", x, a);
  N.convrelu(q + ".conv2", a.slice(C - S, S), s1);
  { ConvEpi e; e.act1 = ACT_PRELU; e.slope1 = N.V(q + ".conv3.1.weight"); N.conv_e(q + ".conv3.0", a.slice(0, C - S), s1, b, e); }
  N.convrelu(q + ".conv4", b.slice(C - S, S), s2);
  { ConvEpi e; e.res = x; e.act2 = ACT_PRELU; e.slope2 = N.V(q + ".prelu.weight"); N.conv_e(q + ".conv5", b.slice(0, C - S), s2, x, e); }
  A.release(mk);
}

struct RaftPrepK {  // image = 2 * ((255 * x) / 255) - 1   (gimmvfi_r.py:349 + raft/raft.py:111-112)
  TV src, dst;
  GV_HD void operator()(int64_t i) const {
    float v = src.p[i];
    dst.p[i] = 2.f * ((255.f * v) / 255.0f) - 1.0f;
  }
};
struct ScaleShiftK {  // y = a*x + b on a dense tensor
  const float* src; float* dst; float a, b;
  GV_HD void operator()(int64_t i) const { dst[i] = a * src[i] - b; }
};

static void check_problem(const Problem& p) {
  if (p.B < 1 || p.T < 1) throw std::runtime_error("gimmvfi: B and T must be >= 1");
  const int H = p.H(), W = p.W();
  if (H % 8 || W % 8 || H < 128 || W < 128)
    throw std::runtime_error("gimmvfi: network resolution must be a multiple of 8 and >= 128 (got " + std::to_string(H) + "x" + std::to_string(W) + ")");
  if (p.Hc != H || p.Wc != W)
    throw std::runtime_error("gimmvfi: coordinate grid must match the network resolution (frame_synthesize consumes flow_t at that size)");
}

void Engine::run(Ctx& cx, const Problem& P, const IO& io, const FlowInputs* fin) {
  check_problem(P);
  Net N{*this, cx};
  Arena& A = cx.arena;
  const int knobs = precise_;   // which post-RAFT stages run in 3xTF32 (fp32-class) arithmetic: see set_precise()
  const int B = P.B, Hf = P.Hf, Wf = P.Wf, H = P.H(), W = P.W(), T = P.T;
  const int h = H / 8, w = W / 8, H4 = H / 4, W4 = W / 4;
  const bool ds = P.ds > 0.f;
  const float rds = ds ? (float)(1.0 / (double)P.ds) : 1.f;       // resize(x, ds): src = rds*(dst+.5)-.5
  const double inv_d = ds ? (double)Hf / (double)H : 1.0;          // gimmvfi_r.py:297
  const float inv = (float)inv_d, rinv = (float)(1.0 / inv_d);

  // ------------------------------------------------------------ inputs
  // frames as NHWC, stacked [I0 batch ; I1 batch]
  TV full01 = A.tensor(2 * B, Hf, Wf, 3, 4);
  for (int f = 0; f < 2; ++f)
    nchw_to_nhwc(cx, io.img_xs + (int64_t)f * Hf * Wf, (int64_t)6 * Hf * Wf, (int64_t)2 * Hf * Wf, full01.batch(f * B, B), 1.f, 0.f);
  TV img01 = full01;
  if (ds) {
    img01 = A.tensor(2 * B, H, W, 3, 4);
    resize_bilinear(cx, full01, img01, rds, rds, 1.f, 0, ACT_NONE);  // gimmvfi_r.py:329-337
  }
  TV raft_in = A.tensor(2 * B, H, W, 3, 4);
  TV syn = A.tensor(2 * B, H, W, 3, 4);       // 2*x - 1  (gimmvfi_r.py:230-231)
  {
    // dense views (ld == 4): run over the padded buffers, padding lane included (harmless)
    int64_t cnt = (int64_t)2 * B * H * W * 4;
    TV s = img01; s.c = 4; TV d = raft_in; d.c = 4;
    parallel_for(cx, cnt, RaftPrepK{s, d}, "raft_prep");
    parallel_for(cx, cnt, ScaleShiftK{img01.p, syn.p, 2.f, 1.0f}, "img_to_pm1");
  }
  TV synf = syn;
  if (ds) {
    synf = A.tensor(2 * B, Hf, Wf, 3, 4);
    parallel_for(cx, (int64_t)2 * B * Hf * Wf * 4, ScaleShiftK{full01.p, synf.p, 2.f, 1.0f}, "img_to_pm1_full");
  }

  // persistent across stages
  TV flow_up = A.tensor(2 * B, H, W, 2);            // [f01 ; f10]
  TV fmap = A.tensor(2 * B, h, w, 256);             // fnet features of [I0 ; I1]
  TV feat4 = A.tensor(2 * B, H4, W4, 128), feat8 = A.tensor(2 * B, h, w, 256);  // projected context features
  float* scaler = A.alloc_f(std::max(B, 64));

  // ------------------------------------------------------------ RAFT (both directions batched)
  // The recurrence amplifies operand rounding -> fp32-class arithmetic only: 3xTF32 tensor cores or CUDA cores.
  cx.tc = tc_mode_ >= 2; cx.tc_split = true;
  TV fproj = A.tensor(2 * B, h, w, 256);            // the maps BidirCorrBlock correlates (R: amt_fproj(fnet map); F: the estimator's own)
  if (fin) {
    // external flow estimator (GIMM-VFI-F): NCHW -> the engine's NHWC buffers, samples [frame-0 batch ; frame-1 batch]
    for (int j = 0; j < 2; ++j) {
      nchw_to_nhwc(cx, fin->flows + (int64_t)j * H * W, (int64_t)4 * H * W, (int64_t)2 * H * W, flow_up.batch(j * B, B), 1.f, 0.f);
      nchw_to_nhwc(cx, fin->feat4[j], (int64_t)128 * H4 * W4, (int64_t)H4 * W4, feat4.batch(j * B, B), 1.f, 0.f);
      nchw_to_nhwc(cx, fin->feat8[j], (int64_t)256 * h * w, (int64_t)h * w, feat8.batch(j * B, B), 1.f, 0.f);
      nchw_to_nhwc(cx, fin->fnet[j], (int64_t)256 * h * w, (int64_t)h * w, fproj.batch(j * B, B), 1.f, 0.f);
    }
  } else if (ff_) {
    // GIMM-VFI-F: the native FlowFormer estimator fills the same four products (gimmvfi_f.py:114-138); F has no feature projections
    run_flowformer(cx, N, B, raft_in, flow_up, feat4, feat8, fproj);
  } else {
    const size_t mk = A.mark();
    TV hx = A.tensor(2 * B, h, w, 384);              // [h | inp | motion]
    TV c4, c8;
    // frame cache (set_frame_cache): with `load` the encoders only see the second frames, samples [B, 2B)
    const bool fc_on = fc_ != nullptr && !cx.dry;
    if (fc_on && fc_bytes_ < frame_cache_bytes(P)) throw std::runtime_error("gimmvfi: frame cache smaller than frame_cache_bytes()");
    const bool fload = fc_on && fc_load_, fstore = fc_on && fc_store_;
    const auto fc_it = fc_valid_.find(fc_);
    if (fload && (fc_it == fc_valid_.end() || fc_it->second.B != B || fc_it->second.H != H || fc_it->second.W != W || fc_it->second.mode != tc_mode_))
      throw std::runtime_error("gimmvfi: frame cache load requested, but this buffer does not hold a frame stored by a forward of the same "
                               "problem size and precision mode");
    const int e0 = fload ? B : 0, en = fload ? B : 2 * B;
    {
      TV rin = raft_in.batch(e0, en);
      TV raft_pad = A.tensor(en, H + 6, W + 6, 3, 4);
      pad_image4(cx, rin, raft_pad, 3);
      if (precise_ & 32) cx.tc_split = false;   // experiment: RAFT encoders on plain TF32
      raft_encoder(N, "flow_estimator.fnet", true, rin, raft_pad, fmap.batch(e0, en), nullptr, nullptr, nullptr, TV(), TV());
      // the encoder's temporaries stay allocated until `mk` is released (bump allocator)
      const ConvW& wc = N.W("flow_estimator.cnet.conv2");
      raft_encoder(N, "flow_estimator.cnet", false, rin, raft_pad, TV(), &c4, &c8, &wc, hx.batch(e0, en).slice(0, 128), hx.batch(e0, en).slice(128, 128));
    }
    cx.tc_split = true;
    // context features for the synthesis net (gimmvfi_r.py:134-141)
    N.conv("amt_second_last_cproj", c4, feat4.batch(e0, en));
    N.conv("amt_last_cproj", c8, feat8.batch(e0, en));
    if (fc_on) {
      float* q = fc_;
      TV c_fmap = make_tv(q, B, h, w, 256); q += (size_t)B * h * w * 256;
      TV c_ni = make_tv(q, B, h, w, 256); q += (size_t)B * h * w * 256;
      TV c_f4 = make_tv(q, B, H4, W4, 128); q += (size_t)B * H4 * W4 * 128;
      TV c_f8 = make_tv(q, B, h, w, 256);
      if (fload) {
        copy_channels(cx, c_fmap, fmap.batch(0, B)); copy_channels(cx, c_ni, hx.batch(0, B).slice(0, 256));
        copy_channels(cx, c_f4, feat4.batch(0, B)); copy_channels(cx, c_f8, feat8.batch(0, B));
      }
      if (fstore) fc_valid_[fc_] = FcRec{B, H, W, tc_mode_};
      else if (!fload) fc_valid_.erase(fc_);
      if (fstore) {   // (before the GRU overwrites `net` in place)
        copy_channels(cx, fmap.batch(B, B), c_fmap); copy_channels(cx, hx.batch(B, B).slice(0, 256), c_ni);
        copy_channels(cx, feat4.batch(B, B), c_f4); copy_channels(cx, feat8.batch(B, B), c_f8);
      }
    }
    tap("raft.fmap", fmap);

    // RAFT's correlation: the all-pairs volume pyramid (fp32-class GEMMs; 8.3 ms of GEMM + 20 lookups at 1080p) - or, when that
    // pyramid would not fit the memory budget (a 4K pair without ds_factor: 2 x 78 GB at level 0 alone), NO volume: every lookup
    // computes the 4 x 100 dot products its window needs from the other frame's (pooled, half-precision) features, like
    // BidirCorrBlock's (the B200 successor of alt_cuda_corr).  ~5x the time of the volume path, 0 bytes instead of N^2.
    const double pyr_bytes = 2.0 * B * (double)h * w * (double)h * w * 4.0 * (4.0 / 3.0);
    const bool raft_direct = (raft_direct_ == 1 || (raft_direct_ < 0 && pyr_bytes > raft_direct_auto_bytes_)) && fmap.c == 256 && h >= 16 && w >= 16;
    Pyramid pyr{}; CorrFeat rfeat{};
    if (raft_direct) {
      TV Fl = fmap;
      rfeat.c = 256; rfeat.scale = 1.0f / std::sqrt((float)fmap.c);
      for (int l = 0; l < 4; ++l) {
        rfeat.h[l] = Fl.h; rfeat.w[l] = Fl.w;
        void* hp = A.alloc_f(((size_t)Fl.pixels() * Fl.c + 1) / 2);
        features_to_half(cx, Fl, hp);
        rfeat.lvl[l] = static_cast<const uint16_t*>(hp);
        if (l < 3) { TV nx = A.tensor(2 * B, Fl.h / 2, Fl.w / 2, Fl.c); avgpool2_features(cx, Fl, nx); Fl = nx; }
      }
    } else {
      pyr = build_pyramid(cx, fmap, B, tc_mode_ >= 2 ? 2 : 0);   // fp32-class only
    }
    auto rfeat_of = [&](int sample0) {
      CorrFeat f = rfeat;
      for (int l = 0; l < 4; ++l) f.lvl[l] = rfeat.lvl[l] + (int64_t)sample0 * f.h[l] * f.w[l] * f.c;
      return f;
    };
    const std::string u = "flow_estimator.update_block";
    TV coords1 = A.tensor(2 * B, h, w, 2);
    TV flow = A.tensor(2 * B, h, w, 2, 4);   // ld 4: 16-byte pixel stride so the 7x7 2->128 conv can use TMA
    TV corr = A.tensor(2 * B, h, w, 324), cor1 = A.tensor(2 * B, h, w, 256), corflo = A.tensor(2 * B, h, w, 256);
    TV flo1 = A.tensor(2 * B, h, w, 128), zb = A.tensor(2 * B, h, w, 128), rh = A.tensor(2 * B, h, w, 128);
    TV fh = A.tensor(2 * B, h, w, 256), mask = A.tensor(2 * B, h, w, 576);
    init_coords(cx, coords1);
    TV hcur = hx.slice(0, 128), xin = hx.slice(128, 256), mot = hx.slice(256, 128);
    // iteration-invariant share of the gate convolutions (tensor-core modes; see finalize_weights)
    const bool hoist = cx.tc && gru_hoist_;
    TV Pzr[2], Pq[2];
    if (hoist)
      for (int s = 0; s < 2; ++s) {
        const std::string sfx = s == 0 ? "1" : "2";
        Pzr[s] = A.tensor(2 * B, h, w, 256); Pq[s] = A.tensor(2 * B, h, w, 128);
        N.conv(u + ".gru.convzr" + sfx + "_inp", hx.slice(128, 128), Pzr[s]);
        N.conv(u + ".gru.convq" + sfx + "_inp", hx.slice(128, 128), Pq[s]);
      }
    for (int it = 0; it < raft_iters; ++it) {
      if (raft_direct) {   // direction 0's sources against frame 1's features and vice versa
        corr_lookup_direct(cx, fmap.batch(0, B), rfeat_of(B), coords1.batch(0, B), corr.batch(0, B));
        corr_lookup_direct(cx, fmap.batch(B, B), rfeat_of(0), coords1.batch(B, B), corr.batch(B, B));
      } else {
        corr_lookup(cx, pyr.view(0), coords1, corr);
      }
      coords_minus_grid(cx, coords1, flow, hx.slice(382, 2));
      if (it == 0) tap("raft.corr_it0", corr);
      // BasicMotionEncoder raft/update.py:94-112
      N.conv(u + ".encoder.convc1", corr, cor1, ACT_RELU);
      N.conv(u + ".encoder.convc2", cor1, corflo.slice(0, 192), ACT_RELU);
      N.conv7x(u + ".encoder.convf1", flow, flo1, ACT_RELU);
      N.conv(u + ".encoder.convf2", flo1, corflo.slice(192, 64), ACT_RELU);
      N.conv(u + ".encoder.conv", corflo, hx.slice(256, 126), ACT_RELU);
      // SepConvGRU raft/update.py:35-73 (horizontal 1x5 then vertical 5x1)
      for (const char* sfx : {"1", "2"}) {
        if (hoist) {
          const int s = sfx[0] - '1';
          ConvEpi ezr; ezr.res = Pzr[s]; ezr.act2 = ACT_SIGMOID; ezr.mul = hcur; ezr.out2 = rh; ezr.split_c = 128;
          N.conv_e(u + ".gru.convzr" + sfx + "_hm", hcur, mot, zb, ezr);
          ConvEpi eq; eq.res = Pq[s]; eq.act2 = ACT_TANH; eq.gru_z = zb; eq.gru_h = hcur;
          N.conv_e(u + ".gru.convq" + sfx + "_hm", rh, mot, hcur, eq);
          continue;
        }
        if (cx.tc) {   // one launch for both gates: 1020 tiles instead of 2 x 510 (3.45 waves of 148 SMs each)
          ConvEpi ezr; ezr.act1 = ACT_SIGMOID; ezr.mul = hcur; ezr.out2 = rh; ezr.split_c = 128;
          N.conv_e(u + ".gru.convzr" + sfx, hx, TV(), zb, ezr);
        } else {
          ConvEpi ez; ez.act1 = ACT_SIGMOID;
          N.conv_e(u + ".gru.convz" + sfx, hx, TV(), zb, ez);
          ConvEpi er; er.act1 = ACT_SIGMOID; er.mul = hcur;
          N.conv_e(u + ".gru.convr" + sfx, hx, TV(), rh, er);
        }
        ConvEpi eq; eq.act1 = ACT_TANH; eq.gru_z = zb; eq.gru_h = hcur;
        N.conv_e(u + ".gru.convq" + sfx, rh, xin, hcur, eq);
      }
      // FlowHead raft/update.py:6-14; coords1 += delta_flow (raft/raft.py:153)
      N.conv(u + ".flow_head.conv1", hcur, fh, ACT_RELU);
      { ConvEpi e; e.res = coords1; N.conv_e(u + ".flow_head.conv2", fh, TV(), coords1, e); }
      if (it == 0 || it == 4) { /* debug snapshots are cheap copies */
        if (debug_) { TV snap = A.tensor(2 * B, h, w, 2); coords_minus_grid(cx, coords1, snap, TV()); tap("raft.lowres_flow_it" + std::to_string(it), snap); }
      }
      if (it == raft_iters - 1) {
        // only the last iteration's mask / upsample is consumed (raft/raft.py:159-167)
        N.conv(u + ".mask.0", hcur, fh, ACT_RELU);
        N.conv(u + ".mask.2", fh, mask);
        coords_minus_grid(cx, coords1, flow, TV());
        tap("raft.lowres_flow_final", flow);
        convex_upsample(cx, flow, mask, flow_up);
      }
    }
    tap("raft.net_final", hcur);
    if (!debug_) A.release(mk);
  }
  if (io.raft_flow)
    for (int j = 0; j < 2; ++j)
      nhwc_to_nchw(cx, flow_up.batch(j * B, B), io.raft_flow + (int64_t)j * H * W, (int64_t)4 * H * W, (int64_t)2 * H * W, 1.f, 0.f, 0);

  // Everything downstream of RAFT tolerates TF32 operands (DESIGN.md precision plan).
  cx.tc = tc_mode_ >= 1; cx.tc_split = false;
  // ------------------------------------------------------------ bidirectional volume on projected features
  if (!fin && !ff_) N.conv("amt_fproj", fmap, fproj);
  // the bidirectional volume only feeds TF32 layers (AMT update blocks) -> plain TF32 is at their input precision
  // Few interpolated frames per pair (T <= corr_direct_max_t_): no volume at all - the lookup computes the dot products its window needs
  // (corr.cu "volume-free lookup").  Otherwise the all-pairs pyramid, looked up T times.
  Pyramid bpyr{}; CorrFeat bfeat{};
  const bool corr_direct = tc_mode_ >= 1 && T <= corr_direct_max_t_ && fproj.c == 256 && h >= 16 && w >= 16;
  if (corr_direct) {
    TV Fl = fproj;
    bfeat.c = 256; bfeat.scale = 1.0f / std::sqrt((float)fproj.c);
    for (int l = 0; l < 4; ++l) {
      bfeat.h[l] = Fl.h; bfeat.w[l] = Fl.w;
      void* hp = A.alloc_f(((size_t)Fl.pixels() * Fl.c + 1) / 2);
      features_to_half(cx, Fl, hp);
      bfeat.lvl[l] = static_cast<const uint16_t*>(hp);
      if (l < 3) { TV nx = A.tensor(2 * B, Fl.h / 2, Fl.w / 2, Fl.c); avgpool2_features(cx, Fl, nx); Fl = nx; }
    }
  } else {
    bpyr = build_pyramid(cx, fproj, B, tc_mode_ >= 1 ? 1 : 0);   // gimmvfi_r.py:133, raft/corr.py:23-44
  }
  auto bfeat_of = [&](int sample0) {
    CorrFeat f = bfeat;
    for (int l = 0; l < 4; ++l) f.lvl[l] = bfeat.lvl[l] + (int64_t)sample0 * f.h[l] * f.w[l] * f.c;
    return f;
  };

  // ------------------------------------------------------------ hoisted (t-independent) decoder feature upsampling
  TV fup4 = A.tensor(2 * B, H4, W4, 128);   // NewInitDecoder.upsample   fi_components.py:234-244
  TV fup1 = half_chains(cx) ? A.tensor_h(2 * B, H, W, 64) : A.tensor(2 * B, H, W, 64);      // NewMultiFlowDecoder.upsample fi_components.py:284-295
  {
    const size_t mk = A.mark();
    const std::string p = "amt_init_decoder.upsample.";
    TV a = A.tensor(2 * B, H4, W4, 64), b = A.tensor(2 * B, H4, W4, 64), c = A.tensor(2 * B, H4, W4, 128);
    pixel_shuffle(cx, feat8, a, 1);
    N.convrelu(p + "1", a, b); N.convrelu(p + "2", b, a); N.convrelu(p + "3", a, b); N.convrelu(p + "4", b, a);
    N.convrelu(p + "5", a, c);
    N.conv(p + "6", c, fup4, ACT_RELU);  // 1x1 + folded BatchNorm + ReLU
    A.release(mk);
  }
  {
    const size_t mk = A.mark();
    const std::string p = "amt_final_decoder.upsample.";
    const bool hs = half_chains(cx);
    TV a = A.tensor(2 * B, H, W, 8);
    TV b = hs ? A.tensor_h(2 * B, H, W, 32) : A.tensor(2 * B, H, W, 32), c = A.tensor_like(b, 32), d = A.tensor_like(b, 64);
    pixel_shuffle(cx, feat4, a, 2);
    if (cx.tc) N.conv7x(p + "2.0", a, b, ACT_PRELU, N.V(p + "2.1.weight"));   // x-packed 5x5 (pack_xpacked)
    else N.convrelu(p + "2", a, b);
    N.convrelu(p + "3", b, c); N.convrelu(p + "4", c, b); N.convrelu(p + "5", b, c);
    N.convrelu(p + "6", c, d);
    N.conv(p + "7", d, fup1, ACT_RELU);
    A.release(mk);
  }

  // ------------------------------------------------------------ GIMM (t-independent part)  gimmvfi_r.py:158-168
  TV f01 = flow_up.batch(0, B), f10 = flow_up.batch(B, B);
  {
    float* sc = A.alloc_f((size_t)absmax_scratch_floats(f01));
    absmax_per_sample(cx, f01, f10, scaler, sc);
  }
  flow_absmax_ = scaler;   // per-sample max |flow| over both directions: also bounds the forward splat's reach (gimm_decode)
  TV nf = A.tensor(2 * B, H, W, 2);
  normalize_flow_pair(cx, f01, f10, scaler, nf.batch(0, B), nf.batch(B, B));
  if (io.nflow)
    for (int j = 0; j < 2; ++j)
      nhwc_to_nchw(cx, nf.batch(j * B, B), io.nflow + (int64_t)j * H * W, (int64_t)4 * H * W, (int64_t)2 * H * W, 1.f, 0.f, 0);
  TV wts = A.tensor(2 * B, H, W, 1);
  TV X64 = A.tensor(B, H, W, 64);   // [lat0 | lat1 | splat0 | splat1]
  cx.tc_split = (knobs & 1) != 0;
  gimm_encode(N, nf, f01, f10, wts, X64);
  cx.tc_split = false;

  // ------------------------------------------------------------ per-timestep: GIMM decode + frame synthesis
  TV grid_flow = A.tensor(B, h, w, 4);   // flow_4_lr = cat(fl0, fl1)
  const size_t t_mark = A.mark();
  for (int ti = 0; ti < T; ++ti) {
    A.release(t_mark);
    const float* tdev = io.t + (int64_t)ti * B;
    TV ninr = A.tensor(B, H, W, 2);
    cx.tc_split = (knobs & 1) != 0;
    gimm_decode(N, X64, f01, f10, wts, tdev, io.coords + (int64_t)ti * B * P.Hc * P.Wc * 3, ninr, ti == 0);
    cx.tc_split = (knobs & 4) != 0;   // init decoder + update blocks
    if (io.ninrflow) nhwc_to_nchw(cx, ninr, io.ninrflow + (int64_t)ti * B * 2 * H * W, (int64_t)2 * H * W, (int64_t)H * W, 1.f, 0.f, 0);
    TV flow_t = A.tensor(B, H, W, 2);
    unnormalize_flow(cx, ninr, scaler, flow_t);
    if (io.flowt) nhwc_to_nchw(cx, flow_t, io.flowt + (int64_t)ti * B * 2 * H * W, (int64_t)2 * H * W, (int64_t)H * W, 1.f, 0.f, 0);

    // ---- frame_synthesize (gimmvfi_r.py:222-322)
    TV s0 = syn.batch(0, B), s1 = syn.batch(B, B);
    TV ft0 = A.tensor(B, H, W, 2), ft1 = A.tensor(B, H, W, 2);
    scale_flow_t(cx, flow_t, tdev, ft0, ft1);
    // NewInitDecoder (fi_components.py:255-276)
    TV fin = A.tensor(B, H4, W4, 272);
    TV f0in = fin.slice(256, 2), f1in = fin.slice(258, 2);
    resize_bilinear(cx, ft0, f0in, 4.f, 4.f, 0.25f, 0, ACT_NONE);
    resize_bilinear(cx, ft1, f1in, 4.f, 4.f, 0.25f, 0, ACT_NONE);
    backwarp(cx, fup4.batch(0, B), f0in, fin.slice(0, 128));
    backwarp(cx, fup4.batch(B, B), f1in, fin.slice(128, 128));
    resize_bilinear(cx, s0, fin.slice(260, 3), 4.f, 4.f, 1.f, 0, ACT_NONE);
    resize_bilinear(cx, s1, fin.slice(263, 3), 4.f, 4.f, 1.f, 0, ACT_NONE);
    backwarp(cx, fin.slice(260, 3), f0in, fin.slice(266, 3));
    backwarp(cx, fin.slice(263, 3), f1in, fin.slice(269, 3));
    TV x4 = half_chains(cx) ? A.tensor_h(B, H4, W4, 128) : A.tensor(B, H4, W4, 128);   // (mode 4: the H/4 residual trunk in half too)
    N.convrelu("amt_init_decoder.convblock.0", fin, x4);
    for (int k = 1; k <= 3; ++k) resblock(N, "amt_init_decoder.convblock." + std::to_string(k), x4, 128, 64);
    TV o136 = A.tensor(B, H4, W4, 133, 136);
    N.conv("amt_init_decoder.convblock.4", x4, o136);
    TV ft_4 = o136.slice(0, 128);
    TV fl4 = A.tensor(B, H4, W4, 4), mask4 = A.tensor(B, H4, W4, 1);
    flow_mask_split(cx, o136, f0in, f1in, fl4.slice(0, 2), fl4.slice(2, 2), mask4);
    // warp_w_mask(scale=4) -> img_warp_4 (aux output, gimmvfi_r.py:259-261)
    if (io.img_warp_4) {
      const size_t mk = A.mark();
      TV a0 = A.tensor(B, H, W, 2), a1 = A.tensor(B, H, W, 2), am = A.tensor(B, H, W, 1), wb = A.tensor(B, H, W, 3, 4);
      resize_bilinear(cx, fl4.slice(0, 2), a0, 0.25f, 0.25f, 4.f, 0, ACT_NONE);
      resize_bilinear(cx, fl4.slice(2, 2), a1, 0.25f, 0.25f, 4.f, 0, ACT_NONE);
      resize_bilinear(cx, mask4, am, 0.25f, 0.25f, 1.f, 0, ACT_SIGMOID);
      warp_blend(cx, s0, s1, a0, a1, am, wb);
      nhwc_to_nchw(cx, wb, io.img_warp_4 + (int64_t)ti * B * 3 * H * W, (int64_t)3 * H * W, (int64_t)H * W, 0.5f, 1.0f, 1);
      A.release(mk);
    }
    // _amt_corr_scale_lookup(downsample=2)  (gimmvfi_r.py:494-507)
    TV corr648 = A.tensor(B, h, w, 648);
    {
      resize_bilinear(cx, fl4, grid_flow, 2.f, 2.f, 0.5f, 0, ACT_NONE);
      TV cA = A.tensor(B, h, w, 2), cB = A.tensor(B, h, w, 2);
      lookup_coords(cx, grid_flow.slice(2, 2), tdev, 0, cA);  // coord + flow1 * 1/(1-t) -> volume
      lookup_coords(cx, grid_flow.slice(0, 2), tdev, 1, cB);  // coord + flow0 * 1/t     -> transposed volume
      if (corr_direct) {   // rows = frame 0's pixels against frame 1's features, and the transposed volume's counterpart
        corr_lookup_direct(cx, fproj.batch(0, B), bfeat_of(B), cA, corr648.slice(0, 324));
        corr_lookup_direct(cx, fproj.batch(B, B), bfeat_of(0), cB, corr648.slice(324, 324));
      } else {
        corr_lookup(cx, bpyr.view(0), cA, corr648.slice(0, 324));
        corr_lookup(cx, bpyr.view(B), cB, corr648.slice(324, 324));
      }
    }
    amt_update(N, "amt_update4_low", true, ft_4, fl4, grid_flow, corr648);
    {
      TV corr_hi = A.tensor(B, H4, W4, 648);
      resize_bilinear(cx, corr648, corr_hi, 0.5f, 0.5f, 1.f, 0, ACT_NONE);  // gimmvfi_r.py:274
      amt_update(N, "amt_update4_high", false, ft_4, fl4, fl4, corr_hi);
      if (ti == 0) tap("syn.corr_4_hi", corr_hi);
    }
    if (ti == 0) { tap("syn.ft_4", ft_4); tap("syn.fl4", fl4); }
    if (io.flowt0_4) nhwc_to_nchw(cx, fl4.slice(0, 2), io.flowt0_4 + (int64_t)ti * B * 2 * H4 * W4, (int64_t)2 * H4 * W4, (int64_t)H4 * W4, 1.f, 0.f, 0);
    if (io.flowt1_4) nhwc_to_nchw(cx, fl4.slice(2, 2), io.flowt1_4 + (int64_t)ti * B * 2 * H4 * W4, (int64_t)2 * H4 * W4, (int64_t)H4 * W4, 1.f, 0.f, 0);

    // NewMultiFlowDecoder (fi_components.py:307-340)
    cx.tc_split = (knobs & 8) != 0;
    TV F0 = A.tensor(B, H, W, 6, 8), F1 = A.tensor(B, H, W, 6, 8), Mk = A.tensor(B, H, W, 3, 4), Rs = A.tensor(B, H, W, 9, 12);
    {
      const size_t mk = A.mark();
      // precision mode 4: the 273-channel concat buffer itself is half precision (written by the half-aware resize / backwarp /
      // copy kernels); the up-sampled flows and mask are additionally kept in fp32 (they drive the warps and the output heads)
      const bool hs = half_chains(cx);
      TV fin1 = hs ? A.tensor_h(B, H, W, 273, 280) : A.tensor(B, H, W, 273, 276);
      TV aux5 = hs ? A.tensor(B, H, W, 5, 8) : TV();
      TV fl0 = hs ? aux5.slice(0, 2) : fin1.slice(256, 2), fl1 = hs ? aux5.slice(2, 2) : fin1.slice(258, 2), mk1 = hs ? aux5.slice(4, 1) : fin1.slice(260, 1);
      resize_bilinear(cx, fl4.slice(0, 2), fl0, 0.25f, 0.25f, 4.f, 0, ACT_NONE);
      resize_bilinear(cx, fl4.slice(2, 2), fl1, 0.25f, 0.25f, 4.f, 0, ACT_NONE);
      resize_bilinear(cx, ft_4, fin1.slice(0, 128), 0.25f, 0.25f, 1.f, 0, ACT_NONE);
      resize_bilinear(cx, mask4, mk1, 0.25f, 0.25f, 1.f, 0, ACT_NONE);
      if (hs) copy_channels(cx, aux5.slice(0, 5), fin1.slice(256, 5));
      backwarp(cx, fup1.batch(0, B), fl0, fin1.slice(128, 64));
      backwarp(cx, fup1.batch(B, B), fl1, fin1.slice(192, 64));
      copy_channels(cx, s0, fin1.slice(261, 3));
      copy_channels(cx, s1, fin1.slice(264, 3));
      backwarp(cx, s0, fl0, fin1.slice(267, 3));
      backwarp(cx, s1, fl1, fin1.slice(270, 3));
      // precision mode 3: the 256-channel residual trunk (9 of the 10 heaviest launches) is stored in fp16 and runs on the
      // kind::f16 tensor-core path (same 10-bit mantissa as TF32, half the bytes per operand, twice the MMA rate)
      TV x1 = (cx.tc && tc_mode_ >= 3 && !(knobs & 8)) ? A.tensor_h(B, H, W, 256) : A.tensor(B, H, W, 256);
      N.convrelu("amt_final_decoder.convblock.0", fin1, x1);
      for (int k = 1; k <= 3; ++k) resblock(N, "amt_final_decoder.convblock." + std::to_string(k), x1, 256, 64);
      TV o24 = A.tensor(B, H, W, 24);
      N.conv("amt_final_decoder.convblock.4", x1, o24);
      final_heads(cx, o24, fl0, fl1, mk1, F0, F1, Mk, Rs);
      A.release(mk);
    }
    cx.tc_split = (knobs & 16) != 0;   // comb block
    TV F0f = F0, F1f = F1, Mkf = Mk, Rsf = Rs;
    if (ds) {  // gimmvfi_r.py:294-303
      F0f = A.tensor(B, Hf, Wf, 6, 8); F1f = A.tensor(B, Hf, Wf, 6, 8); Mkf = A.tensor(B, Hf, Wf, 3, 4); Rsf = A.tensor(B, Hf, Wf, 9, 12);
      // (views widened to the pixel stride: the padding lanes ride along so that the 16-byte form of the kernel applies - 6 / 3 / 9
      //  channels are not multiples of 4; the scalar form was 12.6 ms of the 4K configuration's 208)
      auto wide = [](TV t) { t.c = t.ld; return t; };
      resize_bilinear(cx, wide(F0), wide(F0f), rinv, rinv, inv, 0, ACT_NONE);
      resize_bilinear(cx, wide(F1), wide(F1f), rinv, rinv, inv, 0, ACT_NONE);
      resize_bilinear(cx, wide(Mk), wide(Mkf), rinv, rinv, 1.f, 0, ACT_NONE);
      resize_bilinear(cx, wide(Rs), wide(Rsf), rinv, rinv, 1.f, 0, ACT_NONE);
    }
    if (io.flowt0_1) nhwc_to_nchw(cx, F0f, io.flowt0_1 + (int64_t)ti * B * 6 * Hf * Wf, (int64_t)6 * Hf * Wf, (int64_t)Hf * Wf, 1.f, 0.f, 0);
    if (io.flowt1_1) nhwc_to_nchw(cx, F1f, io.flowt1_1 + (int64_t)ti * B * 6 * Hf * Wf, (int64_t)6 * Hf * Wf, (int64_t)Hf * Wf, 1.f, 0.f, 0);
    // multi_flow_combine + amt_comb_block at FULL resolution (fi_components.py:57-94)
    {
      TV w9 = A.tensor(B, Hf, Wf, 9, 12), mean3 = A.tensor(B, Hf, Wf, 3, 4), c18 = A.tensor(B, Hf, Wf, 18, 20), c3 = A.tensor(B, Hf, Wf, 3, 4);
      multi_flow_blend(cx, synf.batch(0, B), synf.batch(B, B), F0f, F1f, Mkf, Rsf, w9, mean3);
      N.conv7x("amt_comb_block.0", w9, c18, ACT_PRELU, N.V("amt_comb_block.1.weight"));
      N.conv7x("amt_comb_block.2", c18, c3);
      combine_output(cx, mean3, c3, io.imgt_pred + (int64_t)ti * B * 3 * Hf * Wf);
    }
    cx.tc_split = false;
  }
}


// ---------------------------------------------------------------------------
// GIMM stages (gimm.py:129-214 == gimmvfi_r.py:158-211), shared by run() and run_gimm()
// ---------------------------------------------------------------------------
// t-independent part: splatting metrics (gimm.py:82-127) and the motion encoder on both normalised flows -> X64[0:32]
void Engine::gimm_encode(Net& N, const TV& nf /*2B: [nf01; nf10]*/, const TV& f01, const TV& f10, const TV& wts /*2B,1*/, const TV& X64) {
  Ctx& cx = N.cx; Arena& A = cx.arena;
  const int B = f01.n, H = f01.h, W = f01.w;
  splat_weights(cx, f01, f10, g9_, alpha_fe_, alpha_v_, wts.batch(0, B));
  splat_weights(cx, f10, f01, g9_, alpha_fe_, alpha_v_, wts.batch(B, B));
  tap("gimm.splat_w", wts);
  {
    const size_t mk = A.mark();
    const bool hs = half_chains(cx);   // precision mode 4: the 32-channel full-resolution chain lives in half precision
    TV a = A.tensor(2 * B, H, W, 16);
    TV x = hs ? A.tensor_h(2 * B, H, W, 32) : A.tensor(2 * B, H, W, 32), y = A.tensor_like(x, 32), m = A.tensor_like(x, 32);
    N.conv("cnn_encoder.0", nf, a);
    N.conv("cnn_encoder.1", a, x, ACT_LRELU);
    for (int i = 3; i <= 5; ++i) {  // LateralBlock fi_components.py:17-29 (+ the LeakyReLU after the last one)
      const std::string q = "cnn_encoder." + std::to_string(i);
      N.conv(q + ".layers.0", x, m, ACT_LRELU);
      ConvEpi e; e.res = x; e.act2 = (i == 5) ? ACT_LRELU : ACT_NONE;
      N.conv_e(q + ".layers.2", m, TV(), y, e);
      std::swap(x, y);
    }
    N.conv("cnn_encoder.7", x.batch(0, B), X64.slice(0, 16), ACT_NONE, nullptr, 1, true);
    N.conv("cnn_encoder.7", x.batch(B, B), X64.slice(16, 16), ACT_NONE, nullptr, 1, true);
    A.release(mk);
  }
  tap("gimm.lat0", X64.slice(0, 16));
}

// one timestep: forward splat of both latents to time t, latent refiner, HypoNet -> normalised flow (B,H,W,2)
void Engine::gimm_decode(Net& N, const TV& X64, const TV& f01, const TV& f10, const TV& wts, const float* tdev, const float* coords_t,
                         const TV& ninr, bool tap_it) {
  Ctx& cx = N.cx; Arena& A = cx.arena;
  const int B = f01.n, H = f01.h, W = f01.w;
  const size_t mk0 = A.mark();
  // ---- forward splat of both latents to time t (gimmvfi_r.py:171-193)
  // one pass per direction (target tiles in shared memory, bounded by the per-sample max |flow|); else memset + vector reductions + normalise
  const bool one_pass = flow_absmax_ != nullptr &&
                        softsplat_fused(cx, X64.slice(0, 16), f01, wts.batch(0, B), tdev, 0, flow_absmax_, X64.slice(32, 16)) &&
                        softsplat_fused(cx, X64.slice(16, 16), f10, wts.batch(B, B), tdev, 1, flow_absmax_, X64.slice(48, 16));
  if (!one_pass) {
    TV acc = A.tensor(2 * B, H, W, 17, 20);
    if (!cx.dry) dev_memset(acc.p, 0, (size_t)2 * B * H * W * 20 * sizeof(float), cx.stream);
    softsplat_accumulate(cx, X64.slice(0, 16), f01, wts.batch(0, B), tdev, 0, acc.batch(0, B));
    softsplat_accumulate(cx, X64.slice(16, 16), f10, wts.batch(B, B), tdev, 1, acc.batch(B, B));
    softsplat_normalize(cx, acc.batch(0, B), X64.slice(32, 16));
    softsplat_normalize(cx, acc.batch(B, B), X64.slice(48, 16));
  }
  if (tap_it) tap("gimm.splat0", X64.slice(32, 16));
  // HypoNet (hyponet.py:71-146).  Tensor-core modes: ONE fused kernel (hyponet.cu) that takes the refined latent and the caller's
  // coordinate tensor directly; fp32 mode: five 1x1 convolutions on a packed [latent32 | t,y,x] input.
  const bool fused = cx.tc && hypo_blob_ != nullptr && !(precise_ & 2);
  const bool split_saved = cx.tc_split;
  TV hin = fused ? A.tensor(B, H, W, 32) : A.tensor(B, H, W, 35, 36);
  {
    const size_t mk = A.mark();
    const bool hs = half_chains(cx);
    TV a = hs ? A.tensor_h(B, H, W, 32) : A.tensor(B, H, W, 32);
    TV x = A.tensor_like(a, 64), m = A.tensor_like(a, 64), y = A.tensor_like(a, 64);
    N.conv("res_conv.0", X64, a);
    N.conv("res_conv.1", a, x, ACT_LRELU);
    N.conv("res_conv.3.layers.0", x, m, ACT_LRELU);
    { ConvEpi e; e.res = x; e.act2 = ACT_LRELU; N.conv_e("res_conv.3.layers.2", m, TV(), y, e); }
    { ConvEpi e; e.res = X64.slice(32, 32); N.conv_e("res_conv.5", y, TV(), hin.slice(0, 32), e, 1, true); }
    A.release(mk);
  }
  if (tap_it) tap("gimm.latent", hin.slice(0, 32));
  if (fused) {
    // fp32-class by default: the flow is (2 o - 1) * max|flow|, so operand rounding at 2^-11 becomes ~1e-2 px at 40 px motion and
    // moves real frames by > 1e-3 (tests/test_bench_parity_gpu.py, demo frames); hypo_fast_ selects the TF32 / half-operand kernel
    if (hypo_fast_) hyponet_fused(cx, hin, coords_t, hypo_blob_, ninr);
    else hyponet_fused3(cx, hin, coords_t, hypo_blob3_, ninr);
  } else {
    cx.tc_split = (precise_ & 2) != 0;
    hypo_pack_input(cx, coords_t, hin.slice(32, 3));
    const size_t mk = A.mark();
    TV a = A.tensor(B, H, W, 128), b = A.tensor(B, H, W, 128);
    N.conv("hyponet.params_dict.linear_wb0", hin, a, ACT_SIN);
    N.conv("hyponet.params_dict.linear_wb1", a, b, ACT_SIN);
    N.conv("hyponet.params_dict.linear_wb2", b, a, ACT_SIN);
    N.conv("hyponet.params_dict.linear_wb3", a, b, ACT_SIN);
    N.conv("hyponet.params_dict.linear_wb4", b, ninr);
    A.release(mk);
    cx.tc_split = split_saved;
  }
  if (!debug_) A.release(mk0);   // (debug taps keep pointing into this region)
}

// (B,2,2,H,W) flow pair, reference layout [b][c][j][y][x] -> NHWC (2B,H,W,2), sample j*B+b
struct FlowPairLoadK {
  const float* src; TV dst; int B;
  GV_HD void operator()(int64_t i) const {
    const int c = (int)(i & 1); int64_t r = i >> 1;
    const int x = (int)(r % dst.w); r /= dst.w; const int y = (int)(r % dst.h); const int n = (int)(r / dst.h);
    const int j = n / B, b = n - j * B;
    const int64_t hw = (int64_t)dst.h * dst.w;
    dst.p[dst.off(n, y, x) + c] = src[(((int64_t)b * 2 + c) * 2 + j) * hw + (int64_t)y * dst.w + x];
  }
};

// GIMM.forward (gimm.py:129-214): normalised flows xs + raw flows ori_flow + timesteps + coords -> normalised flows at t
void Engine::run_gimm(Ctx& cx, const Problem& P, const GimmIO& io) {
  if (P.B < 1 || P.T < 1) throw std::runtime_error("gimmvfi: B and T must be >= 1");
  const int B = P.B, H = P.Hf, W = P.Wf, T = P.T;
  if (P.ds > 0.f) throw std::runtime_error("gimm: ds_factor does not exist on this path");
  if (H < 8 || W < 8) throw std::runtime_error("gimm: H and W must be >= 8");
  if (P.Hc != H || P.Wc != W) throw std::runtime_error("gimm: the coordinate grid must match the flow resolution");
  Net N{*this, cx};
  Arena& A = cx.arena;
  cx.tc = tc_mode_ >= 1; cx.tc_split = false;   // downstream-of-RAFT arithmetic (DESIGN.md precision plan)
  TV nf = A.tensor(2 * B, H, W, 2), fl = A.tensor(2 * B, H, W, 2);
  parallel_for(cx, nf.pixels() * 2, FlowPairLoadK{io.xs, nf, B}, "flow_pair_load");
  parallel_for(cx, fl.pixels() * 2, FlowPairLoadK{io.ori_flow, fl, B}, "flow_pair_load");
  TV f01 = fl.batch(0, B), f10 = fl.batch(B, B);
  TV wts = A.tensor(2 * B, H, W, 1);
  TV X64 = A.tensor(B, H, W, 64);
  {   // bound of the splat's reach (the caller's normalisation scale is not passed in: recompute max |ori_flow| per sample)
    float* amax = A.alloc_f(std::max(B, 64));
    float* sc = A.alloc_f((size_t)absmax_scratch_floats(f01));
    absmax_per_sample(cx, f01, f10, amax, sc);
    flow_absmax_ = amax;
  }
  gimm_encode(N, nf, f01, f10, wts, X64);
  TV ninr = A.tensor(B, H, W, 2);
  for (int ti = 0; ti < T; ++ti) {
    gimm_decode(N, X64, f01, f10, wts, io.t + (int64_t)ti * B, io.coords + (int64_t)ti * B * P.Hc * P.Wc * 3, ninr, ti == 0);
    nhwc_to_nchw(cx, ninr, io.out + (int64_t)ti * B * 2 * H * W, (int64_t)2 * H * W, (int64_t)H * W, 1.f, 0.f, 0);
  }
}

size_t Engine::plan_gimm(const Problem& p) {
  if (!finalized_) throw std::runtime_error("gimmvfi: finalize_weights() has not been called");
  Ctx cx; cx.dry = true; cx.arena.dry = true; cx.sm_count = sm_count_;
  GimmIO io; float* fake = reinterpret_cast<float*>(size_t(64));
  io.xs = io.ori_flow = io.coords = io.t = fake; io.out = fake;
  const bool dbg = debug_; debug_ = false;
  run_gimm(cx, p, io);
  debug_ = dbg; taps_.clear();
  return cx.arena.peak + 256;
}

void Engine::forward_gimm(const Problem& p, const GimmIO& io, void* workspace, size_t workspace_bytes, gvStream_t stream) {
  if (!finalized_) throw std::runtime_error("gimmvfi: finalize_weights() has not been called");
  if (!io.xs || !io.ori_flow || !io.coords || !io.t || !io.out) throw std::runtime_error("gimm: xs, ori_flow, coords, t and out are required");
  DeviceGuard dg(device_);
  Ctx cx; cx.stream = stream; cx.sm_count = sm_count_;
  prof_.reset();
  cx.prof = profile_ ? &prof_ : nullptr;
  uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
  cx.arena.base = reinterpret_cast<char*>(base);
  cx.arena.cap = workspace_bytes - (base - reinterpret_cast<uintptr_t>(workspace));
  taps_.clear();
  run_gimm(cx, p, io);
  launches_ = cx.launches;
}

size_t Engine::frame_cache_bytes(const Problem& p) {
  const size_t h = (size_t)p.H() / 8, w = (size_t)p.W() / 8, H4 = (size_t)p.H() / 4, W4 = (size_t)p.W() / 4;
  return ((size_t)p.B * h * w * 256 * 3 + (size_t)p.B * H4 * W4 * 128) * sizeof(float);
}

size_t Engine::plan_from_flow(const Problem& p) {
  if (!finalized_) throw std::runtime_error("gimmvfi: finalize_weights*() has not been called");
  if (gimm_only_) throw std::runtime_error("gimmvfi: only GIMM's weights were loaded");
  Ctx cx; cx.dry = true; cx.arena.dry = true; cx.sm_count = sm_count_;
  IO io; float* fake = reinterpret_cast<float*>(size_t(64));
  io.img_xs = io.coords = io.t = fake;
  io.imgt_pred = io.img_warp_4 = io.flowt0_1 = io.flowt1_1 = io.flowt0_4 = io.flowt1_4 = io.raft_flow = io.nflow = io.ninrflow = io.flowt = fake;
  FlowInputs fin; fin.flows = fake;
  for (int j = 0; j < 2; ++j) fin.feat4[j] = fin.feat8[j] = fin.fnet[j] = fake;
  run(cx, p, io, &fin);
  taps_.clear();
  return cx.arena.peak + 256;
}

void Engine::forward_from_flow(const Problem& p, const IO& io, const FlowInputs& fin, void* workspace, size_t workspace_bytes, gvStream_t stream) {
  if (!finalized_) throw std::runtime_error("gimmvfi: finalize_weights*() has not been called");
  if (gimm_only_) throw std::runtime_error("gimmvfi: only GIMM's weights were loaded (finalize_weights_gimm)");
  if (!io.img_xs || !io.coords || !io.t || !io.imgt_pred) throw std::runtime_error("gimmvfi: img_xs, coords, t and imgt_pred are required");
  if (!fin.flows || !fin.feat4[0] || !fin.feat4[1] || !fin.feat8[0] || !fin.feat8[1] || !fin.fnet[0] || !fin.fnet[1])
    throw std::runtime_error("gimmvfi: forward_from_flow needs flows, feat4[2], feat8[2] and fnet[2]");
  DeviceGuard dg(device_);
  Ctx cx; cx.stream = stream; cx.sm_count = sm_count_;
  prof_.reset();
  cx.prof = profile_ ? &prof_ : nullptr;
  uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
  cx.arena.base = reinterpret_cast<char*>(base);
  cx.arena.cap = workspace_bytes - (base - reinterpret_cast<uintptr_t>(workspace));
  taps_.clear();
  run(cx, p, io, &fin);
  launches_ = cx.launches;
}

size_t Engine::plan(const Problem& p) {
  if (!finalized_) throw std::runtime_error("gimmvfi: finalize_weights() has not been called");
  if (synth_only_) throw std::runtime_error("gimmvfi: no flow estimator weights were loaded (finalize_weights_synthesis); use plan_from_flow");
  Ctx cx; cx.dry = true; cx.arena.dry = true; cx.sm_count = sm_count_;
  IO io;  // null pointers; dry run never dereferences them, but optional outputs must be "present"
  float* fake = reinterpret_cast<float*>(size_t(64));
  io.img_xs = io.coords = io.t = fake;
  io.imgt_pred = io.img_warp_4 = io.flowt0_1 = io.flowt1_1 = io.flowt0_4 = io.flowt1_4 = io.raft_flow = io.nflow = io.ninrflow = io.flowt = fake;
  bool dbg = debug_;
  run(cx, p, io);
  debug_ = dbg;
  taps_.clear();
  return cx.arena.peak + 256;
}

void Engine::clear_graphs() {
#ifndef GV_HOSTSIM
  DeviceGuard dg(device_);
  for (GraphEntry& g : graphs_) if (g.exec) cudaGraphExecDestroy((cudaGraphExec_t)g.exec);
#endif
  graphs_.clear();
}

void Engine::forward(const Problem& p, const IO& io, void* workspace, size_t workspace_bytes, gvStream_t stream) {
  if (!finalized_) throw std::runtime_error("gimmvfi: finalize_weights() has not been called");
  if (gimm_only_) throw std::runtime_error("gimmvfi: only GIMM's weights were loaded (finalize_weights_gimm); use gimm_forward");
  if (synth_only_) throw std::runtime_error("gimmvfi: no flow estimator weights were loaded (finalize_weights_synthesis); use forward_from_flow");
  if (!io.img_xs || !io.coords || !io.t || !io.imgt_pred) throw std::runtime_error("gimmvfi: img_xs, coords, t and imgt_pred are required");
  DeviceGuard dg(device_);
  Ctx cx; cx.stream = stream; cx.sm_count = sm_count_;
  prof_.reset();
  cx.prof = profile_ ? &prof_ : nullptr;
  uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
  cx.arena.base = reinterpret_cast<char*>(base);
  cx.arena.cap = workspace_bytes - (base - reinterpret_cast<uintptr_t>(workspace));
  taps_.clear();
#ifndef GV_HOSTSIM
  // (the legacy default stream - handles 0 / 1 / 2 - cannot be captured: such callers run eagerly)
  if (use_graph_ && !profile_ && !debug_ && fc_ == nullptr && reinterpret_cast<uintptr_t>(stream) > 2) {
    // everything a recorded launch sequence depends on: problem, every caller pointer, the workspace, the arithmetic mode, the weights
    std::vector<uint64_t> key = {(uint64_t)p.B, (uint64_t)p.Hf, (uint64_t)p.Wf, (uint64_t)p.T, (uint64_t)p.Hc, (uint64_t)p.Wc, 0, (uint64_t)tc_mode_, (uint64_t)precise_,
                                 (uint64_t)hypo_fast_, (uint64_t)(raft_direct_ + 2), (uint64_t)corr_direct_max_t_, (uint64_t)gru_hoist_, (uint64_t)weights_version_, (uint64_t)raft_iters, (uint64_t)workspace, (uint64_t)workspace_bytes, (uint64_t)(uintptr_t)stream};
    std::memcpy(&key[6], &p.ds, sizeof(float));
    const void* ptrs[] = {io.img_xs, io.coords, io.t, io.imgt_pred, io.img_warp_4, io.flowt0_1, io.flowt1_1, io.flowt0_4, io.flowt1_4, io.raft_flow, io.nflow, io.ninrflow, io.flowt};
    for (const void* q : ptrs) key.push_back((uint64_t)(uintptr_t)q);
    GraphEntry* hit = nullptr;
    for (GraphEntry& g : graphs_) if (g.key == key) { hit = &g; break; }
    if (hit && hit->exec) {
      cuda_ok(cudaGraphLaunch((cudaGraphExec_t)hit->exec, stream), "cudaGraphLaunch");
      launches_ = hit->launches; ++graph_replays_;
      return;
    }
    if (hit && hit->seen >= 1) {   // second sighting: record.  (The first call ran eagerly: lazy one-time host work - function attributes - is done.)
      cudaGraph_t graph = nullptr;
      cuda_ok(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal), "cudaStreamBeginCapture");
      try {
        run(cx, p, io);
      } catch (...) {
        cudaStreamEndCapture(stream, &graph);
        if (graph) cudaGraphDestroy(graph);
        throw;
      }
      cuda_ok(cudaStreamEndCapture(stream, &graph), "cudaStreamEndCapture");
      cudaGraphExec_t exec = nullptr;
      cuda_ok(cudaGraphInstantiate(&exec, graph, 0), "cudaGraphInstantiate");
      cudaGraphDestroy(graph);
      hit->exec = exec; hit->launches = cx.launches;
      cuda_ok(cudaGraphLaunch(exec, stream), "cudaGraphLaunch");
      launches_ = cx.launches; ++graph_replays_;
      return;
    }
    if (!hit) {
      if (graphs_.size() >= 8) { if (graphs_.front().exec) cudaGraphExecDestroy((cudaGraphExec_t)graphs_.front().exec); graphs_.erase(graphs_.begin()); }
      GraphEntry g; g.key = key; g.seen = 1;
      graphs_.push_back(g);
    }
  }
#endif
  run(cx, p, io);
  launches_ = cx.launches;
}

}  // namespace gv
