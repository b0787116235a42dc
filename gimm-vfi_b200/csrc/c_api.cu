// extern "C" boundary (include/gimmvfi_b200.h).  No exceptions cross it.
#include "../../include/gimmvfi_b200.h"
#include "engine.h"

using namespace gv;

struct gimmvfi_engine {
  Engine eng;
  std::string err, prof_json;
  explicit gimmvfi_engine(int dev) : eng(dev) {}
};

static thread_local std::string g_static_err;

#define GV_TRY(e_, body)                                          \
  try { body; return 0; }                                         \
  catch (const std::exception& ex) { if (e_) (e_)->err = ex.what(); else g_static_err = ex.what(); return 1; } \
  catch (...) { if (e_) (e_)->err = "unknown error"; else g_static_err = "unknown error"; return 1; }

static TV to_tv(const gimmvfi_view* v) {
  TV t;
  if (!v) return t;
  t.p = v->data; t.n = v->n; t.h = v->h; t.w = v->w; t.c = v->c; t.ld = v->pixel_stride; t.sn = v->batch_stride;
  return t;
}
static Ctx op_ctx(void* stream) {
  Ctx cx; cx.stream = (gvStream_t)stream;
#ifndef GV_HOSTSIM
  int dev = 0; cudaGetDevice(&dev); int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); cx.sm_count = sms;
#endif
  return cx;
}
static Problem to_problem(const gimmvfi_problem* p) {
  Problem q; q.B = p->batch; q.Hf = p->height; q.Wf = p->width; q.T = p->timesteps; q.ds = p->ds_factor; q.Hc = p->coord_height; q.Wc = p->coord_width;
  return q;
}

extern "C" {

int gimmvfi_create(int device, gimmvfi_engine** out) {
  gimmvfi_engine* none = nullptr;
  GV_TRY(none, { *out = new gimmvfi_engine(device); })
}
void gimmvfi_destroy(gimmvfi_engine* e) { delete e; }

int gimmvfi_load_weight(gimmvfi_engine* e, const char* key, const float* host_data, const int64_t* shape, int ndim) {
  GV_TRY(e, { e->eng.load_weight(key, host_data, shape, ndim); })
}
int gimmvfi_finalize_weights(gimmvfi_engine* e) { GV_TRY(e, { e->eng.finalize_weights(); }) }

int gimmvfi_plan(gimmvfi_engine* e, const gimmvfi_problem* p, size_t* workspace_bytes) {
  GV_TRY(e, { *workspace_bytes = e->eng.plan(to_problem(p)); })
}

int gimmvfi_forward(gimmvfi_engine* e, const gimmvfi_problem* p, const gimmvfi_io* io, void* workspace, size_t workspace_bytes,
                    void* cuda_stream) {
  GV_TRY(e, {
    IO q;
    q.img_xs = io->img_xs; q.coords = io->coords; q.t = io->t; q.imgt_pred = io->imgt_pred; q.img_warp_4 = io->img_warp_4;
    q.flowt0_1 = io->flowt0_1; q.flowt1_1 = io->flowt1_1; q.flowt0_4 = io->flowt0_4; q.flowt1_4 = io->flowt1_4;
    q.raft_flow = io->raft_flow; q.nflow = io->nflow; q.ninrflow = io->ninrflow; q.flowt = io->flowt;
    e->eng.forward(to_problem(p), q, workspace, workspace_bytes, (gvStream_t)cuda_stream);
  })
}

static IO to_io(const gimmvfi_io* io) {
  IO q;
  q.img_xs = io->img_xs; q.coords = io->coords; q.t = io->t; q.imgt_pred = io->imgt_pred; q.img_warp_4 = io->img_warp_4;
  q.flowt0_1 = io->flowt0_1; q.flowt1_1 = io->flowt1_1; q.flowt0_4 = io->flowt0_4; q.flowt1_4 = io->flowt1_4;
  q.raft_flow = io->raft_flow; q.nflow = io->nflow; q.ninrflow = io->ninrflow; q.flowt = io->flowt;
  return q;
}
int gimmvfi_finalize_weights_synthesis(gimmvfi_engine* e) { GV_TRY(e, { e->eng.finalize_weights_synthesis(); }) }
int gimmvfi_finalize_weights_f(gimmvfi_engine* e) { GV_TRY(e, { e->eng.finalize_weights_f(); }) }
int gimmvfi_set_flowformer_iters(gimmvfi_engine* e, int iters) { GV_TRY(e, { if (iters < 1) throw std::runtime_error("iters must be >= 1"); e->eng.ff_iters = iters; }) }
int gimmvfi_plan_from_flow(gimmvfi_engine* e, const gimmvfi_problem* p, size_t* workspace_bytes) {
  GV_TRY(e, { *workspace_bytes = e->eng.plan_from_flow(to_problem(p)); })
}
int gimmvfi_forward_from_flow(gimmvfi_engine* e, const gimmvfi_problem* p, const gimmvfi_io* io, const gimmvfi_flow_inputs* fin,
                              void* workspace, size_t workspace_bytes, void* cuda_stream) {
  GV_TRY(e, {
    if (!io || !fin) throw std::runtime_error("gimmvfi_forward_from_flow: io and fin are required");
    FlowInputs f; f.flows = fin->flows;
    for (int j = 0; j < 2; ++j) { f.feat4[j] = fin->feat4[j]; f.feat8[j] = fin->feat8[j]; f.fnet[j] = fin->fnet[j]; }
    e->eng.forward_from_flow(to_problem(p), to_io(io), f, workspace, workspace_bytes, (gvStream_t)cuda_stream);
  })
}
int gimmvfi_finalize_weights_gimm(gimmvfi_engine* e) { GV_TRY(e, { e->eng.finalize_weights_gimm(); }) }
int gimmvfi_gimm_plan(gimmvfi_engine* e, const gimmvfi_problem* p, size_t* workspace_bytes) {
  GV_TRY(e, { *workspace_bytes = e->eng.plan_gimm(to_problem(p)); })
}
int gimmvfi_gimm_forward(gimmvfi_engine* e, const gimmvfi_problem* p, const float* xs, const float* ori_flow, const float* coords, const float* t,
                         float* out, void* workspace, size_t workspace_bytes, void* cuda_stream) {
  GV_TRY(e, {
    GimmIO q; q.xs = xs; q.ori_flow = ori_flow; q.coords = coords; q.t = t; q.out = out;
    e->eng.forward_gimm(to_problem(p), q, workspace, workspace_bytes, (gvStream_t)cuda_stream);
  })
}

const char* gimmvfi_last_error(gimmvfi_engine* e) { return e ? e->err.c_str() : g_static_err.c_str(); }
int64_t gimmvfi_last_launches(gimmvfi_engine* e) { return e->eng.last_launches(); }
int64_t gimmvfi_weights_version(gimmvfi_engine* e) { return e->eng.weights_version(); }
int gimmvfi_set_raft_iters(gimmvfi_engine* e, int iters) { GV_TRY(e, { if (iters < 1) throw std::runtime_error("iters must be >= 1"); e->eng.raft_iters = iters; }) }
int gimmvfi_set_debug(gimmvfi_engine* e, int on) { GV_TRY(e, { e->eng.set_debug(on != 0); }) }
int gimmvfi_get_tap(gimmvfi_engine* e, const char* name, gimmvfi_view* out) {
  GV_TRY(e, {
    auto it = e->eng.taps().find(name);
    if (it == e->eng.taps().end()) throw std::runtime_error(std::string("no tap named '") + name + "'");
    const TV& t = it->second;
    out->data = t.p; out->n = t.n; out->h = t.h; out->w = t.w; out->c = t.c; out->pixel_stride = t.ld; out->batch_stride = t.sn;
  })
}
size_t gimmvfi_frame_cache_bytes(const gimmvfi_problem* p) { return Engine::frame_cache_bytes(to_problem(p)); }
int gimmvfi_set_frame_cache(gimmvfi_engine* e, void* cache, size_t bytes, int load, int store) {
  GV_TRY(e, { e->eng.set_frame_cache(static_cast<float*>(cache), bytes, load != 0, store != 0); })
}
int gimmvfi_set_tensor_cores(gimmvfi_engine* e, int mode) { GV_TRY(e, { e->eng.set_tensor_cores(mode); }) }
int gimmvfi_set_cuda_graph(gimmvfi_engine* e, int on) { GV_TRY(e, { e->eng.set_cuda_graph(on != 0); }) }
int64_t gimmvfi_graph_replays(gimmvfi_engine* e) { return e->eng.graph_replays(); }
int gimmvfi_set_profile(gimmvfi_engine* e, int on) { GV_TRY(e, { e->eng.set_profile(on != 0); }) }
const char* gimmvfi_profile_json(gimmvfi_engine* e, void* stream) {
  try { e->prof_json = e->eng.profile_json((gvStream_t)stream); } catch (const std::exception& ex) { e->err = ex.what(); e->prof_json = "{}"; }
  return e->prof_json.c_str();
}
const char* gimmvfi_build_info(void) {
#ifdef GV_HOSTSIM
  return "gimmvfi_b200 HOSTSIM (test-only CPU emulation of the kernels; not a product build)";
#else
  return "gimmvfi_b200 sm_100a CUDA build";
#endif
}

// ------------------------------------------------------------ per-kernel entry points
int gimmvfi_op_softsplat(const gimmvfi_view* lat, const gimmvfi_view* flow, const gimmvfi_view* metric, const float* t, int t_mode,
                         const gimmvfi_view* scratch, const gimmvfi_view* out, void* stream) {
  gimmvfi_engine* e = nullptr;
  GV_TRY(e, {
    Ctx cx = op_ctx(stream);
    TV acc = to_tv(scratch); acc.c = 17;
    if (acc.ld < 17 || to_tv(lat).c != 16) throw std::runtime_error("softsplat: lat must have 16 channels and scratch pixel_stride >= 17");
    for (int n = 0; n < acc.n; ++n) dev_memset(acc.p + (int64_t)n * acc.sn, 0, (size_t)acc.h * acc.w * acc.ld * sizeof(float), cx.stream);
    softsplat_accumulate(cx, to_tv(lat), to_tv(flow), to_tv(metric), t, t_mode, acc);
    softsplat_normalize(cx, acc, to_tv(out));
  })
}
int gimmvfi_op_softsplat_fused(const gimmvfi_view* lat, const gimmvfi_view* flow, const gimmvfi_view* metric, const float* t, int t_mode,
                               const float* flow_absmax, const gimmvfi_view* out, void* stream) {
  gimmvfi_engine* e = nullptr;
  GV_TRY(e, {
    Ctx cx = op_ctx(stream);
    if (!softsplat_fused(cx, to_tv(lat), to_tv(flow), to_tv(metric), t, t_mode, flow_absmax, to_tv(out), /*force=*/true))
      throw std::runtime_error("softsplat_fused: 16-channel 16-byte-aligned latent / output, 8-byte-aligned flow and a flow bound are required");
  })
}
int gimmvfi_op_backwarp(const gimmvfi_view* src, const gimmvfi_view* flow, const gimmvfi_view* dst, void* stream) {
  gimmvfi_engine* e = nullptr;
  GV_TRY(e, { Ctx cx = op_ctx(stream); backwarp(cx, to_tv(src), to_tv(flow), to_tv(dst)); })
}
int gimmvfi_op_resize(const gimmvfi_view* src, const gimmvfi_view* dst, float scale_factor, float mult, void* stream) {
  gimmvfi_engine* e = nullptr;
  GV_TRY(e, {
    Ctx cx = op_ctx(stream);
    float r = (float)(1.0 / (double)scale_factor);
    resize_bilinear(cx, to_tv(src), to_tv(dst), r, r, mult, 0, ACT_NONE);
  })
}
// ---- FlowFormer / Twins token-side kernels (ops_tokens.cu), NHWC views
int gimmvfi_op_layernorm(const gimmvfi_view* x, const float* gamma, const float* beta, float eps, const gimmvfi_view* out, float pe_scale, int pe_dim,
                         void* stream) {
  gimmvfi_engine* e = nullptr;
  GV_TRY(e, { Ctx cx = op_ctx(stream); layernorm(cx, to_tv(x), gamma, beta, eps, to_tv(out), pe_scale, pe_dim); })
}
int gimmvfi_op_window_attention(const gimmvfi_view* q, const gimmvfi_view* k, const gimmvfi_view* v, const gimmvfi_view* out, int heads, int ws,
                                void* stream) {
  gimmvfi_engine* e = nullptr;
  GV_TRY(e, { Ctx cx = op_ctx(stream); window_attention(cx, to_tv(q), to_tv(k), to_tv(v), to_tv(out), heads, ws); })
}
int gimmvfi_op_global_attention(const gimmvfi_view* q, const gimmvfi_view* k, const gimmvfi_view* v, const gimmvfi_view* out, int heads, void* stream) {
  gimmvfi_engine* e = nullptr;
  GV_TRY(e, {
    Ctx cx = op_ctx(stream);
    const TV Q = to_tv(q); const TV K = to_tv(k); const TV V = to_tv(v); const TV O = to_tv(out);
    if (Q.c != O.c || K.c != Q.c || V.c != Q.c || Q.c % heads || Q.n != K.n || K.h != V.h || K.w != V.w) throw std::runtime_error("global_attention: shape mismatch");
    AttnDims a{};
    a.nb1 = Q.n; a.nb2 = 1; a.nq = (int64_t)Q.h * Q.w; a.nk = (int64_t)K.h * K.w; a.heads = heads;
    a.q_s1 = Q.sn; a.q_si = Q.ld; a.k_s1 = K.sn; a.k_sj = K.ld; a.v_s1 = V.sn; a.v_sj = V.ld; a.o_s1 = O.sn; a.o_si = O.ld;
    strided_attention(cx, Q.p, K.p, V.p, O.p, a, Q.c / heads);
  })
}
int gimmvfi_op_patchify(const gimmvfi_view* src, const gimmvfi_view* dst, int k, void* stream) {
  gimmvfi_engine* e = nullptr;
  GV_TRY(e, { Ctx cx = op_ctx(stream); patchify(cx, to_tv(src), to_tv(dst), k); })
}
int gimmvfi_op_cost_conv1(const float* vol, int64_t maps, int h, int w, const float* w_host, const float* b_host, const gimmvfi_view* out_padded,
                          void* stream) {
  gimmvfi_engine* e = nullptr;
  GV_TRY(e, {
    Ctx cx = op_ctx(stream);
    const TV o = to_tv(out_padded);
    cost_conv1(cx, vol, maps, h, w, w_host, b_host, o, o.h - 4, o.w - 4);
  })
}
int gimmvfi_op_conv7x7_small_cout(const gimmvfi_view* in, const float* w_tap_cin_4, const float* bias, int cout, const gimmvfi_view* out, void* stream) {
  gimmvfi_engine* e = nullptr;
  GV_TRY(e, {
    Ctx cx = op_ctx(stream);
    ConvW w; w.w = w_tap_cin_4; w.b = bias; w.cin = to_tv(in).c; w.cout = cout; w.kh = w.kw = 7; w.cout_ld = 4;
    if (!conv7x7_small_cout(cx, to_tv(in), w, ACT_NONE, nullptr, to_tv(out))) throw std::runtime_error("conv7x7_small_cout: layer / tensors not of that class (or host simulation)");
  })
}
int gimmvfi_op_corr_volume(const gimmvfi_view* fa, const gimmvfi_view* fb, float* vol, void* stream) {
  gimmvfi_engine* e = nullptr;
  GV_TRY(e, { Ctx cx = op_ctx(stream); TV a = to_tv(fa); corr_volume(cx, a, to_tv(fb), vol, 1.0f / std::sqrt((float)a.c)); })
}
int gimmvfi_op_corr_volume_tc(const gimmvfi_view* fa, const gimmvfi_view* fb, float* scratch, float* vol, int split, void* stream) {
  gimmvfi_engine* e = nullptr;
  GV_TRY(e, {
#ifdef GV_HOSTSIM
    throw std::runtime_error("corr_volume_tc is a tcgen05 kernel; not available in the host simulation");
#else
    Ctx cx = op_ctx(stream);
    TV a = to_tv(fa); TV b = to_tv(fb);
    if (a.n != 1 || b.n != 1 || a.c % 32 || a.ld != a.c || b.ld != b.c) throw std::runtime_error("corr_volume_tc: one dense sample with C % 32 == 0");
    const int64_t N = (int64_t)a.h * a.w;
    float* planes = scratch; float* zeros = scratch + 2 * N * a.c;
    dev_memset(zeros, 0, (size_t)(((N + 255) / 256) * 256 + 512) * sizeof(float), cx.stream);
    if (split && corr_volume_tc_wants_f16_planes()) split_planes_f16(cx, b, planes);
    else if (split) split_planes(cx, b, planes);
    corr_volume_tc(cx, a, split ? planes : b.p, zeros, vol, 1.0f / std::sqrt((float)a.c), split != 0);
#endif
  })
}
int gimmvfi_op_corr_pool(const float* src, float* dst, int64_t rows, int h, int w, void* stream) {
  gimmvfi_engine* e = nullptr;
  GV_TRY(e, { Ctx cx = op_ctx(stream); corr_pool(cx, src, dst, rows, h, w); })
}
int gimmvfi_op_corr_pool_pyramid(const float* l0, float* l1, float* l2, float* l3, int64_t rows, int h, int w, void* stream) {
  gimmvfi_engine* e = nullptr;
  GV_TRY(e, { Ctx cx = op_ctx(stream); corr_pool_pyramid(cx, l0, l1, l2, l3, rows, h, w); })
}
int gimmvfi_op_corr_lookup(const float* const lvl[4], const int32_t lvl_h[4], const int32_t lvl_w[4], const gimmvfi_view* coords,
                           const gimmvfi_view* out, void* stream) {
  gimmvfi_engine* e = nullptr;
  GV_TRY(e, {
    Ctx cx = op_ctx(stream);
    TV c = to_tv(coords);
    CorrPyr p;
    for (int l = 0; l < 4; ++l) { p.lvl[l] = lvl[l]; p.h[l] = lvl_h[l]; p.w[l] = lvl_w[l]; }
    p.rows_per_sample = (int64_t)c.h * c.w;
    corr_lookup(cx, p, c, to_tv(out));
  })
}
int gimmvfi_op_corr_lookup_direct(const gimmvfi_view* src, const void* const tgt_half[4], const int32_t lvl_h[4], const int32_t lvl_w[4], float scale,
                                  const gimmvfi_view* coords, const gimmvfi_view* out, void* stream) {
  gimmvfi_engine* e = nullptr;
  GV_TRY(e, {
    Ctx cx = op_ctx(stream);
    TV s = to_tv(src);
    CorrFeat f;
    for (int l = 0; l < 4; ++l) { f.lvl[l] = static_cast<const uint16_t*>(tgt_half[l]); f.h[l] = lvl_h[l]; f.w[l] = lvl_w[l]; }
    f.c = s.c; f.scale = scale;
    corr_lookup_direct(cx, s, f, to_tv(coords), to_tv(out));
  })
}
int gimmvfi_op_conv2d(const gimmvfi_view* in0, const gimmvfi_view* in1, const float* w_packed, const float* bias, int cin, int cout,
                      int cout_ld, int kh, int kw, int stride, int pad_h, int pad_w, int reflect, int act, const float* slope,
                      const gimmvfi_view* residual, const gimmvfi_view* out, void* stream) {
  gimmvfi_engine* e = nullptr;
  GV_TRY(e, {
    Ctx cx = op_ctx(stream);
    ConvW w; w.w = w_packed; w.b = bias; w.cin = cin; w.cout = cout; w.cout_ld = cout_ld; w.kh = kh; w.kw = kw;
    ConvGeom g; g.stride = stride; g.ph = pad_h; g.pw = pad_w; g.reflect = reflect;
    ConvEpi ep; ep.act1 = act; ep.slope1 = slope; ep.res = to_tv(residual);
    conv2d(cx, to_tv(in0), to_tv(in1), w, g, ep, to_tv(out));
  })
}
int gimmvfi_op_conv2d_tc(const gimmvfi_view* in0, const gimmvfi_view* in1, const float* w_tc, const float* bias, int cin, int cout, int kh,
                         int kw, int act1, const float* slope1, const gimmvfi_view* residual, int act2, const float* slope2,
                         const gimmvfi_view* mul, const gimmvfi_view* gru_z, const gimmvfi_view* gru_h, int split,
                         const gimmvfi_view* out, const void* w_tc_s, float w_scale, void* stream) {
  gimmvfi_engine* e = nullptr;
  GV_TRY(e, {
#ifdef GV_HOSTSIM
    throw std::runtime_error("conv2d_tc is a tcgen05 kernel; not available in the host simulation");
#else
    Ctx cx = op_ctx(stream);
    ConvW w; w.b = bias; w.cin = cin; w.cout = cout; w.kh = kh; w.kw = kw; w.w_tc = w_tc; w.has_lo = true;
    w.cout_pad = tc_cout_pad(cout); w.cin_pad = (cin + 31) & ~31;
    if (w_tc_s) { w.w_tc_s = w_tc_s; w.cin_pad_s = (cin + 63) & ~63; w.w_scale = w_scale; }   // 3xF16 form of the split kernel
    ConvGeom g; g.stride = 1; g.ph = kh / 2; g.pw = kw / 2;
    ConvEpi ep; ep.act1 = act1; ep.slope1 = slope1; ep.res = to_tv(residual); ep.act2 = act2; ep.slope2 = slope2;
    ep.mul = to_tv(mul); ep.gru_z = to_tv(gru_z); ep.gru_h = to_tv(gru_h);
    if (!conv2d_tc_supported(to_tv(in0), to_tv(in1), w, g, ep, to_tv(out), split != 0)) throw std::runtime_error("conv2d_tc: unsupported configuration");
    conv2d_tc(cx, to_tv(in0), to_tv(in1), w, g, ep, to_tv(out), split != 0);
#endif
  })
}
int gimmvfi_op_conv2d_tc_strided(const gimmvfi_view* in0, const float* w_tc, const float* bias, int cin, int cout, int kh, int kw, int stride,
                                 int act1, int split, const gimmvfi_view* out, void* stream) {
  gimmvfi_engine* e = nullptr;
  GV_TRY(e, {
#ifdef GV_HOSTSIM
    throw std::runtime_error("conv2d_tc is a tcgen05 kernel; not available in the host simulation");
#else
    Ctx cx = op_ctx(stream);
    ConvW w; w.b = bias; w.cin = cin; w.cout = cout; w.kh = kh; w.kw = kw; w.w_tc = w_tc; w.has_lo = true;
    w.cout_pad = tc_cout_pad(cout); w.cin_pad = (cin + 31) & ~31;
    ConvGeom g; g.stride = stride; g.ph = kh / 2; g.pw = kw / 2;
    ConvEpi ep; ep.act1 = act1;
    if (!conv2d_tc_supported(to_tv(in0), TV(), w, g, ep, to_tv(out), split != 0)) throw std::runtime_error("conv2d_tc_strided: unsupported configuration");
    conv2d_tc(cx, to_tv(in0), TV(), w, g, ep, to_tv(out), split != 0);
#endif
  })
}
int gimmvfi_op_conv2d_tc_f16(const gimmvfi_view* in0, const gimmvfi_view* in1, const void* w_tc_h, const float* w_tc, const float* bias, int cin,
                             int cout, int kh, int kw, int act1, const float* slope1, const gimmvfi_view* residual, int act2,
                             const float* slope2, int half_mask, const gimmvfi_view* out, void* stream) {
  gimmvfi_engine* e = nullptr;
  GV_TRY(e, {
#ifdef GV_HOSTSIM
    throw std::runtime_error("conv2d_tc is a tcgen05 kernel; not available in the host simulation");
#else
    Ctx cx = op_ctx(stream);
    ConvW w; w.b = bias; w.cin = cin; w.cout = cout; w.kh = kh; w.kw = kw; w.w_tc = w_tc; w.has_lo = true;
    w.cout_pad = tc_cout_pad(cout); w.cin_pad = (cin + 31) & ~31; w.w_tc_h = w_tc_h; w.cin_pad_h = (cin + 63) & ~63;
    ConvGeom g; g.stride = 1; g.ph = kh / 2; g.pw = kw / 2;
    TV a0 = to_tv(in0); TV a1 = to_tv(in1); TV r = to_tv(residual); TV o = to_tv(out);
    a0.f16 = half_mask & 1; if (a1.p) a1.f16 = half_mask & 1; o.f16 = (half_mask >> 1) & 1; if (r.p) r.f16 = (half_mask >> 2) & 1;
    ConvEpi ep; ep.act1 = act1; ep.slope1 = slope1; ep.res = r; ep.act2 = act2; ep.slope2 = slope2;
    if (!conv2d_tc_supported(a0, a1, w, g, ep, o, false)) throw std::runtime_error("conv2d_tc_f16: unsupported configuration");
    conv2d_tc(cx, a0, a1, w, g, ep, o, false);
#endif
  })
}
int gimmvfi_op_conv2d_halo(const gimmvfi_view* in0, const void* w_tc_h, const float* w_tc, const float* bias, int cin, int cout, int act1,
                           const float* slope1, const gimmvfi_view* residual, int act2, const float* slope2, int half_mask, int prepadded,
                           const gimmvfi_view* out, void* stream) {
  gimmvfi_engine* e = nullptr;
  GV_TRY(e, {
#ifdef GV_HOSTSIM
    throw std::runtime_error("conv2d_halo is a tcgen05 kernel; not available in the host simulation");
#else
    Ctx cx = op_ctx(stream);
    const int k = prepadded == 2 ? 1 : 3;   // prepadded: 0 = 3x3 zero padding 1, 1 = 3x3 on a pre-padded input, 2 = 1x1
    ConvW w; w.b = bias; w.cin = cin; w.cout = cout; w.kh = k; w.kw = k; w.w_tc = w_tc; w.has_lo = true;
    w.cout_pad = tc_cout_pad(cout); w.cin_pad = (cin + 31) & ~31; w.w_tc_h = w_tc_h; w.cin_pad_h = (cin + 63) & ~63;
    ConvGeom g; g.stride = 1; g.ph = prepadded == 1 ? 0 : k / 2; g.pw = g.ph; g.loose_w = prepadded == 1 ? 1 : 0;
    TV a0 = to_tv(in0); TV r = to_tv(residual); TV o = to_tv(out);
    a0.f16 = half_mask & 1; o.f16 = (half_mask >> 1) & 1; if (r.p) r.f16 = (half_mask >> 2) & 1;
    ConvEpi ep; ep.act1 = act1; ep.slope1 = slope1; ep.res = r; ep.act2 = act2; ep.slope2 = slope2;
    if (!conv2d_halo_supported(a0, TV(), w, g, ep, o)) throw std::runtime_error("conv2d_halo: unsupported configuration");
    conv2d_halo(cx, a0, w, g, ep, o);
#endif
  })
}
int gimmvfi_op_hyponet(gimmvfi_engine* e, const gimmvfi_view* latent, const float* coords, const gimmvfi_view* out, int fp32_class, void* stream) {
  GV_TRY(e, {
    if (!e->eng.finalized() || !e->eng.hyponet_blob(fp32_class != 0)) throw std::runtime_error("hyponet: finalize_weights() first");
    Ctx cx = op_ctx(stream);
    if (!hyponet_fused_supported(to_tv(latent), to_tv(out))) throw std::runtime_error("hyponet: latent (n,h,w,32) dense fp32, out (n,h,w,2)");
    if (fp32_class) hyponet_fused3(cx, to_tv(latent), coords, e->eng.hyponet_blob(true), to_tv(out));
    else hyponet_fused(cx, to_tv(latent), coords, e->eng.hyponet_blob(false), to_tv(out));
  })
}
int64_t gimmvfi_instnorm_scratch_floats(int n, int c) {
  TV t; t.n = n; t.c = c;
  return instnorm_scratch_floats(t) + (int64_t)n * c * 2;
}
int gimmvfi_op_instnorm(const gimmvfi_view* x, int relu, float* scratch, const gimmvfi_view* out, void* stream) {
  gimmvfi_engine* e = nullptr;
  GV_TRY(e, {
    Ctx cx = op_ctx(stream);
    TV t = to_tv(x);
    float* mr = scratch + instnorm_scratch_floats(t);
    instnorm_stats(cx, t, mr, scratch, 0);
    instnorm_apply(cx, t, mr, relu ? ACT_RELU : ACT_NONE, TV(), ACT_NONE, to_tv(out));
  })
}
int gimmvfi_op_convex_upsample(const gimmvfi_view* flow, const gimmvfi_view* mask, const gimmvfi_view* out, void* stream) {
  gimmvfi_engine* e = nullptr;
  GV_TRY(e, { Ctx cx = op_ctx(stream); convex_upsample(cx, to_tv(flow), to_tv(mask), to_tv(out)); })
}
int gimmvfi_op_frames_u8_to_padded_f32(const uint8_t* frames, int n, int h, int w, float* dst_nchw, int H, int W, int pad_top, int pad_left,
                                       void* stream) {
  gimmvfi_engine* e = nullptr;
  GV_TRY(e, {
    if (H < h || W < w || pad_top < 0 || pad_left < 0 || pad_top + h > H || pad_left + w > W) throw std::runtime_error("bad padding geometry");
    Ctx cx = op_ctx(stream); frames_u8_to_padded_f32(cx, frames, n, h, w, dst_nchw, H, W, pad_top, pad_left);
  })
}
int gimmvfi_op_pred_to_u8(const float* pred_nchw, int n, int H, int W, uint8_t* dst, int h, int w, int pad_top, int pad_left, int bgr,
                          void* stream) {
  gimmvfi_engine* e = nullptr;
  GV_TRY(e, {
    if (H < h || W < w || pad_top < 0 || pad_left < 0 || pad_top + h > H || pad_left + w > W) throw std::runtime_error("bad padding geometry");
    Ctx cx = op_ctx(stream); pred_to_u8(cx, pred_nchw, n, H, W, dst, h, w, pad_top, pad_left, bgr);
  })
}
int gimmvfi_op_pixel_shuffle(const gimmvfi_view* src, const gimmvfi_view* dst, int times, void* stream) {
  gimmvfi_engine* e = nullptr;
  GV_TRY(e, { Ctx cx = op_ctx(stream); pixel_shuffle(cx, to_tv(src), to_tv(dst), times); })
}

}  // extern "C"
