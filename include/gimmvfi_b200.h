/* gimmvfi_b200 — C ABI of the B200-native GIMM-VFI-R inference path.
 *
 * The reference (GSeanCDAT/GIMM-VFI) is pure Python; it has no FFI for this path.
 * Its only native boundaries are
 *   - the CuPy-JIT'd splat kernel launched with raw data_ptr()s on the current
 *     torch stream (src/models/generalizable_INR/modules/softsplat.py:358-446), and
 *   - the (unused) pybind extension alt_cuda_corr.forward(fmap1, fmap2, coords, r)
 *     (.../flowformer/alt_cuda_corr/correlation.cpp:19-53).
 * This header is what a maintainer binds (ctypes) to replace the
 * body of GIMMVFI_R.forward() (src/models/generalizable_INR/gimmvfi_r.py:324-407):
 * raw device pointers + sizes + a cudaStream_t, no torch types, no exceptions across
 * the boundary (every call returns 0 on success; gimmvfi_last_error() explains).
 *
 * Threading / ownership: one engine per GPU and per host thread; all work is
 * enqueued on the caller's stream; the caller owns inputs, outputs and the
 * workspace; forward() performs no allocation and no host synchronisation.
 */
#ifndef GIMMVFI_B200_H_
#define GIMMVFI_B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gimmvfi_engine gimmvfi_engine;

/* One call of GIMMVFI_R.forward(img_xs, coord, t, ds_factor=ds). */
typedef struct gimmvfi_problem {
  int32_t batch;       /* B: frame pairs                                            */
  int32_t height;      /* Hf, Wf: caller-padded (x32) input size,                    */
  int32_t width;       /*   src/utils/utils.py:156-185 InputPadder                   */
  int32_t timesteps;   /* T = len(t) = len(coord)                                    */
  float ds_factor;     /* 0 -> None; else gimmvfi_r.py:329-337                       */
  int32_t coord_height; /* Hc, Wc of the coord grids = int(H*ds), coord_sampler.py:35 */
  int32_t coord_width;
} gimmvfi_problem;

/* Device pointers (fp32).  Optional outputs may be NULL.  Layouts are the
 * reference's, with the python list index (timestep) as the outermost dim.
 * H, W = network resolution (= Hf,Wf, or floor(Hf*ds), floor(Wf*ds)). */
typedef struct gimmvfi_io {
  const float* img_xs;   /* (B,3,2,Hf,Wf) in [0,1]                      gimmvfi_r.py:324 */
  const float* coords;   /* (T,B,1,Hc,Wc,3) last dim (t,y,x)     coord_sampler.py:21-43 */
  const float* t;        /* (T,B)                                                        */
  float* imgt_pred;      /* (T,B,3,Hf,Wf)   required                    gimmvfi_r.py:399 */
  float* img_warp_4;     /* (T,B,3,H,W)     other_pred[i][0]                         :400 */
  float* flowt0_1;       /* (T,B,3,2,Hf,Wf) flowt0_pred[i][0]                        :401 */
  float* flowt1_1;       /* (T,B,3,2,Hf,Wf) flowt1_pred[i][0]                        :402 */
  float* flowt0_4;       /* (T,B,2,H/4,W/4) flowt0_pred[i][1]                             */
  float* flowt1_4;       /* (T,B,2,H/4,W/4) flowt1_pred[i][1]                             */
  float* raft_flow;      /* (B,2,2,H,W)                                              :403 */
  float* nflow;          /* (B,2,2,H,W)                                              :405 */
  float* ninrflow;       /* (T,B,2,1,Hc,Wc)                                          :404 */
  float* flowt;          /* (T,B,2,Hc,Wc)                                            :406 */
} gimmvfi_io;

/* NHWC fp32 view used by the per-kernel entry points and debug taps:
 * element (n,y,x,c) = data[n*batch_stride + (y*w + x)*pixel_stride + c]. */
typedef struct gimmvfi_view {
  float* data;
  int32_t n, h, w, c;
  int32_t pixel_stride;
  int64_t batch_stride;
} gimmvfi_view;

/* ---- engine life cycle ---- */
int gimmvfi_create(int device, gimmvfi_engine** out);
void gimmvfi_destroy(gimmvfi_engine* e);
/* state_dict entry (HOST pointer, fp32; int64 buffers are skipped by the caller).
 * Keys are the reference's 414 state_dict keys (gimmvfi_r.py:37-111). */
int gimmvfi_load_weight(gimmvfi_engine* e, const char* key, const float* host_data, const int64_t* shape, int ndim);
/* Packs (layout change, BatchNorm folding, HypoNet column normalisation) and uploads. */
int gimmvfi_finalize_weights(gimmvfi_engine* e);
int gimmvfi_plan(gimmvfi_engine* e, const gimmvfi_problem* p, size_t* workspace_bytes);
int gimmvfi_forward(gimmvfi_engine* e, const gimmvfi_problem* p, const gimmvfi_io* io, void* workspace, size_t workspace_bytes,
                    void* cuda_stream);
/* ---- GIMM-VFI-F (gimmvfi_f.py): everything downstream of the flow estimator ----
 * GIMMVFI_F.forward = FlowFormer (cal_bidirection_flow, gimmvfi_f.py:114-138) + exactly the GIMM / synthesis stages of GIMM-VFI-R
 * without the feature projections.  Those stages run natively on the outputs of an EXTERNAL estimator, given in the reference's
 * NCHW layouts at the network resolution H x W (= Hf x Wf, or floor(Hf*ds) x floor(Wf*ds)):                                      */
typedef struct gimmvfi_flow_inputs {
  const float* flows;      /* (B,2,2,H,W)       [f01 | f10] on dim 2 = `ori_flows` of cal_bidirection_flow        gimmvfi_f.py:127 */
  const float* feat4[2];   /* (B,128,H/4,W/4)   features0[0], features1[0]  (context-encoder stage 1)       encoders.py:22-48 */
  const float* feat8[2];   /* (B,256,H/8,W/8)   features0[1], features1[1]  (context-encoder stage 2)                         */
  const float* fnet[2];    /* (B,256,H/8,W/8)   fnet0, fnet1 -> BidirCorrBlock                               gimmvfi_f.py:121 */
} gimmvfi_flow_inputs;
/* Weights: the GIMM-VFI-F state_dict minus flow_estimator.* (gimmvfi_load_weight each, then this instead of gimmvfi_finalize_weights).
 * A full GIMM-VFI-R engine (gimmvfi_finalize_weights) accepts gimmvfi_forward_from_flow as well (its RAFT is then skipped). */
int gimmvfi_finalize_weights_synthesis(gimmvfi_engine* e);
/* The COMPLETE GIMM-VFI-F state_dict (639 tensors, flow_estimator.* = FlowFormer included; gimmvfi_f.py:27-111): gimmvfi_forward then
 * runs the native FlowFormer estimator (Twins-SVT-L x2, cost-perceiver memory encoder, 32-iteration GMA memory decoder; replaces
 * flowformer/core/FlowFormer/LatentCostFormer/transformer.py:45-74 as called by gimmvfi_f.py:114-138) followed by the synthesis half.
 * The network resolution (H, W after ds_factor) must be a multiple of 32. */
int gimmvfi_finalize_weights_f(gimmvfi_engine* e);
/* decoder_depth of the FlowFormer memory decoder (default 32, flowformer/configs/submission.py:50) */
int gimmvfi_set_flowformer_iters(gimmvfi_engine* e, int iters);
int gimmvfi_plan_from_flow(gimmvfi_engine* e, const gimmvfi_problem* p, size_t* workspace_bytes);
/* io->raft_flow (optional) receives a copy of `flows` (the reference returns them as "raft_flow", gimmvfi_f.py:382). */
int gimmvfi_forward_from_flow(gimmvfi_engine* e, const gimmvfi_problem* p, const gimmvfi_io* io, const gimmvfi_flow_inputs* fin,
                              void* workspace, size_t workspace_bytes, void* cuda_stream);

/* ---- GIMM standalone: GIMM.forward (gimm.py:129-214), the motion-modelling network alone ----
 * Weights: either the full GIMM-VFI-R state_dict (gimmvfi_finalize_weights) or a GIMM checkpoint holding only gimm.py's module
 * tree (cnn_encoder.*, res_conv.*, hyponet.*, g_filter, alpha_v, alpha_fe) followed by gimmvfi_finalize_weights_gimm.
 * problem: batch, height, width = flow resolution (any size >= 8), timesteps = T, ds_factor = 0, coord grid = height x width.
 * xs (B,2,2,H,W): flows normalised as in fi_utils.py:52-60, [f01 | f10] on dim 2; ori_flow (B,2,2,H,W): the raw flows;
 * coords (T,B,1,H,W,3); t (T,B); out (T,B,2,1,H,W) = the list of keep_xs_shape=True outputs. */
int gimmvfi_finalize_weights_gimm(gimmvfi_engine* e);
int gimmvfi_gimm_plan(gimmvfi_engine* e, const gimmvfi_problem* p, size_t* workspace_bytes);
int gimmvfi_gimm_forward(gimmvfi_engine* e, const gimmvfi_problem* p, const float* xs, const float* ori_flow, const float* coords,
                         const float* t, float* out, void* workspace, size_t workspace_bytes, void* cuda_stream);

const char* gimmvfi_last_error(gimmvfi_engine* e);
int64_t gimmvfi_last_launches(gimmvfi_engine* e);
/* bumped by every gimmvfi_finalize_weights*: callers that keep derived state (the frame cache below) key it on this */
int64_t gimmvfi_weights_version(gimmvfi_engine* e);
int gimmvfi_set_raft_iters(gimmvfi_engine* e, int iters);
/* debug taps: intermediate tensors of the last forward (views into the workspace) */
int gimmvfi_set_debug(gimmvfi_engine* e, int on);
int gimmvfi_get_tap(gimmvfi_engine* e, const char* name, gimmvfi_view* out);
/* Video callers (src/video_Nx.py:134-216) walk consecutive pairs (j, j+1), (j+1, j+2), ...  A caller-owned device buffer of
 * gimmvfi_frame_cache_bytes() keeps the RAFT encoder products of a call's SECOND frame (raft/raft.py:118-136 fnet map and cnet
 * net/inp, gimmvfi_r.py:134-141 projected context features); with load != 0 the next forward takes its FIRST frame's products
 * from it instead of recomputing them (the caller guarantees it is the same frame, same problem), with store != 0 it writes the
 * second frame's products.  cache == NULL switches the mechanism off.  Results are bit-identical to an uncached forward. */
size_t gimmvfi_frame_cache_bytes(const gimmvfi_problem* p);
int gimmvfi_set_frame_cache(gimmvfi_engine* e, void* cache, size_t bytes, int load, int store);
/* 0: fp32 CUDA cores everywhere; 1: post-RAFT convolutions on the tcgen05 TF32 path (fp32 accumulate);
 * 2: additionally the RAFT convolutions on tcgen05 with 3xTF32 operand splitting (fp32-class accuracy) */
int gimmvfi_set_tensor_cores(gimmvfi_engine* e, int mode);
/* CUDA graphs (off by default): gimmvfi_forward records its launch sequence the second time it is called with the same problem, the
 * same caller pointers (inputs, outputs, workspace), stream and mode, and replays the instantiated graph from then on - one graph
 * launch instead of ~550 kernel launches + ~1000 host-side tensor-map encodes.  Calls with other pointers run eagerly (and start
 * their own record; 8 graphs are kept).  gimmvfi_graph_replays: forwards served by a replay so far. */
int gimmvfi_set_cuda_graph(gimmvfi_engine* e, int on);
int64_t gimmvfi_graph_replays(gimmvfi_engine* e);
/* per-kernel CUDA-event timing of subsequent forwards; profile_json() synchronises the stream and
 * returns {"kernel": {"ms": total, "work": flops-or-elements, "launches": n}, ...} for the LAST forward */
int gimmvfi_set_profile(gimmvfi_engine* e, int on);
const char* gimmvfi_profile_json(gimmvfi_engine* e, void* cuda_stream);
const char* gimmvfi_build_info(void);

/* ---- per-kernel entry points (unit tests; all NHWC fp32 device views) ---- */
/* softsplat "linear-zeroeps": modules/softsplat.py:286-352 (kernel :371-421).
 * lat (n,h,w,16) flow (n,h,w,2) metric (n,h,w,1) t (n) scratch (n,h,w,>=17 ch, pixel_stride>=17) out (n,h,w,16);
 * the splat flow is flow*t (t_mode 0) or flow*(1-t) (t_mode 1). */
int gimmvfi_op_softsplat(const gimmvfi_view* lat, const gimmvfi_view* flow, const gimmvfi_view* metric, const float* t, int t_mode,
                         const gimmvfi_view* scratch, const gimmvfi_view* out, void* stream);
/* backward warp: modules/fi_utils.py:19-49 */
/* the same splat in ONE pass (target tiles accumulated in shared memory); flow_absmax[n] (device) >= max |flow| of sample n */
int gimmvfi_op_softsplat_fused(const gimmvfi_view* lat, const gimmvfi_view* flow, const gimmvfi_view* metric, const float* t_per_sample, int t_mode,
                               const float* flow_absmax, const gimmvfi_view* out, void* stream);
int gimmvfi_op_backwarp(const gimmvfi_view* src, const gimmvfi_view* flow, const gimmvfi_view* dst, void* stream);
/* F.interpolate(bilinear, align_corners=False): modules/fi_utils.py:67-70; dst = mult * resize(src) */
int gimmvfi_op_resize(const gimmvfi_view* src, const gimmvfi_view* dst, float scale_factor, float mult, void* stream);
/* all-pairs correlation raft/corr.py:167-175: vol[n][i][j] = <fa[n,i,:], fb[n,j,:]> / sqrt(C) */
/* FlowFormer / Twins-SVT token-side kernels (GIMM-VFI-F's estimator; csrc/ops_tokens.cu).  A token sequence (B, H*W, C) is an NHWC view.
 *   layernorm        nn.LayerNorm over the channels; `out` may be larger than `x` (zero rows / columns: the padding Twins applies after
 *                    the norm, LatentCostFormer/twins.py:835-842); pe_dim > 0 adds LinearPositionEmbeddingSine(x*pe_scale, y*pe_scale)
 *                    (attention.py:170-182)
 *   window_attention LocallyGroupedAttn core (twins.py:846-860): q, k, v on the padded map, ws x ws windows, softmax(q k^T / sqrt(d)) v
 *   global_attention GlobalSubSampleAttn core (twins.py:898-921): every token of q against all tokens of the (sub-sampled) k / v map
 *   patchify         space-to-depth: a k x k stride-k Conv2d (PatchEmbed twins.py:1142, `sr` twins.py:890) = this + a 1x1 convolution
 *   cost_conv1       Conv2d(1, 16, 6, stride 2, padding 2) + ReLU over `maps` cost maps (encoder.py:38-48), zero-extended to a multiple of 8,
 *                    written with a zero border of 2 (out_padded: (maps, oh+4, ow+4, 16)); weights [36][16] / bias [16] are HOST pointers
 *   conv7x7_small_cout  7x7 stride-1 zero-padded Conv2d with <= 4 output channels (amt_comb_block.2, gimmvfi_r.py:60-64), exact fp32;
 *                    weights [49][cin][4] / bias device pointers */
int gimmvfi_op_layernorm(const gimmvfi_view* x, const float* gamma, const float* beta, float eps, const gimmvfi_view* out, float pe_scale, int pe_dim,
                         void* stream);
int gimmvfi_op_window_attention(const gimmvfi_view* q, const gimmvfi_view* k, const gimmvfi_view* v, const gimmvfi_view* out, int heads, int ws,
                                void* stream);
int gimmvfi_op_global_attention(const gimmvfi_view* q, const gimmvfi_view* k, const gimmvfi_view* v, const gimmvfi_view* out, int heads, void* stream);
int gimmvfi_op_patchify(const gimmvfi_view* src, const gimmvfi_view* dst, int k, void* stream);
int gimmvfi_op_cost_conv1(const float* vol, int64_t maps, int h, int w, const float* w_host, const float* b_host, const gimmvfi_view* out_padded,
                          void* stream);
int gimmvfi_op_conv7x7_small_cout(const gimmvfi_view* in, const float* w_tap_cin_4, const float* bias, int cout, const gimmvfi_view* out, void* stream);
int gimmvfi_op_corr_volume(const gimmvfi_view* fa, const gimmvfi_view* fb, float* vol, void* stream);
/* the same volume for ONE sample (n == 1, dense views, C % 32 == 0) as a tcgen05 GEMM: split != 0 -> 3xTF32 (RAFT's volume),
 * else TF32 (the bidirectional volume).  scratch >= 2*h*w*C + h*w + 1024 floats */
int gimmvfi_op_corr_volume_tc(const gimmvfi_view* fa, const gimmvfi_view* fb, float* scratch, float* vol, int split, void* stream);
/* 2x2 average pooling of every row's (h,w) image: raft/corr.py:139-142 */
int gimmvfi_op_corr_pool(const float* src, float* dst, int64_t rows, int h, int w, void* stream);
/* levels 1..3 from level 0 in one pass (three successive corr_pool's, raft/corr.py:139-142) */
int gimmvfi_op_corr_pool_pyramid(const float* l0, float* l1, float* l2, float* l3, int64_t rows, int h, int w, void* stream);
/* 4-level 9x9 lookup raft/corr.py:144-165: lvl[k] = (n*h*w) x (h_k*w_k) pyramids; out (n,h,w,324) */
int gimmvfi_op_corr_lookup(const float* const lvl[4], const int32_t lvl_h[4], const int32_t lvl_w[4], const gimmvfi_view* coords,
                           const gimmvfi_view* out, void* stream);
/* Volume-free form of the same lookup (reference: raft/corr.py:23-93 BidirCorrBlock.__call__, looked up once per interpolated frame):
 * src = the source frame's fp32 features (n,h,w,256); tgt_half[l] = the other frame's features, 2^l x 2^l average-pooled, IEEE half,
 * dense NHWC (n, lvl_h[l], lvl_w[l], 256); out[..., l*81 + a*9 + b] = scale * <src, bilinear sample of level l>, as gimmvfi_op_corr_lookup
 * returns from the volume pyramid. */
int gimmvfi_op_corr_lookup_direct(const gimmvfi_view* src, const void* const tgt_half[4], const int32_t lvl_h[4], const int32_t lvl_w[4], float scale,
                                  const gimmvfi_view* coords, const gimmvfi_view* out, void* stream);
/* conv2d on NHWC: weight packed [kh*kw][cin][cout_ld], act: 0 none,1 relu,2 lrelu(0.1),3 prelu,4 sigmoid,5 tanh,6 sin */
int gimmvfi_op_conv2d(const gimmvfi_view* in0, const gimmvfi_view* in1_or_null, const float* w_packed, const float* bias, int cin,
                      int cout, int cout_ld, int kh, int kw, int stride, int pad_h, int pad_w, int reflect, int act,
                      const float* slope, const gimmvfi_view* residual_or_null, const gimmvfi_view* out, void* stream);
/* tcgen05/TMA implicit-GEMM conv (stride 1, "same" zero padding).  w_tc packed [2][kh*kw][cout_pad][cin_pad32]
 * (plane 0 = TF32(w), plane 1 = TF32(w - plane 0); cout_pad = N tiling of cout, see tc_tile_n); bias padded to cout_pad.
 * y = gru( act2(residual + act1(conv + bias)) * mul );  split != 0 -> three-term operand split (fp32-class accuracy, unrounded
 * output): 3xTF32, or — when w_tc_s is given — 3xF16: w_tc_s = half [2][kh*kw][cout_pad][cin_pad64] holding the fp16 hi / lo planes of
 * w * w_scale (w_scale a power of two), activations split on the fly, kind::f16 MMAs at twice the TF32 rate */
int gimmvfi_op_conv2d_tc(const gimmvfi_view* in0, const gimmvfi_view* in1_or_null, const float* w_tc, const float* bias, int cin,
                         int cout, int kh, int kw, int act1, const float* slope1, const gimmvfi_view* residual_or_null, int act2,
                         const float* slope2, const gimmvfi_view* mul_or_null, const gimmvfi_view* gru_z_or_null,
                         const gimmvfi_view* gru_h_or_null, int split, const gimmvfi_view* out, const void* w_tc_s_or_null, float w_scale,
                         void* stream);
/* the same kernel at stride 1 or 2 with "same"-style padding k/2 (RAFT encoder down-sampling convs, raft/extractor.py:42-48,140):
 * out is (n, (h + 2*(kh/2) - kh)/stride + 1, ...); TMA element strides pick every stride-th input pixel */
int gimmvfi_op_conv2d_tc_strided(const gimmvfi_view* in0, const float* w_tc, const float* bias, int cin, int cout, int kh, int kw,
                                 int stride, int act1, int split, const gimmvfi_view* out, void* stream);
/* the same kernel with half-precision storage (precision mode 3: the final decoder's residual trunk, fi_components.py:97-154,
 * 299-305).  half_mask: bit 0 = in0/in1 hold IEEE half (kind::f16 MMAs; weights = w_tc_h [kh*kw][cout_pad][cin_pad64] half),
 * bit 1 = out is half, bit 2 = residual is half; views of half tensors give strides in ELEMENTS.  With bit 0 clear the operands are
 * fp32/TF32 (w_tc as above) and only the store / residual formats change. */
int gimmvfi_op_conv2d_tc_f16(const gimmvfi_view* in0, const gimmvfi_view* in1_or_null, const void* w_tc_h, const float* w_tc,
                             const float* bias, int cin, int cout, int kh, int kw, int act1, const float* slope1,
                             const gimmvfi_view* residual_or_null, int act2, const float* slope2, int half_mask,
                             const gimmvfi_view* out, void* stream);
/* nn.InstanceNorm2d + optional relu: raft/extractor.py:133-134; scratch >= gimmvfi_instnorm_scratch_floats */
/* 3x3 (or 1x1) stride-1 convolution of a K-poor layer (cin <= 32 fp32 or <= 64 half, cout <= 64): halo tile loaded once, the 9 taps as
 * shifted UMMA operand views, weights resident in shared memory (csrc/conv_halo.cu).  Weight packing as gimmvfi_op_conv2d_tc (plane 0) /
 * gimmvfi_op_conv2d_tc_f16; half_mask bit 0: input half, bit 1: output half, bit 2: residual half; prepadded: 0 = 3x3 with zero padding 1,
 * 1 = 3x3, `in0` is (h+2, w+2) and carries its own padding (valid conv), 2 = 1x1. */
int gimmvfi_op_conv2d_halo(const gimmvfi_view* in0, const void* w_tc_h_or_null, const float* w_tc, const float* bias, int cin, int cout,
                           int act1, const float* slope1, const gimmvfi_view* residual_or_null, int act2, const float* slope2,
                           int half_mask, int prepadded, const gimmvfi_view* out, void* stream);
/* HypoNet.forward (modules/hyponet.py:71-146) as ONE fused tcgen05 kernel with the engine's loaded weights: latent (n,h,w,32),
 * coords n*h*w x (t,y,x) as the caller's coordinate tensor holds them -> out (n,h,w,2) = normalised flow (output_bias included).
 * fp32_class != 0: the default kernel of the forward pass (fp16 hi/lo operand pairs, fp32-class result); 0: TF32 / half operands */
int gimmvfi_op_hyponet(gimmvfi_engine* e, const gimmvfi_view* latent, const float* coords, const gimmvfi_view* out, int fp32_class,
                       void* cuda_stream);
int64_t gimmvfi_instnorm_scratch_floats(int n, int c);
int gimmvfi_op_instnorm(const gimmvfi_view* x, int relu, float* scratch, const gimmvfi_view* out, void* stream);
/* convex x8 upsampling raft/raft.py:86-97: flow (n,h,w,2), mask (n,h,w,576), out (n,8h,8w,2) */
int gimmvfi_op_convex_upsample(const gimmvfi_view* flow, const gimmvfi_view* mask, const gimmvfi_view* out, void* stream);
/* Driver-side pre/post-processing on the GPU (reference src/video_Nx.py:40-50,152-153,182-202):
 * uint8 HWC RGB frames (n,h,w,3) -> float32 (n,3,H,W) = x/255.0, replicate-padded like InputPadder(shape, 32).pad
 * (src/utils/utils.py:156-174: pad_left = pad_w//2, pad_top = pad_h//2) */
int gimmvfi_op_frames_u8_to_padded_f32(const uint8_t* frames, int n, int h, int w, float* dst_nchw, int H, int W, int pad_top,
                                       int pad_left, void* stream);
/* float32 (n,3,H,W) prediction -> unpadded uint8 (n,h,w,3): (x*255.0).astype(uint8) with optional RGB->BGR flip
 * (InputPadder.unpad + video_Nx.py:190-196) */
int gimmvfi_op_pred_to_u8(const float* pred_nchw, int n, int H, int W, uint8_t* dst, int h, int w, int pad_top, int pad_left, int bgr,
                          void* stream);
/* nn.PixelShuffle(2) applied `times` times */
int gimmvfi_op_pixel_shuffle(const gimmvfi_view* src, const gimmvfi_view* dst, int times, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GIMMVFI_B200_H_ */
