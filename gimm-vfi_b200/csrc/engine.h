// Engine: weights, workspace planning and the stream-ordered forward pass of
// GIMM-VFI-R (reference hot path: gimmvfi_r.py:324-407).  See DESIGN.md.
#pragma once
#include "common.h"

namespace gv {

struct HostTensor {
  std::vector<int64_t> shape;
  std::vector<float> data;
};

struct Problem {
  int B = 1;        // pairs in this call
  int Hf = 0, Wf = 0;  // full (caller-padded) resolution
  int T = 1;        // timesteps
  float ds = 0.f;   // ds_factor (0 -> None)
  int Hc = 0, Wc = 0;  // HypoNet coordinate grid (== network resolution at inference)
  int H() const { return ds > 0.f ? (int)std::floor((double)Hf * (double)ds) : Hf; }
  int W() const { return ds > 0.f ? (int)std::floor((double)Wf * (double)ds) : Wf; }
};

// Device pointers, caller-owned.  Layouts are the reference's (NCHW-style), one
// contiguous block per output with the timestep as the outermost dimension.
struct IO {
  const float* img_xs = nullptr;   // (B,3,2,Hf,Wf) in [0,1]
  const float* coords = nullptr;   // (T,B,1,Hc,Wc,3)  last dim (t,y,x)
  const float* t = nullptr;        // (T,B)
  float* imgt_pred = nullptr;      // (T,B,3,Hf,Wf)
  float* img_warp_4 = nullptr;     // (T,B,3,H,W)
  float* flowt0_1 = nullptr;       // (T,B,3,2,Hf,Wf)
  float* flowt1_1 = nullptr;
  float* flowt0_4 = nullptr;       // (T,B,2,H/4,W/4)
  float* flowt1_4 = nullptr;
  float* raft_flow = nullptr;      // (B,2,2,H,W)
  float* nflow = nullptr;          // (B,2,2,H,W)
  float* ninrflow = nullptr;       // (T,B,2,1,Hc,Wc)
  float* flowt = nullptr;          // (T,B,2,Hc,Wc)
};

// Outputs of an EXTERNAL bidirectional flow estimator at the network resolution (GIMM-VFI-F: FlowFormer, gimmvfi_f.py:114-138) in the
// reference's NCHW layouts; with these the engine runs everything downstream of cal_bidirection_flow (forward_from_flow)
struct FlowInputs {
  const float* flows = nullptr;                  // (B,2,2,H,W): [f01 | f10] on dim 2
  const float* feat4[2] = {nullptr, nullptr};    // (B,128,H/4,W/4): features0[0], features1[0]
  const float* feat8[2] = {nullptr, nullptr};    // (B,256,H/8,W/8): features0[1], features1[1]
  const float* fnet[2] = {nullptr, nullptr};     // (B,256,H/8,W/8): the maps BidirCorrBlock correlates
};

struct DebugTap { TV tv; };

// GIMM.forward (gimm.py:129-214) device pointers, reference layouts
struct GimmIO {
  const float* xs = nullptr;        // (B,2,2,H,W) normalised flows [f01 | f10] on dim 2
  const float* ori_flow = nullptr;  // (B,2,2,H,W) raw flows
  const float* coords = nullptr;    // (T,B,1,H,W,3)
  const float* t = nullptr;         // (T,B)
  float* out = nullptr;             // (T,B,2,1,H,W) normalised flow at t (keep_xs_shape=True)
};

struct Net;

class Engine {
 public:
  explicit Engine(int device);
  ~Engine();
  void load_weight(const std::string& key, const float* host, const int64_t* shape, int ndim);
  void finalize_weights();
  void finalize_weights_gimm();                        // a standalone GIMM checkpoint (gimm.py's module tree only)
  void finalize_weights_synthesis();                   // everything downstream of the flow estimator (GIMM-VFI-F's tree minus flow_estimator.*)
  void finalize_weights_f();                           // the complete GIMM-VFI-F tree: synthesis half + the native FlowFormer estimator (flowformer.cu)
  size_t plan_from_flow(const Problem& p);
  void forward_from_flow(const Problem& p, const IO& io, const FlowInputs& fin, void* workspace, size_t workspace_bytes, gvStream_t stream);
  size_t plan_gimm(const Problem& p);
  void forward_gimm(const Problem& p, const GimmIO& io, void* workspace, size_t workspace_bytes, gvStream_t stream);
  size_t plan(const Problem& p);                       // dry run -> workspace bytes
  void forward(const Problem& p, const IO& io, void* workspace, size_t workspace_bytes, gvStream_t stream);
  int64_t last_launches() const { return launches_; }
  // CUDA graphs: forward() records the launch sequence of a (problem, buffers, mode) combination the SECOND time it sees it and
  // replays the instantiated graph afterwards - one cudaGraphLaunch instead of ~550 kernel launches and ~1000 host-side tensor-map
  // encodes per forward.  The replay is exactly the recorded stream of kernels with the recorded parameters, so it is only used
  // when every pointer the launches captured is unchanged (key below); profiling, debug taps and the frame cache bypass it.
  void set_cuda_graph(bool on) { use_graph_ = on; if (!on) clear_graphs(); }
  int64_t graph_replays() const { return graph_replays_; }
  // debug taps: name -> NHWC view inside the workspace of the last forward
  void set_debug(bool on) { debug_ = on; }
  // per-kernel CUDA-event timing of the next forward(s): {"name": {"ms", "work", "launches"}}
  void set_profile(bool on) { profile_ = on; }
  // 0: fp32 CUDA cores everywhere; 1: post-RAFT convs on TF32 tensor cores; 2: + RAFT convs on 3xTF32 tensor cores;
  // 3: + final-decoder residual trunk in fp16; 4: + the 32/64-channel full-resolution chains in fp16 (experimental)
  void set_tensor_cores(int mode) { tc_mode_ = mode; }
  // precision mode 4 (experimental): the 32/64-channel full-resolution chains (motion encoder laterals, latent
  // refiner, final-decoder upsample branch) are stored in half precision like the residual trunk of mode 3
  bool half_chains(const Ctx& cx) const { return cx.tc && tc_mode_ >= 4; }
  std::string profile_json(gvStream_t stream);
  // Video callers (src/video_Nx.py:134-216) walk consecutive pairs (j, j+1), (j+1, j+2), ...: the RAFT encoder products of a
  // call's SECOND frame (fnet map, cnet net/inp, projected context features) can be kept in a caller-owned device buffer and
  // re-used as the next call's FIRST frame.  load: frame 0 comes from the cache (the caller guarantees it
  // is the frame the cache was stored from); store: frame 1's products are written to it.  Results are bit-identical.
  static size_t frame_cache_bytes(const Problem& p);
  void set_frame_cache(float* cache, size_t bytes, bool load, bool store) { fc_ = cache; fc_bytes_ = bytes; fc_load_ = load; fc_store_ = store; }
  const std::map<std::string, TV>& taps() const { return taps_; }
  std::string last_error;
  int raft_iters = 20;  // GIMMVFI_R hard-codes iters=20 (gimmvfi_r.py:126-132)
  int ff_iters = 32;    // FlowFormer decoder_depth (flowformer/configs/submission.py:50; gimmvfi_f.py:115 passes iters=None)
  bool is_f() const { return ff_; }
  const float* ff_c1w() const { return ff_c1w_.data(); }   // first cost-map patch-embedding layer, [36 taps][16] and [16] (host copies)
  const float* ff_c1b() const { return ff_c1b_.data(); }
  // debug taps from the estimator's helpers (flowformer.cu)
  bool debug_on() const { return debug_; }
  void tap_pub(const std::string& name, const TV& tv) { tap(name, tv); }
  void tap_copy(Ctx& cx, const std::string& name, const TV& tv);   // a snapshot (the tensor is overwritten later in the forward)
  int device() const { return device_; }
  int64_t weights_version() const { return weights_version_; }
  const void* hyponet_blob(bool fp32_class) const { return fp32_class ? hypo_blob3_ : hypo_blob_; }   // bumped by every finalize_weights*: callers key caches on it
  bool finalized() const { return finalized_; }

 private:
  struct Impl;
  void run(Ctx& cx, const Problem& p, const IO& io, const FlowInputs* fin = nullptr);
  void finalize_decoders();
  void finalize_flowformer();
  void pack_patch_conv(const std::string& name);
  void pack_twins(const std::string& prefix);
  void run_flowformer(Ctx& cx, Net& N, int B, const TV& img, const TV& flow_up, const TV& feat4, const TV& feat8, const TV& fproj);
  void run_gimm(Ctx& cx, const Problem& p, const GimmIO& io);
  void finalize_gimm_part();
  void gimm_encode(Net& N, const TV& nf, const TV& f01, const TV& f10, const TV& wts, const TV& X64);
  void gimm_decode(Net& N, const TV& X64, const TV& f01, const TV& f10, const TV& wts, const float* tdev, const float* coords_t, const TV& ninr,
                   bool tap_it);
  ConvW pack_conv(const std::string& name, const std::string& bn = "", float out_scale = 1.f, const std::vector<int>* perm = nullptr);
  const float* upload(const std::vector<float>& v);
  const float* vec(const std::string& key);
  const HostTensor& raw(const std::string& key) const;
  void tap(const std::string& name, const TV& tv) { if (debug_) taps_[name] = tv; }

  int device_ = 0;
  bool finalized_ = false, debug_ = false, profile_ = false, gimm_only_ = false, synth_only_ = false, ff_ = false;
  int tc_mode_ = 0;
  int64_t weights_version_ = 0;
  bool use_graph_ = false;
  int64_t graph_replays_ = 0;
  struct GraphEntry { std::vector<uint64_t> key; void* exec = nullptr; int64_t launches = 0; int seen = 0; };
  std::vector<GraphEntry> graphs_;
  void clear_graphs();
  int precise_ = 0;   // bit mask of post-RAFT stages in 3xTF32: 1 GIMM encoders / latent refiner, 2 HypoNet, 4 init decoder + update blocks, 8 final decoder, 16 combine
  std::vector<float> ff_c1w_, ff_c1b_;
  float* fc_ = nullptr; size_t fc_bytes_ = 0; bool fc_load_ = false, fc_store_ = false;
  // what the cache holds (host-side bookkeeping of the last store): a load with a different buffer / problem is refused
  // (one record per cache buffer, so several video streams can share an engine; cleared when the weights change)
  struct FcRec { int B, H, W, mode; };
  std::map<const float*, FcRec> fc_valid_;
  void pack_tc(ConvW& c, const std::vector<float>& packed);
  void pack_tc_f16(ConvW& c, const std::vector<float>& packed);
  void pack_stem(const std::string& name, const std::string& bn);
  void pack_xpacked(const std::string& name, int ldp);
  Profiler prof_;
  int64_t launches_ = 0;
  int sm_count_ = 148;
  std::map<std::string, HostTensor> raw_;
  std::map<std::string, ConvW> conv_;
  std::map<std::string, const float*> vec_;
  std::vector<void*> dev_allocs_;
  std::map<std::string, TV> taps_;
  const float* flow_absmax_ = nullptr;   // device, per sample: max |flow| over both directions of the forward in flight
  const void* hypo_blob_ = nullptr;   // packed parameters of the fused HypoNet kernel (common.h hypo::)
  const void* hypo_blob3_ = nullptr;  // ... of its fp32-class variant (common.h hypo3::), the default
  bool hypo_fast_ = false;            // true: TF32 / half-operand HypoNet kernel (0.56 ms vs the fp32-class one; fails the 1e-3 bound on real frames)
  bool gru_hoist_ = true;             // SepConvGRU: the context input's share of the gate convolutions is computed once per pair (GIMMVFI_GRU_HOIST=0: every iteration)
  int raft_direct_ = -1;              // RAFT's own lookups without a volume: -1 = when the volume pyramid exceeds raft_direct_auto_bytes_ (GIMMVFI_RAFT_CORR_DIRECT)
  double raft_direct_auto_bytes_ = 64e9;
  int corr_direct_max_t_ = 2;         // BidirCorrBlock: volume-free lookup when a pair is interpolated at <= this many timesteps (GIMMVFI_CORR_DIRECT; 0: always the volume)
  const float* g9_ = nullptr; const float* alpha_fe_ = nullptr; const float* alpha_v_ = nullptr;

  // op helpers used by run()
  friend struct Net;
};

}  // namespace gv
