// Engine: owns folded/packed weights, activation workspaces and the per-volume job state.
// Host-side mirror of reference inference.py:29-126 (predict) + unet3d.py:77-105 (forward),
// scheduling the sm_100a kernels layer by layer over batches of patches.
#pragma once

#include <map>
#include <string>
#include <vector>

#include "../../include/exaspim_b200.h"
#include "kernels.h"

namespace exa {

struct HostTensor {
  std::vector<int64_t> shape;
  int dtype = EXA_DTYPE_F32;
  std::vector<float> f32;
  std::vector<int64_t> i64;
};

struct ConvLayer {
  std::string conv_key, bn_key;
  int cin = 0, cout = 0;
  __nv_bfloat16* w_bf16 = nullptr;  // [27][cout][cin]   (tcgen05 path)
  __nv_bfloat16* w_zfold = nullptr; // [9 (ky,kx)][3 (kz=2,1,0)][cout][cin]  (z-folded path)
  float* w_f32 = nullptr;           // [27][cin][cout]   (fp32 validation path)
  float* bias = nullptr;            // [cout] folded
};

struct Plan {
  int D = 0, H = 0, W = 0;
  AxisGeom az, ay, ax;
  int n_patches = 0;
};

// pure host helpers (also exported through the C ABI)
std::vector<int> axis_starts(int dim, int patch, int overlap);
Status make_plan(int D, int H, int W, const exa_predict_params& p, Plan* plan);
Status plan_slab(const Plan& plan, int row_begin, int row_end, exa_slab_plan* out);
Status percentiles_from_hist(const uint64_t* hist, int nbins, double q_lo, double q_hi, double* mn,
                             double* mx);

class Engine {
 public:
  Engine(int device, int precision) : device_(device), precision_(precision) {}
  ~Engine();
  Status init();
  Status load_weight(const char* name, const void* data, const int64_t* shape, int ndim, int dtype);
  Status finalize_weights();
  int out_channels() const { return out_channels_; }

  Status forward(const float* x, float* logits, int batch, const int32_t patch[3], cudaStream_t s);
  Status predict_host(const uint16_t* vol, int D, int H, int W, const exa_predict_params& p,
                      float* out);
  Status predict_device(const uint16_t* vol_dev, int D, int H, int W, const exa_predict_params& p,
                        float* out_dev, cudaStream_t s);
  Status histogram(const uint16_t* vol_dev, int64_t n, int clip, uint64_t* hist_dev,
                   cudaStream_t s);
  Status set_normalization(double mn, double mx, int clip);
  Status set_normalization_table(const double* values, int n, double mn, double mx);
  Status slab_run(const uint16_t* slab_dev, int D, int H, int W, const exa_predict_params& p,
                  int row_begin, int row_end, cudaStream_t s);
  Status slab_partial(float* halo_dev, cudaStream_t s);
  // channel_stride = 0: dense (C, out planes, H, W); otherwise elements between channels of out_dev
  Status slab_stitch(const float* seed_dev, float* out_dev, int64_t channel_stride, cudaStream_t s);
  // rows [row_begin,row_end) as a row-group pipeline with overlapped D2H (out_host may be null);
  // slab_finish() completes the planes that need the previous rank's partial sums
  Status slab_predict(const uint16_t* slab_dev, int D, int H, int W, const exa_predict_params& p,
                      int row_begin, int row_end, float* out_dev, int64_t channel_stride,
                      float* out_host, int64_t host_channel_stride, float* halo_dev, cudaStream_t s);
  Status slab_finish(const float* seed_dev, float* out_dev, int64_t channel_stride, float* out_host,
                     int64_t host_channel_stride, cudaStream_t s);

  // fused gather: while set, every stitch whose destination lies inside [local_base, +elems)
  // also stores to the same offset of each peer base (peer-mapped copies of the same array)
  Status set_peer_outputs(float* local_base, int64_t elems, float* const* peer_bases, int n_peers);

  // per-category kernel timing with CUDA events on the launching stream (bench / roofline)
  enum Category { CAT_HIST = 0, CAT_STEM, CAT_CONV, CAT_POOL, CAT_UPSAMPLE, CAT_HEAD, CAT_STITCH,
                  CAT_COUNT };
  Status profile_begin();
  Status profile_end(double* ms_by_cat, int64_t* launches_by_cat, int n);
  // per conv layer (index = position in unet3d.py:64-74 order, 0 = stem) of the last profile_end;
  // kind: 0 none, 1 K1 (conv_umma), 2 K1z (conv_zfold), 3 K1z2 (conv_zfold2, CTA pairs)
  Status profile_layers(double* ms, int64_t* launches, int32_t* kind, int n) const;

  // progress of the current predict call: cb(user, done, total) from a CUDA host function each
  // time a wave of patches has finished on the device (tqdm bar of predict(verbose=True))
  void set_progress(exa_progress_fn cb, void* user) {
    progress_cb_ = cb;
    progress_user_ = user;
  }

  std::string last_error;
  int64_t launches = 0;

 private:
  Status ensure_workspace(int batch, int pz, int py, int px);
  Status predict_pipeline(const uint16_t* vol_dev, int D, int H, int W, const exa_predict_params& p,
                          float* out_dev, float* out_host, cudaStream_t s);
  Status pipeline_rows(const uint16_t* vol_dev, int vol_z0, int D, int H, int W,
                       const exa_predict_params& p, int R0, int R1, float* out_dev,
                       size_t out_cstride, float* out_host, size_t host_cstride, int out_zbase,
                       float* halo_out, bool defer_seed, cudaStream_t s);
  Status pipeline_finish(const float* seed_in, float* out_dev, size_t out_cstride, float* out_host,
                         size_t host_cstride, int out_zbase, cudaStream_t s);
  Status copy_planes_to_host(const float* out_dev, size_t out_cstride, float* out_host,
                             size_t host_cstride, int out_zbase, int z0, int z1, size_t plane,
                             cudaStream_t s, int y0 = 0, int y1 = 0);
  Status stitch_planes(const float* seed_dev, float* out_dev, size_t out_cstride, int z0, int z1,
                       cudaStream_t s, int y0 = 0, int y1 = 0);
  Status stream_band(int patches_done, int n_slab, cudaStream_t s);
  Status run_network(const PatchSource& src, int batch, int pz, int py, int px,
                     const HeadParams& head, cudaStream_t s);
  Status conv(const ConvLayer& L, const Act& in, const Act& out, const HeadParams* head,
              const ConvRegion* region, const Act* pool_out, cudaStream_t s);

  exa_progress_fn progress_cb_ = nullptr;
  void* progress_user_ = nullptr;
  int64_t progress_done_ = 0, progress_total_ = 0;
  struct ProfRec {
    int cat;
    int tag;  // conv: layer index 1..17; other categories: -1 (EXA_LAYER_PROF=1 prints per-tag sums)
    cudaEvent_t start, stop;
  };
  int cur_tag_ = -1;
  double layer_ms_[18] = {0};
  int64_t layer_n_[18] = {0};
  int32_t layer_kind_[18] = {0};
  struct Scope {  // counts the launch and, when profiling, brackets it with two events
    Engine* e;
    cudaStream_t s;
    cudaEvent_t stop = nullptr;
    Scope(Engine* eng, int cat, cudaStream_t st);
    ~Scope();
  };
  bool use_pair_ = true;   // EXA_NO_PAIR=1: one CTA per MMA (conv_zfold.cuh) instead of CTA pairs
  bool use_zfold_ = true;  // EXA_NO_ZFOLD=1 selects the plain per-tap kernel everywhere (A/B tests)
  bool prof_on_ = false;
  std::vector<ProfRec> prof_;

  int device_, precision_;
  int num_sms_ = 148;
  bool finalized_ = false;
  int out_channels_ = 0;
  std::map<std::string, HostTensor> raw_;
  ConvLayer layers_[18];
  // model configuration, read off the state_dict in finalize_weights (unet3d.py:37-75)
  int chan_[5] = {32, 64, 128, 256, 512};
  bool trilinear_ = true;
  struct UpConv {  // ConvTranspose3d(k=2, s=2) of an Up block when trilinear=False
    int cin = 0, cout = 0;
    float* w = nullptr;     // [8 taps][cin][cout] float32 (bf16-rounded values in bf16 mode)
    float* bias = nullptr;  // [cout]
  } upconv_[4];
  std::vector<StemWeights> stem_;  // one group of 32 output channels each
  __nv_bfloat16* stem_band_ = nullptr;  // [9 (kz,ky)][128 (xo,c)][16 (x',hi|lo)] Toeplitz weights (conv_stem.cuh)
  float* stem_bias_ = nullptr;          // [32]
  bool use_tc_stem_ = true;             // EXA_NO_TC_STEM=1: SIMT fp32 stem also in bf16 mode
  float* head_w_ = nullptr;
  float* head_b_ = nullptr;
  std::vector<float> head_w_host_, head_b_host_;

  // activation workspace
  void* ws_ = nullptr;
  size_t ws_bytes_ = 0;
  int ws_batch_ = 0, ws_p_[3] = {0, 0, 0};
  // normalisation LUT
  float* lut_ = nullptr;
  int lut_clip_ = -1;
  bool norm_set_ = false;
  // per-volume job
  Plan plan_;
  exa_slab_plan slab_{};
  int row_begin_ = 0, row_end_ = 0;
  bool job_ready_ = false;
  float* probs_ = nullptr;
  size_t probs_bytes_ = 0;
  std::vector<int> starts_host_;
  int* starts_dev_ = nullptr;
  size_t starts_cap_ = 0;
  unsigned long long* hist_dev_ = nullptr;
  // row-group pipeline (predict_pipeline)
  float* seed_ = nullptr;
  size_t seed_bytes_ = 0;
  void* vol_stage_ = nullptr;
  size_t vol_stage_bytes_ = 0;
  void* out_stage_ = nullptr;
  size_t out_stage_bytes_ = 0;
  cudaStream_t copy_stream_ = nullptr;
  cudaEvent_t copy_event_ = nullptr;
  // first row group of a deferred pipeline: its patches wait here for the previous rank's seed
  float* probs_hold_ = nullptr;
  size_t probs_hold_bytes_ = 0;
  bool held_valid_ = false;
  exa_slab_plan held_slab_{};
  Plan held_plan_;
  int held_rows_[2] = {0, 0};
  int pipe_zbase_ = 0;
  // y-band streaming inside a one-row group: rows of the group's planes are stitched and copied
  // to the host as soon as the patches covering them are done (set by pipeline_rows)
  struct BandSink {
    bool on = false;
    const float* seed = nullptr;
    float* out_dev = nullptr;
    size_t out_cstride = 0;
    float* out_host = nullptr;
    size_t host_cstride = 0;
    int out_zbase = 0, z0 = 0, z1 = 0, y_done = 0;
  } band_;
  // stitch stream: group k's stitch overlaps group k+1's convolutions (EXA_STITCH_OVERLAP=1: on)
  bool overlap_stitch_ = false;
  cudaStream_t stitch_stream_ = nullptr;
  cudaEvent_t conv_done_ = nullptr, probs_free_ = nullptr, alt_free_ = nullptr;
  float* probs_alt_ = nullptr;
  size_t probs_alt_bytes_ = 0;
  float* peer_local_ = nullptr;
  int64_t peer_elems_ = 0;
  std::vector<float*> peer_bases_;
  // finished planes go to the peers' copies by copy-engine transfers on side streams (no SM time,
  // overlapped with the next waves); EXA_GATHER=store: by stores from the stitch kernel instead
  bool peer_ce_ = true;
  static constexpr int kPeerStreams = 8;   // created; n_peer_streams_ of them are used
  int n_peer_streams_ = 7;                 // EXA_PEER_STREAMS (1..8); 7 = one per peer at 8 GPUs
  cudaStream_t peer_stream_[kPeerStreams] = {};
  cudaEvent_t peer_ready_ = nullptr, peer_done_[kPeerStreams] = {};
  bool peer_pending_ = false;
  Status copy_planes_to_peers(const float* out_dev, size_t out_cstride, int nz, int y0, int y1,
                              cudaStream_t s);
  Status join_peer_copies(cudaStream_t s);
};

}  // namespace exa
