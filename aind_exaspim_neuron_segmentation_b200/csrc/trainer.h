// Trainer: one training step of the reference U-Net on the device.
// Host-side mirror of reference machine_learning/train.py:123-157 (train_step) and 200-223
// (forward_pass) for `model.train(); hat_y = model(x); loss.backward()`: the forward pass of
// unet3d.py:77-105 with BatchNorm3d in training mode (batch statistics, running statistics
// updated in place) and the backward pass down to every parameter gradient.
//
// Parameters are NOT copied: the caller binds the device storage of every state_dict entry
// (torch parameters / buffers) once; each forward re-packs the conv weights from them.  The
// gradients of one backward land in one flat float32 buffer, one slot per parameter.
#pragma once

#include <map>
#include <string>
#include <vector>

#include "../../include/exaspim_b200.h"
#include "train_kernels.h"

namespace exa {

class Trainer {
 public:
  Trainer(int device, int precision) : device_(device), precision_(precision) {}
  ~Trainer();
  Status init();
  Status bind(const char* name, float* dev_ptr, const int64_t* shape, int ndim);
  Status grad_slot(const char* name, int64_t* offset, int64_t* numel);
  Status grad_elems(int64_t* n);
  int out_channels() const { return out_channels_; }
  // x: (B,1,D,H,W) float32, logits: (B,C,D,H,W) float32, both on the device
  Status forward(const float* x, int batch, const int32_t patch[3], float* logits, cudaStream_t s);
  // x: the tensor of the matching forward; dlogits: (B,C,D,H,W); grads: grad_elems() floats
  Status backward(const float* x, const float* dlogits, float* grads, cudaStream_t s);
  size_t workspace_bytes() const { return ws_bytes_; }
  // per-category kernel timing with CUDA events on the launching stream (bench / roofline)
  enum Category { CAT_PACK = 0, CAT_FPROP, CAT_BN_FWD, CAT_MISC_FWD, CAT_BN_BWD, CAT_WGRAD,
                  CAT_DGRAD, CAT_HEAD_BWD, CAT_WGRAD_STEM, CAT_WGRAD_REDUCE, CAT_UPSAMPLE_BWD,
                  CAT_POOL_BWD, CAT_COUNT };
  Status profile_begin();
  Status profile_end(double* ms_by_cat, int64_t* launches_by_cat, int n);

  std::string last_error;
  int64_t launches = 0;

 private:
  struct Bound {
    float* ptr = nullptr;
    std::vector<int64_t> shape;
  };
  struct Layer {
    std::string conv_key, bn_key;
    int cin = 0, cout = 0, lvl = 0;
    float *w = nullptr, *b = nullptr, *gamma = nullptr, *beta = nullptr, *rmean = nullptr,
          *rvar = nullptr;
    int64_t gw = 0, gb = 0, ggamma = 0, gbeta = 0;  // offsets in the flat gradient buffer
    void *w_fwd = nullptr, *w_fwd_zf = nullptr, *w_bwd = nullptr, *w_bwd_zf = nullptr;
    float *mean = nullptr, *rstd = nullptr, *scale = nullptr, *shift = nullptr, *coef = nullptr;
    double* sums = nullptr;  // [2 * cout]
    Act x, z, a, gx;         // conv input, raw conv output (encoded), activation, gradient of x
  };
  Status resolve();
  Status ensure_workspace(int batch, int pz, int py, int px);
  Status layer_forward(int l, const float* x, cudaStream_t s);
  Status layer_backward(int l, const TView& grad_a, const float* x, float* grads, cudaStream_t s);
  Status conv_any(const Act& in, const Act& out, const void* w_plain, const void* w_zf,
                  const float* bias, cudaStream_t s);

  struct ProfRec {
    int cat;
    cudaEvent_t start, stop;
  };
  struct Scope {  // counts the launch and, when profiling, brackets it with two events
    Trainer* t;
    cudaStream_t s;
    cudaEvent_t stop = nullptr;
    Scope(Trainer* tr, int cat, cudaStream_t st);
    ~Scope();
  };
  bool prof_on_ = false;
  std::vector<ProfRec> prof_;

  int device_, precision_;
  int num_sms_ = 148;
  bool resolved_ = false;
  std::map<std::string, Bound> bound_;
  std::map<std::string, std::pair<int64_t, int64_t>> slots_;
  int64_t grad_elems_ = 0;
  int chan_[5] = {32, 64, 128, 256, 512};
  int out_channels_ = 0;
  Layer layers_[18];
  float *head_w_ = nullptr, *head_b_ = nullptr;
  int64_t g_head_w_ = 0, g_head_b_ = 0;
  void* small_ = nullptr;  // per-layer statistics, packed weights
  size_t small_bytes_ = 0;
  float* zero_bias_ = nullptr;
  double* head_sums_ = nullptr;
  // workspace
  void* ws_ = nullptr;
  size_t ws_bytes_ = 0;
  int ws_batch_ = 0, ws_p_[3] = {0, 0, 0};
  Act a0_, cat4_, x1_, p1_, d1a_, cat3_, x2_, p2_, d2a_, cat2_, x3_, p3_, d3a_, cat1_, x4_, p4_,
      d4a_, x5_, u1a_, u1_, u2a_, u2_, u3a_, u3_, u4a_, u4_;
  Act up1_slot_, up2_slot_, up3_slot_, up4_slot_;
  Act g_u4_, g_u4a_, g_cat4_, g_u3_, g_u3a_, g_cat3_, g_u2_, g_u2a_, g_cat2_, g_u1_, g_u1a_,
      g_cat1_, g_x5_, g_d4a_, g_p4_, g_x4_, g_d3a_, g_p3_, g_x3_, g_d2a_, g_p2_, g_x2_, g_d1a_,
      g_p1_, g_x1_, g_a0_;
  Act dz_;                 // scratch for the largest raw-output gradient
  void* stem_in_ = nullptr;  // bf16 (hi, lo) split or fp32 16-channel expansion of the input
  float* partial_ = nullptr;
  bool forward_valid_ = false;
};

// Operator-level entry points of the two tensor-core pieces of the backward pass, for parity
// tests against torch.nn.grad.conv3d_weight / conv3d_input on identical operands.  Tensors are
// NDHWC device arrays of the precision's element type (bf16 or fp32); w and dw are float32 in
// the reference's (Cout, Cin, 3, 3, 3) layout.  Temporary buffers are allocated per call.
Status conv3d_weight_grad(int device, int precision, const void* x, const void* dz, int B, int D,
                          int H, int W, int cin, int cout, float* dw, cudaStream_t s);
Status conv3d_data_grad(int device, int precision, const void* dz, const float* w, int B, int D,
                        int H, int W, int cin, int cout, void* dx, cudaStream_t s);

}  // namespace exa
