// Launch API of the training-step kernels: the parts of
// reference machine_learning/train.py:123-157,200-223 (Trainer.train_step / forward_pass:
// model.train(); hat_y = model(x); loss.backward()) that are not already covered by the
// inference kernels.  Training-mode BatchNorm3d (batch statistics, unet3d.py:144,147), the
// backward of BatchNorm + LeakyReLU, MaxPool3d, trilinear Upsample, the 1x1x1 head, and the
// weight gradients of the 3x3x3 convolutions.  The forward convolutions and the data gradients
// (a 3x3x3 convolution with the flipped, transposed weights) run on the tcgen05 kernels of the
// inference path.  Everything is stream-ordered; nothing here synchronises.
#pragma once

#include "kernels.h"

namespace exa {

// The conv kernels' epilogue is bias + LeakyReLU(0.01) (unet3d.py:145).  Training needs the raw
// convolution, and LeakyReLU is a bijection: a tensor marked `enc` holds leaky(value) and its
// readers undo it (v < 0 ? 100 v : v).  In bf16 the round trip costs nothing beyond the one
// rounding of the output (same relative precision on both branches).
struct TView {
  Act a;
  bool enc = false;
};

// ---- weights: master fp32 (Cout, Cin, 3,3,3) -> operand layouts of the conv kernels ----------
// rows = Cout (row_is_cout) or Cin; flip: tap' = 26 - tap (data gradient); zfold: the
// [9 (ky,kx)][3 (kz=2,1,0)][rows][cols] form of conv_zfold2.cuh; fp32: float output
Status launch_pack_conv_weights(const float* w, void* out, int cout, int cin, bool row_is_cout,
                                bool flip, bool zfold, bool out_fp32, cudaStream_t s);
// Toeplitz band matrices of the tensor-core stem (conv_stem.cuh) from (32g, 1, 3,3,3) weights
Status launch_pack_stem_band(const float* w, __nv_bfloat16* band, int groups, cudaStream_t s);
// fp32 validation mode: stem weights as a 16-input-channel conv [27][16][Cout] (channel 0 = w)
Status launch_pack_stem_fp32(const float* w, float* out, int cout, cudaStream_t s);
// x (B,1,D,H,W) float32 -> [voxel][16] float32 with the value in channel 0
Status launch_expand_input16(const float* x, float* out, size_t voxels, cudaStream_t s);

// ---- BatchNorm3d, training mode (unet3d.py:144,147; eps 1e-5, momentum 0.1) --------------------
// sums[0..C) += sum z, sums[C..2C) += sum z^2 (double; zero them first)
Status launch_bn_stats(const TView& z, double* sums, cudaStream_t s);
// batch mean / biased variance -> mean, rstd, scale = gamma*rstd, shift = beta - mean*scale;
// running_mean / running_var updated in place (unbiased variance), like nn.BatchNorm3d
Status launch_bn_finalize(const double* sums, int C, double count, const float* gamma,
                          const float* beta, float* running_mean, float* running_var, float* mean,
                          float* rstd, float* scale, float* shift, cudaStream_t s);
// a = LeakyReLU(scale * z + shift)
Status launch_bn_apply(const TView& z, const float* scale, const float* shift, const Act& a,
                       cudaStream_t s);
// backward of BatchNorm + LeakyReLU: g = dA * (scale z + shift > 0 ? 1 : 0.01) -- the mask is
// recomputed from the raw conv output exactly as bn_apply computed the pre-activation, which
// saves reading the activation again;
// sums[0..C) += sum g, sums[C..2C) += sum g * xhat      (xhat = (z - mean) * rstd)
Status launch_bn_bwd_reduce(const TView& grad_a, const float* scale, const float* shift,
                            const TView& z, const float* mean, const float* rstd, double* sums,
                            cudaStream_t s);
// dgamma, dbeta -> gradient slots; coef[0..C) = gamma*rstd, [C..2C) = sum g / n, [2C..3C) = sum g xhat / n
Status launch_bn_bwd_finalize(const double* sums, int C, double count, const float* gamma,
                              const float* rstd, float* dgamma, float* dbeta, float* coef,
                              cudaStream_t s);
// dz = k (g - c1 - xhat c2), dense, plain; bias_sums[0..C) += sum dz (double; the conv bias gradient)
Status launch_bn_bwd_apply(const TView& grad_a, const float* scale, const float* shift,
                           const TView& z, const float* mean, const float* rstd, const float* coef,
                           const Act& dz, double* bias_sums, cudaStream_t s);
Status launch_double_to_float(const double* in, float* out, int n, cudaStream_t s);
// out = plain values of an (encoded) tensor
Status launch_decode(const TView& in, const Act& out, cudaStream_t s);

// ---- MaxPool3d(2) backward merged with the skip connection's gradient --------------------------
// out[v] = skip[v] + (v is the first maximum of its 2x2x2 window of `a` ? pooled[window] : 0)
Status launch_pool_bwd_merge(const TView& skip, const TView& pooled, const Act& a, const Act& out,
                             cudaStream_t s);
// ---- trilinear x2 upsample (align_corners=True) backward: the adjoint, in gather form ----------
Status launch_upsample_bwd(const TView& grad_out, const Act& grad_in, cudaStream_t s);

// ---- 1x1x1 head (unet3d.py:318) backward --------------------------------------------------------
// dlogits: (B, C, D, H, W) float32; du = W^T dlogits (dense, plain)
Status launch_head_bwd_dx(const float* dlogits, const float* hw, int C, const Act& du,
                          cudaStream_t s);
// sums[0 .. C*cin) += dW, sums[C*cin .. C*cin + C) += db   (double; zero them first)
Status launch_head_bwd_dw(const float* dlogits, const Act& u, int C, double* sums, cudaStream_t s);

// ---- weight gradient of a 3x3x3 convolution -----------------------------------------------------
// dW[co][ci][tap] = sum_v dz[v][co] * x[v + off(tap)][ci]  (zero padding), written as `splits`
// partial sums [splits][cout][cin][27] that launch_wgrad_reduce adds up (deterministic).
int wgrad_splits(const Act& x, int cout, int num_sms);
size_t wgrad_partial_elems(const Act& x, int cout, int num_sms);
// bf16: tcgen05 kernel of train_wgrad.cu (EXA_WGRAD=mma: the warp-level m16n8k16 kernel of
// train_kernels.cu instead), fp32 accumulation; fp32: SIMT
bool wgrad_tc_enabled();
int wgrad_tc_splits(const Act& x, int cout, int num_sms);
Status launch_wgrad_tc(const Act& x, const Act& dz, float* partial, int num_sms, cudaStream_t s);
Status launch_wgrad(const Act& x, const Act& dz, float* partial, int num_sms, cudaStream_t s);
// stem (Cin = 1): x is the raw (B,1,D,H,W) float32 input
Status launch_wgrad_stem(const float* x, const Act& dz, float* partial, int num_sms,
                         cudaStream_t s);
int wgrad_stem_splits(const Act& dz, int num_sms);
Status launch_wgrad_reduce(const float* partial, int splits, size_t elems, float* dw,
                           cudaStream_t s);

// ---- BCEWithLogitsLoss (train.py:76,222), mean reduction, with its gradient ---------------------
// loss_sum[0] += sum of the per-element losses (double; zero it first); grad = grad_scale *
// (sigmoid(x) - y) / n when grad is not null
Status launch_bce_with_logits(const float* logits, const float* target, size_t n, float grad_scale,
                              double* loss_sum, float* grad, cudaStream_t s);

}  // namespace exa
