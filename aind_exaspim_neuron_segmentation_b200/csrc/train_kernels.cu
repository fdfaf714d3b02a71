// Kernels of the training step (reference machine_learning/train.py:123-157,
// 200-223 with the model of unet3d.py:77-105 in train() mode).  See train_kernels.h for what
// each launch computes.  Activations and gradients are NDHWC (bf16, or fp32 in validation mode);
// per-channel reductions accumulate short fp32 runs into double atomics.
//
//   T1  weight packing (fp32 master -> bf16 operand layouts, forward and flipped/transposed)
//   T2  BatchNorm batch statistics / apply / backward (two reductions + one apply)
//   T3  max-pool backward merged with the skip gradient, upsample adjoint, head backward
//   T4  conv weight gradient: mma.sync m16n8k16 bf16 with fp32 accumulators held in registers
//       for all 27 taps of a 32x32 (Cout x Cin) block, deterministic split-K reduction
//   T5  BCEWithLogitsLoss + gradient
#include "train_kernels.h"

namespace exa {

namespace {

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&f)[8]);
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 raw = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&f)[8]);
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 raw;
  raw.x = pack_bf16x2(f[0], f[1]);
  raw.y = pack_bf16x2(f[2], f[3]);
  raw.z = pack_bf16x2(f[4], f[5]);
  raw.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = raw;
}
template <>
__device__ __forceinline__ void store8<float>(float* p, const float (&f)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}
// undo the conv epilogue's LeakyReLU(0.01) (train_kernels.h, TView)
__device__ __forceinline__ float dec(float v, bool enc) { return (enc && v < 0.f) ? v * 100.f : v; }

template <typename T>
__device__ __forceinline__ void load8_dec(const T* p, bool enc, float (&f)[8]) {
  load8<T>(p, f);
  if (enc) {
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = dec(f[j], true);
  }
}

bool vec_ok(const Act& a) { return a.C % 8 == 0 && a.cstride % 8 == 0 && a.coff % 8 == 0; }
bool same_grid(const Act& a, const Act& b) {
  return a.B == b.B && a.D == b.D && a.H == b.H && a.W == b.W;
}

constexpr int RED_THREADS = 256;
constexpr int RED_ITER = 32;  // voxels per thread and block: fp32 partial sums stay short

}  // namespace

// Raw 8-channel vectors: the loads of several voxels are issued before any is converted, so
// that enough bytes are in flight per thread (these kernels hold ~70 per-channel constants in
// registers, which limits occupancy).
template <typename T>
struct Raw8;
template <>
struct Raw8<__nv_bfloat16> {
  uint4 v;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { v = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void get(float (&f)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
};
template <>
struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void get(float (&f)[8]) const {
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
};
constexpr int BWD_UNROLL = 4;

// ---------------------------------------------------------------------------
// T1: weight packing
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
pack_conv_weights_kernel(const float* __restrict__ w, T* __restrict__ out, int cout, int cin,
                         int row_is_cout, int flip, int zfold) {
  const int rows = row_is_cout ? cout : cin, cols = row_is_cout ? cin : cout;
  const size_t n = (size_t)27 * rows * cols;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % cols);
  size_t r = i / cols;
  const int row = (int)(r % rows);
  const int t = (int)(r / rows);  // plain: tap'; zfold: t9 * 3 + kzr
  int tap = t;
  if (zfold) {
    const int t9 = t / 3, kzr = t % 3;
    tap = (2 - kzr) * 9 + t9;
  }
  if (flip) tap = 26 - tap;
  const int co = row_is_cout ? row : c, ci = row_is_cout ? c : row;
  const float v = w[((size_t)co * cin + ci) * 27 + tap];
  out[i] = (T)v;
}

Status launch_pack_conv_weights(const float* w, void* out, int cout, int cin, bool row_is_cout,
                                bool flip, bool zfold, bool out_fp32, cudaStream_t s) {
  EXA_CHECK(w && out && cout > 0 && cin > 0, "pack_conv_weights: bad arguments");
  const size_t n = (size_t)27 * cout * cin;
  const int blocks = (int)ceil_div64((int64_t)n, 256);
  if (out_fp32)
    pack_conv_weights_kernel<float><<<blocks, 256, 0, s>>>(w, (float*)out, cout, cin, row_is_cout,
                                                            flip, zfold);
  else
    pack_conv_weights_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(
        w, (__nv_bfloat16*)out, cout, cin, row_is_cout, flip, zfold);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

// band[g][t9][n = xo*32 + c][k = 2*x' + part] = w[32g + c][t9*3 + (x' - xo)] (both parts), else 0
__global__ void __launch_bounds__(256)
pack_stem_band_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ band, int groups) {
  const int per_group = 9 * 128 * 16;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per_group * groups) return;
  const int g = i / per_group;
  int r = i % per_group;
  const int k = r % 16;
  r /= 16;
  const int n = r % 128, t9 = r / 128;
  const int xo = n / 32, c = n % 32;
  const int kx = k / 2 - xo;
  float v = 0.f;
  if (kx >= 0 && kx < 3) v = w[(size_t)(g * 32 + c) * 27 + t9 * 3 + kx];
  band[i] = __float2bfloat16_rn(v);
}

Status launch_pack_stem_band(const float* w, __nv_bfloat16* band, int groups, cudaStream_t s) {
  const int n = 9 * 128 * 16 * groups;
  pack_stem_band_kernel<<<ceil_div(n, 256), 256, 0, s>>>(w, band, groups);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

__global__ void __launch_bounds__(256)
pack_stem_fp32_kernel(const float* __restrict__ w, float* __restrict__ out, int cout) {
  const int n = 27 * 16 * cout;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int co = i % cout;
  const int ci = (i / cout) % 16;
  const int tap = i / (cout * 16);
  out[i] = ci == 0 ? w[(size_t)co * 27 + tap] : 0.f;
}

Status launch_pack_stem_fp32(const float* w, float* out, int cout, cudaStream_t s) {
  const int n = 27 * 16 * cout;
  pack_stem_fp32_kernel<<<ceil_div(n, 256), 256, 0, s>>>(w, out, cout);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

__global__ void __launch_bounds__(256)
expand_input16_kernel(const float* __restrict__ x, float* __restrict__ out, size_t voxels) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // one float4 of the output
  if (i >= voxels * 4) return;
  const size_t v = i / 4;
  const int q = (int)(i % 4);
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  if (q == 0) o.x = x[v];
  *reinterpret_cast<float4*>(out + v * 16 + q * 4) = o;
}

Status launch_expand_input16(const float* x, float* out, size_t voxels, cudaStream_t s) {
  expand_input16_kernel<<<(unsigned)ceil_div64((int64_t)voxels * 4, 256), 256, 0, s>>>(x, out,
                                                                                        voxels);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

// ---------------------------------------------------------------------------
// T2: BatchNorm3d in training mode
// ---------------------------------------------------------------------------
// Shared tail of the per-channel reductions: every thread holds two 8-channel partial sums of
// its lane; lanes are added in double and one atomic per channel and block goes to global.
__device__ __forceinline__ void block_channel_reduce(const float (&p0)[8], const float (&p1)[8],
                                                     int cv, int C, double* __restrict__ sums) {
  __shared__ float sh[2][RED_THREADS * 8];
  const int tid = threadIdx.x;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sh[0][tid * 8 + j] = p0[j];
    sh[1][tid * 8 + j] = p1[j];
  }
  __syncthreads();
  const int lanes = RED_THREADS / cv;
  for (int ch = tid; ch < C; ch += RED_THREADS) {
    const int c8 = ch / 8, j = ch % 8;
    double a0 = 0.0, a1 = 0.0;
    for (int l = 0; l < lanes; ++l) {
      a0 += (double)sh[0][(l * cv + c8) * 8 + j];
      a1 += (double)sh[1][(l * cv + c8) * 8 + j];
    }
    atomicAdd(sums + ch, a0);
    atomicAdd(sums + C + ch, a1);
  }
}

template <typename T>
__global__ void __launch_bounds__(RED_THREADS)
bn_stats_kernel(const T* __restrict__ z, int cstride, int coff, int C, bool enc, size_t voxels,
                double* __restrict__ sums) {
  const int cv = C / 8, lanes = RED_THREADS / cv;
  const int c8 = threadIdx.x % cv, lane = threadIdx.x / cv;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  const size_t v0 = (size_t)blockIdx.x * lanes * RED_ITER;
  for (int k = 0; k < RED_ITER; ++k) {
    const size_t v = v0 + (size_t)k * lanes + lane;
    if (v >= voxels) break;
    float f[8];
    load8_dec<T>(z + v * cstride + coff + 8 * c8, enc, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j] += f[j];
      q[j] = fmaf(f[j], f[j], q[j]);
    }
  }
  block_channel_reduce(s, q, cv, C, sums);
}

static Status check_reduce_shape(const Act& a, const char* what) {
  EXA_CHECK(vec_ok(a) && RED_THREADS % (a.C / 8) == 0 && a.C <= 2048,
            std::string(what) + ": channels must be a multiple of 8 dividing 2048");
  return Status::OK();
}

Status launch_bn_stats(const TView& z, double* sums, cudaStream_t s) {
  EXA_TRY(check_reduce_shape(z.a, "bn_stats"));
  const size_t vox = z.a.voxels();
  const int lanes = RED_THREADS / (z.a.C / 8);
  const unsigned blocks = (unsigned)ceil_div64((int64_t)vox, (int64_t)lanes * RED_ITER);
  if (z.a.fp32)
    bn_stats_kernel<float><<<blocks, RED_THREADS, 0, s>>>((const float*)z.a.ptr, z.a.cstride,
                                                           z.a.coff, z.a.C, z.enc, vox, sums);
  else
    bn_stats_kernel<__nv_bfloat16><<<blocks, RED_THREADS, 0, s>>>(
        (const __nv_bfloat16*)z.a.ptr, z.a.cstride, z.a.coff, z.a.C, z.enc, vox, sums);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, int C, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ mean, float* __restrict__ rstd,
                                   float* __restrict__ scale, float* __restrict__ shift) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    const double m = sums[c] / count;
    double var = sums[C + c] / count - m * m;  // biased (what normalises the batch)
    if (var < 0.0) var = 0.0;
    const double r = 1.0 / sqrt(var + 1e-5);
    mean[c] = (float)m;
    rstd[c] = (float)r;
    const double sc = (double)gamma[c] * r;
    scale[c] = (float)sc;
    shift[c] = (float)((double)beta[c] - m * sc);
    if (running_mean) {
      // nn.BatchNorm3d: momentum 0.1, running_var takes the unbiased estimate
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      running_mean[c] = (float)(0.9 * (double)running_mean[c] + 0.1 * m);
      running_var[c] = (float)(0.9 * (double)running_var[c] + 0.1 * unbiased);
    }
  }
}

Status launch_bn_finalize(const double* sums, int C, double count, const float* gamma,
                          const float* beta, float* running_mean, float* running_var, float* mean,
                          float* rstd, float* scale, float* shift, cudaStream_t s) {
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, s>>>(sums, C, count, gamma, beta, running_mean,
                                                     running_var, mean, rstd, scale, shift);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const T* __restrict__ z, int z_cstride, int z_coff, bool enc,
                const float* __restrict__ scale, const float* __restrict__ shift,
                T* __restrict__ a, int a_cstride, int a_coff, int C, size_t voxels) {
  const int cv = C / 8;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= voxels * cv) return;
  const int c8 = (int)(i % cv);
  const size_t v = i / cv;
  float f[8];
  load8_dec<T>(z + v * z_cstride + z_coff + 8 * c8, enc, f);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    f[j] = leaky_relu(fmaf(f[j], __ldg(scale + 8 * c8 + j), __ldg(shift + 8 * c8 + j)));
  store8<T>(a + v * a_cstride + a_coff + 8 * c8, f);
}

Status launch_bn_apply(const TView& z, const float* scale, const float* shift, const Act& a,
                       cudaStream_t s) {
  EXA_CHECK(vec_ok(z.a) && vec_ok(a) && z.a.C == a.C && same_grid(z.a, a) && z.a.fp32 == a.fp32,
            "bn_apply: shape mismatch");
  const size_t vox = a.voxels();
  const unsigned blocks = (unsigned)ceil_div64((int64_t)(vox * (a.C / 8)), 256);
  if (a.fp32)
    bn_apply_kernel<float><<<blocks, 256, 0, s>>>((const float*)z.a.ptr, z.a.cstride, z.a.coff,
                                                   z.enc, scale, shift, (float*)a.ptr, a.cstride,
                                                   a.coff, a.C, vox);
  else
    bn_apply_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(
        (const __nv_bfloat16*)z.a.ptr, z.a.cstride, z.a.coff, z.enc, scale, shift,
        (__nv_bfloat16*)a.ptr, a.cstride, a.coff, a.C, vox);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

template <typename T>
__global__ void __launch_bounds__(RED_THREADS)
bn_bwd_reduce_kernel(const T* __restrict__ g, int g_cstride, int g_coff, bool g_enc,
                     const float* __restrict__ scale, const float* __restrict__ shift,
                     const T* __restrict__ z, int z_cstride, int z_coff, bool z_enc,
                     const float* __restrict__ mean, const float* __restrict__ rstd, int C,
                     size_t voxels, double* __restrict__ sums) {
  const int cv = C / 8, lanes = RED_THREADS / cv;
  const int c8 = threadIdx.x % cv, lane = threadIdx.x / cv;
  float m[8], r[8], sc[8], sh[8], s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    m[j] = __ldg(mean + 8 * c8 + j);
    r[j] = __ldg(rstd + 8 * c8 + j);
    sc[j] = __ldg(scale + 8 * c8 + j);
    sh[j] = __ldg(shift + 8 * c8 + j);
    s[j] = q[j] = 0.f;
  }
  const size_t v0 = (size_t)blockIdx.x * lanes * RED_ITER;
  for (int k = 0; k < RED_ITER; k += BWD_UNROLL) {
    Raw8<T> rg[BWD_UNROLL], rz[BWD_UNROLL];
#pragma unroll
    for (int u = 0; u < BWD_UNROLL; ++u) {
      const size_t v = v0 + (size_t)(k + u) * lanes + lane;
      if (v < voxels) {
        rg[u].load(g + v * g_cstride + g_coff + 8 * c8);
        rz[u].load(z + v * z_cstride + z_coff + 8 * c8);
      }
    }
#pragma unroll
    for (int u = 0; u < BWD_UNROLL; ++u) {
      const size_t v = v0 + (size_t)(k + u) * lanes + lane;
      if (v < voxels) {
        float fg[8], fz[8];
        rg[u].get(fg);
        rz[u].get(fz);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float gv = dec(fg[j], g_enc), zv = dec(fz[j], z_enc);
          // LeakyReLU mask from the pre-activation, recomputed exactly as bn_apply computed it
          const float pre = fmaf(zv, sc[j], sh[j]);
          const float gg = pre > 0.f ? gv : 0.01f * gv;
          const float xh = (zv - m[j]) * r[j];
          s[j] += gg;
          q[j] = fmaf(gg, xh, q[j]);
        }
      }
    }
  }
  block_channel_reduce(s, q, cv, C, sums);
}

Status launch_bn_bwd_reduce(const TView& grad_a, const float* scale, const float* shift,
                            const TView& z, const float* mean, const float* rstd, double* sums,
                            cudaStream_t s) {
  EXA_TRY(check_reduce_shape(z.a, "bn_bwd_reduce"));
  EXA_CHECK(vec_ok(grad_a.a) && grad_a.a.C == z.a.C && same_grid(grad_a.a, z.a) &&
                grad_a.a.fp32 == z.a.fp32,
            "bn_bwd_reduce: shape mismatch");
  const size_t vox = z.a.voxels();
  const int lanes = RED_THREADS / (z.a.C / 8);
  const unsigned blocks = (unsigned)ceil_div64((int64_t)vox, (int64_t)lanes * RED_ITER);
  if (z.a.fp32)
    bn_bwd_reduce_kernel<float><<<blocks, RED_THREADS, 0, s>>>(
        (const float*)grad_a.a.ptr, grad_a.a.cstride, grad_a.a.coff, grad_a.enc, scale, shift,
        (const float*)z.a.ptr, z.a.cstride, z.a.coff, z.enc, mean, rstd, z.a.C, vox, sums);
  else
    bn_bwd_reduce_kernel<__nv_bfloat16><<<blocks, RED_THREADS, 0, s>>>(
        (const __nv_bfloat16*)grad_a.a.ptr, grad_a.a.cstride, grad_a.a.coff, grad_a.enc, scale,
        shift, (const __nv_bfloat16*)z.a.ptr, z.a.cstride, z.a.coff, z.enc, mean, rstd, z.a.C, vox,
        sums);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

__global__ void bn_bwd_finalize_kernel(const double* __restrict__ sums, int C, double count,
                                       const float* __restrict__ gamma,
                                       const float* __restrict__ rstd, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, float* __restrict__ coef) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    const double sg = sums[c], sgx = sums[C + c];
    dbeta[c] = (float)sg;
    dgamma[c] = (float)sgx;
    coef[c] = gamma[c] * rstd[c];
    coef[C + c] = (float)(sg / count);
    coef[2 * C + c] = (float)(sgx / count);
  }
}

Status launch_bn_bwd_finalize(const double* sums, int C, double count, const float* gamma,
                              const float* rstd, float* dgamma, float* dbeta, float* coef,
                              cudaStream_t s) {
  bn_bwd_finalize_kernel<<<ceil_div(C, 128), 128, 0, s>>>(sums, C, count, gamma, rstd, dgamma,
                                                         dbeta, coef);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

template <typename T>
__global__ void __launch_bounds__(RED_THREADS)
bn_bwd_apply_kernel(const T* __restrict__ g, int g_cstride, int g_coff, bool g_enc,
                    const float* __restrict__ scale, const float* __restrict__ shift,
                    const T* __restrict__ z, int z_cstride, int z_coff, bool z_enc,
                    const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ coef, T* __restrict__ dz, int C, size_t voxels,
                    double* __restrict__ bias_sums) {
  const int cv = C / 8, lanes = RED_THREADS / cv;
  const int c8 = threadIdx.x % cv, lane = threadIdx.x / cv;
  // dz = k0 (g - c1 - xhat c2) with xhat = (z - m) r  ==  k0 g - z (k0 c2 r) - k0 (c1 - m r c2)
  float sc[8], sh[8], k0[8], kz[8], kc[8], s[8], unused[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float m = __ldg(mean + 8 * c8 + j), r = __ldg(rstd + 8 * c8 + j);
    const float c1 = __ldg(coef + C + 8 * c8 + j), c2 = __ldg(coef + 2 * C + 8 * c8 + j);
    sc[j] = __ldg(scale + 8 * c8 + j);
    sh[j] = __ldg(shift + 8 * c8 + j);
    k0[j] = __ldg(coef + 8 * c8 + j);
    kz[j] = k0[j] * c2 * r;
    kc[j] = k0[j] * (c1 - m * r * c2);
    s[j] = 0.f;
    unused[j] = 0.f;
  }
  const size_t v0 = (size_t)blockIdx.x * lanes * RED_ITER;
  for (int k = 0; k < RED_ITER; k += BWD_UNROLL) {
    Raw8<T> rg[BWD_UNROLL], rz[BWD_UNROLL];
#pragma unroll
    for (int u = 0; u < BWD_UNROLL; ++u) {
      const size_t v = v0 + (size_t)(k + u) * lanes + lane;
      if (v < voxels) {
        rg[u].load(g + v * g_cstride + g_coff + 8 * c8);
        rz[u].load(z + v * z_cstride + z_coff + 8 * c8);
      }
    }
#pragma unroll
    for (int u = 0; u < BWD_UNROLL; ++u) {
      const size_t v = v0 + (size_t)(k + u) * lanes + lane;
      if (v < voxels) {
        float fg[8], fz[8], o[8];
        rg[u].get(fg);
        rz[u].get(fz);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float gv = dec(fg[j], g_enc), zv = dec(fz[j], z_enc);
          const float pre = fmaf(zv, sc[j], sh[j]);
          const float gg = pre > 0.f ? gv : 0.01f * gv;
          o[j] = fmaf(k0[j], gg, -fmaf(zv, kz[j], kc[j]));
          s[j] += o[j];
        }
        store8<T>(dz + v * C + 8 * c8, o);
      }
    }
  }
  // sums[0..C) take the bias gradient; the second half of the reduction is unused (zeros)
  block_channel_reduce(s, unused, cv, C, bias_sums);
}

Status launch_bn_bwd_apply(const TView& grad_a, const float* scale, const float* shift,
                           const TView& z, const float* mean, const float* rstd, const float* coef,
                           const Act& dz, double* bias_sums, cudaStream_t s) {
  EXA_TRY(check_reduce_shape(z.a, "bn_bwd_apply"));
  EXA_CHECK(vec_ok(grad_a.a) && grad_a.a.C == z.a.C && same_grid(grad_a.a, z.a) &&
                same_grid(dz, z.a) && dz.C == z.a.C && dz.cstride == dz.C && dz.coff == 0 &&
                dz.fp32 == z.a.fp32 && grad_a.a.fp32 == z.a.fp32,
            "bn_bwd_apply: shape mismatch (dz must be dense)");
  const size_t vox = z.a.voxels();
  const int lanes = RED_THREADS / (z.a.C / 8);
  const unsigned blocks = (unsigned)ceil_div64((int64_t)vox, (int64_t)lanes * RED_ITER);
  if (z.a.fp32)
    bn_bwd_apply_kernel<float><<<blocks, RED_THREADS, 0, s>>>(
        (const float*)grad_a.a.ptr, grad_a.a.cstride, grad_a.a.coff, grad_a.enc, scale, shift,
        (const float*)z.a.ptr, z.a.cstride, z.a.coff, z.enc, mean, rstd, coef, (float*)dz.ptr,
        z.a.C, vox, bias_sums);
  else
    bn_bwd_apply_kernel<__nv_bfloat16><<<blocks, RED_THREADS, 0, s>>>(
        (const __nv_bfloat16*)grad_a.a.ptr, grad_a.a.cstride, grad_a.a.coff, grad_a.enc, scale,
        shift, (const __nv_bfloat16*)z.a.ptr, z.a.cstride, z.a.coff, z.enc, mean, rstd, coef,
        (__nv_bfloat16*)dz.ptr, z.a.C, vox, bias_sums);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

__global__ void double_to_float_kernel(const double* __restrict__ in, float* __restrict__ out,
                                       int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}

Status launch_double_to_float(const double* in, float* out, int n, cudaStream_t s) {
  double_to_float_kernel<<<ceil_div(n, 256), 256, 0, s>>>(in, out, n);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

template <typename T>
__global__ void __launch_bounds__(256)
decode_kernel(const T* __restrict__ in, int i_cstride, int i_coff, bool enc, T* __restrict__ out,
              int o_cstride, int o_coff, int C, size_t voxels) {
  const int cv = C / 8;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= voxels * cv) return;
  const int c8 = (int)(i % cv);
  const size_t v = i / cv;
  float f[8];
  load8_dec<T>(in + v * i_cstride + i_coff + 8 * c8, enc, f);
  store8<T>(out + v * o_cstride + o_coff + 8 * c8, f);
}

Status launch_decode(const TView& in, const Act& out, cudaStream_t s) {
  EXA_CHECK(vec_ok(in.a) && vec_ok(out) && in.a.C == out.C && same_grid(in.a, out) &&
                in.a.fp32 == out.fp32,
            "decode: shape mismatch");
  const size_t vox = out.voxels();
  const unsigned blocks = (unsigned)ceil_div64((int64_t)(vox * (out.C / 8)), 256);
  if (out.fp32)
    decode_kernel<float><<<blocks, 256, 0, s>>>((const float*)in.a.ptr, in.a.cstride, in.a.coff,
                                                 in.enc, (float*)out.ptr, out.cstride, out.coff,
                                                 out.C, vox);
  else
    decode_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(
        (const __nv_bfloat16*)in.a.ptr, in.a.cstride, in.a.coff, in.enc, (__nv_bfloat16*)out.ptr,
        out.cstride, out.coff, out.C, vox);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

// ---------------------------------------------------------------------------
// T3: max-pool backward (+ skip gradient), upsample adjoint, head backward
// ---------------------------------------------------------------------------
// One thread: one 2x2x2 window x 8 channels.  The gradient of a window goes to its FIRST maximum
// in (z, y, x) scan order -- the index ATen's max_pool3d forward records (strict '>' update).
template <typename T>
__global__ void __launch_bounds__(256)
pool_bwd_merge_kernel(const T* __restrict__ skip, int s_cstride, int s_coff, bool s_enc,
                      const T* __restrict__ pooled, int p_cstride, int p_coff, bool p_enc,
                      const T* __restrict__ a, int a_cstride, int a_coff, T* __restrict__ out,
                      int o_cstride, int o_coff, int B, int Do, int Ho, int Wo, int C) {
  const int cv = C / 8;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * Do * Ho * Wo * cv;
  if (i >= total) return;
  const int c8 = (int)(i % cv);
  size_t w = i / cv;  // window index == pooled voxel index
  const size_t pv = w;
  const int xo = (int)(w % Wo); w /= Wo;
  const int yo = (int)(w % Ho); w /= Ho;
  const int zo = (int)(w % Do);
  const int b = (int)(w / Do);
  const int Di = 2 * Do, Hi = 2 * Ho, Wi = 2 * Wo;
  size_t vox[8];
  float best[8];
  int arg[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    best[j] = -INFINITY;
    arg[j] = 0;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int dz = k >> 2, dy = (k >> 1) & 1, dx = k & 1;
    vox[k] = (((size_t)b * Di + 2 * zo + dz) * Hi + 2 * yo + dy) * Wi + 2 * xo + dx;
    float f[8];
    load8<T>(a + vox[k] * a_cstride + a_coff + 8 * c8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (f[j] > best[j] || f[j] != f[j]) {
        best[j] = f[j];
        arg[j] = k;
      }
  }
  float gp[8];
  load8_dec<T>(pooled + pv * p_cstride + p_coff + 8 * c8, p_enc, gp);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float gs[8];
    load8_dec<T>(skip + vox[k] * s_cstride + s_coff + 8 * c8, s_enc, gs);
#pragma unroll
    for (int j = 0; j < 8; ++j) gs[j] += (arg[j] == k) ? gp[j] : 0.f;
    store8<T>(out + vox[k] * o_cstride + o_coff + 8 * c8, gs);
  }
}

Status launch_pool_bwd_merge(const TView& skip, const TView& pooled, const Act& a, const Act& out,
                             cudaStream_t s) {
  EXA_CHECK(vec_ok(skip.a) && vec_ok(pooled.a) && vec_ok(a) && vec_ok(out) && skip.a.C == a.C &&
                pooled.a.C == a.C && out.C == a.C && same_grid(skip.a, a) && same_grid(out, a) &&
                pooled.a.B == a.B && 2 * pooled.a.D == a.D && 2 * pooled.a.H == a.H &&
                2 * pooled.a.W == a.W,
            "pool_bwd_merge: shape mismatch");
  EXA_CHECK(skip.a.fp32 == a.fp32 && pooled.a.fp32 == a.fp32 && out.fp32 == a.fp32,
            "pool_bwd_merge: type mismatch");
  const size_t total = pooled.a.voxels() * (a.C / 8);
  const unsigned blocks = (unsigned)ceil_div64((int64_t)total, 256);
  if (a.fp32)
    pool_bwd_merge_kernel<float><<<blocks, 256, 0, s>>>(
        (const float*)skip.a.ptr, skip.a.cstride, skip.a.coff, skip.enc,
        (const float*)pooled.a.ptr, pooled.a.cstride, pooled.a.coff, pooled.enc,
        (const float*)a.ptr, a.cstride, a.coff, (float*)out.ptr, out.cstride, out.coff, a.B,
        pooled.a.D, pooled.a.H, pooled.a.W, a.C);
  else
    pool_bwd_merge_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(
        (const __nv_bfloat16*)skip.a.ptr, skip.a.cstride, skip.a.coff, skip.enc,
        (const __nv_bfloat16*)pooled.a.ptr, pooled.a.cstride, pooled.a.coff, pooled.enc,
        (const __nv_bfloat16*)a.ptr, a.cstride, a.coff, (__nv_bfloat16*)out.ptr, out.cstride,
        out.coff, a.B, pooled.a.D, pooled.a.H, pooled.a.W, a.C);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

// Weight with which output index o (0 <= o < 2n) of the forward interpolation reads input i:
// src = o (n-1)/(2n-1), i0 = floor(src), i1 = min(i0+1, n-1), weights (1 - frac, frac)
// -- the arithmetic of upsample_kernel / ATen's area_pixel_compute_source_index in fp32.
__device__ __forceinline__ float upsample_weight(int o, int i, int n) {
  if (o < 0 || o >= 2 * n) return 0.f;
  const float scale = (2 * n > 1) ? (float)(n - 1) / (float)(2 * n - 1) : 0.f;
  const float src = scale * (float)o;
  const int i0 = (int)src;
  const int i1 = i0 + (i0 < n - 1 ? 1 : 0);
  const float w1 = src - (float)i0;
  float w = 0.f;
  if (i0 == i) w += 1.f - w1;
  if (i1 == i) w += w1;
  return w;
}

// One thread: one INPUT voxel x 8 channels; gathers from the <= 4 outputs per axis that read it
// (candidates 2i-2 .. 2i+3 are evaluated with the forward formula, so rounding at the interval
// ends cannot lose or double a contribution).
template <typename T>
__global__ void __launch_bounds__(256)
upsample_bwd_kernel(const T* __restrict__ go, int go_cstride, int go_coff, bool go_enc,
                    T* __restrict__ gi, int gi_cstride, int gi_coff, int B, int Di, int Hi, int Wi,
                    int C) {
  const int cv = C / 8;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * Di * Hi * Wi * cv;
  if (i >= total) return;
  const int c8 = (int)(i % cv);
  size_t v = i / cv;
  const size_t iv = v;
  const int x = (int)(v % Wi); v /= Wi;
  const int y = (int)(v % Hi); v /= Hi;
  const int z = (int)(v % Di);
  const int b = (int)(v / Di);
  const int Do = 2 * Di, Ho = 2 * Hi, Wo = 2 * Wi;
  float wz[6], wy[6], wx[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    wz[k] = upsample_weight(2 * z - 2 + k, z, Di);
    wy[k] = upsample_weight(2 * y - 2 + k, y, Hi);
    wx[k] = upsample_weight(2 * x - 2 + k, x, Wi);
  }
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int kz = 0; kz < 6; ++kz) {
    if (wz[kz] == 0.f) continue;
    const int oz = 2 * z - 2 + kz;
    for (int ky = 0; ky < 6; ++ky) {
      if (wy[ky] == 0.f) continue;
      const int oy = 2 * y - 2 + ky;
      const float wzy = wz[kz] * wy[ky];
      for (int kx = 0; kx < 6; ++kx) {
        if (wx[kx] == 0.f) continue;
        const int ox = 2 * x - 2 + kx;
        const size_t ov = (((size_t)b * Do + oz) * Ho + oy) * Wo + ox;
        float f[8];
        load8_dec<T>(go + ov * go_cstride + go_coff + 8 * c8, go_enc, f);
        const float w = wzy * wx[kx];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(w, f[j], acc[j]);
      }
    }
  }
  store8<T>(gi + iv * gi_cstride + gi_coff + 8 * c8, acc);
}

Status launch_upsample_bwd(const TView& grad_out, const Act& grad_in, cudaStream_t s) {
  const Act& o = grad_out.a;
  EXA_CHECK(vec_ok(o) && vec_ok(grad_in) && o.C == grad_in.C && o.B == grad_in.B &&
                o.D == 2 * grad_in.D && o.H == 2 * grad_in.H && o.W == 2 * grad_in.W &&
                o.fp32 == grad_in.fp32,
            "upsample_bwd: shape mismatch");
  const size_t total = grad_in.voxels() * (grad_in.C / 8);
  const unsigned blocks = (unsigned)ceil_div64((int64_t)total, 256);
  if (o.fp32)
    upsample_bwd_kernel<float><<<blocks, 256, 0, s>>>(
        (const float*)o.ptr, o.cstride, o.coff, grad_out.enc, (float*)grad_in.ptr, grad_in.cstride,
        grad_in.coff, grad_in.B, grad_in.D, grad_in.H, grad_in.W, grad_in.C);
  else
    upsample_bwd_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(
        (const __nv_bfloat16*)o.ptr, o.cstride, o.coff, grad_out.enc, (__nv_bfloat16*)grad_in.ptr,
        grad_in.cstride, grad_in.coff, grad_in.B, grad_in.D, grad_in.H, grad_in.W, grad_in.C);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

template <typename T>
__global__ void __launch_bounds__(256)
head_bwd_dx_kernel(const float* __restrict__ dl, const float* __restrict__ hw, int C, int cin,
                   T* __restrict__ du, int du_cstride, int du_coff, int B, size_t vb) {
  const int cv = cin / 8;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * vb * cv) return;
  const int c8 = (int)(i % cv);
  const size_t v = i / cv;
  const size_t b = v / vb, sp = v % vb;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int k = 0; k < C; ++k) {
    const float d = __ldg(dl + (b * C + k) * vb + sp);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(d, __ldg(hw + k * cin + 8 * c8 + j), acc[j]);
  }
  store8<T>(du + v * du_cstride + du_coff + 8 * c8, acc);
}

Status launch_head_bwd_dx(const float* dlogits, const float* hw, int C, const Act& du,
                          cudaStream_t s) {
  EXA_CHECK(vec_ok(du) && C >= 1 && C <= 8, "head_bwd_dx: bad arguments");
  const size_t vb = (size_t)du.D * du.H * du.W;
  const unsigned blocks = (unsigned)ceil_div64((int64_t)(du.voxels() * (du.C / 8)), 256);
  if (du.fp32)
    head_bwd_dx_kernel<float><<<blocks, 256, 0, s>>>(dlogits, hw, C, du.C, (float*)du.ptr,
                                                      du.cstride, du.coff, du.B, vb);
  else
    head_bwd_dx_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(
        dlogits, hw, C, du.C, (__nv_bfloat16*)du.ptr, du.cstride, du.coff, du.B, vb);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

// dW[k][c] = sum_v dl[b][k][sp] u[v][c], db[k] = sum_v dl[b][k][sp]; thread = (lane, 8 channels);
// CT = number of head outputs (compile time: 1 foreground, 3 affinities, 8 generic upper bound)
template <typename T, int CT>
__global__ void __launch_bounds__(RED_THREADS)
head_bwd_dw_kernel(const float* __restrict__ dl, const T* __restrict__ u, int u_cstride,
                   int u_coff, int C, int cin, int B, size_t vb, double* __restrict__ sums) {
  __shared__ float sh[RED_THREADS * 8];
  const int cv = cin / 8, lanes = RED_THREADS / cv;
  const int c8 = threadIdx.x % cv, lane = threadIdx.x / cv;
  const size_t voxels = (size_t)B * vb;
  float acc[CT][8];
  float db[CT];
#pragma unroll
  for (int k = 0; k < CT; ++k) {
    db[k] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;
  }
  const size_t v0 = (size_t)blockIdx.x * lanes * RED_ITER;
  for (int it = 0; it < RED_ITER; it += BWD_UNROLL) {
    Raw8<T> ru[BWD_UNROLL];
    float d[BWD_UNROLL][CT];
#pragma unroll
    for (int q = 0; q < BWD_UNROLL; ++q) {
      const size_t v = v0 + (size_t)(it + q) * lanes + lane;
      if (v < voxels) {
        const size_t b = v / vb, sp = v % vb;
        ru[q].load(u + v * u_cstride + u_coff + 8 * c8);
#pragma unroll
        for (int k = 0; k < CT; ++k) d[q][k] = k < C ? __ldg(dl + (b * C + k) * vb + sp) : 0.f;
      }
    }
#pragma unroll
    for (int q = 0; q < BWD_UNROLL; ++q) {
      const size_t v = v0 + (size_t)(it + q) * lanes + lane;
      if (v < voxels) {
        float f[8];
        ru[q].get(f);
#pragma unroll
        for (int k = 0; k < CT; ++k) {
          db[k] += d[q][k];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[k][j] = fmaf(d[q][k], f[j], acc[k][j]);
        }
      }
    }
  }
  // per output channel k: reduce the lanes through shared memory, one double atomic per entry
#pragma unroll
  for (int k = 0; k < CT; ++k) {
    if (k < C) {
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 8; ++j) sh[threadIdx.x * 8 + j] = acc[k][j];
      __syncthreads();
      for (int ch = threadIdx.x; ch < cin; ch += RED_THREADS) {
        const int cc = ch / 8, j = ch % 8;
        double a0 = 0.0;
        for (int l = 0; l < lanes; ++l) a0 += (double)sh[(l * cv + cc) * 8 + j];
        atomicAdd(sums + (size_t)k * cin + ch, a0);
      }
    }
  }
  __syncthreads();
  if (c8 == 0) {
#pragma unroll
    for (int k = 0; k < CT; ++k) sh[lane * 8 + k] = db[k];
  }
  __syncthreads();
  if (threadIdx.x < C) {
    double a0 = 0.0;
    for (int l = 0; l < lanes; ++l) a0 += (double)sh[l * 8 + threadIdx.x];
    atomicAdd(sums + (size_t)C * cin + threadIdx.x, a0);
  }
}

template <typename T>
static void head_bwd_dw_dispatch(const float* dlogits, const Act& u, int C, double* sums,
                                 unsigned blocks, size_t vb, cudaStream_t s) {
  const T* up = (const T*)u.ptr;
  if (C == 1)
    head_bwd_dw_kernel<T, 1><<<blocks, RED_THREADS, 0, s>>>(dlogits, up, u.cstride, u.coff, C, u.C,
                                                             u.B, vb, sums);
  else if (C <= 3)
    head_bwd_dw_kernel<T, 3><<<blocks, RED_THREADS, 0, s>>>(dlogits, up, u.cstride, u.coff, C, u.C,
                                                             u.B, vb, sums);
  else
    head_bwd_dw_kernel<T, 8><<<blocks, RED_THREADS, 0, s>>>(dlogits, up, u.cstride, u.coff, C, u.C,
                                                             u.B, vb, sums);
}

Status launch_head_bwd_dw(const float* dlogits, const Act& u, int C, double* sums, cudaStream_t s) {
  EXA_TRY(check_reduce_shape(u, "head_bwd_dw"));
  EXA_CHECK(C >= 1 && C <= 8, "head_bwd_dw: at most 8 output channels");
  const size_t vb = (size_t)u.D * u.H * u.W;
  const int lanes = RED_THREADS / (u.C / 8);
  const unsigned blocks = (unsigned)ceil_div64((int64_t)u.voxels(), (int64_t)lanes * RED_ITER);
  if (u.fp32)
    head_bwd_dw_dispatch<float>(dlogits, u, C, sums, blocks, vb, s);
  else
    head_bwd_dw_dispatch<__nv_bfloat16>(dlogits, u, C, sums, blocks, vb, s);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

// ---------------------------------------------------------------------------
// T4: conv weight gradient
// ---------------------------------------------------------------------------
// bf16: one CTA owns a 32 (Cout) x 32 (Cin) block of dW for ALL 27 taps and walks voxel tiles of
// one z plane, 8 (y) x 16 (x) = 128 voxels = the K extent of 8 m16n8k16 steps.  Per tile the dz
// rows [128][32] and the input halo [3][10][18][32] are staged in shared memory (80-byte voxel
// rows: ldmatrix reads are conflict-free); both operands are K-major in memory, so both fragments
// come from ldmatrix.trans.  Warp w: Cin columns 8 (w & 3) .. +8, taps 14 (w >> 2) .. +14, both
// 16-row halves of Cout: 14 x 2 x 4 = 112 fp32 accumulators per thread, written once at the end.
namespace {
constexpr int WG_TY = 8, WG_TX = 16;
constexpr int WG_VS = 40;  // bf16 elements per staged voxel row (32 channels + 8 pad = 80 B)
constexpr int WG_HALO = 3 * (WG_TY + 2) * (WG_TX + 2);  // 540 voxels
constexpr int WG_SMEM = (WG_HALO + WG_TY * WG_TX) * WG_VS * 2;
constexpr int WG_TAPS_PER_GROUP = 14;

struct WgArgs {
  const __nv_bfloat16* x;
  int x_cstride, x_coff;
  const __nv_bfloat16* dz;  // dense [voxel][cout]
  float* partial;           // [splits][cout][cin][27]
  int B, D, H, W, cin, cout;
  int nty, ntx, tiles_total;
};

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t addr, uint32_t (&r)[2]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
               : "=r"(r[0]), "=r"(r[1])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4],
                                               const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
      "{%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
}  // namespace

__global__ void __launch_bounds__(256, 1)
wgrad_mma_kernel(const WgArgs a) {
  extern __shared__ __align__(16) unsigned char wg_smem[];
  __nv_bfloat16* Xs = reinterpret_cast<__nv_bfloat16*>(wg_smem);  // [540][40]
  __nv_bfloat16* Ds = Xs + WG_HALO * WG_VS;                       // [128][40]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cib = blockIdx.y, cob = blockIdx.z;
  const int nt = warp & 3, tg = warp >> 2;
  const int tap0 = tg * WG_TAPS_PER_GROUP;

  float acc[WG_TAPS_PER_GROUP][2][4];
#pragma unroll
  for (int j = 0; j < WG_TAPS_PER_GROUP; ++j)
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[j][m][e] = 0.f;

  // ldmatrix row addresses of this lane
  //   A (dz^T, m16 x k16): matrix j = lane / 8: voxel row +8 for j >= 2, Cout column +8 for odd j
  const int a_row = ((lane >> 4) & 1) * 8 + (lane & 7);
  const int a_col = ((lane >> 3) & 1) * 8;
  //   B (x, k16 x n8): matrix j = (lane / 8) & 1: voxel row +8; lanes 16..31 repeat valid addresses
  const int b_row = ((lane >> 3) & 1) * 8 + (lane & 7);
  const uint32_t xs_base = smem_u32(Xs), ds_base = smem_u32(Ds);

  for (int tile = blockIdx.x; tile < a.tiles_total; tile += gridDim.x) {
    int r = tile;
    const int tx = r % a.ntx; r /= a.ntx;
    const int ty = r % a.nty; r /= a.nty;
    const int z = r % a.D;
    const int b = r / a.D;
    const int y0 = ty * WG_TY, x0 = tx * WG_TX;
    __syncthreads();  // the previous tile's fragments have been read
    // input halo: voxels (z-1.., y0-1.., x0-1..), 32 channels = four 16-byte chunks each
    for (int i = tid; i < WG_HALO * 4; i += 256) {
      const int chunk = i & 3, hv = i >> 2;
      const int hx = hv % (WG_TX + 2);
      const int hy = (hv / (WG_TX + 2)) % (WG_TY + 2);
      const int hz = hv / ((WG_TX + 2) * (WG_TY + 2));
      const int gz = z - 1 + hz, gy = y0 - 1 + hy, gx = x0 - 1 + hx;
      uint4 q = make_uint4(0u, 0u, 0u, 0u);
      if (gz >= 0 && gz < a.D && gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) {
        const size_t gv = (((size_t)b * a.D + gz) * a.H + gy) * a.W + gx;
        q = *reinterpret_cast<const uint4*>(a.x + gv * a.x_cstride + a.x_coff + cib * 32 + chunk * 8);
      }
      *reinterpret_cast<uint4*>(Xs + hv * WG_VS + chunk * 8) = q;
    }
    // dz rows of the tile (zero outside the volume: they contribute nothing)
    for (int i = tid; i < WG_TY * WG_TX * 4; i += 256) {
      const int chunk = i & 3, tv = i >> 2;
      const int gy = y0 + tv / WG_TX, gx = x0 + tv % WG_TX;
      uint4 q = make_uint4(0u, 0u, 0u, 0u);
      if (gy < a.H && gx < a.W) {
        const size_t gv = (((size_t)b * a.D + z) * a.H + gy) * a.W + gx;
        q = *reinterpret_cast<const uint4*>(a.dz + gv * a.cout + cob * 32 + chunk * 8);
      }
      *reinterpret_cast<uint4*>(Ds + tv * WG_VS + chunk * 8) = q;
    }
    __syncthreads();

#pragma unroll 1
    for (int ks = 0; ks < WG_TY; ++ks) {  // k16 step = tile row ks, x = 0..15
      uint32_t af[2][4];
#pragma unroll
      for (int m = 0; m < 2; ++m)
        ldmatrix_x4_trans(ds_base + (uint32_t)(((ks * WG_TX + a_row) * WG_VS + m * 16 + a_col) * 2),
                          af[m]);
#pragma unroll
      for (int j = 0; j < WG_TAPS_PER_GROUP; ++j) {
        const int tap = tap0 + j;
        if (tap < 27) {
          const int kz = tap / 9, ky = (tap / 3) % 3, kx = tap % 3;
          const int hv = (kz * (WG_TY + 2) + ks + ky) * (WG_TX + 2) + kx + b_row;
          uint32_t bf[2];
          ldmatrix_x2_trans(xs_base + (uint32_t)((hv * WG_VS + nt * 8) * 2), bf);
          mma_bf16_16816(acc[j][0], af[0], bf);
          mma_bf16_16816(acc[j][1], af[1], bf);
        }
      }
    }
  }

  // accumulator (row g / g+8, cols 2t, 2t+1) -> partial[split][co][ci][tap]
  const int g = lane >> 2, t = lane & 3;
  float* dst = a.partial + (size_t)blockIdx.x * a.cout * a.cin * 27;
#pragma unroll
  for (int j = 0; j < WG_TAPS_PER_GROUP; ++j) {
    const int tap = tap0 + j;
    if (tap < 27) {
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int co = cob * 32 + m * 16 + g + (e >> 1) * 8;
          const int ci = cib * 32 + nt * 8 + 2 * t + (e & 1);
          dst[((size_t)co * a.cin + ci) * 27 + tap] = acc[j][m][e];
        }
    }
  }
}

// fp32 validation mode: one block owns a 32 (Cout) x 16 (Cin) block for all 27 taps and walks
// runs of 128 voxels of the flattened index; thread = (ci, 2 co), 54 accumulators.
__global__ void __launch_bounds__(256)
wgrad_fp32_kernel(const float* __restrict__ x, int x_cstride, int x_coff,
                  const float* __restrict__ dz, float* __restrict__ partial, int B, int D, int H,
                  int W, int cin, int cout, int chunks_total) {
  __shared__ float As[16][128 + 4];
  __shared__ __align__(16) float Dsh[128][32];
  const int t = threadIdx.x;
  const int cib = blockIdx.y, cob = blockIdx.z;
  const size_t nvox = (size_t)B * D * H * W;
  const int ci = t >> 4, co2 = (t & 15) * 2;
  float acc[27][2];
#pragma unroll
  for (int k = 0; k < 27; ++k) acc[k][0] = acc[k][1] = 0.f;
  const int lv = t >> 1, lh = (t & 1) * 8;  // loader role: voxel, channel half
  for (int chunk = blockIdx.x; chunk < chunks_total; chunk += gridDim.x) {
    const size_t v0 = (size_t)chunk * 128;
    const size_t lvox = v0 + lv;
    const bool lvalid = lvox < nvox;
    int lx = 0, ly = 0, lz = 0, lb = 0;
    if (lvalid) {
      size_t r = lvox;
      lx = (int)(r % W); r /= W;
      ly = (int)(r % H); r /= H;
      lz = (int)(r % D);
      lb = (int)(r / D);
    }
    __syncthreads();
    // dz rows: 128 voxels x 32 channels, 16 floats per thread
    {
      const int dv = t >> 1, dh = (t & 1) * 16;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 d4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (v0 + dv < nvox)
          d4 = *reinterpret_cast<const float4*>(dz + (v0 + dv) * cout + cob * 32 + dh + q * 4);
        *reinterpret_cast<float4*>(&Dsh[dv][dh + q * 4]) = d4;
      }
    }
#pragma unroll
    for (int tap = 0; tap < 27; ++tap) {
      const int kz = tap / 9 - 1, ky = (tap / 3) % 3 - 1, kx = tap % 3 - 1;
      const int sx = lx + kx, sy = ly + ky, sz = lz + kz;
      const bool inb = lvalid && sx >= 0 && sx < W && sy >= 0 && sy < H && sz >= 0 && sz < D;
      float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0;
      if (inb) {
        const float* src =
            x + ((((size_t)lb * D + sz) * H + sy) * W + sx) * x_cstride + x_coff + cib * 16 + lh;
        q0 = *reinterpret_cast<const float4*>(src);
        q1 = *reinterpret_cast<const float4*>(src + 4);
      }
      __syncthreads();
      As[lh + 0][lv] = q0.x; As[lh + 1][lv] = q0.y; As[lh + 2][lv] = q0.z; As[lh + 3][lv] = q0.w;
      As[lh + 4][lv] = q1.x; As[lh + 5][lv] = q1.y; As[lh + 6][lv] = q1.z; As[lh + 7][lv] = q1.w;
      __syncthreads();
      float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
      for (int v = 0; v < 128; ++v) {
        const float av = As[ci][v];
        const float2 dv = *reinterpret_cast<const float2*>(&Dsh[v][co2]);
        s0 = fmaf(av, dv.x, s0);
        s1 = fmaf(av, dv.y, s1);
      }
      acc[tap][0] += s0;
      acc[tap][1] += s1;
    }
  }
  float* dst = partial + (size_t)blockIdx.x * cout * cin * 27;
#pragma unroll
  for (int tap = 0; tap < 27; ++tap) {
    const int c = cib * 16 + ci;
    dst[((size_t)(cob * 32 + co2) * cin + c) * 27 + tap] = acc[tap][0];
    dst[((size_t)(cob * 32 + co2 + 1) * cin + c) * 27 + tap] = acc[tap][1];
  }
}

static int wgrad_units(const Act& x) {  // tiles (bf16) or 128-voxel chunks (fp32)
  if (x.fp32) return (int)ceil_div64((int64_t)x.voxels(), 128);
  return x.B * x.D * ceil_div(x.H, WG_TY) * ceil_div(x.W, WG_TX);
}

int wgrad_splits(const Act& x, int cout, int num_sms) {
  if (!x.fp32 && wgrad_tc_enabled()) return wgrad_tc_splits(x, cout, num_sms);
  const int blocks = (x.fp32 ? x.C / 16 : x.C / 32) * (cout / 32);
  int s = (2 * num_sms + blocks - 1) / blocks;
  const int units = wgrad_units(x);
  if (s > units) s = units;
  return s < 1 ? 1 : s;
}

size_t wgrad_partial_elems(const Act& x, int cout, int num_sms) {
  return (size_t)wgrad_splits(x, cout, num_sms) * cout * x.C * 27;
}

Status launch_wgrad(const Act& x, const Act& dz, float* partial, int num_sms, cudaStream_t s) {
  EXA_CHECK(same_grid(x, dz) && x.fp32 == dz.fp32 && dz.cstride == dz.C && dz.coff == 0 &&
                dz.C % 32 == 0,
            "wgrad: dz must be dense with a multiple of 32 channels on the input's grid");
  const int splits = wgrad_splits(x, dz.C, num_sms);
  if (x.fp32) {
    EXA_CHECK(x.C % 16 == 0 && x.cstride % 4 == 0 && x.coff % 4 == 0, "wgrad: Cin % 16 (fp32)");
    dim3 grid((unsigned)splits, (unsigned)(x.C / 16), (unsigned)(dz.C / 32));
    wgrad_fp32_kernel<<<grid, 256, 0, s>>>((const float*)x.ptr, x.cstride, x.coff,
                                           (const float*)dz.ptr, partial, x.B, x.D, x.H, x.W, x.C,
                                           dz.C, wgrad_units(x));
    EXA_CUDA(cudaGetLastError());
    return Status::OK();
  }
  EXA_CHECK(x.C % 32 == 0 && x.cstride % 8 == 0 && x.coff % 8 == 0, "wgrad: Cin % 32 (bf16)");
  if (wgrad_tc_enabled()) return launch_wgrad_tc(x, dz, partial, num_sms, s);
  static bool configured = false;
  if (!configured) {
    EXA_CUDA(cudaFuncSetAttribute(wgrad_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  WG_SMEM));
    configured = true;
  }
  WgArgs a{};
  a.x = (const __nv_bfloat16*)x.ptr; a.x_cstride = x.cstride; a.x_coff = x.coff;
  a.dz = (const __nv_bfloat16*)dz.ptr;
  a.partial = partial;
  a.B = x.B; a.D = x.D; a.H = x.H; a.W = x.W; a.cin = x.C; a.cout = dz.C;
  a.nty = ceil_div(x.H, WG_TY); a.ntx = ceil_div(x.W, WG_TX);
  a.tiles_total = wgrad_units(x);
  dim3 grid((unsigned)splits, (unsigned)(x.C / 32), (unsigned)(dz.C / 32));
  wgrad_mma_kernel<<<grid, 256, WG_SMEM, s>>>(a);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

// Stem (Cin = 1): thread = (co of 32, kz): nine accumulators (ky, kx); a block walks rows
// (b, z, y) of W voxels staged in shared memory (dz row as floats, the 3x3 input rows around it).
// Four voxels per iteration: 4 dz loads + 3 x (float4 + float2) input loads feed 36 FMAs.
template <typename T>
__global__ void __launch_bounds__(96)
wgrad_stem_kernel(const float* __restrict__ x, const T* __restrict__ dz, float* __restrict__ partial,
                  int B, int D, int H, int W, int cout, int rows_total) {
  extern __shared__ __align__(16) float st_smem[];
  const int Wp = (W + 3) & ~3, XS = Wp + 4;
  float* Dsh = st_smem;            // [Wp][32], zero beyond W
  float* Xsh = st_smem + Wp * 32;  // [9][XS]: column c <-> x = c - 1, zero outside the volume
  const int t = threadIdx.x, co = t & 31, kz = t >> 5;
  const int cob = blockIdx.y;
  float acc[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) acc[i][j] = 0.f;
  for (int row = blockIdx.x; row < rows_total; row += gridDim.x) {
    const int y = row % H, z = (row / H) % D, b = row / (H * D);
    __syncthreads();
    for (int i = t; i < Wp * 4; i += 96) {  // 8 channels per item
      const int xv = i >> 2, c8 = i & 3;
      float f[8];
      if (xv < W) {
        const size_t gv = (((size_t)b * D + z) * H + y) * W + xv;
        load8<T>(dz + gv * cout + cob * 32 + c8 * 8, f);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) Dsh[xv * 32 + c8 * 8 + j] = f[j];
    }
    for (int i = t; i < 9 * XS; i += 96) {
      const int xx = i % XS - 1, r9 = i / XS;
      const int gz = z + r9 / 3 - 1, gy = y + r9 % 3 - 1;
      float v = 0.f;
      if (gz >= 0 && gz < D && gy >= 0 && gy < H && xx >= 0 && xx < W)
        v = x[(((size_t)b * D + gz) * H + gy) * W + xx];
      Xsh[i] = v;
    }
    __syncthreads();
    for (int xv = 0; xv < Wp; xv += 4) {
      float d[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) d[i] = Dsh[(xv + i) * 32 + co];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const float* rowp = Xsh + (kz * 3 + ky) * XS + xv;
        const float4 p = *reinterpret_cast<const float4*>(rowp);
        const float2 q = *reinterpret_cast<const float2*>(rowp + 4);
        const float v[6] = {p.x, p.y, p.z, p.w, q.x, q.y};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) acc[ky][kx] = fmaf(d[i], v[i + kx], acc[ky][kx]);
      }
    }
  }
  float* dst = partial + (size_t)blockIdx.x * cout * 27 + (size_t)(cob * 32 + co) * 27 + kz * 9;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) dst[ky * 3 + kx] = acc[ky][kx];
}

int wgrad_stem_splits(const Act& dz, int num_sms) {
  const int rows = dz.B * dz.D * dz.H;
  int s = 12 * num_sms / (dz.C / 32);  // 96-thread blocks, ~16 KB of shared memory each
  if (s > rows) s = rows;
  return s < 1 ? 1 : s;
}

Status launch_wgrad_stem(const float* x, const Act& dz, float* partial, int num_sms,
                         cudaStream_t s) {
  EXA_CHECK(dz.cstride == dz.C && dz.coff == 0 && dz.C % 32 == 0 && dz.W <= 1024,
            "wgrad_stem: dz must be dense with a multiple of 32 channels");
  const int rows = dz.B * dz.D * dz.H;
  const int splits = wgrad_stem_splits(dz, num_sms);
  const int wp = (dz.W + 3) & ~3;
  const size_t smem = (size_t)(wp * 32 + 9 * (wp + 4)) * 4;
  dim3 grid((unsigned)splits, (unsigned)(dz.C / 32));
  if (dz.fp32) {
    if (smem > 48 * 1024)
      EXA_CUDA(cudaFuncSetAttribute(wgrad_stem_kernel<float>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wgrad_stem_kernel<float><<<grid, 96, smem, s>>>(x, (const float*)dz.ptr, partial, dz.B, dz.D,
                                                      dz.H, dz.W, dz.C, rows);
  } else {
    if (smem > 48 * 1024)
      EXA_CUDA(cudaFuncSetAttribute(wgrad_stem_kernel<__nv_bfloat16>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wgrad_stem_kernel<__nv_bfloat16><<<grid, 96, smem, s>>>(x, (const __nv_bfloat16*)dz.ptr,
                                                              partial, dz.B, dz.D, dz.H, dz.W, dz.C,
                                                              rows);
  }
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, int splits, size_t elems,
                    float* __restrict__ dw) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= elems) return;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += partial[(size_t)k * elems + i];  // fixed order
  dw[i] = s;
}

Status launch_wgrad_reduce(const float* partial, int splits, size_t elems, float* dw,
                           cudaStream_t s) {
  wgrad_reduce_kernel<<<(unsigned)ceil_div64((int64_t)elems, 256), 256, 0, s>>>(partial, splits,
                                                                                 elems, dw);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

// ---------------------------------------------------------------------------
// T5: BCEWithLogitsLoss (mean) and its gradient
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bce_with_logits_kernel(const float* __restrict__ x, const float* __restrict__ y, size_t n,
                       float gscale, double* __restrict__ loss_sum, float* __restrict__ grad) {
  __shared__ float sh[256];
  float local = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float xv = x[i], yv = y[i];
    // max(x, 0) - x y + log(1 + exp(-|x|)): the stable form ATen uses
    local += fmaxf(xv, 0.f) - xv * yv + log1pf(expf(-fabsf(xv)));
    if (grad) grad[i] = gscale * (1.f / (1.f + expf(-xv)) - yv);
  }
  sh[threadIdx.x] = local;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(loss_sum, (double)sh[0]);
}

Status launch_bce_with_logits(const float* logits, const float* target, size_t n, float grad_scale,
                              double* loss_sum, float* grad, cudaStream_t s) {
  EXA_CHECK(logits && target && loss_sum && n > 0, "bce_with_logits: bad arguments");
  // at most 64 elements per thread: short fp32 runs, then doubles
  const unsigned blocks = (unsigned)ceil_div64((int64_t)n, 256 * 64);
  bce_with_logits_kernel<<<blocks, 256, 0, s>>>(logits, target, n, grad_scale / (float)n, loss_sum,
                                                grad);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

}  // namespace exa
