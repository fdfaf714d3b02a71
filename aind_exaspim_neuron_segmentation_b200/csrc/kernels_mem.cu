// Memory-bound kernels of the affinity-prediction path (all HBM-bound byte/elementwise
// work: coalesced, 16-byte vectorised, no tensor cores) plus the fp32 validation conv.
//
//   K0   histogram of min(v, clip)                    reference inference.py:79, img_util.py:526
//   K0b  patch gather + LUT normalise + reflect pad,
//        fused with the Cin=1 stem conv               inference.py:188-191, img_util.py:378-379,
//                                                     424-428, 527-531; unet3d.py:143-145 (inc.0)
//   K3   2x2x2 max-pool                               unet3d.py:195
//   K4   trilinear x2 upsample (align_corners=True)
//        written into the concat slot                 unet3d.py:248-250, 288
//   K5'  1x1x1 head + sigmoid + trim (fp32 mode)      unet3d.py:318, inference.py:158-162
//   K6   overlap stitch + coverage normalisation      inference.py:99-125
//   Kf   fp32 SIMT 3x3x3 conv (validation mode)       unet3d.py:143-148
#include "kernels.h"

namespace exa {

// ---------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------
template <typename T>
struct Vec8;  // 8 channels
template <>
struct Vec8<__nv_bfloat16> {
  uint4 raw;
  __device__ void load(const __nv_bfloat16* p) { raw = *reinterpret_cast<const uint4*>(p); }
  __device__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = raw; }
  __device__ void to_float(float (&f)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  __device__ void from_float(const float (&f)[8]) {
    raw.x = pack_bf16x2(f[0], f[1]);
    raw.y = pack_bf16x2(f[2], f[3]);
    raw.z = pack_bf16x2(f[4], f[5]);
    raw.w = pack_bf16x2(f[6], f[7]);
  }
};
template <>
struct Vec8<float> {
  float4 a, b;
  __device__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = a;
    *reinterpret_cast<float4*>(p + 4) = b;
  }
  __device__ void to_float(float (&f)[8]) const {
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  __device__ void from_float(const float (&f)[8]) {
    a = make_float4(f[0], f[1], f[2], f[3]);
    b = make_float4(f[4], f[5], f[6], f[7]);
  }
};

// np.pad(mode="reflect") index for position p >= 0 in a box of length L (edge not repeated).
__device__ __forceinline__ int reflect_index(int p, int L) {
  if (L <= 1) return 0;
  const int period = 2 * (L - 1);
  p %= period;
  return p < L ? p : period - p;
}

// ---------------------------------------------------------------------------
// K0: histogram
// ---------------------------------------------------------------------------
constexpr int HIST_SMEM_BINS = 4096;

__global__ void __launch_bounds__(512)
histogram_kernel(const uint16_t* __restrict__ vol, size_t n, int clip,
                 unsigned long long* __restrict__ hist) {
  extern __shared__ unsigned int sh[];
  const int bins = clip + 1;
  const bool use_smem = bins <= HIST_SMEM_BINS;
  if (use_smem) {
    for (int i = threadIdx.x; i < bins; i += blockDim.x) sh[i] = 0;
    __syncthreads();
  }
  // scalar head up to the first 16-byte boundary, 16-byte vector body, scalar tail
  size_t head = ((16 - ((uintptr_t)vol & 15)) & 15) / 2;
  if (head > n) head = n;
  const size_t n8 = (n - head) / 8;
  const uint4* v8 = reinterpret_cast<const uint4*>(vol + head);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const uint4 q = __ldg(v8 + i);
    const unsigned int w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int lo = min((int)(w[k] & 0xFFFFu), clip);
      const int hi = min((int)(w[k] >> 16), clip);
      if (use_smem) {
        atomicAdd(&sh[lo], 1u);
        atomicAdd(&sh[hi], 1u);
      } else {
        atomicAdd(&hist[lo], 1ull);
        atomicAdd(&hist[hi], 1ull);
      }
    }
  }
  if (blockIdx.x == 0) {
    const size_t tail0 = head + n8 * 8;
    for (size_t j = threadIdx.x; j < head + (n - tail0); j += blockDim.x) {
      const size_t i = j < head ? j : tail0 + (j - head);
      const int v = min((int)vol[i], clip);
      if (use_smem) atomicAdd(&sh[v], 1u);
      else atomicAdd(&hist[v], 1ull);
    }
  }
  if (use_smem) {
    __syncthreads();
    for (int i = threadIdx.x; i < bins; i += blockDim.x) {
      const unsigned int c = sh[i];
      if (c) atomicAdd(&hist[i], (unsigned long long)c);
    }
  }
}

Status launch_histogram(const uint16_t* vol, size_t n, int clip, unsigned long long* hist,
                        cudaStream_t s) {
  EXA_CHECK(clip >= 0 && clip <= 65535, "brightness_clip must be in [0, 65535]");
  EXA_CHECK(((uintptr_t)vol & 1) == 0, "volume pointer must be 2-byte aligned");
  EXA_CUDA(cudaMemsetAsync(hist, 0, sizeof(unsigned long long) * (clip + 1), s));
  if (n == 0) return Status::OK();
  const int bins = clip + 1;
  const size_t smem = bins <= HIST_SMEM_BINS ? sizeof(unsigned int) * bins : 0;
  int blocks = (int)std::min<size_t>((n / 8 + 511) / 512 + 1, 148 * 4);
  histogram_kernel<<<blocks, 512, smem, s>>>(vol, n, clip, hist);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

// ---------------------------------------------------------------------------
// K0b + stem conv (Cin = 1 -> 32), SIMT fp32
// ---------------------------------------------------------------------------
struct StemArgs {
  PatchSource src;
  int B, Pz, Py, Px;
  void* out;
  int cstride, coff;  // the 32 output channels sit at offset coff of voxels cstride elements apart
};

constexpr int ST_TX = 16, ST_TY = 8, ST_TZ = 8;

template <typename T, bool FROM_VOLUME>
__global__ void __launch_bounds__(128)
stem_kernel(const __grid_constant__ StemWeights wt, const StemArgs a) {
  __shared__ float tile[ST_TZ + 2][ST_TY + 2][ST_TX + 2];
  const int ntx = a.Px / ST_TX;
  const int x0 = (blockIdx.x % ntx) * ST_TX;
  const int y0 = (blockIdx.x / ntx) * ST_TY;
  const int z0 = blockIdx.y * ST_TZ;
  const int b = blockIdx.z;

  int sz = 0, sy = 0, sx = 0, Lz = 0, Ly = 0, Lx = 0;
  if (FROM_VOLUME) {
    sz = a.src.starts[3 * b + 0];
    sy = a.src.starts[3 * b + 1];
    sx = a.src.starts[3 * b + 2];
    Lz = min(a.Pz, a.src.gD - sz);  // clipped box lengths (img_util.py:424-428)
    Ly = min(a.Py, a.src.gH - sy);
    Lx = min(a.Px, a.src.gW - sx);
  }
  constexpr int TILE_N = (ST_TZ + 2) * (ST_TY + 2) * (ST_TX + 2);
  for (int i = threadIdx.x; i < TILE_N; i += blockDim.x) {
    const int lx = i % (ST_TX + 2);
    const int ly = (i / (ST_TX + 2)) % (ST_TY + 2);
    const int lz = i / ((ST_TX + 2) * (ST_TY + 2));
    const int px = x0 + lx - 1, py = y0 + ly - 1, pz = z0 + lz - 1;
    float v = 0.f;  // conv zero padding outside the patch (unet3d.py:143, padding=1)
    if (px >= 0 && px < a.Px && py >= 0 && py < a.Py && pz >= 0 && pz < a.Pz) {
      if (FROM_VOLUME) {
        const int gz = sz + (pz < Lz ? pz : reflect_index(pz, Lz));
        const int gy = sy + (py < Ly ? py : reflect_index(py, Ly));
        const int gx = sx + (px < Lx ? px : reflect_index(px, Lx));
        const size_t idx = ((size_t)(gz - a.src.vz0) * a.src.gH + gy) * a.src.gW + gx;
        const int raw = min((int)__ldg(a.src.vol + idx), a.src.clip);
        v = __ldg(a.src.lut + raw);
      } else {
        v = __ldg(a.src.x + (((size_t)b * a.Pz + pz) * a.Py + py) * a.Px + px);
      }
    }
    tile[lz][ly][lx] = v;
  }
  __syncthreads();

  const int tx = threadIdx.x % ST_TX, ty = threadIdx.x / ST_TX;
  T* outp = reinterpret_cast<T*>(a.out);
  // Two output planes per pass, channel pairs in packed fp32 (FFMA2): the weight pairs come
  // from the constant bank through uniform registers and are shared by both planes.
  const f32x2* wpair = reinterpret_cast<const f32x2*>(&wt.w[0][0]);
  const f32x2* bpair = reinterpret_cast<const f32x2*>(&wt.b[0]);
#pragma unroll 1
  for (int tz = 0; tz < ST_TZ; tz += 2) {
    f32x2 acc0[16], acc1[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc0[c] = acc1[c] = bpair[c];
#pragma unroll
    for (int kz = 0; kz < 3; ++kz)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float v0 = tile[tz + kz][ty + ky][tx + kx];
          const float v1 = tile[tz + 1 + kz][ty + ky][tx + kx];
          const f32x2 vv0 = f2_pack(v0, v0), vv1 = f2_pack(v1, v1);
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const f32x2 w = wpair[(kz * 9 + ky * 3 + kx) * 16 + c];
            acc0[c] = f2_fma(vv0, w, acc0[c]);
            acc1[c] = f2_fma(vv1, w, acc1[c]);
          }
        }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const size_t vox =
          (((size_t)b * a.Pz + (z0 + tz + h)) * a.Py + (y0 + ty)) * a.Px + (x0 + tx);
      T* dst = outp + vox * a.cstride + a.coff;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float f[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float lo, hi;
          f2_unpack(h == 0 ? acc0[4 * g + j] : acc1[4 * g + j], lo, hi);
          f[2 * j] = leaky_relu(lo);
          f[2 * j + 1] = leaky_relu(hi);
        }
        Vec8<T> o;
        o.from_float(f);
        o.store(dst + 8 * g);
      }
    }
  }
}

Status launch_stem(const PatchSource& src, const StemWeights& w, const Act& out, cudaStream_t s) {
  EXA_CHECK(out.C == 32 && out.cstride % 8 == 0 && out.coff % 8 == 0 && out.coff + 32 <= out.cstride,
            "stem output must be a 32-channel slot aligned to 8 channels");
  EXA_CHECK(out.W % ST_TX == 0 && out.H % ST_TY == 0 && out.D % ST_TZ == 0,
            "patch dims must be multiples of 16");
  StemArgs a;
  a.src = src;
  a.B = out.B;
  a.Pz = out.D;
  a.Py = out.H;
  a.Px = out.W;
  a.out = out.ptr;
  a.cstride = out.cstride;
  a.coff = out.coff;
  dim3 grid((out.W / ST_TX) * (out.H / ST_TY), out.D / ST_TZ, out.B);
  const bool from_vol = src.vol != nullptr;
  EXA_CHECK(from_vol || src.x != nullptr, "stem: no input source");
  if (out.fp32) {
    if (from_vol) stem_kernel<float, true><<<grid, 128, 0, s>>>(w, a);
    else stem_kernel<float, false><<<grid, 128, 0, s>>>(w, a);
  } else {
    if (from_vol) stem_kernel<__nv_bfloat16, true><<<grid, 128, 0, s>>>(w, a);
    else stem_kernel<__nv_bfloat16, false><<<grid, 128, 0, s>>>(w, a);
  }
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

// K0b for the tensor-core stem: gather + clip + normalise + far-end reflect padding
// (inference.py:79-80,188-191; img_util.py:378-379,424-428,526-531), written as an interleaved
// bf16 (hi, lo) pair per voxel (x = hi + lo to 2^-17 relative), x innermost: [B][Pz][Py][Px+8].
// Voxel x is stored at index x + 1: element 0 and the tail are zeros (the conv's zero padding),
// so that every 8-voxel window the tensor-core stem fetches starts at a multiple of 4 voxels
// (TMA needs 16 B aligned innermost offsets).
template <bool FROM_VOLUME>
__global__ void __launch_bounds__(256)
stem_split_kernel(const StemArgs a, __nv_bfloat162* __restrict__ xs) {
  // grid: x = (py, group of 4 row elements), y = pz, z = b; one 16 B store per thread
  const int Wp = a.Px + 8;
  const unsigned gw = (unsigned)Wp / 4;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (unsigned)a.Py * gw) return;
  const int xp = 4 * (int)(i % gw), py = (int)(i / gw), pz = blockIdx.y, b = blockIdx.z;
  int sx = 0, Lx = 0;
  size_t rowi = 0;
  if (FROM_VOLUME) {
    const int sz = a.src.starts[3 * b + 0], sy = a.src.starts[3 * b + 1];
    sx = a.src.starts[3 * b + 2];
    const int Lz = min(a.Pz, a.src.gD - sz), Ly = min(a.Py, a.src.gH - sy);
    Lx = min(a.Px, a.src.gW - sx);
    const int gz = sz + (pz < Lz ? pz : reflect_index(pz, Lz));
    const int gy = sy + (py < Ly ? py : reflect_index(py, Ly));
    rowi = ((size_t)(gz - a.src.vz0) * a.src.gH + gy) * a.src.gW;
  } else {
    rowi = (((size_t)b * a.Pz + pz) * a.Py + py) * a.Px;
  }
  uint32_t w[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int px = xp + k - 1;
    float v = 0.f;
    if (px >= 0 && px < a.Px) {
      if (FROM_VOLUME) {
        const int gx = sx + (px < Lx ? px : reflect_index(px, Lx));
        v = __ldg(a.src.lut + min((int)__ldg(a.src.vol + rowi + gx), a.src.clip));
      } else {
        v = __ldg(a.src.x + rowi + px);
      }
    }
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    __nv_bfloat162 hl;
    hl.x = h;
    hl.y = l;
    w[k] = *reinterpret_cast<uint32_t*>(&hl);
  }
  const size_t o = (((size_t)b * a.Pz + pz) * a.Py + py) * Wp + xp;
  *reinterpret_cast<uint4*>(xs + o) = make_uint4(w[0], w[1], w[2], w[3]);
}

Status launch_stem_split(const PatchSource& src, int B, int Pz, int Py, int Px, __nv_bfloat16* xs,
                         cudaStream_t s) {
  EXA_CHECK(Px % 4 == 0 && Pz <= 65535 && B <= 65535, "stem_split: patch dims");
  StemArgs a;
  a.src = src;
  a.B = B; a.Pz = Pz; a.Py = Py; a.Px = Px;
  a.out = nullptr;
  const dim3 grid((unsigned)ceil_div(Py * ((Px + 8) / 4), 256), (unsigned)Pz, (unsigned)B);
  const bool from_vol = src.vol != nullptr;
  EXA_CHECK(from_vol || src.x != nullptr, "stem_split: no input source");
  if (from_vol) stem_split_kernel<true><<<grid, 256, 0, s>>>(a, (__nv_bfloat162*)xs);
  else stem_split_kernel<false><<<grid, 256, 0, s>>>(a, (__nv_bfloat162*)xs);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

// ---------------------------------------------------------------------------
// K3: 2x2x2 max-pool, NDHWC, 8 channels (16 B for bf16) per thread
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
maxpool_kernel(const T* __restrict__ in, int in_cstride, int in_coff, T* __restrict__ out,
               int out_cstride, int out_coff, int B, int Do, int Ho, int Wo, int C) {
  // grid: x covers (yo, xo, c8) of one output plane, y = zo, z = b -> 32-bit index math only
  const unsigned cv = (unsigned)C / 8;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (unsigned)Ho * Wo * cv) return;
  const int c8 = (int)(i % cv);
  const unsigned v = i / cv;
  const int xo = (int)(v % (unsigned)Wo);
  const int yo = (int)(v / (unsigned)Wo);
  const int zo = blockIdx.y;
  const int b = blockIdx.z;
  const int Di = 2 * Do, Hi = 2 * Ho, Wi = 2 * Wo;
  float m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
  for (int dz = 0; dz < 2; ++dz)
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const size_t vox = (((size_t)b * Di + 2 * zo + dz) * Hi + 2 * yo + dy) * Wi + 2 * xo + dx;
        Vec8<T> q;
        q.load(in + vox * in_cstride + in_coff + 8 * c8);
        float f[8];
        q.to_float(f);
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], f[j]);
      }
  const size_t ovox = (((size_t)b * Do + zo) * Ho + yo) * Wo + xo;
  Vec8<T> o;
  o.from_float(m);  // max of bf16 values is a bf16 value: no extra rounding
  o.store(out + ovox * out_cstride + out_coff + 8 * c8);
}

Status launch_maxpool(const Act& in, const Act& out, cudaStream_t s) {
  EXA_CHECK(in.fp32 == out.fp32 && in.C == out.C && in.C % 8 == 0, "maxpool: type/channel mismatch");
  EXA_CHECK(in.D == 2 * out.D && in.H == 2 * out.H && in.W == 2 * out.W && in.B == out.B,
            "maxpool: shape mismatch");
  EXA_CHECK(out.B <= 65535 && out.D <= 65535, "maxpool: batch/depth too large for the grid");
  const dim3 blocks((unsigned)ceil_div(out.H * out.W * (out.C / 8), 256), (unsigned)out.D,
                    (unsigned)out.B);
  if (in.fp32)
    maxpool_kernel<float><<<blocks, 256, 0, s>>>((const float*)in.ptr, in.cstride, in.coff,
                                                  (float*)out.ptr, out.cstride, out.coff, out.B,
                                                  out.D, out.H, out.W, out.C);
  else
    maxpool_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(
        (const __nv_bfloat16*)in.ptr, in.cstride, in.coff, (__nv_bfloat16*)out.ptr, out.cstride,
        out.coff, out.B, out.D, out.H, out.W, out.C);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

// ---------------------------------------------------------------------------
// K4: trilinear x2 upsample, align_corners=True, into a concat slot
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
upsample_kernel(const T* __restrict__ in, int in_cstride, int in_coff, T* __restrict__ out,
                int out_cstride, int out_coff, int B, int Di, int Hi, int Wi, int C,
                const ConvRegion rg) {
  const int Do = 2 * Di, Ho = 2 * Hi, Wo = 2 * Wi;
  // grid: x covers (yo, xo, c8) of one plane of the output region, y = zo, z = b
  // -> 32-bit index math only
  const unsigned cv = (unsigned)C / 8;
  const unsigned rw = (unsigned)(rg.hi[2] - rg.lo[2]), rh = (unsigned)(rg.hi[1] - rg.lo[1]);
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rh * rw * cv) return;
  const int c8 = (int)(i % cv);
  const unsigned v = i / cv;
  const int xo = rg.lo[2] + (int)(v % rw);
  const int yo = rg.lo[1] + (int)(v / rw);
  const int zo = rg.lo[0] + blockIdx.y;
  const int b = blockIdx.z;
  // src = dst * (in-1)/(out-1), computed in fp32 like ATen's area_pixel_compute_source_index
  const float sz = (Do > 1) ? (float)(Di - 1) / (float)(Do - 1) : 0.f;
  const float sy = (Ho > 1) ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f;
  const float sx = (Wo > 1) ? (float)(Wi - 1) / (float)(Wo - 1) : 0.f;
  const float fz = sz * zo, fy = sy * yo, fx = sx * xo;
  const int z0 = (int)fz, y0 = (int)fy, x0 = (int)fx;
  const int z1 = z0 + (z0 < Di - 1 ? 1 : 0), y1 = y0 + (y0 < Hi - 1 ? 1 : 0),
            x1 = x0 + (x0 < Wi - 1 ? 1 : 0);
  const float wz1 = fz - z0, wy1 = fy - y0, wx1 = fx - x0;
  const float wz0 = 1.f - wz1, wy0 = 1.f - wy1, wx0 = 1.f - wx1;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const int zs[2] = {z0, z1}, ys[2] = {y0, y1}, xs[2] = {x0, x1};
  const float wz[2] = {wz0, wz1}, wy[2] = {wy0, wy1}, wx[2] = {wx0, wx1};
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int bb = 0; bb < 2; ++bb)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const size_t vox = (((size_t)b * Di + zs[a]) * Hi + ys[bb]) * Wi + xs[c];
        Vec8<T> q;
        q.load(in + vox * in_cstride + in_coff + 8 * c8);
        float f[8];
        q.to_float(f);
        const float w = wz[a] * wy[bb] * wx[c];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(w, f[j], acc[j]);
      }
  const size_t ovox = (((size_t)b * Do + zo) * Ho + yo) * Wo + xo;
  Vec8<T> o;
  o.from_float(acc);
  o.store(out + ovox * out_cstride + out_coff + 8 * c8);
}

// bf16 fast path: one thread produces a 2x2x2 output block for 8 channels from the 3x3x3 input
// window w = (max(j-1,0), j, min(j+1,n-1)), separably: 27 vector loads per 8 outputs instead
// of 64.  With align_corners=True and scale (n-1)/(2n-1), output 2j interpolates window
// elements (w0, w1) and output 2j+1 elements (w1, w2); l = src - floor index.  Channel pairs
// are processed as packed fp32 (FFMA2).
struct AxisTaps {
  int w[3];
  float a0, a1;  // output 2j   = a0*in[w0] + a1*in[w1]
  float b0, b1;  // output 2j+1 = b0*in[w1] + b1*in[w2]
};
__device__ __forceinline__ AxisTaps axis_taps(int j, int n) {
  AxisTaps t;
  t.w[0] = max(j - 1, 0);
  t.w[1] = j;
  t.w[2] = min(j + 1, n - 1);
  const float scale = (float)(n - 1) / (float)(2 * n - 1);
  const float sa = scale * (float)(2 * j), sb = scale * (float)(2 * j + 1);
  t.a1 = sa - (float)t.w[0];   // j = 0: sa = 0 and w0 = w1 = 0, any split of the weight is exact
  t.a0 = 1.f - t.a1;
  t.b1 = sb - (float)j;
  t.b0 = 1.f - t.b1;
  return t;
}

__global__ void __launch_bounds__(128)
upsample2_bf16_kernel(const __nv_bfloat16* __restrict__ in, int in_cstride, int in_coff,
                      __nv_bfloat16* __restrict__ out, int out_cstride, int out_coff, int Di, int Hi,
                      int Wi, int C, const ConvRegion rg, int jz0, int jy0, int jx0, int njy,
                      int njx) {
  const unsigned cv = (unsigned)C / 8;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (unsigned)njy * njx * cv) return;
  const int c8 = (int)(i % cv);
  const unsigned v = i / cv;
  const int xj = jx0 + (int)(v % (unsigned)njx);
  const int yj = jy0 + (int)(v / (unsigned)njx);
  const int zj = jz0 + blockIdx.y;
  const int b = blockIdx.z;
  const AxisTaps tz = axis_taps(zj, Di), ty = axis_taps(yj, Hi), tx = axis_taps(xj, Wi);
  const f32x2 xa0 = f2_pack(tx.a0, tx.a0), xa1 = f2_pack(tx.a1, tx.a1);
  const f32x2 xb0 = f2_pack(tx.b0, tx.b0), xb1 = f2_pack(tx.b1, tx.b1);
  const f32x2 zero = f2_pack(0.f, 0.f);

  f32x2 acc[2][2][2][4];  // [z out][y out][x out][channel pair]
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int bb = 0; bb < 2; ++bb)
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[a][bb][c][j] = zero;

  const __nv_bfloat16* base = in + in_coff + 8 * c8;
#pragma unroll
  for (int dz = 0; dz < 3; ++dz) {
    f32x2 py[2][2][4];  // [y out][x out]
#pragma unroll
    for (int bb = 0; bb < 2; ++bb)
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) py[bb][c][j] = zero;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const size_t rowv = (((size_t)b * Di + tz.w[dz]) * Hi + ty.w[dy]) * Wi;
      f32x2 f[3][4];
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const uint4 q = *reinterpret_cast<const uint4*>(base + (rowv + tx.w[dx]) * in_cstride);
        f[dx][0] = bf16x2_to_f2(q.x);
        f[dx][1] = bf16x2_to_f2(q.y);
        f[dx][2] = bf16x2_to_f2(q.z);
        f[dx][3] = bf16x2_to_f2(q.w);
      }
      f32x2 pxa[4], pxb[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        pxa[j] = f2_fma(xa1, f[1][j], f2_fma(xa0, f[0][j], zero));
        pxb[j] = f2_fma(xb1, f[2][j], f2_fma(xb0, f[1][j], zero));
      }
      // y pass: output 2yj uses window rows (0,1), output 2yj+1 rows (1,2)
      const float ca = dy == 0 ? ty.a0 : (dy == 1 ? ty.a1 : 0.f);
      const float cb = dy == 1 ? ty.b0 : (dy == 2 ? ty.b1 : 0.f);
      const f32x2 ca2 = f2_pack(ca, ca), cb2 = f2_pack(cb, cb);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (dy < 2) {
          py[0][0][j] = f2_fma(ca2, pxa[j], py[0][0][j]);
          py[0][1][j] = f2_fma(ca2, pxb[j], py[0][1][j]);
        }
        if (dy > 0) {
          py[1][0][j] = f2_fma(cb2, pxa[j], py[1][0][j]);
          py[1][1][j] = f2_fma(cb2, pxb[j], py[1][1][j]);
        }
      }
    }
    const float za = dz == 0 ? tz.a0 : (dz == 1 ? tz.a1 : 0.f);
    const float zb = dz == 1 ? tz.b0 : (dz == 2 ? tz.b1 : 0.f);
    const f32x2 za2 = f2_pack(za, za), zb2 = f2_pack(zb, zb);
#pragma unroll
    for (int bb = 0; bb < 2; ++bb)
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (dz < 2) acc[0][bb][c][j] = f2_fma(za2, py[bb][c][j], acc[0][bb][c][j]);
          if (dz > 0) acc[1][bb][c][j] = f2_fma(zb2, py[bb][c][j], acc[1][bb][c][j]);
        }
  }
  const int Ho = 2 * Hi, Wo = 2 * Wi, Do = 2 * Di;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int bb = 0; bb < 2; ++bb)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int zo = 2 * zj + a, yo = 2 * yj + bb, xo = 2 * xj + c;
        if (zo >= rg.lo[0] && zo < rg.hi[0] && yo >= rg.lo[1] && yo < rg.hi[1] && xo >= rg.lo[2] &&
            xo < rg.hi[2]) {
          const size_t ovox = (((size_t)b * Do + zo) * Ho + yo) * Wo + xo;
          uint4 o;
          float lo, hi;
          f2_unpack(acc[a][bb][c][0], lo, hi); o.x = pack_bf16x2(lo, hi);
          f2_unpack(acc[a][bb][c][1], lo, hi); o.y = pack_bf16x2(lo, hi);
          f2_unpack(acc[a][bb][c][2], lo, hi); o.z = pack_bf16x2(lo, hi);
          f2_unpack(acc[a][bb][c][3], lo, hi); o.w = pack_bf16x2(lo, hi);
          *reinterpret_cast<uint4*>(out + ovox * out_cstride + out_coff + 8 * c8) = o;
        }
      }
}

// bf16 streaming path: one thread owns a (yj, xj) column of 2x2 output pixels for 8 channels and
// marches along z.  Per step it loads the 3x3 in-plane window of ONE new input plane (9 x 16 B),
// interpolates it in x and y (4 output pixels) and blends it with the two previous planes'
// results into two output planes: 9 loads and ~150 packed FMAs per 8 output voxels instead of 27
// loads and ~230, and the z neighbours never leave registers.
// NJ = channel pairs per thread: 4 (8 channels, 16-byte accesses) or 2 (4 channels, 8-byte accesses:
// about half the registers, twice the resident warps).
template <int NJ>
struct PlaneXY {
  f32x2 v[2][2][NJ];  // [y out][x out][channel pair]
};
template <int NJ>
struct PlaneLoads {
  uint32_t q[3][3][NJ];
};
template <int NJ>
__device__ __forceinline__ void upsample_load(const __nv_bfloat16* __restrict__ plane_base,
                                              const int (&off)[3][3], PlaneLoads<NJ>& l) {
#pragma unroll
  for (int dy = 0; dy < 3; ++dy)
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      if (NJ == 4) {
        const uint4 t = *reinterpret_cast<const uint4*>(plane_base + off[dy][dx]);
        l.q[dy][dx][0] = t.x; l.q[dy][dx][1] = t.y; l.q[dy][dx][NJ - 2] = t.z; l.q[dy][dx][NJ - 1] = t.w;
      } else {
        const uint2 t = *reinterpret_cast<const uint2*>(plane_base + off[dy][dx]);
        l.q[dy][dx][0] = t.x; l.q[dy][dx][1] = t.y;
      }
    }
}
template <int NJ>
__device__ __forceinline__ void upsample_plane_xy(const PlaneLoads<NJ>& l, const AxisTaps& ty,
                                                  const AxisTaps& tx, PlaneXY<NJ>& o) {
  const f32x2 zero = f2_pack(0.f, 0.f);
  const f32x2 xa0 = f2_pack(tx.a0, tx.a0), xa1 = f2_pack(tx.a1, tx.a1);
  const f32x2 xb0 = f2_pack(tx.b0, tx.b0), xb1 = f2_pack(tx.b1, tx.b1);
  f32x2 pxa[3][NJ], pxb[3][NJ];  // x-interpolated rows: output 2xj / 2xj+1
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const f32x2 f0 = bf16x2_to_f2(l.q[dy][0][j]), f1 = bf16x2_to_f2(l.q[dy][1][j]),
                  f2 = bf16x2_to_f2(l.q[dy][2][j]);
      pxa[dy][j] = f2_fma(xa1, f1, f2_fma(xa0, f0, zero));
      pxb[dy][j] = f2_fma(xb1, f2, f2_fma(xb0, f1, zero));
    }
  }
  const f32x2 ya0 = f2_pack(ty.a0, ty.a0), ya1 = f2_pack(ty.a1, ty.a1);
  const f32x2 yb0 = f2_pack(ty.b0, ty.b0), yb1 = f2_pack(ty.b1, ty.b1);
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    o.v[0][0][j] = f2_fma(ya1, pxa[1][j], f2_fma(ya0, pxa[0][j], zero));
    o.v[0][1][j] = f2_fma(ya1, pxb[1][j], f2_fma(ya0, pxb[0][j], zero));
    o.v[1][0][j] = f2_fma(yb1, pxa[2][j], f2_fma(yb0, pxa[1][j], zero));
    o.v[1][1][j] = f2_fma(yb1, pxb[2][j], f2_fma(yb0, pxb[1][j], zero));
  }
}

// Along z every output plane blends two consecutive input planes: output 2j = a0(j)*p[j-1] +
// a1(j)*p[j], output 2j+1 = b0(j)*p[j] + b1(j)*p[j+1] (indices clamped).  So only TWO
// x/y-interpolated planes are live: when plane j+1 arrives, outputs 2j+1 and 2j+2 are emitted
// from (p[j], p[j+1]).  The loads of plane j+2 are issued before that blend, so their latency
// hides behind the packed FMAs and the stores.
template <int NJ>
__global__ void __launch_bounds__(128, NJ == 2 ? 5 : 3)
upsample_march_bf16_kernel(const __nv_bfloat16* __restrict__ in, int in_cstride, int in_coff,
                           __nv_bfloat16* __restrict__ out, int out_cstride, int out_coff, int Di,
                           int Hi, int Wi, int C, const ConvRegion rg, int jz0, int jy0, int jx0,
                           int njz, int njy, int njx) {
  constexpr int CPT = 2 * NJ;  // channels per thread
  const unsigned cv = (unsigned)C / CPT;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (unsigned)njy * njx * cv) return;
  const int cg = (int)(i % cv);
  const unsigned v = i / cv;
  const int xj = jx0 + (int)(v % (unsigned)njx);
  const int yj = jy0 + (int)(v / (unsigned)njx);
  const int b = blockIdx.z;
  const int zj_begin = jz0 + (int)blockIdx.y * njz, zj_end = min(zj_begin + njz, (rg.hi[0] + 1) / 2);
  if (zj_begin >= zj_end) return;
  const AxisTaps ty = axis_taps(yj, Hi), tx = axis_taps(xj, Wi);
  // all in-plane offsets are fixed for the thread (32-bit; a patch plane is far below 2^31 elements)
  int off[3][3];
#pragma unroll
  for (int dy = 0; dy < 3; ++dy)
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) off[dy][dx] = (ty.w[dy] * Wi + tx.w[dx]) * in_cstride;
  const int Ho = 2 * Hi, Wo = 2 * Wi, Do = 2 * Di;
  int ooff[2][2];
  bool ok[2][2];
#pragma unroll
  for (int bb = 0; bb < 2; ++bb)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int yo = 2 * yj + bb, xo = 2 * xj + c;
      ooff[bb][c] = (yo * Wo + xo) * out_cstride;
      ok[bb][c] = yo >= rg.lo[1] && yo < rg.hi[1] && xo >= rg.lo[2] && xo < rg.hi[2];
    }
  const size_t in_plane = (size_t)Hi * Wi * in_cstride;
  const size_t out_plane = (size_t)Ho * Wo * out_cstride;
  const __nv_bfloat16* ibase = in + in_coff + CPT * cg + (size_t)b * Di * in_plane;
  __nv_bfloat16* obase = out + out_coff + CPT * cg + (size_t)b * Do * out_plane;
  const f32x2 zero = f2_pack(0.f, 0.f);

  // one output plane zo = w_lo * lo + w_hi * hi, stored for the (up to) four pixels of the column
  auto emit = [&](int zo, float w_lo, float w_hi, const PlaneXY<NJ>& lo, const PlaneXY<NJ>& hi) {
    if (zo < rg.lo[0] || zo >= rg.hi[0]) return;
    const f32x2 wl = f2_pack(w_lo, w_lo), wh = f2_pack(w_hi, w_hi);
    __nv_bfloat16* oplane = obase + (size_t)zo * out_plane;
#pragma unroll
    for (int bb = 0; bb < 2; ++bb)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (!ok[bb][c]) continue;
        uint32_t o[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const f32x2 r = f2_fma(wh, hi.v[bb][c][j], f2_fma(wl, lo.v[bb][c][j], zero));
          float a, d;
          f2_unpack(r, a, d);
          o[j] = pack_bf16x2(a, d);
        }
        if (NJ == 4)
          *reinterpret_cast<uint4*>(oplane + ooff[bb][c]) = make_uint4(o[0], o[1], o[NJ - 2], o[NJ - 1]);
        else
          *reinterpret_cast<uint2*>(oplane + ooff[bb][c]) = make_uint2(o[0], o[1]);
      }
  };

  PlaneLoads<NJ> ld;
  PlaneXY<NJ> pc, pn;  // interpolated planes zj and zj+1
  {
    // prologue: output 2*zj_begin blends planes (zj_begin-1, zj_begin)
    PlaneXY<NJ> pm;
    upsample_load<NJ>(ibase + (size_t)max(zj_begin - 1, 0) * in_plane, off, ld);
    upsample_plane_xy<NJ>(ld, ty, tx, pm);
    upsample_load<NJ>(ibase + (size_t)zj_begin * in_plane, off, ld);
    upsample_plane_xy<NJ>(ld, ty, tx, pc);
    const AxisTaps tz = axis_taps(zj_begin, Di);
    emit(2 * zj_begin, tz.a0, tz.a1, pm, pc);
  }
  upsample_load<NJ>(ibase + (size_t)min(zj_begin + 1, Di - 1) * in_plane, off, ld);
  for (int zj = zj_begin; zj < zj_end; ++zj) {
    upsample_plane_xy<NJ>(ld, ty, tx, pn);  // plane min(zj+1, Di-1) (clamped: window element 2)
    if (zj + 1 < zj_end) upsample_load<NJ>(ibase + (size_t)min(zj + 2, Di - 1) * in_plane, off, ld);
    const AxisTaps tz = axis_taps(zj, Di);
    emit(2 * zj + 1, tz.b0, tz.b1, pc, pn);
    if (zj + 1 < zj_end) {
      const AxisTaps tn = axis_taps(zj + 1, Di);
      emit(2 * zj + 2, tn.a0, tn.a1, pc, pn);
    }
    pc = pn;
  }
}

Status launch_upsample(const Act& in, const Act& out, const ConvRegion* region, cudaStream_t s) {
  EXA_CHECK(in.fp32 == out.fp32 && in.C == out.C && in.C % 8 == 0, "upsample: type/channel mismatch");
  EXA_CHECK(out.D == 2 * in.D && out.H == 2 * in.H && out.W == 2 * in.W && in.B == out.B,
            "upsample: shape mismatch");
  EXA_CHECK(out.B <= 65535 && out.D <= 65535, "upsample: batch/depth too large for the grid");
  ConvRegion rg;
  rg.lo[0] = rg.lo[1] = rg.lo[2] = 0;
  rg.hi[0] = out.D; rg.hi[1] = out.H; rg.hi[2] = out.W;
  if (region) {
    rg = *region;
    EXA_CHECK(rg.lo[0] >= 0 && rg.lo[1] >= 0 && rg.lo[2] >= 0 && rg.hi[0] <= out.D &&
                  rg.hi[1] <= out.H && rg.hi[2] <= out.W && rg.lo[0] < rg.hi[0] &&
                  rg.lo[1] < rg.hi[1] && rg.lo[2] < rg.hi[2],
              "upsample: bad output region");
  }
  const dim3 blocks(
      (unsigned)ceil_div((rg.hi[1] - rg.lo[1]) * (rg.hi[2] - rg.lo[2]) * (out.C / 8), 256),
      (unsigned)(rg.hi[0] - rg.lo[0]), (unsigned)out.B);
  if (in.fp32) {
    upsample_kernel<float><<<blocks, 256, 0, s>>>((const float*)in.ptr, in.cstride, in.coff,
                                                   (float*)out.ptr, out.cstride, out.coff, in.B,
                                                   in.D, in.H, in.W, in.C, rg);
  } else {
    // 2x2x2 output blocks: block index ranges covering the region
    const int jz0 = rg.lo[0] / 2, jy0 = rg.lo[1] / 2, jx0 = rg.lo[2] / 2;
    const int njz = (rg.hi[0] + 1) / 2 - jz0, njy = (rg.hi[1] + 1) / 2 - jy0,
              njx = (rg.hi[2] + 1) / 2 - jx0;
    // channels per thread: 8 (16-byte accesses).  The 4-channel form (EXA_UP_CPT=4: 93 instead of
    // 149 registers, 5 instead of 3 resident blocks) was measured SLOWER on a B200, 12.8 vs 10.3 ms
    // per 512^3: the kernel is bound by its 8-byte accesses then, not by occupancy
    // (profiles/r02g_ab_upsample_cpt.txt)
    static const int cpt = []() {
      const char* e = getenv("EXA_UP_CPT");
      return e && atoi(e) == 4 ? 4 : 8;
    }();
    // z is marched inside the kernel; split it only as far as needed to fill the GPU
    const int threads_per_col = njy * njx * (out.C / cpt) * out.B;
    int zsplit = 1;
    while (zsplit < njz && (long long)threads_per_col * zsplit < 148LL * 2048) zsplit *= 2;
    const int zchunk = ceil_div(njz, zsplit);
    const dim3 blocks2((unsigned)ceil_div(njy * njx * (out.C / cpt), 128), (unsigned)ceil_div(njz, zchunk),
                       (unsigned)out.B);
    if (cpt == 8) {
      upsample_march_bf16_kernel<4><<<blocks2, 128, 0, s>>>(
          (const __nv_bfloat16*)in.ptr, in.cstride, in.coff, (__nv_bfloat16*)out.ptr, out.cstride,
          out.coff, in.D, in.H, in.W, in.C, rg, jz0, jy0, jx0, zchunk, njy, njx);
    } else {
      upsample_march_bf16_kernel<2><<<blocks2, 128, 0, s>>>(
          (const __nv_bfloat16*)in.ptr, in.cstride, in.coff, (__nv_bfloat16*)out.ptr, out.cstride,
          out.coff, in.D, in.H, in.W, in.C, rg, jz0, jy0, jx0, zchunk, njy, njx);
    }
  }
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

// ---------------------------------------------------------------------------
// K5' (fp32 validation mode, and bf16 models wider than 32 channels where the head cannot ride
// in the last conv's epilogue): 1x1x1 head + sigmoid + trim  (unet3d.py:318, inference.py:158-162)
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
head_kernel(const T* __restrict__ in, int cstride, int coff, int cin, int B, int D, int H, int W,
            const float* __restrict__ hw, const float* __restrict__ hb, float* __restrict__ out,
            int C, int trim, int apply_sigmoid) {
  const int Dz = D - 2 * trim, Hy = H - 2 * trim, Wx = W - 2 * trim;
  const size_t total = (size_t)B * Dz * Hy * Wx;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % Wx);
  size_t v = i / Wx;
  const int y = (int)(v % Hy);
  v /= Hy;
  const int z = (int)(v % Dz);
  const int b = (int)(v / Dz);
  const size_t vox = (((size_t)b * D + z + trim) * H + y + trim) * W + x + trim;
  const T* src = in + vox * cstride + coff;
  float s[8];
#pragma unroll
  for (int oc = 0; oc < 8; ++oc) s[oc] = oc < C ? __ldg(hb + oc) : 0.f;
  // channels in ascending order, eight at a time (the accumulation order of the fp32 reference
  // is not specified; 32-term fp32 sums agree to ~1e-7)
  for (int c8 = 0; c8 < cin; c8 += 8) {
    float f[8];
    Vec8<T> q;
    q.load(src + c8);
    q.to_float(f);
    for (int oc = 0; oc < C; ++oc) {
      float acc = s[oc];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(__ldg(hw + oc * cin + c8 + j), f[j], acc);
      s[oc] = acc;
    }
  }
  for (int oc = 0; oc < C; ++oc) {
    float r = s[oc];
    if (apply_sigmoid) r = 1.f / (1.f + expf(-r));
    out[((((size_t)b * C + oc) * Dz + z) * Hy + y) * Wx + x] = r;
  }
}

Status launch_head(const Act& in, const HeadParams& h, cudaStream_t s) {
  EXA_CHECK(in.C % 8 == 0 && in.cstride % 8 == 0 && in.coff % 8 == 0 && h.C >= 1 && h.C <= 8,
            "head: channels must be multiples of 8 and at most 8 outputs");
  const size_t total =
      (size_t)in.B * (in.D - 2 * h.trim) * (in.H - 2 * h.trim) * (in.W - 2 * h.trim);
  const int blocks = (int)ceil_div64((int64_t)total, 256);
  if (in.fp32) {
    head_kernel<float><<<blocks, 256, 0, s>>>((const float*)in.ptr, in.cstride, in.coff, in.C, in.B,
                                               in.D, in.H, in.W, h.w, h.b, h.out, h.C, h.trim,
                                               h.apply_sigmoid);
  } else {
    head_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>((const __nv_bfloat16*)in.ptr, in.cstride,
                                                       in.coff, in.C, in.B, in.D, in.H, in.W, h.w,
                                                       h.b, h.out, h.C, h.trim, h.apply_sigmoid);
  }
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

// ---------------------------------------------------------------------------
// ConvTranspose3d(k = 2, s = 2) of an Up block built with trilinear=False (unet3d.py:254-256):
// out[b, 2z+dz, 2y+dy, 2x+dx, co] = bias[co] + sum_ci in[b, z, y, x, ci] * w[ci][co][dz][dy][dx],
// written into the concat slot.  fp32 accumulation in ascending ci; bf16 mode rounds the result
// once.  One thread: one output voxel x 8 output channels.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
upconv_kernel(const T* __restrict__ in, int in_cstride, int in_coff, int cin, T* __restrict__ out,
              int out_cstride, int out_coff, int cout, int B, int Di, int Hi, int Wi,
              const float* __restrict__ w, const float* __restrict__ bias, const ConvRegion rg) {
  const int c8n = cout / 8;
  const int rz = rg.hi[0] - rg.lo[0], ry = rg.hi[1] - rg.lo[1], rx = rg.hi[2] - rg.lo[2];
  const size_t total = (size_t)B * rz * ry * rx * c8n;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c8 = (int)(i % c8n);
  size_t v = i / c8n;
  const int xo = rg.lo[2] + (int)(v % rx);
  v /= rx;
  const int yo = rg.lo[1] + (int)(v % ry);
  v /= ry;
  const int zo = rg.lo[0] + (int)(v % rz);
  const int b = (int)(v / rz);
  const int tap = (zo & 1) * 4 + (yo & 1) * 2 + (xo & 1);
  const T* src = in + in_coff +
                 ((((size_t)b * Di + (zo >> 1)) * Hi + (yo >> 1)) * Wi + (xo >> 1)) * in_cstride;
  const float* wt = w + (size_t)tap * cin * cout + 8 * c8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = __ldg(bias + 8 * c8 + j);
  for (int ci8 = 0; ci8 < cin; ci8 += 8) {
    float f[8];
    Vec8<T> q;
    q.load(src + ci8);
    q.to_float(f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(wt + (size_t)(ci8 + k) * cout));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(wt + (size_t)(ci8 + k) * cout + 4));
      acc[0] = fmaf(f[k], w0.x, acc[0]); acc[1] = fmaf(f[k], w0.y, acc[1]);
      acc[2] = fmaf(f[k], w0.z, acc[2]); acc[3] = fmaf(f[k], w0.w, acc[3]);
      acc[4] = fmaf(f[k], w1.x, acc[4]); acc[5] = fmaf(f[k], w1.y, acc[5]);
      acc[6] = fmaf(f[k], w1.z, acc[6]); acc[7] = fmaf(f[k], w1.w, acc[7]);
    }
  }
  const int Ho = 2 * Hi, Wo = 2 * Wi, Do = 2 * Di;
  T* dst = out + out_coff + ((((size_t)b * Do + zo) * Ho + yo) * Wo + xo) * out_cstride + 8 * c8;
  Vec8<T> o;
  o.from_float(acc);
  o.store(dst);
}

Status launch_upconv(const Act& in, const Act& out, const float* w, const float* bias,
                     const ConvRegion* region, cudaStream_t s) {
  EXA_CHECK(w && bias, "upconv: the model has no transposed-conv weights");
  EXA_CHECK(in.fp32 == out.fp32 && in.C % 8 == 0 && out.C % 8 == 0 && in.cstride % 8 == 0 &&
                in.coff % 8 == 0 && out.cstride % 8 == 0 && out.coff % 8 == 0,
            "upconv: type/channel mismatch");
  EXA_CHECK(out.D == 2 * in.D && out.H == 2 * in.H && out.W == 2 * in.W && in.B == out.B,
            "upconv: shape mismatch");
  ConvRegion rg;
  rg.lo[0] = rg.lo[1] = rg.lo[2] = 0;
  rg.hi[0] = out.D; rg.hi[1] = out.H; rg.hi[2] = out.W;
  if (region) {
    rg = *region;
    EXA_CHECK(rg.lo[0] >= 0 && rg.lo[1] >= 0 && rg.lo[2] >= 0 && rg.hi[0] <= out.D &&
                  rg.hi[1] <= out.H && rg.hi[2] <= out.W && rg.lo[0] < rg.hi[0] &&
                  rg.lo[1] < rg.hi[1] && rg.lo[2] < rg.hi[2],
              "upconv: bad output region");
  }
  const size_t total = (size_t)out.B * (rg.hi[0] - rg.lo[0]) * (rg.hi[1] - rg.lo[1]) *
                       (rg.hi[2] - rg.lo[2]) * (out.C / 8);
  const unsigned blocks = (unsigned)ceil_div64((int64_t)total, 256);
  if (in.fp32) {
    upconv_kernel<float><<<blocks, 256, 0, s>>>((const float*)in.ptr, in.cstride, in.coff, in.C,
                                                 (float*)out.ptr, out.cstride, out.coff, out.C, in.B,
                                                 in.D, in.H, in.W, w, bias, rg);
  } else {
    upconv_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(
        (const __nv_bfloat16*)in.ptr, in.cstride, in.coff, in.C, (__nv_bfloat16*)out.ptr,
        out.cstride, out.coff, out.C, in.B, in.D, in.H, in.W, w, bias, rg);
  }
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

// ---------------------------------------------------------------------------
// K6: overlap stitch (gather form) + coverage normalisation
// ---------------------------------------------------------------------------
// For an output coordinate p along an axis, windows k with
//   k*stride + trim <= p < min(k*stride + trim + keep, dim)      (inference.py:101-105)
// cover it; keep = patch - 2*trim.  [k_lo, k_hi] is that index range clamped to [0, n).
__device__ __forceinline__ void cover_range(const AxisGeom& g, int p, int& k_lo, int& k_hi) {
  const int keep = g.patch - 2 * g.trim;
  const int q = p - g.trim;  // need k*stride <= q  and  q < k*stride + keep
  if (q < 0) {
    k_lo = 0;
    k_hi = -1;
    return;
  }
  k_hi = min(q / g.stride, g.n - 1);
  const int t = q - keep + 1;  // need k*stride >= t
  k_lo = t <= 0 ? 0 : (t + g.stride - 1) / g.stride;
}

__global__ void __launch_bounds__(256)
stitch_kernel(const StitchArgs a) {
  const int H = a.ay.dim, W = a.ax.dim;
  // grid: x covers one (y, x) plane, y = plane index -> 32-bit index math only
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (unsigned)(a.y_end - a.y_begin) * W) return;
  const int x = (int)(i % (unsigned)W);
  const int y = a.y_begin + (int)(i / (unsigned)W);
  const int zl = blockIdx.y;
  const int z = a.z_begin + zl;

  int kz0, kz1, ky0, ky1, kx0, kx1;
  cover_range(a.az, z, kz0, kz1);
  cover_range(a.ay, y, ky0, ky1);
  cover_range(a.ax, x, kx0, kx1);
  // global coverage count (all rows, also those owned by other slabs)
  const int cnt = max(kz1 - kz0 + 1, 0) * max(ky1 - ky0 + 1, 0) * max(kx1 - kx0 + 1, 0);
  // restrict z rows to the resident slab
  const int rz0 = max(kz0, a.row_begin), rz1 = min(kz1, a.row_end - 1);

  const int Pt_z = a.az.patch - 2 * a.az.trim, Pt_y = a.ay.patch - 2 * a.ay.trim,
            Pt_x = a.ax.patch - 2 * a.ax.trim;
  const size_t chan = (size_t)Pt_z * Pt_y * Pt_x;
  for (int c = 0; c < a.C; ++c) {
    float acc = 0.f;
    if (a.seed != nullptr && z >= a.seed_z0 && z < a.seed_z1) {
      acc = a.seed[(((size_t)c * (a.seed_z1 - a.seed_z0) + (z - a.seed_z0)) * H + y) * W + x];
    }
    // same order as the reference's patch loop: z-major, then y, then x (inference.py:396,115)
    for (int kz = rz0; kz <= rz1; ++kz) {
      const int lz = z - (kz * a.az.stride + a.az.trim);
      for (int ky = ky0; ky <= ky1; ++ky) {
        const int ly = y - (ky * a.ay.stride + a.ay.trim);
        for (int kx = kx0; kx <= kx1; ++kx) {
          const int lx = x - (kx * a.ax.stride + a.ax.trim);
          const size_t slot = ((size_t)(kz - a.row_begin) * a.ay.n + ky) * a.ax.n + kx;
          acc += __ldg(a.probs + (slot * a.C + c) * chan + ((size_t)lz * Pt_y + ly) * Pt_x + lx);
        }
      }
    }
    if (a.finalize && cnt > 0) acc = acc / (float)cnt;
    const size_t o = (size_t)c * a.out_cstride + ((size_t)zl * H + y) * W + x;
    a.out[o] = acc;
    for (int p = 0; p < a.n_peers; ++p) a.peer_out[p][o] = acc;
  }
}

// Same, four consecutive x per thread (16 B loads / stores).  Valid when the x geometry is a
// multiple of 4 (trim, stride, kept width, volume width) -- then the set of covering windows is
// the same for the four voxels and every access is 16 B aligned.  Same summation order.
__global__ void __launch_bounds__(256)
stitch_kernel_x4(const StitchArgs a) {
  const int H = a.ay.dim, W = a.ax.dim, W4 = W >> 2;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (unsigned)(a.y_end - a.y_begin) * W4) return;
  const int x = 4 * (int)(i % (unsigned)W4);
  const int y = a.y_begin + (int)(i / (unsigned)W4);
  const int zl = blockIdx.y;
  const int z = a.z_begin + zl;
  int kz0, kz1, ky0, ky1, kx0, kx1;
  cover_range(a.az, z, kz0, kz1);
  cover_range(a.ay, y, ky0, ky1);
  cover_range(a.ax, x, kx0, kx1);
  const int cnt = max(kz1 - kz0 + 1, 0) * max(ky1 - ky0 + 1, 0) * max(kx1 - kx0 + 1, 0);
  const int rz0 = max(kz0, a.row_begin), rz1 = min(kz1, a.row_end - 1);
  const int Pt_z = a.az.patch - 2 * a.az.trim, Pt_y = a.ay.patch - 2 * a.ay.trim,
            Pt_x = a.ax.patch - 2 * a.ax.trim;
  const size_t chan = (size_t)Pt_z * Pt_y * Pt_x;
  const float inv = 1.f / (float)max(cnt, 1);
  for (int c = 0; c < a.C; ++c) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.seed != nullptr && z >= a.seed_z0 && z < a.seed_z1) {
      acc = *reinterpret_cast<const float4*>(
          a.seed + (((size_t)c * (a.seed_z1 - a.seed_z0) + (z - a.seed_z0)) * H + y) * W + x);
    }
    for (int kz = rz0; kz <= rz1; ++kz) {
      const int lz = z - (kz * a.az.stride + a.az.trim);
      for (int ky = ky0; ky <= ky1; ++ky) {
        const int ly = y - (ky * a.ay.stride + a.ay.trim);
        for (int kx = kx0; kx <= kx1; ++kx) {
          const int lx = x - (kx * a.ax.stride + a.ax.trim);
          const size_t slot = ((size_t)(kz - a.row_begin) * a.ay.n + ky) * a.ax.n + kx;
          const float4 v = __ldg(reinterpret_cast<const float4*>(
              a.probs + (slot * a.C + c) * chan + ((size_t)lz * Pt_y + ly) * Pt_x + lx));
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
      }
    }
    if (a.finalize && cnt > 0) {
      // counts are 1, 2, 4 or 8: multiplying by the exact reciprocal equals the division
      if ((cnt & (cnt - 1)) == 0) { acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv; }
      else { acc.x /= (float)cnt; acc.y /= (float)cnt; acc.z /= (float)cnt; acc.w /= (float)cnt; }
    }
    const size_t o = (size_t)c * a.out_cstride + ((size_t)zl * H + y) * W + x;
    *reinterpret_cast<float4*>(a.out + o) = acc;
    // the gather, fused: one 16 B store per peer (a warp writes 512 contiguous bytes per peer)
    for (int p = 0; p < a.n_peers; ++p) *reinterpret_cast<float4*>(a.peer_out[p] + o) = acc;
  }
}

Status launch_stitch(const StitchArgs& a_in, cudaStream_t s) {
  StitchArgs a = a_in;
  if (a.y_begin == 0 && a.y_end == 0) a.y_end = a.ay.dim;
  const int nz = a.z_end - a.z_begin, ny = a.y_end - a.y_begin;
  EXA_CHECK(a.y_begin >= 0 && a.y_end <= a.ay.dim, "stitch: row range out of bounds");
  if (nz <= 0 || ny <= 0 || a.ax.dim <= 0) return Status::OK();
  EXA_CHECK(nz <= 65535, "stitch: too many planes for one launch");
  const int keep_x = a.ax.patch - 2 * a.ax.trim;
  const size_t chan = (size_t)(a.az.patch - 2 * a.az.trim) * (a.ay.patch - 2 * a.ay.trim) * keep_x;
  const bool x4 = a.ax.dim % 4 == 0 && a.ax.trim % 4 == 0 && a.ax.stride % 4 == 0 && keep_x % 4 == 0 &&
                  chan % 4 == 0 && a.out_cstride % 4 == 0 && ((uintptr_t)a.out & 15) == 0 &&
                  ((uintptr_t)a.probs & 15) == 0 && ((uintptr_t)a.seed & 15) == 0;
  for (int p = 0; p < a.n_peers; ++p)
    EXA_CHECK(a.peer_out[p] != nullptr && ((uintptr_t)a.peer_out[p] & 15) == ((uintptr_t)a.out & 15),
              "stitch: peer output misaligned");
  if (x4) {
    const dim3 blocks4((unsigned)ceil_div64((int64_t)ny * (a.ax.dim / 4), 256), (unsigned)nz);
    stitch_kernel_x4<<<blocks4, 256, 0, s>>>(a);
    EXA_CUDA(cudaGetLastError());
    return Status::OK();
  }
  const dim3 blocks((unsigned)ceil_div64((int64_t)ny * a.ax.dim, 256), (unsigned)nz);
  stitch_kernel<<<blocks, 256, 0, s>>>(a);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

// ---------------------------------------------------------------------------
// Kf: fp32 SIMT 3x3x3 conv + bias + LeakyReLU (validation mode)
//   block = 128 voxels (linear run of the flattened B*D*H*W index) x 32 output channels,
//   256 threads, each 4 voxels x 4 channels; K loop = 27 taps x 16-channel chunks.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
conv3x3_fp32_kernel(const float* __restrict__ in, int in_cstride, int in_coff,
                    const float* __restrict__ w,  // [27][Cin][Cout]
                    const float* __restrict__ bias, float* __restrict__ out, int out_cstride,
                    int out_coff, int B, int D, int H, int W, int Cin, int Cout) {
  __shared__ float As[16][128 + 4];
  __shared__ __align__(16) float Ws[16][32];
  const size_t nvox = (size_t)B * D * H * W;
  const size_t v0 = (size_t)blockIdx.x * 128;
  const int n0 = blockIdx.y * 32;
  const int t = threadIdx.x;
  // loader role: voxel lv, channel half lh (8 channels)
  const int lv = t >> 1, lh = (t & 1) * 8;
  const size_t lvox = v0 + lv;
  int lx = 0, ly = 0, lz = 0, lb = 0;
  const bool lvalid = lvox < nvox;
  if (lvalid) {
    size_t r = lvox;
    lx = (int)(r % W); r /= W;
    ly = (int)(r % H); r /= H;
    lz = (int)(r % D);
    lb = (int)(r / D);
  }
  // compute role
  const int vg = t & 31, cg = t >> 5;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int tap = 0; tap < 27; ++tap) {
    const int kz = tap / 9 - 1, ky = (tap / 3) % 3 - 1, kx = tap % 3 - 1;
    const int sx = lx + kx, sy = ly + ky, sz = lz + kz;
    const bool inb = lvalid && sx >= 0 && sx < W && sy >= 0 && sy < H && sz >= 0 && sz < D;
    const float* src = in + ((((size_t)lb * D + sz) * H + sy) * W + sx) * in_cstride + in_coff + lh;
    for (int c0 = 0; c0 < Cin; c0 += 16) {
      float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0;
      if (inb) {
        q0 = *reinterpret_cast<const float4*>(src + c0);
        q1 = *reinterpret_cast<const float4*>(src + c0 + 4);
      }
      // weights: 16 x 32 floats, 2 per thread
      const int wr = t >> 4, wc = (t & 15) * 2;
      const float2 wq = *reinterpret_cast<const float2*>(w + ((size_t)tap * Cin + c0 + wr) * Cout + n0 + wc);
      __syncthreads();
      As[lh + 0][lv] = q0.x; As[lh + 1][lv] = q0.y; As[lh + 2][lv] = q0.z; As[lh + 3][lv] = q0.w;
      As[lh + 4][lv] = q1.x; As[lh + 5][lv] = q1.y; As[lh + 6][lv] = q1.z; As[lh + 7][lv] = q1.w;
      Ws[wr][wc] = wq.x;
      Ws[wr][wc + 1] = wq.y;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float4 av = *reinterpret_cast<const float4*>(&As[k][vg * 4]);
        const float4 wv = *reinterpret_cast<const float4*>(&Ws[k][cg * 4]);
        const float a4[4] = {av.x, av.y, av.z, av.w};
        const float w4[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], w4[j], acc[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const size_t vox = v0 + vg * 4 + i;
    if (vox < nvox) {
      float4 o;
      o.x = leaky_relu(acc[i][0] + bias[n0 + cg * 4 + 0]);
      o.y = leaky_relu(acc[i][1] + bias[n0 + cg * 4 + 1]);
      o.z = leaky_relu(acc[i][2] + bias[n0 + cg * 4 + 2]);
      o.w = leaky_relu(acc[i][3] + bias[n0 + cg * 4 + 3]);
      *reinterpret_cast<float4*>(out + vox * out_cstride + out_coff + n0 + cg * 4) = o;
    }
  }
}

Status launch_conv_fp32(const Act& in, const Act& out, const float* w, const float* bias,
                        cudaStream_t s) {
  EXA_CHECK(in.fp32 && out.fp32, "conv_fp32 expects fp32 activations");
  EXA_CHECK(in.C % 16 == 0 && out.C % 32 == 0, "conv_fp32: Cin%16 / Cout%32");
  EXA_CHECK(in.B == out.B && in.D == out.D && in.H == out.H && in.W == out.W, "conv_fp32: shape");
  dim3 grid((unsigned)ceil_div64((int64_t)in.voxels(), 128), out.C / 32);
  conv3x3_fp32_kernel<<<grid, 256, 0, s>>>((const float*)in.ptr, in.cstride, in.coff, w, bias,
                                           (float*)out.ptr, out.cstride, out.coff, in.B, in.D,
                                           in.H, in.W, in.C, out.C);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

}  // namespace exa
