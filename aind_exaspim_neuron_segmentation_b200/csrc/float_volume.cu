// Floating-point images (reference inference.py:79-80 accepts any dtype: np.minimum(img, clip), then
// img_util.normalize in float64).  The kernels of the path consume uint16 volumes through a lookup
// table, so a float volume is RANK-COMPRESSED on the GPU: the distinct values of min(x, clip) are
// sorted into a table (at most 65536 of them -- true for ExaSPIM data stored as float, whose
// clipped values are integers; anything else is refused, not approximated), every voxel becomes the
// uint16 index of its value, and the normalisation table is evaluated on the exact values in
// float64 (Engine::set_normalization_table).  Percentiles come from the histogram of the indices
// (percentiles_from_hist_values), so the result equals the reference's bit for bit.
#include <cub/cub.cuh>

#include <vector>

#include "common.cuh"
#include "float_volume.h"

namespace exa {

namespace {

template <typename T>
__global__ void __launch_bounds__(256)
clip_kernel(const T* __restrict__ in, size_t n, T clip, T* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const T x = in[i];
  out[i] = x < clip ? x : (x != x ? x : clip);  // np.minimum: NaN propagates (refused below)
}

// index of min(x, clip) in the sorted table of distinct values
template <typename T>
__global__ void __launch_bounds__(256)
rank_kernel(const T* __restrict__ in, size_t n, T clip, const T* __restrict__ table, int m,
            uint16_t* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const T x = in[i];
  const T y = x < clip ? x : clip;
  int lo = 0, hi = m - 1;  // table[lo] <= y <= table[hi]; every y is in the table
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (table[mid] < y) lo = mid + 1; else hi = mid;
  }
  out[i] = (uint16_t)lo;
}

struct Scratch {
  void* p = nullptr;
  ~Scratch() {
    if (p) cudaFree(p);
  }
  Status alloc(size_t bytes) {
    EXA_CUDA(cudaMalloc(&p, bytes ? bytes : 1));
    return Status::OK();
  }
};

template <typename T>
Status compress_impl(const T* vol, size_t n, double clip, uint16_t* idx, double* table_out,
                     int* n_table, cudaStream_t s) {
  EXA_CHECK(n < ((size_t)1 << 31), "float image: more than 2^31 voxels");
  const unsigned blocks = (unsigned)((n + 255) / 256);
  Scratch a, b, tmp, count;
  EXA_TRY(a.alloc(n * sizeof(T)));
  EXA_TRY(b.alloc(n * sizeof(T)));
  EXA_TRY(count.alloc(sizeof(int)));
  T* clipped = static_cast<T*>(a.p);
  T* sorted = static_cast<T*>(b.p);
  const T clip_t = (T)clip;
  clip_kernel<T><<<blocks, 256, 0, s>>>(vol, n, clip_t, clipped);
  EXA_CUDA(cudaGetLastError());
  size_t bytes = 0;
  EXA_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, clipped, sorted, (int)n, 0, (int)sizeof(T) * 8, s));
  EXA_TRY(tmp.alloc(bytes));
  EXA_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, bytes, clipped, sorted, (int)n, 0, (int)sizeof(T) * 8, s));
  // distinct values (== on floats: -0.0 and 0.0 are one value) into `clipped`, which is free again
  size_t bytes2 = 0;
  EXA_CUDA(cub::DeviceSelect::Unique(nullptr, bytes2, sorted, clipped, static_cast<int*>(count.p), (int)n, s));
  Scratch tmp2;
  EXA_TRY(tmp2.alloc(bytes2));
  EXA_CUDA(cub::DeviceSelect::Unique(tmp2.p, bytes2, sorted, clipped, static_cast<int*>(count.p), (int)n, s));
  int m = 0;
  EXA_CUDA(cudaMemcpyAsync(&m, count.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  EXA_CUDA(cudaStreamSynchronize(s));
  EXA_CHECK(m >= 1, "float image: empty volume");
  EXA_CHECK(m <= 65536,
            "float image: more than 65536 distinct values after the brightness clip (" +
                std::to_string(m) + "); the B200 path handles float images whose clipped values "
                "form a table of at most 65536 entries (e.g. integer data stored as float)");
  std::vector<T> host(m);
  EXA_CUDA(cudaMemcpyAsync(host.data(), clipped, (size_t)m * sizeof(T), cudaMemcpyDeviceToHost, s));
  EXA_CUDA(cudaStreamSynchronize(s));
  EXA_CHECK(host[m - 1] == host[m - 1], "float image: NaN values are not supported");
  for (int i = 0; i < m; ++i) table_out[i] = (double)host[i];
  *n_table = m;
  rank_kernel<T><<<blocks, 256, 0, s>>>(vol, n, clip_t, clipped, m, idx);
  EXA_CUDA(cudaGetLastError());
  EXA_CUDA(cudaStreamSynchronize(s));  // `clipped` (the table) is freed on return
  return Status::OK();
}

// np.percentile(method="linear") on the sorted multiset {values[v] repeated hist[v] times}; numpy
// takes the difference of the two neighbours in the array's own dtype (float32 for float32 images)
// before the float64 interpolation (function_base._lerp), so does this
double percentile_values(const uint64_t* hist, const double* values, int nbins, int is_f32, uint64_t n,
                         double pct) {
  const double q = pct / 100.0;
  const double vidx = (double)(n - 1) * q;
  const double prev = floor(vidx);
  const double gamma = vidx - prev;
  const int64_t last = (int64_t)n - 1;
  int64_t i0 = (int64_t)prev;
  int64_t i1 = i0 + 1;
  if (vidx >= (double)last) i0 = i1 = last;
  if (vidx < 0) i0 = i1 = 0;
  double a = 0, b = 0;
  uint64_t cum = 0;
  bool got_a = false, got_b = false;
  for (int v = 0; v < nbins && !(got_a && got_b); ++v) {
    cum += hist[v];
    if (!got_a && (uint64_t)i0 < cum) {
      a = values[v];
      got_a = true;
    }
    if (!got_b && (uint64_t)i1 < cum) {
      b = values[v];
      got_b = true;
    }
  }
  const double diff = is_f32 ? (double)((float)b - (float)a) : b - a;
  double r = a + diff * gamma;
  if (gamma >= 0.5) r = b - diff * (1.0 - gamma);
  return r;
}

}  // namespace

Status compress_float_volume(const void* vol_dev, int is_double, int64_t n, double clip,
                             uint16_t* idx_dev, double* table_out, int* n_table, cudaStream_t s) {
  EXA_CHECK(vol_dev && idx_dev && table_out && n_table && n > 0, "compress_float_volume: bad arguments");
  if (is_double)
    return compress_impl<double>(static_cast<const double*>(vol_dev), (size_t)n, clip, idx_dev,
                                 table_out, n_table, s);
  return compress_impl<float>(static_cast<const float*>(vol_dev), (size_t)n, clip, idx_dev, table_out,
                              n_table, s);
}

Status percentiles_from_hist_values(const uint64_t* hist, const double* values, int nbins, int is_f32,
                                    double q_lo, double q_hi, double* mn, double* mx) {
  EXA_CHECK(hist && values && nbins > 0 && mn && mx, "percentiles_from_hist_values: bad arguments");
  uint64_t n = 0;
  for (int i = 0; i < nbins; ++i) n += hist[i];
  EXA_CHECK(n > 0, "percentiles_from_hist_values: empty histogram");
  *mn = percentile_values(hist, values, nbins, is_f32, n, q_lo);
  *mx = percentile_values(hist, values, nbins, is_f32, n, q_hi);
  return Status::OK();
}

}  // namespace exa
