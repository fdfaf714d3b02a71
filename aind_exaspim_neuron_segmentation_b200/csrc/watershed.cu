// K7: affinities -> segmentation (reference inference.py:196-237).
//
// The reference hands the float32 (3, D, H, W) affinities to waterz.agglomerate (watershed
// fragments, region graph, hierarchical merging by 1 - mean affinity) and then drops small
// segments (img_util.py:536-559).  Here the voxel-sized steps run on the GPU and only the merge
// queue over the (small) region graph runs on the host:
//
//   K7a ws_best_kernel      strongest incident affinity of every voxel (edges < low removed)
//   K7b ws_union_kernel     keep an edge if it is >= high or the strongest edge of one of its
//                           voxels; lock-free union-find, the root of a fragment is its smallest
//                           voxel index
//   K7c ws_flatten_kernel   root of every voxel + root flags -> (scan) fragment ids 1..n in order
//                           of first appearance
//   K7d ws_count/emit_faces faces between different fragments -> (fragment pair, affinity),
//                           radix sort by pair, segmented sum / count  = the region graph
//   host                    agglomeration: min-heap on (1 - sum/count, a, b, count) with lazy
//                           deletion, statistics of parallel edges added on merge
//   K7e ws_size_kernel      fragment sizes -> segment sizes, small segments dropped, ids in order
//                           of first appearance
//   K7f ws_relabel_kernel   fragment id -> final id
//
// Edge convention (img_util.py:160,207-216): aff[c][z,y,x] is the edge from (z,y,x) to its NEXT
// neighbour along axis c.  Sort, scan and run-length steps use CUB (plumbing, not arithmetic).
#include <cub/cub.cuh>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "watershed.h"
#include "ws_agglomerate.h"

namespace exa {

namespace {

struct DevBuf {  // frees on scope exit: every early return of EXA_CUDA/EXA_TRY stays leak-free
  void* p = nullptr;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  Status alloc(size_t bytes) {
    if (p) cudaFree(p);
    p = nullptr;
    EXA_CUDA(cudaMalloc(&p, bytes ? bytes : 1));
    return Status::OK();
  }
  template <typename T>
  T* as() const {
    return static_cast<T*>(p);
  }
};

struct Vol {
  int D, H, W;
  size_t n, hw;
};

__device__ __forceinline__ void coords(const Vol& g, size_t v, int& z, int& y, int& x) {
  x = (int)(v % (size_t)g.W);
  const size_t r = v / (size_t)g.W;
  y = (int)(r % (size_t)g.H);
  z = (int)(r / (size_t)g.H);
}

// K7a
__global__ void __launch_bounds__(256)
ws_best_kernel(const float* __restrict__ aff, Vol g, float low, float* __restrict__ best) {
  const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= g.n) return;
  int z, y, x;
  coords(g, v, z, y, x);
  float b = 0.f;
  float w;
  if (z + 1 < g.D) { w = aff[v];                 if (w >= low) b = fmaxf(b, w); }
  if (y + 1 < g.H) { w = aff[g.n + v];           if (w >= low) b = fmaxf(b, w); }
  if (x + 1 < g.W) { w = aff[2 * g.n + v];       if (w >= low) b = fmaxf(b, w); }
  if (z > 0)       { w = aff[v - g.hw];          if (w >= low) b = fmaxf(b, w); }
  if (y > 0)       { w = aff[g.n + v - g.W];     if (w >= low) b = fmaxf(b, w); }
  if (x > 0)       { w = aff[2 * g.n + v - 1];   if (w >= low) b = fmaxf(b, w); }
  best[v] = b;
}

__global__ void __launch_bounds__(256)
ws_init_kernel(size_t n, uint32_t* __restrict__ parent, uint8_t* __restrict__ linked) {
  const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  parent[v] = (uint32_t)v;
  linked[v] = 0;
}

__device__ __forceinline__ uint32_t uf_find(uint32_t* parent, uint32_t x) {
  // path halving; plain stores are safe: a non-root only ever gets re-pointed to an ancestor
  uint32_t p = parent[x];
  while (p != x) {
    const uint32_t gp = parent[p];
    if (gp != p) parent[x] = gp;
    x = p;
    p = gp;
  }
  return x;
}

__device__ __forceinline__ void uf_union(uint32_t* parent, uint32_t a, uint32_t b) {
  // the larger root is hooked under the smaller one, so the final root is the minimum index
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a > b) {
      const uint32_t t = a;
      a = b;
      b = t;
    }
    const uint32_t old = atomicCAS(&parent[b], b, a);
    if (old == b) return;
    b = old;
  }
}

// K7b
__global__ void __launch_bounds__(256)
ws_union_kernel(const float* __restrict__ aff, const float* __restrict__ best, Vol g, float low,
                float high, uint32_t* parent, uint8_t* linked) {
  const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= g.n) return;
  int z, y, x;
  coords(g, v, z, y, x);
  const float bv = best[v];
  const bool has[3] = {z + 1 < g.D, y + 1 < g.H, x + 1 < g.W};
  const size_t step[3] = {g.hw, (size_t)g.W, 1};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (!has[c]) continue;
    const float w = aff[(size_t)c * g.n + v];
    const size_t u = v + step[c];
    if (w >= low && (w >= high || w >= bv || w >= best[u])) {
      linked[v] = 1;
      linked[u] = 1;
      uf_union(parent, (uint32_t)v, (uint32_t)u);
    }
  }
}

// K7c: root of every voxel; flag = 1 for the root voxel of every fragment
__global__ void __launch_bounds__(256)
ws_flatten_kernel(size_t n, uint32_t* parent, const uint8_t* __restrict__ linked,
                  uint32_t* __restrict__ root, uint32_t* __restrict__ flag) {
  const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  // the root goes to its own array: parent[] is still being rewritten by the path halving of
  // other threads (always to an ancestor, but not necessarily to the root)
  const uint32_t r = uf_find(parent, (uint32_t)v);
  root[v] = r;
  flag[v] = (linked[v] && r == (uint32_t)v) ? 1u : 0u;
}

// root index -> fragment id, in place
__global__ void __launch_bounds__(256)
ws_assign_kernel(size_t n, const uint8_t* __restrict__ linked, const uint32_t* __restrict__ rank,
                 uint32_t* __restrict__ frag) {
  const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  frag[v] = linked[v] ? rank[frag[v]] + 1u : 0u;
}

// K7d: faces between different (non-background) fragments
__device__ __forceinline__ int face_list(const uint32_t* frag, const float* aff, const Vol& g, size_t v,
                                         unsigned long long* keys, float* vals) {
  int z, y, x;
  coords(g, v, z, y, x);
  const uint32_t a = frag[v];
  if (a == 0) return 0;
  const bool has[3] = {z + 1 < g.D, y + 1 < g.H, x + 1 < g.W};
  const size_t step[3] = {g.hw, (size_t)g.W, 1};
  int m = 0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (!has[c]) continue;
    const uint32_t b = frag[v + step[c]];
    if (b == 0 || b == a) continue;
    if (keys) {
      const uint32_t lo = a < b ? a : b, hi = a < b ? b : a;
      keys[m] = ((unsigned long long)lo << 32) | hi;
      vals[m] = aff[(size_t)c * g.n + v];
    }
    ++m;
  }
  return m;
}

__global__ void __launch_bounds__(256)
ws_count_faces_kernel(const uint32_t* __restrict__ frag, Vol g, uint32_t* __restrict__ cnt) {
  const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= g.n) return;
  cnt[v] = (uint32_t)face_list(frag, nullptr, g, v, nullptr, nullptr);
}

__global__ void __launch_bounds__(256)
ws_emit_faces_kernel(const uint32_t* __restrict__ frag, const float* __restrict__ aff, Vol g,
                     const unsigned long long* __restrict__ offset,
                     unsigned long long* __restrict__ keys, float* __restrict__ vals) {
  const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= g.n) return;
  unsigned long long k[3];
  float w[3];
  const int m = face_list(frag, aff, g, v, k, w);
  const unsigned long long o = offset[v];
  for (int i = 0; i < m; ++i) {
    keys[o + i] = k[i];
    vals[o + i] = w[i];
  }
}

// K7e
__global__ void __launch_bounds__(256)
ws_size_kernel(size_t n, const uint32_t* __restrict__ frag, unsigned long long* __restrict__ sizes) {
  const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  const uint32_t f = frag[v];
  if (f == 0) return;
  // neighbouring voxels mostly share a fragment: one atomic per run inside the warp
  const unsigned mask = __activemask();
  const unsigned same = __match_any_sync(mask, f);
  if ((int)(threadIdx.x & 31) == __ffs(same) - 1) atomicAdd(&sizes[f], (unsigned long long)__popc(same));
}

// K7f
__global__ void __launch_bounds__(256)
ws_relabel_kernel(size_t n, const uint32_t* __restrict__ frag, const uint32_t* __restrict__ lut,
                  uint64_t* __restrict__ seg) {
  const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  seg[v] = (uint64_t)lut[frag[v]];
}

struct ToU64 {
  __host__ __device__ unsigned long long operator()(uint32_t c) const { return c; }
};
struct ToF64 {
  __host__ __device__ double operator()(float w) const { return (double)w; }
};

inline unsigned grid_for(size_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace

Status affinities_to_segmentation_device(const float* aff, int D, int H, int W,
                                         const double* thresholds, int n_thresholds, double aff_low,
                                         double aff_high, int64_t min_segment_size, uint64_t* seg,
                                         int64_t* n_fragments, int64_t* n_segments,
                                         cudaStream_t s) {
  EXA_CHECK(aff && seg, "affinities_to_segmentation: null buffer");
  EXA_CHECK(D > 0 && H > 0 && W > 0, "affinities_to_segmentation: dims must be positive");
  EXA_CHECK(thresholds && n_thresholds > 0, "affinities_to_segmentation: no agglomeration threshold");
  Vol g{D, H, W, (size_t)D * H * W, (size_t)H * W};
  EXA_CHECK(g.n < (1ull << 32) - 1, "affinities_to_segmentation: volume too large for 32-bit voxel ids");
  // thresholds are cumulative: the segmentation of the last one is what the reference keeps
  // (inference.py:232), i.e. merging runs up to the largest
  const double threshold = *std::max_element(thresholds, thresholds + n_thresholds);
  const float low = (float)aff_low, high = (float)aff_high;
  const unsigned blocks = grid_for(g.n);
  // EXA_WS_PROF=1: wall time of the phases on stderr (every phase ends with a stream sync)
  const char* prof_env = getenv("EXA_WS_PROF");
  const bool prof = prof_env && prof_env[0] == '1';
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!prof) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[exa watershed] %-28s %9.3f ms\n", what,
            std::chrono::duration<double, std::milli>(now - t_last).count());
    t_last = now;
  };

  DevBuf best, parent, linked, flag, frag, tmp;
  EXA_TRY(best.alloc(g.n * 4));
  EXA_TRY(parent.alloc(g.n * 4));
  EXA_TRY(linked.alloc(g.n));
  EXA_TRY(flag.alloc(g.n * 4));
  EXA_TRY(frag.alloc(g.n * 4));

  // ---- fragments ----
  ws_best_kernel<<<blocks, 256, 0, s>>>(aff, g, low, best.as<float>());
  ws_init_kernel<<<blocks, 256, 0, s>>>(g.n, parent.as<uint32_t>(), linked.as<uint8_t>());
  ws_union_kernel<<<blocks, 256, 0, s>>>(aff, best.as<float>(), g, low, high, parent.as<uint32_t>(),
                                          linked.as<uint8_t>());
  ws_flatten_kernel<<<blocks, 256, 0, s>>>(g.n, parent.as<uint32_t>(), linked.as<uint8_t>(),
                                            frag.as<uint32_t>(), flag.as<uint32_t>());
  EXA_CUDA(cudaGetLastError());
  uint32_t* rank = best.as<uint32_t>();  // `best` is dead from here on
  size_t tmp_bytes = 0;
  EXA_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, flag.as<uint32_t>(), rank, g.n, s));
  EXA_TRY(tmp.alloc(tmp_bytes));
  EXA_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, flag.as<uint32_t>(), rank, g.n, s));
  uint32_t last_rank = 0, last_flag = 0;
  EXA_CUDA(cudaMemcpyAsync(&last_rank, rank + (g.n - 1), 4, cudaMemcpyDeviceToHost, s));
  EXA_CUDA(cudaMemcpyAsync(&last_flag, flag.as<uint32_t>() + (g.n - 1), 4, cudaMemcpyDeviceToHost, s));
  ws_assign_kernel<<<blocks, 256, 0, s>>>(g.n, linked.as<uint8_t>(), rank, frag.as<uint32_t>());
  EXA_CUDA(cudaGetLastError());
  EXA_CUDA(cudaStreamSynchronize(s));
  const uint32_t n_frag = last_rank + last_flag;
  if (n_fragments) *n_fragments = n_frag;
  lap("fragments (GPU)");

  // ---- region graph ----
  uint32_t* cnt = flag.as<uint32_t>();  // root flags are dead
  ws_count_faces_kernel<<<blocks, 256, 0, s>>>(frag.as<uint32_t>(), g, cnt);
  EXA_CUDA(cudaGetLastError());
  DevBuf offset;
  EXA_TRY(offset.alloc(g.n * 8));
  cub::TransformInputIterator<unsigned long long, ToU64, const uint32_t*> cnt64(cnt, ToU64());
  EXA_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt64, offset.as<unsigned long long>(),
                                         g.n, s));
  EXA_TRY(tmp.alloc(tmp_bytes));
  EXA_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, cnt64, offset.as<unsigned long long>(),
                                         g.n, s));
  unsigned long long last_off = 0;
  uint32_t last_cnt = 0;
  EXA_CUDA(cudaMemcpyAsync(&last_off, offset.as<unsigned long long>() + (g.n - 1), 8,
                           cudaMemcpyDeviceToHost, s));
  EXA_CUDA(cudaMemcpyAsync(&last_cnt, cnt + (g.n - 1), 4, cudaMemcpyDeviceToHost, s));
  EXA_CUDA(cudaStreamSynchronize(s));
  const unsigned long long n_faces = last_off + last_cnt;
  EXA_CHECK(n_faces < (1ull << 31),
            "affinities_to_segmentation: more than 2^31 faces between fragments; split the volume");

  std::vector<unsigned long long> h_keys;
  std::vector<double> h_sums;
  std::vector<int> h_counts;
  if (n_faces > 0) {
    const int m = (int)n_faces;
    DevBuf keys, vals, keys2, vals2, ukeys, usums, ucnts, nruns;
    EXA_TRY(keys.alloc((size_t)m * 8));
    EXA_TRY(vals.alloc((size_t)m * 4));
    EXA_TRY(keys2.alloc((size_t)m * 8));
    EXA_TRY(vals2.alloc((size_t)m * 4));
    ws_emit_faces_kernel<<<blocks, 256, 0, s>>>(frag.as<uint32_t>(), aff, g,
                                                 offset.as<unsigned long long>(),
                                                 keys.as<unsigned long long>(), vals.as<float>());
    EXA_CUDA(cudaGetLastError());
    EXA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.as<unsigned long long>(),
                                             keys2.as<unsigned long long>(), vals.as<float>(),
                                             vals2.as<float>(), m, 0, 64, s));
    EXA_TRY(tmp.alloc(tmp_bytes));
    EXA_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys.as<unsigned long long>(),
                                             keys2.as<unsigned long long>(), vals.as<float>(),
                                             vals2.as<float>(), m, 0, 64, s));
    // segmented sum (fp64) and run lengths; `keys`/`vals` are free again and take the outputs
    EXA_TRY(usums.alloc((size_t)m * 8));
    EXA_TRY(ucnts.alloc((size_t)m * 4));
    EXA_TRY(nruns.alloc(8));
    unsigned long long* ukey = keys.as<unsigned long long>();
    cub::TransformInputIterator<double, ToF64, const float*> w64(vals2.as<float>(), ToF64());
    EXA_CUDA(cub::DeviceReduce::ReduceByKey(nullptr, tmp_bytes, keys2.as<unsigned long long>(), ukey,
                                            w64, usums.as<double>(), nruns.as<int>(), cub::Sum(), m, s));
    EXA_TRY(tmp.alloc(tmp_bytes));
    EXA_CUDA(cub::DeviceReduce::ReduceByKey(tmp.p, tmp_bytes, keys2.as<unsigned long long>(), ukey,
                                            w64, usums.as<double>(), nruns.as<int>(), cub::Sum(), m, s));
    EXA_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, tmp_bytes, keys2.as<unsigned long long>(),
                                                ukey, ucnts.as<int>(), nruns.as<int>() + 1, m, s));
    EXA_TRY(tmp.alloc(tmp_bytes));
    EXA_CUDA(cub::DeviceRunLengthEncode::Encode(tmp.p, tmp_bytes, keys2.as<unsigned long long>(),
                                                ukey, ucnts.as<int>(), nruns.as<int>() + 1, m, s));
    int runs[2] = {0, 0};
    EXA_CUDA(cudaMemcpyAsync(runs, nruns.p, 8, cudaMemcpyDeviceToHost, s));
    EXA_CUDA(cudaStreamSynchronize(s));
    EXA_CHECK(runs[0] == runs[1], "affinities_to_segmentation: region graph run counts disagree");
    h_keys.resize(runs[0]);
    h_sums.resize(runs[0]);
    h_counts.resize(runs[0]);
    EXA_CUDA(cudaMemcpyAsync(h_keys.data(), ukey, (size_t)runs[0] * 8, cudaMemcpyDeviceToHost, s));
    EXA_CUDA(cudaMemcpyAsync(h_sums.data(), usums.p, (size_t)runs[0] * 8, cudaMemcpyDeviceToHost, s));
    EXA_CUDA(cudaMemcpyAsync(h_counts.data(), ucnts.p, (size_t)runs[0] * 4, cudaMemcpyDeviceToHost, s));
    EXA_CUDA(cudaStreamSynchronize(s));
  }

  lap("region graph (GPU)");
  // ---- agglomeration (host) ----
  const std::vector<uint32_t> root = ws::agglomerate(n_frag, h_keys, h_sums, h_counts, threshold);
  if (prof) fprintf(stderr, "[exa watershed] %u fragments, %zu region edges\n", n_frag, h_keys.size());
  lap("merge queue (host)");

  // ---- small segments out, ids in order of first appearance (img_util.py:536-559) ----
  DevBuf sizes;
  EXA_TRY(sizes.alloc(((size_t)n_frag + 1) * 8));
  EXA_CUDA(cudaMemsetAsync(sizes.p, 0, ((size_t)n_frag + 1) * 8, s));
  ws_size_kernel<<<blocks, 256, 0, s>>>(g.n, frag.as<uint32_t>(), sizes.as<unsigned long long>());
  EXA_CUDA(cudaGetLastError());
  std::vector<unsigned long long> frag_size((size_t)n_frag + 1);
  EXA_CUDA(cudaMemcpyAsync(frag_size.data(), sizes.p, ((size_t)n_frag + 1) * 8, cudaMemcpyDeviceToHost, s));
  EXA_CUDA(cudaStreamSynchronize(s));
  std::vector<unsigned long long> seg_size((size_t)n_frag + 1, 0);
  for (uint32_t f = 1; f <= n_frag; ++f) seg_size[root[f]] += frag_size[f];
  // fragment ids ascend with their first voxel, so a segment first appears with its smallest
  // member fragment: walking fragments in id order numbers the kept segments by first appearance
  std::vector<uint32_t> new_id((size_t)n_frag + 1, 0), lut((size_t)n_frag + 1, 0);
  uint32_t next = 0;
  for (uint32_t f = 1; f <= n_frag; ++f) {
    const uint32_t r = root[f];
    if ((long long)seg_size[r] <= (long long)min_segment_size) continue;
    if (new_id[r] == 0) new_id[r] = ++next;
    lut[f] = new_id[r];
  }
  if (n_segments) *n_segments = next;
  uint32_t* lut_dev = parent.as<uint32_t>();  // union-find array is dead; n_frag + 1 <= n
  DevBuf lut_big;
  if ((size_t)n_frag + 1 > g.n) {
    EXA_TRY(lut_big.alloc(((size_t)n_frag + 1) * 4));
    lut_dev = lut_big.as<uint32_t>();
  }
  EXA_CUDA(cudaMemcpyAsync(lut_dev, lut.data(), ((size_t)n_frag + 1) * 4, cudaMemcpyHostToDevice, s));
  ws_relabel_kernel<<<blocks, 256, 0, s>>>(g.n, frag.as<uint32_t>(), lut_dev, seg);
  EXA_CUDA(cudaGetLastError());
  EXA_CUDA(cudaStreamSynchronize(s));  // lut (host vector) and the device buffers go out of scope
  lap("sizes + relabel (GPU)");
  return Status::OK();
}

Status region_agglomerate(uint32_t n_fragments, int64_t n_edges, const uint64_t* pair_keys,
                          const double* sums, const int32_t* counts, double threshold,
                          uint32_t* root_out) {
  EXA_CHECK(n_edges >= 0 && root_out && (n_edges == 0 || (pair_keys && sums && counts)),
            "region_agglomerate: null argument");
  std::vector<unsigned long long> k(pair_keys, pair_keys + n_edges);
  std::vector<double> s(sums, sums + n_edges);
  std::vector<int> c(counts, counts + n_edges);
  for (int64_t i = 0; i < n_edges; ++i) {
    const uint64_t a = k[i] >> 32, b = k[i] & 0xffffffffu;
    EXA_CHECK(a >= 1 && a < b && b <= n_fragments && c[i] > 0, "region_agglomerate: bad edge");
  }
  const std::vector<uint32_t> root = ws::agglomerate(n_fragments, k, s, c, threshold);
  std::copy(root.begin(), root.end(), root_out);
  return Status::OK();
}

Status affinities_to_segmentation_host(int device, const float* aff, int D, int H, int W,
                                       const double* thresholds, int n_thresholds, double aff_low,
                                       double aff_high, int64_t min_segment_size, uint64_t* seg,
                                       int64_t* n_fragments, int64_t* n_segments) {
  EXA_CHECK(aff && seg, "affinities_to_segmentation: null buffer");
  EXA_CHECK(D > 0 && H > 0 && W > 0, "affinities_to_segmentation: dims must be positive");
  EXA_CUDA(cudaSetDevice(device));
  const size_t n = (size_t)D * H * W;
  DevBuf aff_dev, seg_dev;
  EXA_TRY(aff_dev.alloc(n * 12));
  EXA_TRY(seg_dev.alloc(n * 8));
  EXA_CUDA(cudaMemcpy(aff_dev.p, aff, n * 12, cudaMemcpyHostToDevice));
  EXA_TRY(affinities_to_segmentation_device(aff_dev.as<float>(), D, H, W, thresholds, n_thresholds,
                                            aff_low, aff_high, min_segment_size,
                                            seg_dev.as<uint64_t>(), n_fragments, n_segments, nullptr));
  EXA_CUDA(cudaMemcpy(seg, seg_dev.p, n * 8, cudaMemcpyDeviceToHost));
  return Status::OK();
}

}  // namespace exa
