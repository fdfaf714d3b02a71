// K7: affinities -> segmentation (reference inference.py:196-237).
//
// The reference hands the float32 (3, D, H, W) affinities to waterz.agglomerate (watershed
// fragments, region graph, hierarchical merging by 1 - mean affinity) and then drops small
// segments (img_util.py:536-559).  waterz conventions (restated in oracle/ws_ref.cpp, which is the
// checker for everything in this file): aff[c][z,y,x] is the edge between voxel (z,y,x) and its
// PREVIOUS neighbour along axis c; a voxel takes part only if its largest incident affinity is
// > low; it points along every edge that equals that maximum or is >= high; plateaus are divided
// breadth first from their exits; basins are numbered in order of first appearance.
//
//   K7a ws_dirs_kernel      direction bits of every voxel
//   K7b ws_corner_kernel    plateau corners (a set direction the neighbour does not return) ->
//                           level 0 of the division, ranked in scan order
//       ws_claim/count/emit breadth-first levels of the plateau division.  waterz's queue is
//                           first-in-first-out, so a voxel's position in it is (level, position of
//                           its first parent, direction from that parent): every level is ranked
//                           with one atomicMin pass and one scan -- no sequential queue
//   K7c ws_point_kernel     the ONE edge every divided voxel keeps (the last direction, in
//                           waterz's order, that leads to an exit or to a voxel earlier in the
//                           queue); undivided plateaus keep all their edges; lock-free union-find
//                           over the kept edges, the root of a basin is its smallest voxel index
//       ws_flatten/assign   basin ids 1..n in order of first appearance
//   K7d ws_count/emit_faces faces between different fragments -> (fragment pair, affinity in
//                           32.32 fixed point); radix sort by pair; segmented sum and count
//                           = the region graph, exact and independent of summation order
//   K7e agg_* kernels       agglomeration, parallel rounds: every region finds its best edge
//                           (smallest score, ties by original edge rank); an edge is merged when
//                           it is the best edge of both ends, or the best edge of one end X whose
//                           other neighbours all have a later best edge (nothing can then reach X
//                           or change that edge before the sequential queue would merge it).
//                           Mean-affinity linkage is reducible, so these merges give exactly the
//                           partition of waterz's one-at-a-time priority queue.  Parallel edges are
//                           combined through a hash table (integer adds).  When a round merges
//                           too little -- one hub swallowing its neighbours one by one -- the
//                           contracted graph goes to the exact host queue (ws_agglomerate.h).
//   K7f ws_size/.../relabel fragment sizes -> segment sizes, small segments dropped, ids in
//                           order of first appearance, fragment id -> final id
//
// Sort, scan and run-length steps use CUB (plumbing, not arithmetic).
#include <cub/cub.cuh>

#include <algorithm>
#include <chrono>
#include <map>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "watershed.h"
#include "ws_agglomerate.h"

namespace exa {

namespace {

constexpr uint32_t kNone = 0xffffffffu;

double g_alloc_ms = 0.0;  // host time inside the allocator calls (EXA_WS_PROF)
// wall milliseconds of the phases of the last affinities_to_segmentation call of this process
// (exa_ws_last_profile): fragments, region graph, parallel rounds, host queue, sizes + relabel,
// then counts: parallel rounds, region-graph edges, edges handed to the host queue
double g_last_profile[8] = {0, 0, 0, 0, 0, 0, 0, 0};
struct AllocTimer {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  ~AllocTimer() {
    g_alloc_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
};

// Device memory comes from a block cache of our own: cudaMalloc'd blocks in size classes (four per
// octave above 1 MiB) that go back to a per-device free list instead of to the driver, so after
// the first call at a given size the ~30 allocations of a call cost microseconds and nothing
// depends on the driver's pools.  (Measured with stream-ordered pools, default or our own with the
// release threshold at its maximum: now and then 0.1-0.9 s of allocator time per call at
// 256^3-512^3 -- the pool re-maps fragmented memory.)  Everything of one call is queued on ONE
// stream, so a block may be handed out again while work that used it is still queued: the later
// work runs after it.  Every call ends with that stream synchronised (also on errors), so blocks
// are idle when another call, possibly on another stream, takes them.  exa_ws_release_memory()
// gives the cached blocks back to the driver; an allocation that fails does so first and retries.
struct BlockCache {
  std::mutex mu;
  std::multimap<size_t, void*> free_blocks;
};
BlockCache g_cache[64];

size_t size_class(size_t bytes) {
  if (bytes <= ((size_t)1 << 20)) return (bytes + 511) / 512 * 512 + 512;
  size_t step = (size_t)1 << 18;
  while (step * 8 <= bytes) step <<= 1;  // step = 2^(floor(log2 bytes) - 2)
  return (bytes + step - 1) / step * step;
}

void cache_release_all(int dev) {
  BlockCache& c = g_cache[dev];
  std::lock_guard<std::mutex> lock(c.mu);
  for (auto& kv : c.free_blocks) cudaFree(kv.second);
  c.free_blocks.clear();
}

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int dev = -1;
  explicit DevBuf(cudaStream_t = nullptr) {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (!p) return;
    BlockCache& c = g_cache[dev];
    std::lock_guard<std::mutex> lock(c.mu);
    c.free_blocks.emplace(bytes, p);
    p = nullptr;
  }
  Status alloc(size_t want) {
    release();
    AllocTimer t;
    EXA_CUDA(cudaGetDevice(&dev));
    EXA_CHECK(dev >= 0 && dev < 64, "affinities_to_segmentation: device index out of range");
    bytes = size_class(want);
    {
      BlockCache& c = g_cache[dev];
      std::lock_guard<std::mutex> lock(c.mu);
      auto it = c.free_blocks.find(bytes);
      if (it != c.free_blocks.end()) {
        p = it->second;
        c.free_blocks.erase(it);
        return Status::OK();
      }
    }
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaErrorMemoryAllocation) {
      cudaGetLastError();
      cache_release_all(dev);
      e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) {
      p = nullptr;
      return Status::Err(std::string("affinities_to_segmentation: cudaMalloc of ") +
                         std::to_string(bytes >> 20) + " MiB failed: " + cudaGetErrorString(e));
    }
    return Status::OK();
  }
  void swap(DevBuf& o) {  // the block with its size: it returns to the cache under that size
    std::swap(p, o.p);
    std::swap(bytes, o.bytes);
    std::swap(dev, o.dev);
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

// direction d = 0..5: -z, -y, -x, +z, +y, +x (waterz's order); the opposite of d is (d + 3) % 6
__device__ __forceinline__ long long dir_step(const Vol& g, int d) {
  const long long s = d % 3 == 0 ? (long long)g.hw : d % 3 == 1 ? (long long)g.W : 1ll;
  return d < 3 ? -s : s;
}
__device__ __forceinline__ uint32_t opposite_bit(int d) { return 1u << ((d + 3) % 6); }

// K7a
__global__ void __launch_bounds__(256)
ws_dirs_kernel(const float* __restrict__ aff, Vol g, float low, float high,
               uint8_t* __restrict__ bits) {
  const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= g.n) return;
  int z, y, x;
  coords(g, v, z, y, x);
  float w[6];
  w[0] = z > 0 ? aff[v] : low;
  w[1] = y > 0 ? aff[g.n + v] : low;
  w[2] = x > 0 ? aff[2 * g.n + v] : low;
  w[3] = z + 1 < g.D ? aff[v + g.hw] : low;
  w[4] = y + 1 < g.H ? aff[g.n + v + g.W] : low;
  w[5] = x + 1 < g.W ? aff[2 * g.n + v + 1] : low;
  float m = w[0];
#pragma unroll
  for (int d = 1; d < 6; ++d) m = w[d] > m ? w[d] : m;
  uint32_t b = 0;
  if (m > low) {
#pragma unroll
    for (int d = 0; d < 6; ++d)
      if (w[d] == m || w[d] >= high) b |= 1u << d;
  }
  bits[v] = (uint8_t)b;
}

// K7b: corner = a set direction whose neighbour does not point back
__global__ void __launch_bounds__(256)
ws_corner_kernel(const uint8_t* __restrict__ bits, Vol g, uint32_t* __restrict__ flag) {
  const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= g.n) return;
  const uint32_t b = bits[v];
  uint32_t corner = 0;
#pragma unroll
  for (int d = 0; d < 6; ++d)
    if ((b >> d) & 1u) {
      if (!(bits[(long long)v + dir_step(g, d)] & opposite_bit(d))) corner = 1;
    }
  flag[v] = corner;
}

// level 0 of the division: corners in scan order
__global__ void __launch_bounds__(256)
ws_level0_kernel(size_t n, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ rank,
                 uint32_t* __restrict__ pos, uint32_t* __restrict__ front) {
  const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  if (flag[v]) {
    pos[v] = rank[v];
    front[rank[v]] = (uint32_t)v;
  } else {
    pos[v] = kNone;
  }
}

__global__ void __launch_bounds__(256) fill_u32_kernel(size_t n, uint32_t* __restrict__ a, uint32_t val) {
  const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) a[v] = val;
}
__global__ void __launch_bounds__(256) iota_u32_kernel(size_t n, uint32_t* __restrict__ a) {
  const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) a[v] = (uint32_t)v;
}

// is the voxel in direction d of front voxel p a child (plateau neighbour not yet in the queue)?
__device__ __forceinline__ bool plateau_child(const uint8_t* bits, const uint32_t* pos, const Vol& g,
                                              uint32_t p, uint32_t bp, int d, uint32_t* child) {
  if (!((bp >> d) & 1u)) return false;
  const uint32_t j = (uint32_t)((long long)p + dir_step(g, d));
  if (!(bits[j] & opposite_bit(d))) return false;
  if (pos[j] != kNone) return false;
  *child = j;
  return true;
}

// every unreached plateau neighbour of the front is claimed by its first parent
__global__ void __launch_bounds__(256)
ws_claim_kernel(const uint32_t* __restrict__ front, uint32_t n_front, const uint8_t* __restrict__ bits,
                const uint32_t* __restrict__ pos, Vol g, uint32_t* claim, uint32_t* n_claims) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_front) return;
  const uint32_t p = front[r];
  const uint32_t bp = bits[p];
  uint32_t any = 0;
#pragma unroll
  for (int d = 0; d < 6; ++d) {
    uint32_t j;
    if (plateau_child(bits, pos, g, p, bp, d, &j)) {
      atomicMin(&claim[j], r);
      any = 1;
    }
  }
  if (any) atomicAdd(n_claims, 1u);
}

__global__ void __launch_bounds__(256)
ws_count_children_kernel(const uint32_t* __restrict__ front, uint32_t n_front,
                         const uint8_t* __restrict__ bits, const uint32_t* __restrict__ pos, Vol g,
                         const uint32_t* __restrict__ claim, uint32_t* __restrict__ cnt) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_front) return;
  const uint32_t p = front[r];
  const uint32_t bp = bits[p];
  uint32_t c = 0;
#pragma unroll
  for (int d = 0; d < 6; ++d) {
    uint32_t j;
    if (plateau_child(bits, pos, g, p, bp, d, &j) && claim[j] == r) ++c;
  }
  cnt[r] = c;
}

// children in (parent position, direction) order = waterz's queue order; only the one parent
// that won the claim writes pos[child]
__global__ void __launch_bounds__(256)
ws_emit_children_kernel(const uint32_t* __restrict__ front, uint32_t n_front,
                        const uint8_t* __restrict__ bits, uint32_t* pos, Vol g,
                        const uint32_t* __restrict__ claim, const uint32_t* __restrict__ off,
                        uint32_t base, uint32_t* __restrict__ next) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_front) return;
  const uint32_t p = front[r];
  const uint32_t bp = bits[p];
  uint32_t o = off[r];
#pragma unroll
  for (int d = 0; d < 6; ++d) {
    uint32_t j;
    if (plateau_child(bits, pos, g, p, bp, d, &j) && claim[j] == r) {
      next[o] = j;
      pos[j] = base + o;
      ++o;
    }
  }
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

// K7c
__global__ void __launch_bounds__(256)
ws_point_kernel(const uint8_t* __restrict__ bits, const uint32_t* __restrict__ pos, Vol g,
                uint32_t* parent) {
  const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= g.n) return;
  const uint32_t b = bits[v];
  if (b == 0) return;
  const uint32_t pv = pos[v];
  if (pv == kNone) {
    // plateau without an exit: all its edges are returned; the +directions cover every edge once
#pragma unroll
    for (int d = 3; d < 6; ++d)
      if ((b >> d) & 1u) uf_union(parent, (uint32_t)v, (uint32_t)((long long)v + dir_step(g, d)));
    return;
  }
  int keep = -1;
#pragma unroll
  for (int d = 0; d < 6; ++d)
    if ((b >> d) & 1u) {
      const uint32_t j = (uint32_t)((long long)v + dir_step(g, d));
      if (!(bits[j] & opposite_bit(d)) || pos[j] < pv) keep = d;
    }
  if (keep >= 0) uf_union(parent, (uint32_t)v, (uint32_t)((long long)v + dir_step(g, keep)));
}

// root of every voxel; flag = 1 for the root voxel of every fragment
__global__ void __launch_bounds__(256)
ws_flatten_kernel(size_t n, uint32_t* parent, const uint8_t* __restrict__ bits,
                  uint32_t* __restrict__ root, uint32_t* __restrict__ flag) {
  const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  // the root goes to its own array: parent[] is still being rewritten by the path halving of
  // other threads (always to an ancestor, but not necessarily to the root)
  const uint32_t r = uf_find(parent, (uint32_t)v);
  root[v] = r;
  flag[v] = (bits[v] && r == (uint32_t)v) ? 1u : 0u;
}

// root index -> fragment id, in place
__global__ void __launch_bounds__(256)
ws_assign_kernel(size_t n, const uint8_t* __restrict__ bits, const uint32_t* __restrict__ rank,
                 uint32_t* __restrict__ frag) {
  const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  frag[v] = bits[v] ? rank[frag[v]] + 1u : 0u;
}

// K7d: faces between different (non-background) fragments; edge c of voxel v joins v and v - step
__device__ __forceinline__ unsigned long long to_fixed(float a) {
  if (!(a > 0.0f)) return 0ull;  // also NaN
  if (a > 1.0f) a = 1.0f;
  return (unsigned long long)llrint((double)a * 4294967296.0);
}

__device__ __forceinline__ int face_list(const uint32_t* frag, const float* aff, const Vol& g, size_t v,
                                         unsigned long long* keys, unsigned long long* vals) {
  const uint32_t a = frag[v];
  if (a == 0) return 0;
  int z, y, x;
  coords(g, v, z, y, x);
  const bool has[3] = {z > 0, y > 0, x > 0};
  const size_t step[3] = {g.hw, (size_t)g.W, 1};
  int m = 0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (!has[c]) continue;
    const uint32_t b = frag[v - step[c]];
    if (b == 0 || b == a) continue;
    if (keys) {
      const uint32_t lo = a < b ? a : b, hi = a < b ? b : a;
      keys[m] = ((unsigned long long)lo << 32) | hi;
      vals[m] = to_fixed(aff[(size_t)c * g.n + v]);
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
                     unsigned long long* __restrict__ keys, unsigned long long* __restrict__ vals) {
  const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= g.n) return;
  unsigned long long k[3], w[3];
  const int m = face_list(frag, aff, g, v, k, w);
  const unsigned long long o = offset[v];
  for (int i = 0; i < m; ++i) {
    keys[o + i] = k[i];
    vals[o + i] = w[i];
  }
}

__global__ void __launch_bounds__(256)
ws_split_keys_kernel(uint32_t m, const unsigned long long* __restrict__ keys, uint32_t* __restrict__ eu,
                     uint32_t* __restrict__ ev, uint32_t* __restrict__ ek) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  eu[e] = (uint32_t)(keys[e] >> 32);
  ev[e] = (uint32_t)(keys[e] & 0xffffffffu);
  ek[e] = e;
}

// ---- K7e: agglomeration rounds ------------------------------------------------------------
// Edge e: regions eu[e] < ev[e] (eu = 0: dead slot), statistics eq (32.32 fixed-point sum), ec
// (faces), ek (smallest original rank).  T as in ws_agglomerate.h (0 <= T <= 2^32 here).
struct Edges {
  uint32_t* u;
  uint32_t* v;
  unsigned long long* q;
  uint32_t* c;
  uint32_t* k;
};

__device__ __forceinline__ bool edge_before(const Edges& E, uint32_t e1, uint32_t e2) {
  const unsigned long long q1 = E.q[e1], q2 = E.q[e2];
  const unsigned long long c1 = E.c[e1], c2 = E.c[e2];
  const unsigned long long h1 = __umul64hi(q1, c2), l1 = q1 * c2;
  const unsigned long long h2 = __umul64hi(q2, c1), l2 = q2 * c1;
  if (h1 != h2) return h1 > h2;
  if (l1 != l2) return l1 > l2;
  return E.k[e1] < E.k[e2];
}

__device__ __forceinline__ void best_update(uint32_t* best, uint32_t node, uint32_t e, const Edges& E) {
  uint32_t cur = *((volatile uint32_t*)&best[node]);
  while (true) {
    if (cur != kNone && !edge_before(E, e, cur)) return;
    const uint32_t old = atomicCAS(&best[node], cur, e);
    if (old == cur) return;
    cur = old;
  }
}

__global__ void __launch_bounds__(256)
agg_best_kernel(uint32_t m, Edges E, unsigned long long T, int all_below, uint32_t* best) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  const uint32_t u = E.u[e];
  if (u == 0) return;
  if (!all_below && !(E.q[e] > (unsigned long long)E.c[e] * T)) return;
  best_update(best, u, e, E);
  best_update(best, E.v[e], e, E);
}

// X is blocked when a neighbour other than the end of its best edge has a best edge that is not
// later than X's
__global__ void __launch_bounds__(256)
agg_block_kernel(uint32_t m, Edges E, const uint32_t* __restrict__ best, uint8_t* __restrict__ blocked) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  const uint32_t u = E.u[e];
  if (u == 0) return;
  const uint32_t v = E.v[e];
  const uint32_t bu = best[u], bv = best[v];
  if (bu != kNone && bu != e && bv != kNone && !edge_before(E, bu, bv)) blocked[u] = 1;
  if (bv != kNone && bv != e && bu != kNone && !edge_before(E, bv, bu)) blocked[v] = 1;
}

__global__ void __launch_bounds__(256)
agg_merge_kernel(uint32_t m, Edges E, const uint32_t* __restrict__ best,
                 const uint8_t* __restrict__ blocked, uint32_t* parent, uint32_t* __restrict__ stamp,
                 uint32_t round, uint32_t* counters) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  bool merge = false;
  if (e < m) {
    const uint32_t u = E.u[e];
    if (u != 0) {
      const uint32_t v = E.v[e];
      const bool bu = best[u] == e, bv = best[v] == e;
      merge = (bu && bv) || (bu && !blocked[u]) || (bv && !blocked[v]);
      if (merge) {
        uf_union(parent, u, v);
        stamp[u] = round;
        stamp[v] = round;
        E.u[e] = 0;
      }
    }
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, merge);
  if ((threadIdx.x & 31) == 0 && ballot) {
    atomicAdd(&counters[0], (uint32_t)__popc(ballot));   // merges of this round
    atomicAdd(&counters[1], (uint32_t)__popc(ballot));   // dead edge slots
  }
}

__device__ __forceinline__ uint32_t uf_find_ro(const uint32_t* parent, uint32_t x) {
  uint32_t p = parent[x];
  while (p != x) {
    x = p;
    p = parent[x];
  }
  return x;
}

__device__ __forceinline__ uint32_t hash_pair(unsigned long long key, uint32_t mask) {
  key ^= key >> 33;
  key *= 0xff51afd7ed558ccdull;
  key ^= key >> 33;
  key *= 0xc4ceb9fe1a85ec53ull;
  key ^= key >> 33;
  return (uint32_t)key & mask;
}

// edges of regions that can never merge again die; edges that touch a region merged in this round
// are renamed and entered into the hash table: the smallest edge index of a pair owns the pair
__global__ void __launch_bounds__(256)
agg_rename_kernel(uint32_t m, Edges E, const uint32_t* __restrict__ best,
                  const uint32_t* __restrict__ parent, const uint32_t* __restrict__ stamp,
                  uint32_t round, unsigned long long* table_key, uint32_t* table_owner,
                  uint32_t mask, uint32_t* __restrict__ slot_of, uint32_t* counters) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  bool died = false;
  if (e < m) {
    const uint32_t u = E.u[e];
    if (u != 0) {
      const uint32_t v = E.v[e];
      uint32_t slot = kNone;
      if (best[u] == kNone || best[v] == kNone) {
        E.u[e] = 0;
        died = true;
      } else if (stamp[u] == round || stamp[v] == round) {
        const uint32_t ru = uf_find_ro(parent, u), rv = uf_find_ro(parent, v);
        if (ru == rv) {
          E.u[e] = 0;
          died = true;
        } else {
          const uint32_t lo = ru < rv ? ru : rv, hi = ru < rv ? rv : ru;
          const unsigned long long key = ((unsigned long long)lo << 32) | hi;
          slot = hash_pair(key, mask);
          while (true) {
            const unsigned long long old = atomicCAS(&table_key[slot], 0ull, key);
            if (old == 0ull || old == key) break;
            slot = (slot + 1) & mask;
          }
          atomicMin(&table_owner[slot], e);
          E.u[e] = lo;
          E.v[e] = hi;
        }
      }
      slot_of[e] = slot;
    }
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, died);
  if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(&counters[1], (uint32_t)__popc(ballot));
}

// statistics of parallel edges add up in the owner; the others die
__global__ void __launch_bounds__(256)
agg_combine_kernel(uint32_t m, Edges E, const uint32_t* __restrict__ table_owner,
                   const uint32_t* __restrict__ slot_of, uint32_t* counters) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  bool died = false;
  if (e < m && E.u[e] != 0) {
    const uint32_t slot = slot_of[e];
    if (slot != kNone) {
      const uint32_t o = table_owner[slot];
      if (o != e) {
        atomicAdd(&E.q[o], E.q[e]);
        atomicAdd(&E.c[o], E.c[e]);
        atomicMin(&E.k[o], E.k[e]);
        E.u[e] = 0;
        died = true;
      }
    }
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, died);
  if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(&counters[1], (uint32_t)__popc(ballot));
}

__global__ void __launch_bounds__(256)
agg_alive_kernel(uint32_t m, const uint32_t* __restrict__ eu, uint32_t* __restrict__ flag) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < m) flag[e] = eu[e] != 0 ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
agg_compact_kernel(uint32_t m, Edges src, const uint32_t* __restrict__ off, Edges dst) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m || src.u[e] == 0) return;
  const uint32_t o = off[e];
  dst.u[o] = src.u[e];
  dst.v[o] = src.v[e];
  dst.q[o] = src.q[e];
  dst.c[o] = src.c[e];
  dst.k[o] = src.k[e];
}

// hand-over to the host queue: regions that still have edges get dense ids (ascending with the
// region id) on the GPU
__global__ void __launch_bounds__(256)
agg_mark_nodes_kernel(uint32_t m, const uint32_t* __restrict__ eu, const uint32_t* __restrict__ ev,
                      uint32_t* __restrict__ used) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m || eu[e] == 0) return;
  used[eu[e]] = 1;
  used[ev[e]] = 1;
}

__global__ void __launch_bounds__(256)
agg_dense_kernel(uint32_t m, uint32_t* __restrict__ eu, uint32_t* __restrict__ ev,
                 const uint32_t* __restrict__ dense) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  eu[e] = dense[eu[e]] + 1u;
  ev[e] = dense[ev[e]] + 1u;
}

__global__ void __launch_bounds__(256)
agg_dense_ids_kernel(uint32_t n, const uint32_t* __restrict__ used, const uint32_t* __restrict__ dense,
                     uint32_t* __restrict__ ids) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < n && used[f]) ids[dense[f]] = f;
}

__global__ void __launch_bounds__(256)
agg_set_parent_kernel(uint32_t n, const uint32_t* __restrict__ node, const uint32_t* __restrict__ to,
                      uint32_t* __restrict__ parent) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) parent[node[i]] = to[i];
}

__global__ void __launch_bounds__(256)
agg_roots_kernel(uint32_t n, const uint32_t* __restrict__ parent, uint32_t* __restrict__ root) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < n) root[f] = uf_find_ro(parent, f);
}

// ---- K7f ------------------------------------------------------------------------------------
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

__global__ void __launch_bounds__(256)
ws_segment_size_kernel(uint32_t n, const uint32_t* __restrict__ root,
                       const unsigned long long* __restrict__ frag_size,
                       unsigned long long* __restrict__ seg_size) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f == 0 || f >= n) return;
  atomicAdd(&seg_size[root[f]], frag_size[f]);
}

__global__ void __launch_bounds__(256)
ws_keep_kernel(uint32_t n, const uint32_t* __restrict__ root,
               const unsigned long long* __restrict__ seg_size, long long min_size,
               uint32_t* __restrict__ keep) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  keep[f] = (f != 0 && root[f] == f && (long long)seg_size[f] > min_size) ? 1u : 0u;
}

// a segment's root is its smallest fragment and fragment ids ascend with their first voxel, so
// numbering the kept roots in id order numbers the segments by first appearance
__global__ void __launch_bounds__(256)
ws_lut_kernel(uint32_t n, const uint32_t* __restrict__ root, const uint32_t* __restrict__ keep,
              const uint32_t* __restrict__ rank, uint32_t* __restrict__ lut) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  const uint32_t r = root[f];
  lut[f] = (f != 0 && keep[r]) ? rank[r] + 1u : 0u;
}

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

inline unsigned grid_for(size_t n) { return (unsigned)((n + 255) / 256); }

Status exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, DevBuf& tmp, cudaStream_t s) {
  size_t bytes = 0;
  EXA_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, s));
  EXA_TRY(tmp.alloc(bytes));
  EXA_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, n, s));
  return Status::OK();
}

// sum of the last input and the last output of an exclusive scan = the total
Status scan_total_u32(const uint32_t* in, const uint32_t* out, size_t n, uint32_t* total, cudaStream_t s) {
  uint32_t a = 0, b = 0;
  EXA_CUDA(cudaMemcpyAsync(&a, in + (n - 1), 4, cudaMemcpyDeviceToHost, s));
  EXA_CUDA(cudaMemcpyAsync(&b, out + (n - 1), 4, cudaMemcpyDeviceToHost, s));
  EXA_CUDA(cudaStreamSynchronize(s));
  *total = a + b;
  return Status::OK();
}

struct EdgeStore {
  DevBuf u, v, q, c, k;
  explicit EdgeStore(cudaStream_t s) : u(s), v(s), q(s), c(s), k(s) {}
  Status alloc(size_t m) {
    EXA_TRY(u.alloc(m * 4));
    EXA_TRY(v.alloc(m * 4));
    EXA_TRY(q.alloc(m * 8));
    EXA_TRY(c.alloc(m * 4));
    EXA_TRY(k.alloc(m * 4));
    return Status::OK();
  }
  Edges view() const {
    return Edges{u.as<uint32_t>(), v.as<uint32_t>(), q.as<unsigned long long>(), c.as<uint32_t>(),
                 k.as<uint32_t>()};
  }
};

int env_int(const char* name, int fallback) {
  const char* v = getenv(name);
  return v && v[0] ? atoi(v) : fallback;
}

// Agglomeration of the region graph in `es` (m0 edge slots, regions 1..n_frag) up to threshold T.
// root (device, n_frag + 1): the smallest fragment id of the region every fragment ends up in.
//   EXA_WS_GPU_ROUNDS   upper bound on the parallel rounds (default 1 << 30; 0 = host queue only)
//   EXA_WS_HOST_TAIL    1 (default): hand over to the host queue once at most EXA_WS_TAIL_EDGES
//                       (default 32768) edges are alive, or when a round merges fewer than
//                       8 + live / 40000 edges (the host queue is then the cheaper way to finish);
//                       0: parallel rounds to the end
Status agglomerate_rounds(EdgeStore& es, uint32_t m0, uint32_t n_frag, int64_t T, uint32_t* root,
                          cudaStream_t s, bool prof, int* rounds_out, uint32_t* tail_edges_out) {
  const uint32_t n_nodes = n_frag + 1;
  DevBuf parent(s), best(s), blocked(s), stamp(s), counters(s), slot_of(s), tkey(s), towner(s), tmp(s),
      flag(s), off(s);
  EXA_TRY(parent.alloc((size_t)n_nodes * 4));
  iota_u32_kernel<<<grid_for(n_nodes), 256, 0, s>>>(n_nodes, parent.as<uint32_t>());
  EXA_CUDA(cudaGetLastError());
  const int max_rounds = env_int("EXA_WS_GPU_ROUNDS", 1 << 30);
  const bool host_tail = env_int("EXA_WS_HOST_TAIL", 1) != 0;
  const uint32_t tail_edges = (uint32_t)std::max(0, env_int("EXA_WS_TAIL_EDGES", 32768));
  const bool none_below = T > ((int64_t)1 << 32);
  const int all_below = T < 0 ? 1 : 0;
  const unsigned long long Tu = T < 0 ? 0ull : (unsigned long long)T;
  uint32_t m = m0, live = m0;
  int rounds = 0;
  bool finished = none_below || m0 == 0;
  EdgeStore spare(s);
  if (!finished && max_rounds > 0) {
    EXA_TRY(best.alloc((size_t)n_nodes * 4));
    EXA_TRY(blocked.alloc(n_nodes));
    EXA_TRY(stamp.alloc((size_t)n_nodes * 4));
    EXA_TRY(counters.alloc(8));
    EXA_TRY(slot_of.alloc((size_t)m * 4));
    EXA_CUDA(cudaMemsetAsync(stamp.p, 0, (size_t)n_nodes * 4, s));
    EXA_CUDA(cudaMemsetAsync(counters.p, 0, 8, s));
    uint32_t table_slots = 0;
    auto compact = [&]() -> Status {
      EXA_TRY(flag.alloc((size_t)m * 4));
      EXA_TRY(off.alloc((size_t)m * 4));
      const unsigned gb = grid_for(m);
      Edges E = es.view();
      agg_alive_kernel<<<gb, 256, 0, s>>>(m, E.u, flag.as<uint32_t>());
      EXA_CUDA(cudaGetLastError());
      EXA_TRY(exclusive_scan_u32(flag.as<uint32_t>(), off.as<uint32_t>(), m, tmp, s));
      EXA_TRY(spare.alloc(live));
      agg_compact_kernel<<<gb, 256, 0, s>>>(m, E, off.as<uint32_t>(), spare.view());
      EXA_CUDA(cudaGetLastError());
      es.u.swap(spare.u);
      es.v.swap(spare.v);
      es.q.swap(spare.q);
      es.c.swap(spare.c);
      es.k.swap(spare.k);
      m = live;
      return Status::OK();
    };
    // pinned: {merges of the round, dead slots in total}.  Allocated once per process:
    // cudaMallocHost / cudaFreeHost per call cost up to hundreds of milliseconds now and then
    // (they synchronise with the whole device)
    static thread_local uint32_t* h = nullptr;
    if (!h) EXA_CUDA(cudaMallocHost(&h, 8));
    while (rounds < max_rounds) {
      ++rounds;
      Edges E = es.view();
      const unsigned gb = grid_for(m);
      EXA_CUDA(cudaMemsetAsync(best.p, 0xff, (size_t)n_nodes * 4, s));
      EXA_CUDA(cudaMemsetAsync(blocked.p, 0, n_nodes, s));
      EXA_CUDA(cudaMemsetAsync(counters.p, 0, 4, s));
      agg_best_kernel<<<gb, 256, 0, s>>>(m, E, Tu, all_below, best.as<uint32_t>());
      agg_block_kernel<<<gb, 256, 0, s>>>(m, E, best.as<uint32_t>(), blocked.as<uint8_t>());
      agg_merge_kernel<<<gb, 256, 0, s>>>(m, E, best.as<uint32_t>(), blocked.as<uint8_t>(),
                                           parent.as<uint32_t>(), stamp.as<uint32_t>(), (uint32_t)rounds,
                                           counters.as<uint32_t>());
      // rename + combine; the table holds at most `live` pairs (one host round trip per round:
      // the counters are read after the whole round was queued)
      uint32_t want = 1024;
      while ((unsigned long long)want < 2ull * live && want < (1u << 31)) want <<= 1;
      if (want > table_slots) {
        EXA_TRY(tkey.alloc((size_t)want * 8));
        EXA_TRY(towner.alloc((size_t)want * 4));
        table_slots = want;
      }
      EXA_CUDA(cudaMemsetAsync(tkey.p, 0, (size_t)want * 8, s));
      EXA_CUDA(cudaMemsetAsync(towner.p, 0xff, (size_t)want * 4, s));
      agg_rename_kernel<<<gb, 256, 0, s>>>(m, E, best.as<uint32_t>(), parent.as<uint32_t>(),
                                            stamp.as<uint32_t>(), (uint32_t)rounds,
                                            tkey.as<unsigned long long>(), towner.as<uint32_t>(),
                                            want - 1, slot_of.as<uint32_t>(), counters.as<uint32_t>());
      agg_combine_kernel<<<gb, 256, 0, s>>>(m, E, towner.as<uint32_t>(), slot_of.as<uint32_t>(),
                                             counters.as<uint32_t>());
      EXA_CUDA(cudaGetLastError());
      EXA_CUDA(cudaMemcpyAsync(h, counters.p, 8, cudaMemcpyDeviceToHost, s));
      EXA_CUDA(cudaStreamSynchronize(s));
      const uint32_t merges = h[0];
      live = m0 - h[1];
      if (merges == 0 || live == 0) {
        finished = true;  // no edge below the threshold is left
        break;
      }
      if (prof && (rounds <= 4 || (rounds & (rounds - 1)) == 0))
        fprintf(stderr, "[exa watershed]   round %d: %u merges, %u live edges\n", rounds, merges, live);
      // compact the slots when more than half are dead
      if ((m - live) * 2u > m && m > 4096) EXA_TRY(compact());
      // hand-over to the host queue when it is the cheaper way to finish.  Measured on a B200 box
      // (512^3 smooth field, 8-9 M live edges): a round costs about 50 us + 0.19 ns per live edge
      // here, a merge about 6 us there (hub merges move long neighbour lists), so the rounds pay
      // while they merge more than ~8 + live / 40000 edges each; they end in hub chains (one large
      // region swallowing its neighbours a few per round)
      if (host_tail && (live <= tail_edges || (rounds >= 64 && merges < 8u + live / 40000u))) break;
    }
    if (!finished && m != live) EXA_TRY(compact());
  }
  if (rounds_out) *rounds_out = rounds;
  if (tail_edges_out) *tail_edges_out = 0;

  auto t_phase = std::chrono::steady_clock::now();
  if (prof) fprintf(stderr, "[exa watershed]   %d parallel rounds, %u live edges\n", rounds, live);
  auto sublap = [&](const char* what) {
    if (!prof) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[exa watershed]   %-26s %9.3f ms\n", what,
            std::chrono::duration<double, std::milli>(now - t_phase).count());
    t_phase = now;
  };
  if (!finished) {
    // ---- exact host queue on what is left (all m slots are alive here) ----
    Edges E = es.view();
    DevBuf used(s), dense(s), ids_dev(s);
    EXA_TRY(used.alloc((size_t)n_nodes * 4));
    EXA_TRY(dense.alloc((size_t)n_nodes * 4));
    EXA_CUDA(cudaMemsetAsync(used.p, 0, (size_t)n_nodes * 4, s));
    agg_mark_nodes_kernel<<<grid_for(m), 256, 0, s>>>(m, E.u, E.v, used.as<uint32_t>());
    EXA_CUDA(cudaGetLastError());
    EXA_TRY(exclusive_scan_u32(used.as<uint32_t>(), dense.as<uint32_t>(), n_nodes, tmp, s));
    uint32_t n_live = 0;
    EXA_TRY(scan_total_u32(used.as<uint32_t>(), dense.as<uint32_t>(), n_nodes, &n_live, s));
    EXA_TRY(ids_dev.alloc((size_t)n_live * 4));
    agg_dense_ids_kernel<<<grid_for(n_nodes), 256, 0, s>>>(n_nodes, used.as<uint32_t>(),
                                                            dense.as<uint32_t>(), ids_dev.as<uint32_t>());
    agg_dense_kernel<<<grid_for(m), 256, 0, s>>>(m, E.u, E.v, dense.as<uint32_t>());
    EXA_CUDA(cudaGetLastError());
    std::vector<uint32_t> a(m), b(m), c(m), k(m), ids(n_live);
    std::vector<uint64_t> q(m);
    EXA_CUDA(cudaMemcpyAsync(a.data(), E.u, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
    EXA_CUDA(cudaMemcpyAsync(b.data(), E.v, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
    EXA_CUDA(cudaMemcpyAsync(q.data(), E.q, (size_t)m * 8, cudaMemcpyDeviceToHost, s));
    EXA_CUDA(cudaMemcpyAsync(c.data(), E.c, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
    EXA_CUDA(cudaMemcpyAsync(k.data(), E.k, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
    EXA_CUDA(cudaMemcpyAsync(ids.data(), ids_dev.p, (size_t)n_live * 4, cudaMemcpyDeviceToHost, s));
    EXA_CUDA(cudaStreamSynchronize(s));
    if (tail_edges_out) *tail_edges_out = (uint32_t)a.size();
    std::vector<uint32_t> hp((size_t)n_live + 1);
    for (uint32_t i = 0; i <= n_live; ++i) hp[i] = i;
    sublap("tail to the host");
    const auto t_queue = std::chrono::steady_clock::now();
    ws::agglomerate(n_live, a.size(), a.data(), b.data(), q.data(), c.data(), k.data(), T, hp.data());
    g_last_profile[3] =
        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_queue).count();
    sublap("host queue");
    // every merged region points at the smallest region id of its group (dense ids ascend with
    // region ids, and a region id is the smallest fragment id of the region)
    std::vector<uint32_t> top((size_t)n_live + 1), least((size_t)n_live + 1, kNone);
    for (uint32_t i = 1; i <= n_live; ++i) {
      uint32_t x = i;
      while (hp[x] != x) x = hp[x];
      top[i] = x;
      least[x] = std::min(least[x], i);
    }
    std::vector<uint32_t> node, to;
    for (uint32_t i = 1; i <= n_live; ++i)
      if (least[top[i]] != i) {
        node.push_back(ids[i - 1]);
        to.push_back(ids[least[top[i]] - 1]);
      }
    if (!node.empty()) {
      DevBuf dn(s), dt(s);
      EXA_TRY(dn.alloc(node.size() * 4));
      EXA_TRY(dt.alloc(to.size() * 4));
      EXA_CUDA(cudaMemcpyAsync(dn.p, node.data(), node.size() * 4, cudaMemcpyHostToDevice, s));
      EXA_CUDA(cudaMemcpyAsync(dt.p, to.data(), to.size() * 4, cudaMemcpyHostToDevice, s));
      agg_set_parent_kernel<<<grid_for(node.size()), 256, 0, s>>>((uint32_t)node.size(), dn.as<uint32_t>(),
                                                                   dt.as<uint32_t>(), parent.as<uint32_t>());
      EXA_CUDA(cudaGetLastError());
      EXA_CUDA(cudaStreamSynchronize(s));  // node / to are host vectors
    }
  }
  agg_roots_kernel<<<grid_for(n_nodes), 256, 0, s>>>(n_nodes, parent.as<uint32_t>(), root);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

}  // namespace

namespace {
Status segmentation_impl(const float* aff, int D, int H, int W, const double* thresholds,
                         int n_thresholds, double aff_low, double aff_high, int64_t min_segment_size,
                         uint64_t* seg, int64_t* n_fragments, int64_t* n_segments, cudaStream_t s);
}

Status affinities_to_segmentation_device(const float* aff, int D, int H, int W,
                                         const double* thresholds, int n_thresholds, double aff_low,
                                         double aff_high, int64_t min_segment_size, uint64_t* seg,
                                         int64_t* n_fragments, int64_t* n_segments,
                                         cudaStream_t s) {
  Status st = segmentation_impl(aff, D, H, W, thresholds, n_thresholds, aff_low, aff_high,
                                min_segment_size, seg, n_fragments, n_segments, s);
  // also after an error: the blocks of this call are back in the cache and must be idle
  cudaStreamSynchronize(s);
  return st;
}

void ws_release_memory() {
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) {
    cudaDeviceSynchronize();
    cache_release_all(dev);
  }
}

namespace {
Status segmentation_impl(const float* aff, int D, int H, int W, const double* thresholds,
                         int n_thresholds, double aff_low, double aff_high, int64_t min_segment_size,
                         uint64_t* seg, int64_t* n_fragments, int64_t* n_segments, cudaStream_t s) {
  EXA_CHECK(aff && seg, "affinities_to_segmentation: null buffer");
  EXA_CHECK(D > 0 && H > 0 && W > 0, "affinities_to_segmentation: dims must be positive");
  EXA_CHECK(thresholds && n_thresholds > 0, "affinities_to_segmentation: no agglomeration threshold");
  Vol g{D, H, W, (size_t)D * H * W, (size_t)H * W};
  EXA_CHECK(g.n < (1ull << 32) - 1, "affinities_to_segmentation: volume too large for 32-bit voxel ids");
  const float low = (float)aff_low, high = (float)aff_high;
  EXA_CHECK(low < high, "affinities_to_segmentation: aff_threshold_low must be below aff_threshold_high");
  // thresholds are cumulative: the segmentation of the last one is what the reference keeps
  // (inference.py:232), i.e. merging runs up to the largest
  const double threshold = *std::max_element(thresholds, thresholds + n_thresholds);
  const int64_t T = ws::fixed_threshold(threshold);
  const unsigned blocks = grid_for(g.n);
  // EXA_WS_PROF=1: wall time of the phases on stderr (every phase ends with a stream sync)
  const bool prof = env_int("EXA_WS_PROF", 0) == 1;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* what, int slot) {
    cudaStreamSynchronize(s);
    const auto now = std::chrono::steady_clock::now();
    const double ms = std::chrono::duration<double, std::milli>(now - t_last).count();
    g_last_profile[slot] = ms;
    if (prof)
      fprintf(stderr, "[exa watershed] %-28s %9.3f ms  (allocator calls %.3f ms)\n", what, ms,
              g_alloc_ms);
    g_alloc_ms = 0.0;
    t_last = now;
  };
  for (double& v : g_last_profile) v = 0.0;

  auto t_sub = std::chrono::steady_clock::now();
  auto sub = [&](const char* what) {
    if (!prof) return;
    cudaStreamSynchronize(s);
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[exa watershed]   %-26s %9.3f ms\n", what,
            std::chrono::duration<double, std::milli>(now - t_sub).count());
    t_sub = now;
  };

  DevBuf bits(s), pos(s), parent(s), flag(s), rank(s), front(s), next(s), tmp(s), counter(s);
  EXA_TRY(bits.alloc(g.n));
  EXA_TRY(pos.alloc(g.n * 4));
  EXA_TRY(parent.alloc(g.n * 4));
  EXA_TRY(flag.alloc(g.n * 4));
  EXA_TRY(rank.alloc(g.n * 4));
  EXA_TRY(front.alloc(g.n * 4));
  EXA_TRY(counter.alloc(4));
  sub("allocations");

  // ---- fragments ----
  ws_dirs_kernel<<<blocks, 256, 0, s>>>(aff, g, low, high, bits.as<uint8_t>());
  ws_corner_kernel<<<blocks, 256, 0, s>>>(bits.as<uint8_t>(), g, flag.as<uint32_t>());
  EXA_CUDA(cudaGetLastError());
  EXA_TRY(exclusive_scan_u32(flag.as<uint32_t>(), rank.as<uint32_t>(), g.n, tmp, s));
  ws_level0_kernel<<<blocks, 256, 0, s>>>(g.n, flag.as<uint32_t>(), rank.as<uint32_t>(),
                                           pos.as<uint32_t>(), front.as<uint32_t>());
  EXA_CUDA(cudaGetLastError());
  uint32_t n_front = 0;
  EXA_TRY(scan_total_u32(flag.as<uint32_t>(), rank.as<uint32_t>(), g.n, &n_front, s));
  sub("directions + corners");
  {
    // breadth-first levels of the plateau division; `parent` serves as the claim array
    uint32_t* claim = parent.as<uint32_t>();
    fill_u32_kernel<<<blocks, 256, 0, s>>>(g.n, claim, kNone);
    EXA_CUDA(cudaGetLastError());
    uint32_t base = 0;
    int levels = 0;
    while (n_front > 0) {
      const unsigned fb = grid_for(n_front);
      EXA_CUDA(cudaMemsetAsync(counter.p, 0, 4, s));
      ws_claim_kernel<<<fb, 256, 0, s>>>(front.as<uint32_t>(), n_front, bits.as<uint8_t>(),
                                          pos.as<uint32_t>(), g, claim, counter.as<uint32_t>());
      EXA_CUDA(cudaGetLastError());
      uint32_t claims = 0;
      EXA_CUDA(cudaMemcpyAsync(&claims, counter.p, 4, cudaMemcpyDeviceToHost, s));
      EXA_CUDA(cudaStreamSynchronize(s));
      if (claims == 0) break;
      if (!next.p) EXA_TRY(next.alloc(g.n * 4));
      // flag / rank are free again: children per front voxel and their offsets
      ws_count_children_kernel<<<fb, 256, 0, s>>>(front.as<uint32_t>(), n_front, bits.as<uint8_t>(),
                                                   pos.as<uint32_t>(), g, claim, flag.as<uint32_t>());
      EXA_CUDA(cudaGetLastError());
      EXA_TRY(exclusive_scan_u32(flag.as<uint32_t>(), rank.as<uint32_t>(), n_front, tmp, s));
      uint32_t n_next = 0;
      EXA_TRY(scan_total_u32(flag.as<uint32_t>(), rank.as<uint32_t>(), n_front, &n_next, s));
      base += n_front;
      ws_emit_children_kernel<<<fb, 256, 0, s>>>(front.as<uint32_t>(), n_front, bits.as<uint8_t>(),
                                                  pos.as<uint32_t>(), g, claim, rank.as<uint32_t>(),
                                                  base, next.as<uint32_t>());
      EXA_CUDA(cudaGetLastError());
      front.swap(next);
      n_front = n_next;
      ++levels;
    }
    if (prof) fprintf(stderr, "[exa watershed]   plateau division: %d levels\n", levels);
  }
  next.release();
  sub("plateau levels");
  iota_u32_kernel<<<blocks, 256, 0, s>>>(g.n, parent.as<uint32_t>());
  ws_point_kernel<<<blocks, 256, 0, s>>>(bits.as<uint8_t>(), pos.as<uint32_t>(), g, parent.as<uint32_t>());
  uint32_t* frag = front.as<uint32_t>();  // the front list is dead
  ws_flatten_kernel<<<blocks, 256, 0, s>>>(g.n, parent.as<uint32_t>(), bits.as<uint8_t>(), frag,
                                            flag.as<uint32_t>());
  EXA_CUDA(cudaGetLastError());
  EXA_TRY(exclusive_scan_u32(flag.as<uint32_t>(), rank.as<uint32_t>(), g.n, tmp, s));
  ws_assign_kernel<<<blocks, 256, 0, s>>>(g.n, bits.as<uint8_t>(), rank.as<uint32_t>(), frag);
  EXA_CUDA(cudaGetLastError());
  uint32_t n_frag = 0;
  EXA_TRY(scan_total_u32(flag.as<uint32_t>(), rank.as<uint32_t>(), g.n, &n_frag, s));
  if (n_fragments) *n_fragments = n_frag;
  sub("union-find + ids");
  pos.release();
  parent.release();
  bits.release();
  lap("fragments (GPU)", 0);

  // ---- region graph ----
  const uint32_t n_nodes = n_frag + 1;
  DevBuf root(s);
  EXA_TRY(root.alloc((size_t)n_nodes * 4));
  uint32_t n_region_edges = 0;
  int rounds = 0;
  uint32_t tail_edges = 0;
  {
    uint32_t* cnt = flag.as<uint32_t>();  // root flags are dead
    ws_count_faces_kernel<<<blocks, 256, 0, s>>>(frag, g, cnt);
    EXA_CUDA(cudaGetLastError());
    DevBuf offset(s);
    EXA_TRY(offset.alloc(g.n * 8));
    size_t tmp_bytes = 0;
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
    sub("count faces");
    EXA_CHECK(n_faces < (1ull << 32),
              "affinities_to_segmentation: more than 2^32 faces between fragments; split the volume");
    EdgeStore es(s);
    if (n_faces > 0) {
      const long long m = (long long)n_faces;   // 64-bit item counts in the CUB calls below
      DevBuf keys(s), vals(s), keys2(s), vals2(s), ucnt(s), nruns(s);
      EXA_TRY(keys.alloc((size_t)m * 8));
      EXA_TRY(vals.alloc((size_t)m * 8));
      EXA_TRY(keys2.alloc((size_t)m * 8));
      EXA_TRY(vals2.alloc((size_t)m * 8));
      ws_emit_faces_kernel<<<blocks, 256, 0, s>>>(frag, aff, g, offset.as<unsigned long long>(),
                                                   keys.as<unsigned long long>(),
                                                   vals.as<unsigned long long>());
      EXA_CUDA(cudaGetLastError());
      offset.release();
      sub("emit faces");
      int id_bits = 1;
      while ((1ull << id_bits) <= n_frag) ++id_bits;
      EXA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.as<unsigned long long>(),
                                               keys2.as<unsigned long long>(),
                                               vals.as<unsigned long long>(),
                                               vals2.as<unsigned long long>(), m, 0, 32 + id_bits, s));
      EXA_TRY(tmp.alloc(tmp_bytes));
      EXA_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys.as<unsigned long long>(),
                                               keys2.as<unsigned long long>(),
                                               vals.as<unsigned long long>(),
                                               vals2.as<unsigned long long>(), m, 0, 32 + id_bits, s));
      sub("sort faces");
      // segmented sum and run lengths; `keys` / `vals` are free again and take the outputs
      EXA_TRY(ucnt.alloc((size_t)m * 4));
      EXA_TRY(nruns.alloc(16));
      unsigned long long* ukey = keys.as<unsigned long long>();
      unsigned long long* usum = vals.as<unsigned long long>();
      long long* n_runs = nruns.as<long long>();
      EXA_CUDA(cub::DeviceReduce::ReduceByKey(nullptr, tmp_bytes, keys2.as<unsigned long long>(), ukey,
                                              vals2.as<unsigned long long>(), usum, n_runs,
                                              cub::Sum(), m, s));
      EXA_TRY(tmp.alloc(tmp_bytes));
      EXA_CUDA(cub::DeviceReduce::ReduceByKey(tmp.p, tmp_bytes, keys2.as<unsigned long long>(), ukey,
                                              vals2.as<unsigned long long>(), usum, n_runs,
                                              cub::Sum(), m, s));
      // faces per pair: the same reduction over ones (the run-length primitive takes 32-bit counts)
      cub::ConstantInputIterator<uint32_t> ones(1u);
      EXA_CUDA(cub::DeviceReduce::ReduceByKey(nullptr, tmp_bytes, keys2.as<unsigned long long>(), ukey,
                                              ones, ucnt.as<uint32_t>(), n_runs + 1, cub::Sum(), m, s));
      EXA_TRY(tmp.alloc(tmp_bytes));
      EXA_CUDA(cub::DeviceReduce::ReduceByKey(tmp.p, tmp_bytes, keys2.as<unsigned long long>(), ukey,
                                              ones, ucnt.as<uint32_t>(), n_runs + 1, cub::Sum(), m, s));
      long long runs[2] = {0, 0};
      EXA_CUDA(cudaMemcpyAsync(runs, nruns.p, 16, cudaMemcpyDeviceToHost, s));
      EXA_CUDA(cudaStreamSynchronize(s));
      EXA_CHECK(runs[0] == runs[1], "affinities_to_segmentation: region graph run counts disagree");
      EXA_CHECK(runs[0] < (1ll << 32) - 1, "affinities_to_segmentation: more than 2^32 region edges");
      n_region_edges = (uint32_t)runs[0];
      keys2.release();
      vals2.release();
      EXA_TRY(es.u.alloc((size_t)n_region_edges * 4));
      EXA_TRY(es.v.alloc((size_t)n_region_edges * 4));
      EXA_TRY(es.k.alloc((size_t)n_region_edges * 4));
      ws_split_keys_kernel<<<grid_for(n_region_edges), 256, 0, s>>>(
          n_region_edges, ukey, es.u.as<uint32_t>(), es.v.as<uint32_t>(), es.k.as<uint32_t>());
      EXA_CUDA(cudaGetLastError());
      // the sums and counts stay where they are: hand the buffers over
      es.q.swap(vals);
      es.c.swap(ucnt);
    }
    sub("reduce by pair");
    lap("region graph (GPU)", 1);

    // ---- agglomeration ----
    EXA_TRY(agglomerate_rounds(es, n_region_edges, n_frag, T, root.as<uint32_t>(), s, prof, &rounds,
                               &tail_edges));
  }
  if (prof)
    fprintf(stderr, "[exa watershed] %u fragments, %u region edges, %d parallel rounds, %u edges to the host queue\n",
            n_frag, n_region_edges, rounds, tail_edges);
  lap("agglomeration", 2);
  g_last_profile[2] -= g_last_profile[3];   // slot 3 (host queue) was filled by agglomerate_rounds
  g_last_profile[5] = rounds;
  g_last_profile[6] = n_region_edges;
  g_last_profile[7] = tail_edges;

  // ---- small segments out, ids in order of first appearance (img_util.py:536-559) ----
  {
    DevBuf fsize(s), ssize(s), keep(s), krank(s), lut(s);
    EXA_TRY(fsize.alloc((size_t)n_nodes * 8));
    EXA_TRY(ssize.alloc((size_t)n_nodes * 8));
    EXA_TRY(keep.alloc((size_t)n_nodes * 4));
    EXA_TRY(krank.alloc((size_t)n_nodes * 4));
    EXA_TRY(lut.alloc((size_t)n_nodes * 4));
    EXA_CUDA(cudaMemsetAsync(fsize.p, 0, (size_t)n_nodes * 8, s));
    EXA_CUDA(cudaMemsetAsync(ssize.p, 0, (size_t)n_nodes * 8, s));
    const unsigned nb = grid_for(n_nodes);
    ws_size_kernel<<<blocks, 256, 0, s>>>(g.n, frag, fsize.as<unsigned long long>());
    ws_segment_size_kernel<<<nb, 256, 0, s>>>(n_nodes, root.as<uint32_t>(),
                                               fsize.as<unsigned long long>(),
                                               ssize.as<unsigned long long>());
    ws_keep_kernel<<<nb, 256, 0, s>>>(n_nodes, root.as<uint32_t>(), ssize.as<unsigned long long>(),
                                       (long long)min_segment_size, keep.as<uint32_t>());
    EXA_CUDA(cudaGetLastError());
    EXA_TRY(exclusive_scan_u32(keep.as<uint32_t>(), krank.as<uint32_t>(), n_nodes, tmp, s));
    ws_lut_kernel<<<nb, 256, 0, s>>>(n_nodes, root.as<uint32_t>(), keep.as<uint32_t>(),
                                      krank.as<uint32_t>(), lut.as<uint32_t>());
    ws_relabel_kernel<<<blocks, 256, 0, s>>>(g.n, frag, lut.as<uint32_t>(), seg);
    EXA_CUDA(cudaGetLastError());
    uint32_t kept = 0;
    EXA_TRY(scan_total_u32(keep.as<uint32_t>(), krank.as<uint32_t>(), n_nodes, &kept, s));
    if (n_segments) *n_segments = kept;
  }
  EXA_CUDA(cudaStreamSynchronize(s));
  lap("sizes + relabel (GPU)", 4);
  return Status::OK();
}

}  // namespace

void ws_last_profile(double* out, int n) {
  for (int i = 0; i < n && i < 8; ++i) out[i] = g_last_profile[i];
}

Status region_agglomerate(int device, uint32_t n_fragments, int64_t n_edges, const uint32_t* eu,
                          const uint32_t* ev, const uint64_t* qsum, const uint32_t* count,
                          double threshold, uint32_t* root_out) {
  EXA_CHECK(n_edges >= 0 && n_edges < ((int64_t)1 << 31) && root_out &&
                (n_edges == 0 || (eu && ev && qsum && count)),
            "region_agglomerate: bad argument");
  uint64_t faces = 0;
  for (int64_t i = 0; i < n_edges; ++i) {
    EXA_CHECK(eu[i] >= 1 && eu[i] < ev[i] && ev[i] <= n_fragments && count[i] > 0 &&
                  qsum[i] <= ((uint64_t)count[i] << 32),
              "region_agglomerate: bad edge");
    faces += count[i];
  }
  EXA_CHECK(faces < (1ull << 32), "region_agglomerate: more than 2^32 faces in total");
  const int64_t T = ws::fixed_threshold(threshold);
  const size_t m = (size_t)n_edges;
  if (device < 0) {
    // the host queue alone
    std::vector<uint32_t> key(m), parent((size_t)n_fragments + 1);
    for (size_t i = 0; i < m; ++i) key[i] = (uint32_t)i;
    for (uint32_t i = 0; i <= n_fragments; ++i) parent[i] = i;
    ws::agglomerate(n_fragments, m, eu, ev, qsum, count, key.data(), T, parent.data());
    std::vector<uint32_t> top((size_t)n_fragments + 1), least((size_t)n_fragments + 1, kNone);
    for (uint32_t i = 0; i <= n_fragments; ++i) {
      uint32_t x = i;
      while (parent[x] != x) x = parent[x];
      top[i] = x;
      least[x] = std::min(least[x], i);
    }
    for (uint32_t i = 0; i <= n_fragments; ++i) root_out[i] = least[top[i]];
    return Status::OK();
  }
  EXA_CUDA(cudaSetDevice(device));
  cudaStream_t s = nullptr;
  EdgeStore es(s);
  DevBuf root(s);
  EXA_TRY(root.alloc(((size_t)n_fragments + 1) * 4));
  if (m > 0) {
    EXA_TRY(es.alloc(m));
    std::vector<uint32_t> key(m);
    for (size_t i = 0; i < m; ++i) key[i] = (uint32_t)i;
    EXA_CUDA(cudaMemcpyAsync(es.u.p, eu, m * 4, cudaMemcpyHostToDevice, s));
    EXA_CUDA(cudaMemcpyAsync(es.v.p, ev, m * 4, cudaMemcpyHostToDevice, s));
    EXA_CUDA(cudaMemcpyAsync(es.q.p, qsum, m * 8, cudaMemcpyHostToDevice, s));
    EXA_CUDA(cudaMemcpyAsync(es.c.p, count, m * 4, cudaMemcpyHostToDevice, s));
    EXA_CUDA(cudaMemcpyAsync(es.k.p, key.data(), m * 4, cudaMemcpyHostToDevice, s));
    EXA_CUDA(cudaStreamSynchronize(s));
  }
  Status st = agglomerate_rounds(es, (uint32_t)m, n_fragments, T, root.as<uint32_t>(), s, false,
                                 nullptr, nullptr);
  if (st.ok) {
    cudaError_t e = cudaMemcpyAsync(root_out, root.p, ((size_t)n_fragments + 1) * 4,
                                    cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) st = Status::Err(std::string("region_agglomerate: ") + cudaGetErrorString(e));
  }
  cudaStreamSynchronize(s);  // also after an error: the cached blocks must be idle
  return st;
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
