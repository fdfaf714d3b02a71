// CPU restatement of waterz.agglomerate as the reference calls it -- TEST INFRASTRUCTURE ONLY.
//
// Call site: reference src/aind_exaspim_neuron_segmentation/inference.py:224-229
//     waterz.agglomerate(affinities, thresholds, aff_threshold_low=0.1, aff_threshold_high=0.9999)
// with waterz's defaults (scoring_function = OneMinus<MeanAffinity<RegionGraphType, ScoreValue>>,
// fragments=None, discretize_queue=0).  waterz is `git+https://github.com/anna-grim/waterz.git@master`
// (pyproject.toml:32): a fork pinned to a BRANCH, needs boost, absent from /root/reference and from
// this image.  PARITY UNPINNED: nothing here could be checked against waterz output; the functions
// restate the published algorithm (funkey/waterz backend: basic_watershed.hpp, region_graph.hpp,
// IterativeRegionMerging.hpp, MeanAffinityProvider.hpp; Zlateski & Seung 2015) step by step:
//
//  wsref_watershed      steepest-ascent watershed: direction bits per voxel from its six incident
//                       affinities, plateau corners, BFS division of plateaus, basin labelling by
//                       a second BFS in voxel scan order.  aff[c][z][y][x] is the edge between
//                       voxel (z,y,x) and its PREVIOUS neighbour along axis c; index 0 along c has
//                       no such edge.
//  wsref_region_graph   for every voxel and every axis (z, y, x order) the edge to the previous
//                       voxel; faces between two different non-zero fragments add their affinity
//                       to that pair's statistics (sum, count), in voxel scan order.
//  wsref_agglomerate    priority-queue merging: pop the edge with the smallest score
//                       1 - sum/count; stop at the first score >= threshold; merge the two regions;
//                       statistics of parallel edges add up; re-scored edges re-enter the queue.
//
// Two accumulation modes for the edge statistics:
//   mode 0 "float32": float sums in scan / merge order and float scores, as waterz does
//                      (ScoreValue = float).  The result depends on the order of additions.
//   mode 1 "exact":   every affinity is converted once to 32.32 fixed point
//                      (llrint(clamp(a, 0, 1) * 2^32)) and summed as uint64 (q); the score
//                      1 - q / (count * 2^32) is never rounded: two edges are ordered by comparing
//                      q1 * count2 with q2 * count1 in 128-bit integers, and "score < threshold"
//                      is q > count * T with T = llrint((1 - threshold) * 2^32).  Nothing depends
//                      on the order of additions, which is what lets the GPU implementation merge
//                      mutual-best pairs in parallel and still be compared element for element
//                      (mean-affinity linkage is reducible, so with a strict total order on the
//                      edges every merge order yields the same partition).  Differs from mode 0 by
//                      float32 rounding only (tests bound the effect with the adapted-Rand score).
// Ties: waterz's std::priority_queue leaves the order of equal scores unspecified.  Here equal
// scores are ordered by `key`, the smallest lexicographic rank among the original region-graph
// edges an edge is made of (invariant under merging).
// root_out: the smallest fragment id of the region (independent of the merge order).
//
// Build: g++ -O2 -shared -fPIC -o oracle/_build/libwsref.so oracle/ws_ref.cpp  (oracle/build.py)
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <queue>
#include <unordered_map>
#include <utility>
#include <vector>

namespace {

const uint64_t HIGH_BIT = 0x8000000000000000ull;
const uint64_t VISITED = 0x40;

inline uint64_t to_fixed(float a) {
  if (!(a > 0.0f)) return 0;  // also NaN
  if (a > 1.0f) a = 1.0f;
  return (uint64_t)llrint((double)a * 4294967296.0);
}

}  // namespace

extern "C" {

// seg: uint64[D*H*W] out, fragment ids 1..n (0 = background).  Returns n.
int64_t wsref_watershed(const float* aff, int64_t D, int64_t H, int64_t W, float low, float high,
                        uint64_t* seg) {
  const int64_t hw = H * W, n = D * hw;
  const float* az = aff;
  const float* ay = aff + n;
  const float* ax = aff + 2 * n;
  // 1. steepest-ascent direction bits (an edge >= high always counts, the rest only if maximal)
  for (int64_t z = 0; z < D; ++z)
    for (int64_t y = 0; y < H; ++y)
      for (int64_t x = 0; x < W; ++x) {
        const int64_t i = z * hw + y * W + x;
        const float negz = z > 0 ? az[i] : low;
        const float negy = y > 0 ? ay[i] : low;
        const float negx = x > 0 ? ax[i] : low;
        const float posz = z < D - 1 ? az[i + hw] : low;
        const float posy = y < H - 1 ? ay[i + W] : low;
        const float posx = x < W - 1 ? ax[i + 1] : low;
        const float m = std::max({negz, negy, negx, posz, posy, posx});
        uint64_t id = 0;
        if (m > low) {
          if (negz == m || negz >= high) id |= 0x01;
          if (negy == m || negy >= high) id |= 0x02;
          if (negx == m || negx >= high) id |= 0x04;
          if (posz == m || posz >= high) id |= 0x08;
          if (posy == m || posy >= high) id |= 0x10;
          if (posx == m || posx >= high) id |= 0x20;
        }
        seg[i] = id;
      }
  const int64_t dir[6] = {-hw, -W, -1, hw, W, 1};
  const uint64_t dirmask[6] = {0x01, 0x02, 0x04, 0x08, 0x10, 0x20};
  const uint64_t idirmask[6] = {0x08, 0x10, 0x20, 0x01, 0x02, 0x04};

  // 2. plateau corners: voxels with an outgoing edge that is not returned
  std::vector<int64_t> bfs;
  for (int64_t i = 0; i < n; ++i)
    for (int d = 0; d < 6; ++d)
      if (seg[i] & dirmask[d]) {
        if (!(seg[i + dir[d]] & idirmask[d])) {
          seg[i] |= VISITED;
          bfs.push_back(i);
          break;
        }
      }
  // 3. divide the plateaus: breadth first from the corners; every voxel keeps ONE outgoing edge
  //    (the last, in direction order, that leads out of the still undivided part)
  for (size_t k = 0; k < bfs.size(); ++k) {
    const int64_t i = bfs[k];
    uint64_t to_set = 0;
    for (int d = 0; d < 6; ++d)
      if (seg[i] & dirmask[d]) {
        const int64_t j = i + dir[d];
        if (seg[j] & idirmask[d]) {
          if (!(seg[j] & VISITED)) {
            bfs.push_back(j);
            seg[j] |= VISITED;
          }
        } else {
          to_set = dirmask[d];
        }
      }
    seg[i] = to_set;
  }
  bfs.clear();
  // 4. basins: follow the edges from every unlabelled voxel in scan order; reaching a labelled
  //    voxel joins its basin, otherwise the visited set is a new basin
  uint64_t next_id = 1;
  for (int64_t i = 0; i < n; ++i) {
    if (seg[i] == 0) seg[i] |= HIGH_BIT;  // background
    if (!(seg[i] & HIGH_BIT) && seg[i]) {
      bfs.push_back(i);
      seg[i] |= VISITED;
      for (size_t k = 0; k < bfs.size(); ++k) {
        const int64_t me = bfs[k];
        bool joined = false;
        for (int d = 0; d < 6 && !joined; ++d)
          if (seg[me] & dirmask[d]) {
            const int64_t him = me + dir[d];
            if (seg[him] & HIGH_BIT) {
              for (int64_t v : bfs) seg[v] = seg[him];
              bfs.clear();
              joined = true;
            } else if (!(seg[him] & VISITED)) {
              seg[him] |= VISITED;
              bfs.push_back(him);
            }
          }
        if (joined) break;
      }
      if (!bfs.empty()) {
        for (int64_t v : bfs) seg[v] = HIGH_BIT | next_id;
        ++next_id;
        bfs.clear();
      }
    }
  }
  for (int64_t i = 0; i < n; ++i) seg[i] &= ~HIGH_BIT;
  return (int64_t)(next_id - 1);
}

// Region graph.  Call with u == NULL to get the number of edges; then with arrays of that size.
// Edges come out sorted by (u, v), u < v.  fsum: float32 sums in voxel scan order (mode 0);
// qsum: exact 32.32 fixed-point sums (mode 1).
int64_t wsref_region_graph(const float* aff, const uint64_t* seg, int64_t D, int64_t H, int64_t W,
                           uint32_t* u, uint32_t* v, float* fsum, uint64_t* qsum, uint32_t* count) {
  const int64_t hw = H * W, n = D * hw;
  struct St {
    float f;
    uint64_t q;
    uint32_t c;
  };
  std::unordered_map<uint64_t, St> map;
  const int64_t step[3] = {hw, W, 1};
  for (int64_t z = 0; z < D; ++z)
    for (int64_t y = 0; y < H; ++y)
      for (int64_t x = 0; x < W; ++x) {
        const int64_t i = z * hw + y * W + x;
        const uint64_t a = seg[i];
        if (!a) continue;
        const bool has[3] = {z > 0, y > 0, x > 0};
        for (int c = 0; c < 3; ++c) {
          if (!has[c]) continue;
          const uint64_t b = seg[i - step[c]];
          if (!b || b == a) continue;
          const uint64_t lo = std::min(a, b), hi = std::max(a, b);
          St& s = map[(lo << 32) | hi];
          const float w = aff[c * n + i];
          s.f += w;
          s.q += to_fixed(w);
          s.c += 1;
        }
      }
  if (u == nullptr) return (int64_t)map.size();
  std::vector<uint64_t> keys;
  keys.reserve(map.size());
  for (const auto& kv : map) keys.push_back(kv.first);
  std::sort(keys.begin(), keys.end());
  for (size_t k = 0; k < keys.size(); ++k) {
    const St& s = map[keys[k]];
    u[k] = (uint32_t)(keys[k] >> 32);
    v[k] = (uint32_t)(keys[k] & 0xffffffffu);
    fsum[k] = s.f;
    qsum[k] = s.q;
    count[k] = s.c;
  }
  return (int64_t)keys.size();
}

// Sequential agglomeration.  Edges sorted by (u, v) with 1 <= u < v <= n_frag (the rank in that
// order is the tie key).  root_out[0..n_frag]: the smallest fragment id of the region every fragment
// ends up in.  Returns the number of merges.
int64_t wsref_agglomerate(uint32_t n_frag, int64_t n_edges, const uint32_t* eu, const uint32_t* ev,
                          const float* fsum, const uint64_t* qsum, const uint32_t* count,
                          double threshold, int mode, uint32_t* root_out) {
  struct Edge {
    uint32_t a, b;
    float f;
    uint64_t q;
    uint64_t c;
    uint32_t key;
    uint32_t version;
    bool alive;
  };
  std::vector<Edge> edges((size_t)n_edges);
  std::vector<std::unordered_map<uint32_t, uint32_t>> adj((size_t)n_frag + 1);
  struct Item {
    double s;       // mode 0: the float32 score
    uint64_t q, c;  // mode 1: the exact statistics
    uint32_t key, id, version;
  };
  struct Cmp {  // "x comes after y": min-heap on (score, key)
    int mode;
    bool operator()(const Item& x, const Item& y) const {
      if (mode == 0) {
        if (x.s != y.s) return x.s > y.s;
      } else {
        const unsigned __int128 l = (unsigned __int128)x.q * y.c, r = (unsigned __int128)y.q * x.c;
        if (l != r) return l < r;  // smaller mean affinity = larger score = later
      }
      return x.key > y.key;
    }
  };
  auto item = [&](const Edge& e, uint32_t id) {
    return Item{(double)(1.0f - e.f / (float)e.c), e.q, e.c, e.key, id, e.version};
  };
  // "score < threshold"
  const float thr_f = (float)threshold;
  const double t_fixed = (1.0 - threshold) * 4294967296.0;
  const __int128 T = t_fixed > 9.0e18 ? (__int128)9000000000000000000ll
                   : t_fixed < -9.0e18 ? (__int128)-9000000000000000000ll : (__int128)llrint(t_fixed);
  auto below = [&](const Item& it) {
    if (mode == 0) return (float)it.s < thr_f;
    return (__int128)it.q > (__int128)it.c * T;
  };
  std::priority_queue<Item, std::vector<Item>, Cmp> heap(Cmp{mode});
  for (int64_t i = 0; i < n_edges; ++i) {
    Edge& e = edges[i];
    e.a = eu[i];
    e.b = ev[i];
    e.f = fsum[i];
    e.q = qsum[i];
    e.c = count[i];
    e.key = (uint32_t)i;
    e.version = 0;
    e.alive = true;
    adj[e.a][e.b] = (uint32_t)i;
    adj[e.b][e.a] = (uint32_t)i;
    heap.push(item(e, (uint32_t)i));
  }
  std::vector<uint32_t> parent((size_t)n_frag + 1);
  for (uint32_t i = 0; i <= n_frag; ++i) parent[i] = i;
  int64_t merges = 0;
  std::vector<std::pair<uint32_t, uint32_t>> moved;
  while (!heap.empty()) {
    const Item it = heap.top();
    if (!below(it)) break;  // all remaining (valid or stale) entries are at least as expensive
    heap.pop();
    Edge& e = edges[it.id];
    if (!e.alive || e.version != it.version) continue;
    uint32_t a = e.a, b = e.b;
    if (adj[a].size() < adj[b].size()) std::swap(a, b);  // b disappears into a
    parent[b] = a;
    e.alive = false;
    adj[a].erase(b);
    moved.assign(adj[b].begin(), adj[b].end());
    adj[b].clear();
    for (const auto& kv : moved) {
      const uint32_t nb = kv.first, e2 = kv.second;
      if (nb == a) continue;
      adj[nb].erase(b);
      auto f = adj[a].find(nb);
      if (f != adj[a].end()) {  // parallel edge: statistics add up, the edge is re-scored
        Edge& e1 = edges[f->second];
        e1.f += edges[e2].f;
        e1.q += edges[e2].q;
        e1.c += edges[e2].c;
        e1.key = std::min(e1.key, edges[e2].key);
        e1.version += 1;
        edges[e2].alive = false;
        heap.push(item(e1, f->second));
      } else {  // the edge moves from b to a unchanged (its queue entry stays valid)
        Edge& m = edges[e2];
        if (m.a == b) m.a = a; else m.b = a;
        adj[a][nb] = e2;
        adj[nb][a] = e2;
      }
    }
    ++merges;
  }
  std::vector<uint32_t> top((size_t)n_frag + 1), least((size_t)n_frag + 1, 0xffffffffu);
  for (uint32_t i = 0; i <= n_frag; ++i) {
    uint32_t x = i;
    while (parent[x] != x) x = parent[x];
    top[i] = x;
    least[x] = std::min(least[x], i);
  }
  for (uint32_t i = 0; i <= n_frag; ++i) root_out[i] = least[top[i]];
  return merges;
}

}  // extern "C"
