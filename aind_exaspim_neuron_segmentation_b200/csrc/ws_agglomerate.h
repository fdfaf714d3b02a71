// Exact merge queue of K7 (affinities -> segmentation): hierarchical agglomeration of a region
// graph, reference inference.py:224-229 -> waterz.agglomerate with OneMinus<MeanAffinity> scoring.
// Plain C++ (no CUDA).  It finishes what the parallel GPU rounds of watershed.cu leave over (the
// hub chains: one large region swallowing its neighbours one by one) and is exported on its own as
// exa_region_agglomerate for CPU tests.
//
// Edge statistics are exact: q = sum of the affinities in 32.32 fixed point, c = number of faces,
// key = smallest rank of the original region-graph edges the edge is made of.  The score
// 1 - q / (c * 2^32) is never formed:  edge 1 comes before edge 2 iff q1 * c2 > q2 * c1 (128-bit
// products), ties by key; "score < threshold" is q > c * T, T = llrint((1 - threshold) * 2^32).
// Mean-affinity linkage is reducible, so with this strict total order the partition reached at
// the threshold does not depend on the order in which independent merges are carried out.
#pragma once

#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <utility>
#include <vector>

namespace exa {
namespace ws {

struct Stat {  // 16 bytes: the face count of a volume below 2^32 voxels fits 32 bits
  uint64_t q;
  uint32_t c;
  uint32_t key;
};

inline bool stat_before(uint64_t q1, uint64_t c1, uint32_t k1, uint64_t q2, uint64_t c2,
                        uint32_t k2) {
  const unsigned __int128 l = (unsigned __int128)q1 * c2, r = (unsigned __int128)q2 * c1;
  if (l != r) return l > r;
  return k1 < k2;
}

// T of the header comment, clamped to what the comparison needs
inline int64_t fixed_threshold(double threshold) {
  const double t = (1.0 - threshold) * 4294967296.0;
  if (!(t > -1.0)) return -1;                          // every edge is below the threshold
  if (t > 4294967296.0) return (int64_t)1 << 33;       // no edge is (q <= c * 2^32)
  return (int64_t)llrint(t);
}
inline bool below_threshold(uint64_t q, uint64_t c, int64_t T) {
  if (T < 0) return true;
  return (unsigned __int128)q > (unsigned __int128)c * (unsigned __int128)(uint64_t)T;
}

// neighbour table of one region: open addressing on region ids (>= 1), Fibonacci hashing (the
// HIGH bits of k * 2^32/phi: neighbouring fragments have nearby ids), linear probing, at most
// three quarters full including tombstones; 24-byte slots (a region with 9 neighbours: 384 bytes)
class NbrMap {
 public:
  uint32_t size() const { return live_; }
  void reserve(uint32_t n) {
    uint32_t lg = 2;
    while (3ull * (1u << lg) < 4ull * (n + 1)) ++lg;
    if ((1u << lg) > slot_.size()) rehash(lg);
  }
  Stat* find(uint32_t k) {
    if (slot_.empty()) return nullptr;
    const uint32_t mask = (uint32_t)slot_.size() - 1;
    for (uint32_t i = index(k);; i = (i + 1) & mask) {
      if (slot_[i].key == k) return &slot_[i].val;
      if (slot_[i].key == kEmpty) return nullptr;
    }
  }
  void prefetch(uint32_t k) const {
    if (!slot_.empty()) __builtin_prefetch(&slot_[index(k)]);
  }
  void put(uint32_t k, const Stat& v) {
    if (Stat* p = find(k)) {
      *p = v;
      return;
    }
    if (4 * ((uint64_t)used_ + 1) > 3 * (uint64_t)slot_.size()) {
      uint32_t lg = 2;
      while ((1ull << lg) < 2ull * (live_ + 1)) ++lg;  // at most half full after the clean-up
      rehash(lg);
    }
    const uint32_t mask = (uint32_t)slot_.size() - 1;
    for (uint32_t i = index(k);; i = (i + 1) & mask) {
      if (slot_[i].key == kEmpty || slot_[i].key == kTomb) {
        if (slot_[i].key == kEmpty) ++used_;
        slot_[i].key = k;
        slot_[i].val = v;
        ++live_;
        return;
      }
    }
  }
  void erase(uint32_t k) {
    if (slot_.empty()) return;
    const uint32_t mask = (uint32_t)slot_.size() - 1;
    for (uint32_t i = index(k);; i = (i + 1) & mask) {
      if (slot_[i].key == k) {
        slot_[i].key = kTomb;
        --live_;
        return;
      }
      if (slot_[i].key == kEmpty) return;
    }
  }
  template <typename F>
  void for_each(F&& f) const {
    for (const Slot& sl : slot_)
      if (sl.key != kEmpty && sl.key != kTomb) f(sl.key, sl.val);
  }
  void clear() {
    std::vector<Slot>().swap(slot_);
    live_ = used_ = 0;
  }

 private:
  struct Slot {
    Stat val;
    uint32_t key;
  };
  static constexpr uint32_t kEmpty = 0, kTomb = 0xffffffffu;
  uint32_t index(uint32_t k) const { return (k * 2654435769u) >> shift_; }
  void rehash(uint32_t lg) {
    std::vector<Slot> old((size_t)1 << lg, Slot{Stat{0, 0, 0}, kEmpty});
    old.swap(slot_);
    shift_ = 32 - lg;
    live_ = used_ = 0;
    const uint32_t mask = (uint32_t)slot_.size() - 1;
    for (const Slot& sl : old) {
      if (sl.key == kEmpty || sl.key == kTomb) continue;
      uint32_t i = index(sl.key);
      while (slot_[i].key != kEmpty) i = (i + 1) & mask;
      slot_[i] = sl;
      ++live_;
      ++used_;
    }
  }
  std::vector<Slot> slot_;
  uint32_t live_ = 0, used_ = 0, shift_ = 30;
};

struct Entry {  // one queue entry: an edge as it was when pushed (stale once its count changed)
                // a, b: the regions it joined then; they may have been merged into others since
  uint64_t q;
  uint32_t c, key, a, b;
};
struct EntryAfter {  // heap order: the top is the entry that comes first
  bool operator()(const Entry& x, const Entry& y) const {
    return stat_before(y.q, y.c, y.key, x.q, x.c, x.key);
  }
};

// Monotone bucket queue: scores only grow along the merge sequence (a merged edge's mean affinity
// lies between those of its parts, which were both not before the current minimum), so entries are
// binned by floor(q / c) -- integer division, hence exactly monotone in the edge order -- and only
// the small heap of the current bin is ever touched: the pop order is that of one global heap.
// Entries at or above the threshold are never popped before the loop ends and are not stored.
class BucketQueue {
 public:
  explicit BucketQueue(int64_t T) : T_(T), bin_((size_t)1 << kBits) {}
  static size_t bin_of(uint64_t q, uint64_t c) {
    const uint64_t m = std::min<uint64_t>(q / c, 4294967295ull);  // mean affinity, 0.32 fixed point
    return (size_t)((4294967295ull - m) >> (32 - kBits));
  }
  void bulk_load(const std::vector<Entry>& all) {
    std::vector<uint32_t> n(bin_.size(), 0);
    for (const Entry& e : all)
      if (below_threshold(e.q, e.c, T_)) ++n[bin_of(e.q, e.c)];
    for (size_t i = 0; i < bin_.size(); ++i)
      if (n[i]) bin_[i].reserve(n[i] + n[i] / 2);
    for (const Entry& e : all)
      if (below_threshold(e.q, e.c, T_)) bin_[bin_of(e.q, e.c)].push_back(e);
    for (std::vector<Entry>& h : bin_)
      if (!h.empty()) std::make_heap(h.begin(), h.end(), EntryAfter());
    cur_ = 0;
  }
  void push(const Entry& e) {
    if (!below_threshold(e.q, e.c, T_)) return;
    const size_t i = bin_of(e.q, e.c);
    if (i < cur_) cur_ = i;  // cannot happen for a reducible linkage; harmless
    bin_[i].push_back(e);
    std::push_heap(bin_[i].begin(), bin_[i].end(), EntryAfter());
  }
  bool pop(Entry* e) {
    while (cur_ < bin_.size() && bin_[cur_].empty()) ++cur_;
    if (cur_ == bin_.size()) return false;
    std::vector<Entry>& h = bin_[cur_];
    std::pop_heap(h.begin(), h.end(), EntryAfter());
    *e = h.back();
    h.pop_back();
    return true;
  }

 private:
  static constexpr int kBits = 14;
  int64_t T_;
  size_t cur_ = 0;
  std::vector<std::vector<Entry>> bin_;
};

// Edges (a[i], b[i]) between regions 1..n_nodes (a != b, every pair at most once) with statistics
// (q, c, key).  parent[1..n_nodes] must be the identity on entry; on return it is a forest (paths
// partly compressed) whose roots are the regions left at the threshold.  Returns the number of merges.
inline int64_t agglomerate(uint32_t n_nodes, size_t n_edges, const uint32_t* ea, const uint32_t* eb,
                           const uint64_t* eq, const uint32_t* ec, const uint32_t* ekey,
                           int64_t T, uint32_t* parent) {
  std::vector<NbrMap> nbr((size_t)n_nodes + 1);
  {
    std::vector<uint32_t> deg((size_t)n_nodes + 1, 0);
    for (size_t i = 0; i < n_edges; ++i) {
      ++deg[ea[i]];
      ++deg[eb[i]];
    }
    for (uint32_t i = 1; i <= n_nodes; ++i)
      if (deg[i]) nbr[i].reserve(deg[i]);
  }
  BucketQueue heap(T);
  {
    std::vector<Entry> init;
    init.reserve(n_edges);
    for (size_t i = 0; i < n_edges; ++i) {
      const Stat st{eq[i], ec[i], ekey[i]};
      nbr[ea[i]].put(eb[i], st);
      nbr[eb[i]].put(ea[i], st);
      init.push_back(Entry{st.q, st.c, st.key, ea[i], eb[i]});
    }
    heap.bulk_load(init);
  }
  int64_t merges = 0;
  std::vector<std::pair<uint32_t, Stat>> moved;
  // entries name the regions their edge joined when it was pushed; an edge that merely moves to
  // the surviving region keeps its entry (resolved through the forest), so only combined edges
  // are pushed again
  auto find = [&](uint32_t x) {
    while (parent[x] != x) {
      parent[x] = parent[parent[x]];
      x = parent[x];
    }
    return x;
  };
  Entry e;
  while (heap.pop(&e)) {
    uint32_t a = find(e.a), b = find(e.b);
    if (a == b) continue;
    const Stat* cur_ab = nbr[a].find(b);
    if (cur_ab == nullptr || cur_ab->c != e.c) continue;  // counts only grow: c identifies the state
    if (nbr[a].size() < nbr[b].size()) std::swap(a, b);   // the region with fewer neighbours goes
    parent[b] = a;
    ++merges;
    nbr[a].erase(b);
    moved.clear();
    nbr[b].for_each([&](uint32_t nb, const Stat& st) {
      if (nb != a) {
        moved.emplace_back(nb, st);
        __builtin_prefetch(&nbr[nb]);   // the table header first, its slots in the next pass
        nbr[a].prefetch(nb);
      }
    });
    nbr[b].clear();
    for (const auto& kv : moved) nbr[kv.first].prefetch(b);
    for (const auto& kv : moved) {
      const uint32_t nb = kv.first;
      nbr[nb].erase(b);
      Stat cur = kv.second;
      const Stat* f = nbr[a].find(nb);
      const bool combined = f != nullptr;
      if (combined) {
        cur.q += f->q;
        cur.c += f->c;
        cur.key = std::min(cur.key, f->key);
      }
      nbr[a].put(nb, cur);
      nbr[nb].put(a, cur);
      if (combined) heap.push(Entry{cur.q, cur.c, cur.key, a, nb});
    }
  }
  return merges;
}

}  // namespace ws
}  // namespace exa
