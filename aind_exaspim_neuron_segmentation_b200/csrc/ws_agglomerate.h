// Host merge queue of K7 (affinities -> segmentation): hierarchical agglomeration of the region
// graph, reference inference.py:224-229 -> waterz.agglomerate with OneMinus<MeanAffinity> scoring.
// Plain C++ (no CUDA): also exported as exa_region_agglomerate for CPU tests.
#pragma once

#include <stdint.h>

#include <algorithm>
#include <tuple>
#include <vector>

namespace exa {
namespace ws {

// Hierarchical agglomeration with OneMinus<MeanAffinity> scoring (waterz): merge the pair with the
// smallest score while it is below the threshold; statistics of parallel edges add up.  Entries of
// the heap are (score, a, b, count) compared lexicographically; an entry is stale when one of its
// ends was merged away or the edge's count changed since it was pushed.
struct Stat {
  double s;
  long long c;
};
using Entry = std::tuple<double, uint32_t, uint32_t, long long>;

// neighbour table of one region: open addressing on fragment ids (>= 1), Fibonacci hashing
// (the HIGH bits of k * 2^32/phi: neighbouring fragments have nearby ids), linear probing, at most
// half full including tombstones
class NbrMap {
 public:
  uint32_t size() const { return live_; }
  void reserve(uint32_t n) {
    uint32_t lg = 2;
    while ((1u << lg) < 2 * (n + 1)) ++lg;
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
  void put(uint32_t k, const Stat& v) {
    if (Stat* p = find(k)) {
      *p = v;
      return;
    }
    if (2 * ((uint64_t)used_ + 1) > slot_.size()) {
      uint32_t lg = 2;
      while ((1u << lg) < 4 * (live_ + 1)) ++lg;  // a quarter full after the clean-up
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
    std::vector<Slot> old(1u << lg, Slot{Stat{0.0, 0}, kEmpty});
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

// Monotone bucket queue over Entry: scores only grow along the merge sequence (a merged edge's
// mean affinity lies between those of its two parts, which were both >= the current minimum), so
// entries are binned by score and only the small heap of the current bin is ever touched -- the
// pop order is exactly that of one global min-heap.  Entries at or above the threshold are never
// popped before the loop ends and are not stored at all.
class BucketQueue {
 public:
  BucketQueue(double threshold, size_t n_hint) : thr_(threshold) {
    size_t nb = 64;
    while (nb < (1u << 16) && nb * 8 < n_hint) nb <<= 1;
    bin_.resize(nb);
    scale_ = threshold > 0 ? (double)nb / threshold : 0.0;
  }
  // the initial edges at once: bins sized exactly, then heapified (no per-entry reallocation)
  void bulk_load(const std::vector<Entry>& all) {
    std::vector<uint32_t> n(bin_.size(), 0);
    for (const Entry& e : all)
      if (std::get<0>(e) < thr_) ++n[bin_of(std::get<0>(e))];
    for (size_t i = 0; i < bin_.size(); ++i) bin_[i].reserve(n[i] + n[i] / 2);
    for (const Entry& e : all)
      if (std::get<0>(e) < thr_) bin_[bin_of(std::get<0>(e))].push_back(e);
    for (std::vector<Entry>& h : bin_) std::make_heap(h.begin(), h.end(), std::greater<Entry>());
    cur_ = 0;
  }
  void push(const Entry& e) {
    const double sc = std::get<0>(e);
    if (!(sc < thr_)) return;
    const size_t i = bin_of(sc);
    if (i < cur_) cur_ = i;  // rounding put a merged score one ulp under the minimum
    bin_[i].push_back(e);
    std::push_heap(bin_[i].begin(), bin_[i].end(), std::greater<Entry>());
  }
  bool pop(Entry* e) {
    while (cur_ < bin_.size() && bin_[cur_].empty()) ++cur_;
    if (cur_ == bin_.size()) return false;
    std::vector<Entry>& h = bin_[cur_];
    std::pop_heap(h.begin(), h.end(), std::greater<Entry>());
    *e = h.back();
    h.pop_back();
    return true;
  }

 private:
  size_t bin_of(double sc) const {
    return sc <= 0 ? 0 : std::min((size_t)(sc * scale_), bin_.size() - 1);
  }
  double thr_, scale_;
  size_t cur_ = 0;
  std::vector<std::vector<Entry>> bin_;
};

std::vector<uint32_t> agglomerate(uint32_t n_frag, const std::vector<unsigned long long>& keys,
                                  const std::vector<double>& sums, const std::vector<int>& counts,
                                  double threshold) {
  std::vector<uint32_t> parent(n_frag + 1);
  for (uint32_t i = 0; i <= n_frag; ++i) parent[i] = i;
  std::vector<NbrMap> nbr(n_frag + 1);
  {
    std::vector<uint32_t> deg(n_frag + 1, 0);
    for (unsigned long long k : keys) {
      ++deg[(uint32_t)(k >> 32)];
      ++deg[(uint32_t)(k & 0xffffffffu)];
    }
    for (uint32_t i = 1; i <= n_frag; ++i)
      if (deg[i]) nbr[i].reserve(deg[i]);
  }
  BucketQueue heap(threshold, keys.size());
  {
    std::vector<Entry> init;
    init.reserve(keys.size());
    for (size_t i = 0; i < keys.size(); ++i) {
      const uint32_t a = (uint32_t)(keys[i] >> 32), b = (uint32_t)(keys[i] & 0xffffffffu);
      const Stat st{sums[i], (long long)counts[i]};
      nbr[a].put(b, st);
      nbr[b].put(a, st);
      init.emplace_back(1.0 - st.s / (double)st.c, a, b, st.c);
    }
    heap.bulk_load(init);
  }
  std::vector<std::pair<uint32_t, Stat>> moved;
  Entry e;
  while (heap.pop(&e)) {
    uint32_t a = std::get<1>(e), b = std::get<2>(e);
    if (parent[a] != a || parent[b] != b) continue;
    const Stat* cur_ab = nbr[a].find(b);
    if (cur_ab == nullptr || cur_ab->c != std::get<3>(e)) continue;
    if (nbr[a].size() < nbr[b].size()) std::swap(a, b);  // the node with fewer neighbours goes away
    parent[b] = a;
    nbr[a].erase(b);
    moved.clear();
    nbr[b].for_each([&](uint32_t nb, const Stat& st) {
      if (nb != a) moved.emplace_back(nb, st);
    });
    nbr[b].clear();
    for (const auto& kv : moved) {
      const uint32_t nb = kv.first;
      nbr[nb].erase(b);
      Stat cur = kv.second;
      if (const Stat* f = nbr[a].find(nb)) {
        cur.s = f->s + kv.second.s;
        cur.c = f->c + kv.second.c;
      }
      nbr[a].put(nb, cur);
      nbr[nb].put(a, cur);
      heap.push(Entry(1.0 - cur.s / (double)cur.c, std::min(a, nb), std::max(a, nb), cur.c));
    }
  }
  std::vector<uint32_t> root(n_frag + 1);
  for (uint32_t i = 0; i <= n_frag; ++i) {
    uint32_t x = i;
    while (parent[x] != x) x = parent[x];
    root[i] = x;
  }
  return root;
}

}  // namespace ws
}  // namespace exa
