// Times exa::ws::agglomerate (csrc/ws_agglomerate.h, or a variant of it given as -DHDR='"..."') on
// a graph file written by emulate_rounds.py and prints a hash of the partition, so that variants
// can be compared for speed and for identical results without a GPU:
//   g++ -O2 -std=c++17 -DHDR='"../../../aind_exaspim_neuron_segmentation_b200/csrc/ws_agglomerate.h"' \
//       harness.cpp -o replay && ./replay TAIL.bin [threshold] [repetitions]
#include HDR
#include <chrono>
#include <cstdio>
#include <cstdlib>

template <typename V>
static void get(V& v, FILE* f) {
  if (fread(v.data(), sizeof(v[0]), v.size(), f) != v.size()) exit(2);
}

int main(int argc, char** argv) {
  FILE* f = fopen(argv[1], "rb");
  std::vector<uint64_t> hd(2);
  get(hd, f);
  const uint32_t n = (uint32_t)hd[0];
  const size_t m = hd[1];
  std::vector<uint32_t> u(m), v(m), c(m), k(m);
  std::vector<uint64_t> q(m);
  get(u, f), get(v, f), get(q, f), get(c, f), get(k, f);
  fclose(f);
  const int64_t T = exa::ws::fixed_threshold(argc > 2 ? atof(argv[2]) : 0.9);
  for (int r = 0, reps = argc > 3 ? atoi(argv[3]) : 3; r < reps; ++r) {
    std::vector<uint32_t> parent(n + 1);
    for (uint32_t i = 0; i <= n; ++i) parent[i] = i;
    const auto t0 = std::chrono::steady_clock::now();
    const int64_t merges = exa::ws::agglomerate(n, m, u.data(), v.data(), q.data(), c.data(),
                                                k.data(), T, parent.data());
    const auto t1 = std::chrono::steady_clock::now();
    std::vector<uint32_t> top(n + 1), least(n + 1, 0xffffffffu);
    for (uint32_t i = 0; i <= n; ++i) {
      uint32_t x = i;
      while (parent[x] != x) x = parent[x];
      top[i] = x;
      if (i < least[x]) least[x] = i;
    }
    uint64_t h = 1469598103934665603ull;  // FNV-1a over the smallest member of every region
    for (uint32_t i = 0; i <= n; ++i) h = (h ^ least[top[i]]) * 1099511628211ull;
    printf("regions %u edges %zu merges %lld  %.1f ms  partition %016llx\n", n, m, (long long)merges,
           std::chrono::duration<double, std::milli>(t1 - t0).count(), (unsigned long long)h);
  }
}
