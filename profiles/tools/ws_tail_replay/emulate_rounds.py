"""numpy emulation of the parallel agglomeration rounds of csrc/watershed.cu (merge every edge that
is the best edge of both ends, or the best edge of one end whose other neighbours all have a later
best edge), up to the product's hand-over rule (a round with fewer than 8 + live/40000 merges).
Writes the contracted graph the host queue would receive as a flat binary for harness.cpp:
python emulate_rounds.py GRAPH.npz TAIL.bin [threshold].  Scores are compared in float64 here (the
product compares 96-bit products), which is enough for a representative tail; never used by the
product or the tests."""
import sys

import numpy as np

d = np.load(sys.argv[1])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.9
n = int(d["n"])
u, v = d["u"].astype(np.int64), d["v"].astype(np.int64)
q, c, k = d["q"].astype(np.uint64), d["c"].astype(np.uint64), d["k"].astype(np.int64)
T = np.uint64(int(np.rint((1 - thr) * 4294967296.0)))
BIG = np.iinfo(np.int64).max
rounds = 0
while True:
    m = u.size
    order = np.lexsort((k, -(q.astype(np.float64) / c.astype(np.float64))))
    rank = np.empty(m, np.int64)
    rank[order] = np.arange(m)
    cand = q > c * T
    best = np.full(n + 1, BIG, np.int64)
    np.minimum.at(best, u, rank)
    np.minimum.at(best, v, rank)
    bu, bv = best[u] == rank, best[v] == rank
    other = np.full(n + 1, BIG, np.int64)       # earliest best edge among the other neighbours
    np.minimum.at(other, u[~bu], best[v[~bu]])
    np.minimum.at(other, v[~bv], best[u[~bv]])
    both = bu & bv & cand
    one_u = bu & ~bv & cand & (other[u] > rank)
    one_v = bv & ~bu & cand & (other[v] > rank)
    merges = int(both.sum() + one_u.sum() + one_v.sum())
    if merges < 8 + m / 40000:
        break
    parent = np.arange(n + 1, dtype=np.int64)
    parent[u[both]] = v[both]
    parent[u[one_u]] = v[one_u]
    parent[v[one_v]] = u[one_v]
    while True:
        p2 = parent[parent]
        if np.array_equal(p2, parent):
            break
        parent = p2
    nu, nv = parent[u], parent[v]
    keep = nu != nv
    a, b = np.minimum(nu, nv)[keep], np.maximum(nu, nv)[keep]
    pair = a * (n + 1) + b
    o = np.argsort(pair, kind="stable")
    pair, a, b = pair[o], a[o], b[o]
    first = np.flatnonzero(np.r_[True, pair[1:] != pair[:-1]])
    u, v = a[first], b[first]
    q = np.add.reduceat(q[keep][o], first)
    c = np.add.reduceat(c[keep][o], first)
    k = np.minimum.reduceat(k[keep][o], first)
    rounds += 1
ids, inv = np.unique(np.r_[u, v], return_inverse=True)
a, b = inv[:u.size] + 1, inv[u.size:] + 1
print("hand-over after", rounds, "rounds:", ids.size, "regions,", u.size, "edges")
with open(sys.argv[2], "wb") as f:
    np.array([ids.size, u.size], np.uint64).tofile(f)
    np.minimum(a, b).astype(np.uint32).tofile(f)
    np.maximum(a, b).astype(np.uint32).tofile(f)
    q.astype(np.uint64).tofile(f)
    c.astype(np.uint32).tofile(f)
    k.astype(np.uint32).tofile(f)
