"""CPU restatement of the reference's affinity -> segmentation step -- TEST INFRASTRUCTURE ONLY.

Restates ``affinities_to_segmentation`` (REF/inference.py:196-237, REF =
src/aind_exaspim_neuron_segmentation) for BASELINE config 5: the product's affinities and the
oracle's affinities are both pushed through THIS function and the two segmentations are compared
with the adapted-Rand score.  It is never imported by the product (tests/test_layout.py).

**Parity unpinned.**  The arithmetic lives in a third-party dependency that is absent from
/root/reference and from this image: ``waterz @ git+https://github.com/anna-grim/waterz.git@master``
(pyproject.toml:32 -- a fork pinned to a branch, needs boost; call site REF/inference.py:224-229).
Neither the reference's tests nor this container can produce golden vectors for it, so this file
restates the published waterz algorithm:

1. *Fragments* (``waterz.agglomerate`` -> ``watershed(aff, low=0.1, high=0.9999)``, Zlateski &
   Seung's steepest-ascent watershed on the 6-neighbour affinity graph): edges with affinity
   < low are removed; every voxel keeps its strongest incident edge(s); edges >= high are always
   kept; the fragments are the connected components of the kept edges; voxels without any kept
   edge are background (0).  Deviation: exact ties / plateaus are merged into one fragment
   instead of being divided by the BFS of the original -- irrelevant for float32 sigmoid outputs
   and applied identically to both volumes under comparison.
2. *Region graph*: for every pair of touching fragments, the sum and count of the affinities on
   the faces between them.
3. *Agglomeration* with ``OneMinus<MeanAffinity<..>>`` scoring: repeatedly merge the pair with
   the smallest score ``1 - sum/count`` while it is below the threshold, adding up the statistics
   of parallel edges; the thresholds [0.6, 0.8, 0.9] are cumulative, so the last one decides.
4. ``remove_small_segments`` (REF/utils/img_util.py:536-559): keep ids with more than
   ``min_size`` voxels, zero the rest, renumber from 1 in order of first appearance.

Edge orientation: this repository trains ``aff[c][z,y,x]`` as the edge from voxel (z,y,x) to its
NEXT neighbour along axis c (REF/utils/img_util.py:160,207-216); the same convention is used
here.  Upstream waterz reads it as the edge to the PREVIOUS voxel -- a shift of the lattice by
one voxel that is applied identically to both volumes under comparison.
"""

import heapq

import numpy as np
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components


def _edge_lists(aff):
    """Flat (u, v, w) per axis for all in-volume edges u=(z,y,x) -> v=next voxel along the axis."""
    shape = aff.shape[1:]
    idx = np.arange(int(np.prod(shape)), dtype=np.int64).reshape(shape)
    out = []
    for c in range(3):
        sl_u = [slice(None)] * 3
        sl_v = [slice(None)] * 3
        sl_u[c] = slice(0, shape[c] - 1)
        sl_v[c] = slice(1, shape[c])
        u = idx[tuple(sl_u)].ravel()
        v = idx[tuple(sl_v)].ravel()
        w = aff[c][tuple(sl_u)].ravel()
        out.append((u, v, w))
    return out


def watershed_fragments(aff, low=0.1, high=0.9999):
    """Step 1.  aff: float32 (3, D, H, W) -> int64 fragment ids (0 = background), count."""
    aff = np.asarray(aff, dtype=np.float32)
    n = int(np.prod(aff.shape[1:]))
    edges = _edge_lists(aff)
    # strongest incident affinity of every voxel (edges below `low` do not exist)
    best = np.zeros(n, dtype=np.float32)
    for u, v, w in edges:
        wl = np.where(w >= low, w, 0).astype(np.float32)
        np.maximum.at(best, u, wl)
        np.maximum.at(best, v, wl)
    rows, cols = [], []
    for u, v, w in edges:
        ok = w >= low
        keep = ok & ((w >= high) | (w >= best[u]) | (w >= best[v]))
        rows.append(u[keep])
        cols.append(v[keep])
    rows = np.concatenate(rows)
    cols = np.concatenate(cols)
    graph = coo_matrix((np.ones(rows.size, dtype=np.int8), (rows, cols)), shape=(n, n))
    _, comp = connected_components(graph, directed=False)
    linked = np.zeros(n, dtype=bool)
    linked[rows] = True
    linked[cols] = True
    # isolated voxels are background; renumber the rest from 1
    comp = np.where(linked, comp, -1)
    ids, inv = np.unique(comp, return_inverse=True)
    if ids[0] == -1:
        frag = inv.astype(np.int64)          # -1 -> 0, others -> 1..
        count = ids.size - 1
    else:
        frag = inv.astype(np.int64) + 1
        count = ids.size
    return frag.reshape(aff.shape[1:]), count


def region_graph(aff, frag):
    """Step 2.  -> dict {(a, b) with a < b: [sum_affinity, n_faces]} over touching fragments."""
    aff = np.asarray(aff, dtype=np.float32)
    f = frag.ravel()
    stats = {}
    for u, v, w in _edge_lists(aff):
        a, b = f[u], f[v]
        m = (a != b) & (a != 0) & (b != 0)
        a, b, w = a[m], b[m], w[m].astype(np.float64)
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        key = lo * (int(f.max()) + 1) + hi
        order = np.argsort(key, kind="stable")
        key, w = key[order], w[order]
        uniq, start = np.unique(key, return_index=True)
        sums = np.add.reduceat(w, start) if key.size else np.array([])
        cnts = np.diff(np.append(start, key.size))
        base = int(f.max()) + 1
        for k, s, c in zip(uniq.tolist(), sums.tolist(), cnts.tolist()):
            e = (k // base, k % base)
            if e in stats:
                stats[e][0] += s
                stats[e][1] += c
            else:
                stats[e] = [s, c]
    return stats


def agglomerate(n_frag, stats, threshold):
    """Step 3.  Hierarchical merging with score 1 - mean affinity.  -> root id per fragment (1..n)."""
    parent = list(range(n_frag + 1))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    nbr = {i: {} for i in range(1, n_frag + 1)}   # node -> {neighbour: [sum, count]}
    heap = []
    for (a, b), (s, c) in stats.items():
        nbr[a][b] = [s, c]
        nbr[b][a] = nbr[a][b]
        heapq.heappush(heap, (1.0 - s / c, a, b, c))
    while heap:
        score, a, b, c = heapq.heappop(heap)
        if score >= threshold:
            break
        if parent[a] != a or parent[b] != b or b not in nbr[a] or nbr[a][b][1] != c:
            continue  # stale entry (an endpoint was merged away or the edge was updated)
        # merge the node with fewer neighbours into the other
        if len(nbr[a]) < len(nbr[b]):
            a, b = b, a
        parent[b] = a
        del nbr[a][b]
        for nb, st in nbr[b].items():
            if nb == a:
                continue
            del nbr[nb][b]
            if nb in nbr[a]:
                cur = nbr[a][nb]
                cur[0] += st[0]
                cur[1] += st[1]
            else:
                cur = [st[0], st[1]]
                nbr[a][nb] = cur
                nbr[nb][a] = cur
            heapq.heappush(heap, (1.0 - cur[0] / cur[1], min(a, nb), max(a, nb), cur[1]))
        nbr[b] = {}
    roots = np.arange(n_frag + 1, dtype=np.int64)
    for i in range(1, n_frag + 1):
        roots[i] = find(i)
    return roots


def remove_small_segments(seg, min_size):
    """Step 4: REF/utils/img_util.py:536-559 (fastremap restated with numpy)."""
    ids, cnts = np.unique(seg, return_counts=True)
    keep = ids[(cnts > min_size) & (ids != 0)]
    out = np.where(np.isin(seg, keep), seg, 0)
    flat = out.ravel()
    nz = flat[flat != 0]
    _, first = np.unique(nz, return_index=True)
    order = nz[np.sort(first)]                      # ids in order of first appearance
    lut = np.zeros(int(flat.max()) + 1 if flat.size else 1, dtype=np.int64)
    lut[order] = np.arange(1, order.size + 1)
    return lut[out]


def affinities_to_segmentation_ref(affinities, agglomeration_thresholds=(0.6, 0.8, 0.9),
                                   min_segment_size=100):
    """REF/inference.py:196-237 with the waterz call restated (see module docstring)."""
    aff = np.asarray(affinities, dtype=np.float32)
    frag, n = watershed_fragments(aff, low=0.1, high=0.9999)
    stats = region_graph(aff, frag)
    roots = agglomerate(n, stats, max(agglomeration_thresholds))
    seg = roots[frag]
    return remove_small_segments(seg, min_segment_size)


def adapted_rand_agreement(seg, ref):
    """1 - adapted Rand error of `seg` against `ref`, ignoring voxels where `ref` is background:
    2 * sum_ij p_ij^2 / (sum_i a_i^2 + sum_j b_j^2) over the contingency table (SURVEY.md 8c iv)."""
    seg = np.asarray(seg).ravel()
    ref = np.asarray(ref).ravel()
    m = ref != 0
    seg, ref = seg[m], ref[m]
    if seg.size == 0:
        return 1.0
    pair = ref.astype(np.int64) * (int(seg.max()) + 1) + seg.astype(np.int64)
    _, pij = np.unique(pair, return_counts=True)
    _, ai = np.unique(ref, return_counts=True)
    _, bj = np.unique(seg, return_counts=True)
    pij = pij.astype(np.float64)
    return float(2.0 * (pij ** 2).sum() / ((ai.astype(np.float64) ** 2).sum() + (bj.astype(np.float64) ** 2).sum()))
