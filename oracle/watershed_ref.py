"""CPU restatement of the reference's affinity -> segmentation step -- TEST INFRASTRUCTURE ONLY.

Restates ``affinities_to_segmentation`` (REF/inference.py:196-237, REF =
src/aind_exaspim_neuron_segmentation): ``waterz.agglomerate(affinities, thresholds,
aff_threshold_low=0.1, aff_threshold_high=0.9999)`` followed by ``remove_small_segments``
(REF/utils/img_util.py:536-559).  Never imported by the product (tests/test_layout.py).

**Parity unpinned.**  The arithmetic lives in a third-party dependency that is absent from
/root/reference and from this image: ``waterz @ git+https://github.com/anna-grim/waterz.git@master``
(pyproject.toml:32 -- a fork pinned to a branch, needs boost).  Neither the reference's tests nor
this container can produce golden vectors for it, so ``oracle/ws_ref.cpp`` restates the published
waterz algorithm step by step (see its header for the sources followed); this module binds it with
ctypes, adds ``remove_small_segments`` and the adapted-Rand score, and keeps a literal pure-Python
transcription of the watershed (``watershed_fragments_py``, small volumes only) that the compiled
one is checked against.

waterz conventions restated here (they differ from round 1 of this repository, which read
``aff[c][z,y,x]`` as the edge to the NEXT voxel, kept edges with ``>= low`` and merged plateaus):

* ``aff[c][z,y,x]`` is the edge between voxel (z,y,x) and its PREVIOUS neighbour along axis c
  (c = 0, 1, 2 for z, y, x); index 0 along c has no such edge.
* A voxel joins the watershed only if its largest incident affinity is ``> low`` (strict); an
  edge is followed if it equals that maximum or is ``>= high``.
* Plateaus (voxels whose steepest edges point at each other) are divided breadth first from their
  exits, in voxel scan order; every divided voxel keeps exactly one outgoing edge.
* Scores are ``1 - mean affinity``; ``accumulate="float32"`` sums in float32 in scan / merge order
  like waterz's MeanAffinityProvider, ``accumulate="exact"`` (default, what the CUDA path
  implements) sums 32.32 fixed-point values, which makes the result independent of the order of
  additions.  Equal scores are ordered by the smallest rank of the original edges involved
  (waterz leaves ties to std::priority_queue).
"""

import ctypes

import numpy as np

from . import build as _build

_lib = None


def _ws():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(_build.build())
        vp, i64, f32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_float
        lib.wsref_watershed.restype = i64
        lib.wsref_watershed.argtypes = [vp, i64, i64, i64, f32, f32, vp]
        lib.wsref_region_graph.restype = i64
        lib.wsref_region_graph.argtypes = [vp, vp, i64, i64, i64, vp, vp, vp, vp, vp]
        lib.wsref_agglomerate.restype = i64
        lib.wsref_agglomerate.argtypes = [ctypes.c_uint32, i64, vp, vp, vp, vp, vp, ctypes.c_double,
                                          ctypes.c_int, vp]
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def watershed_fragments(aff, low=0.1, high=0.9999):
    """Step 1 (ws_ref.cpp: wsref_watershed).  float32 (3, D, H, W) -> (int64 ids, count)."""
    aff = np.ascontiguousarray(aff, dtype=np.float32)
    d, h, w = aff.shape[1:]
    seg = np.zeros((d, h, w), dtype=np.uint64)
    n = _ws().wsref_watershed(_p(aff), d, h, w, low, high, _p(seg))
    return seg.astype(np.int64), int(n)


def watershed_fragments_py(aff, low=0.1, high=0.9999):
    """The same algorithm transcribed literally in pure Python (small volumes only): the compiled
    restatement is checked against it in tests/test_watershed_oracle.py."""
    aff = np.asarray(aff, dtype=np.float32)
    D, H, W = aff.shape[1:]
    hw, n = H * W, D * H * W
    low, high = np.float32(low), np.float32(high)
    az, ay, ax = (aff[c].ravel() for c in range(3))
    seg = [0] * n
    for z in range(D):
        for y in range(H):
            for x in range(W):
                i = z * hw + y * W + x
                vals = [az[i] if z > 0 else low, ay[i] if y > 0 else low, ax[i] if x > 0 else low,
                        az[i + hw] if z < D - 1 else low, ay[i + W] if y < H - 1 else low,
                        ax[i + 1] if x < W - 1 else low]
                m = max(vals)
                if m > low:
                    for d, v in enumerate(vals):
                        if v == m or v >= high:
                            seg[i] |= 1 << d
    dirs = [-hw, -W, -1, hw, W, 1]
    mask = [1, 2, 4, 8, 16, 32]
    imask = [8, 16, 32, 1, 2, 4]
    VIS, HIGH = 0x40, 1 << 63
    bfs = []
    for i in range(n):
        for d in range(6):
            if seg[i] & mask[d] and not seg[i + dirs[d]] & imask[d]:
                seg[i] |= VIS
                bfs.append(i)
                break
    k = 0
    while k < len(bfs):
        i = bfs[k]
        to_set = 0
        for d in range(6):
            if seg[i] & mask[d]:
                j = i + dirs[d]
                if seg[j] & imask[d]:
                    if not seg[j] & VIS:
                        bfs.append(j)
                        seg[j] |= VIS
                else:
                    to_set = mask[d]
        seg[i] = to_set
        k += 1
    next_id = 1
    for i in range(n):
        if seg[i] == 0:
            seg[i] |= HIGH
        if not seg[i] & HIGH and seg[i]:
            bfs = [i]
            seg[i] |= VIS
            k = 0
            while k < len(bfs):
                me = bfs[k]
                joined = False
                for d in range(6):
                    if seg[me] & mask[d]:
                        him = me + dirs[d]
                        if seg[him] & HIGH:
                            for v in bfs:
                                seg[v] = seg[him]
                            bfs = []
                            joined = True
                            break
                        if not seg[him] & VIS:
                            seg[him] |= VIS
                            bfs.append(him)
                if joined:
                    break
                k += 1
            if bfs:
                for v in bfs:
                    seg[v] = HIGH | next_id
                next_id += 1
    out = np.array([s & ~HIGH for s in seg], dtype=np.int64).reshape(D, H, W)
    return out, next_id - 1


def region_graph(aff, frag):
    """Step 2 (wsref_region_graph).  -> dict of arrays: u, v (u < v, sorted), fsum (float32 sums in
    scan order), qsum (exact 32.32 fixed-point sums), count."""
    aff = np.ascontiguousarray(aff, dtype=np.float32)
    seg = np.ascontiguousarray(frag, dtype=np.uint64)
    d, h, w = seg.shape
    m = _ws().wsref_region_graph(_p(aff), _p(seg), d, h, w, None, None, None, None, None)
    out = dict(u=np.zeros(m, np.uint32), v=np.zeros(m, np.uint32), fsum=np.zeros(m, np.float32),
               qsum=np.zeros(m, np.uint64), count=np.zeros(m, np.uint32))
    got = _ws().wsref_region_graph(_p(aff), _p(seg), d, h, w, _p(out["u"]), _p(out["v"]),
                                   _p(out["fsum"]), _p(out["qsum"]), _p(out["count"]))
    assert got == m
    return out


def agglomerate(n_frag, graph, threshold, accumulate="exact"):
    """Step 3 (wsref_agglomerate).  -> int64 root id per fragment (index 0 = background)."""
    mode = {"float32": 0, "exact": 1}[accumulate]
    m = int(graph["u"].size)
    order = np.lexsort((graph["v"], graph["u"]))
    assert np.array_equal(order, np.arange(m)), "edges must be sorted by (u, v)"
    root = np.zeros(n_frag + 1, dtype=np.uint32)
    _ws().wsref_agglomerate(n_frag, m, _p(graph["u"]), _p(graph["v"]), _p(graph["fsum"]),
                            _p(graph["qsum"]), _p(graph["count"]), float(threshold), mode, _p(root))
    return root.astype(np.int64)


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
                                   min_segment_size=100, accumulate="exact"):
    """REF/inference.py:196-237 with the waterz call restated (see module docstring).  The
    thresholds are cumulative and the segmentation of the LAST one is kept (inference.py:232)."""
    aff = np.asarray(affinities, dtype=np.float32)
    frag, n = watershed_fragments(aff, low=0.1, high=0.9999)
    graph = region_graph(aff, frag)
    roots = agglomerate(n, graph, max(agglomeration_thresholds), accumulate)
    return remove_small_segments(roots[frag], min_segment_size)


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
