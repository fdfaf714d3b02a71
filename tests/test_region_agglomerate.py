"""The exact host merge queue of K7 (exa_region_agglomerate with device = -1, csrc/ws_agglomerate.h)
against the oracle's sequential queue (oracle/ws_ref.cpp: wsref_agglomerate, "exact" statistics) --
no GPU needed.  tests/test_gpu_watershed.py repeats the random-graph cases through the parallel GPU
rounds."""

import ctypes

import numpy as np
import pytest
from scipy.ndimage import gaussian_filter


def native_roots(graph, n, threshold, device=-1):
    from aind_exaspim_neuron_segmentation_b200 import _native

    root = np.zeros(n + 1, dtype=np.uint32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    u, v = np.ascontiguousarray(graph["u"], np.uint32), np.ascontiguousarray(graph["v"], np.uint32)
    q, c = np.ascontiguousarray(graph["qsum"], np.uint64), np.ascontiguousarray(graph["count"], np.uint32)
    code = _native.lib().exa_region_agglomerate(device, n, u.size, p(u), p(v), p(q), p(c),
                                                float(threshold), p(root))
    _native.check(code, None, "exa_region_agglomerate")
    return root


def random_graph(rng, n, m, levels):
    """Arbitrary (non-grid) region graph with many exact score ties: affinities are multiples of
    1/levels (exact in 32.32 fixed point when levels is a power of two, nearly tied otherwise)."""
    pairs = set()
    while len(pairs) < m:
        a, b = (int(x) for x in rng.integers(1, n + 1, 2))
        if a != b:
            pairs.add((min(a, b), max(a, b)))
    pairs = sorted(pairs)
    cnt = rng.integers(1, 6, len(pairs)).astype(np.uint32)
    q = np.zeros(len(pairs), np.uint64)
    f = np.zeros(len(pairs), np.float32)
    for i, c in enumerate(cnt):
        vals = (rng.integers(0, levels + 1, int(c)) / levels).astype(np.float32)
        q[i] = sum(int(np.rint(np.float64(x) * 4294967296.0)) for x in vals)
        f[i] = np.float32(vals.sum())
    return dict(u=np.array([p[0] for p in pairs], np.uint32), v=np.array([p[1] for p in pairs], np.uint32),
                qsum=q, fsum=f, count=cnt)


def smooth_graph(shape, seed, quant):
    from oracle.watershed_ref import region_graph, watershed_fragments

    rng = np.random.default_rng(seed)
    f = np.stack([gaussian_filter(rng.normal(size=shape), 1.5) for _ in range(3)])
    aff = (1.0 / (1.0 + np.exp(-5.0 * f / f.std()))).astype(np.float32)
    if quant:
        aff = (np.round(aff * quant) / quant).astype(np.float32)
    frag, n = watershed_fragments(aff)
    return frag, n, region_graph(aff, frag)


GRID_CASES = [((20, 24, 28), 1, None, 0.9), ((32, 32, 40), 2, None, 0.6), ((24, 24, 24), 3, 20, 0.9),
              ((24, 24, 24), 4, 5, 0.8), ((16, 40, 24), 5, None, 0.3), ((16, 16, 16), 6, None, 2.0),
              ((16, 16, 16), 7, 4, 0.0)]


@pytest.mark.parametrize("shape,seed,quant,threshold", GRID_CASES)
def test_merge_queue_equals_oracle(shape, seed, quant, threshold):
    from oracle.watershed_ref import agglomerate

    frag, n, graph = smooth_graph(shape, seed, quant)
    ref = agglomerate(n, graph, threshold)
    got = native_roots(graph, n, threshold)
    # both report the smallest fragment id of every region: the arrays agree element for element
    assert got[0] == 0 and np.array_equal(got.astype(np.int64), ref)
    if threshold <= 0.0:
        assert np.array_equal(got, np.arange(n + 1))


def test_merge_queue_rejects_bad_edges_and_handles_empty_graph():
    bad = dict(u=np.array([2], np.uint32), v=np.array([1], np.uint32),      # u must be < v
               qsum=np.array([1 << 31], np.uint64), count=np.array([1], np.uint32))
    with pytest.raises(RuntimeError):
        native_roots(bad, 3, 0.9)
    too_big = dict(u=np.array([1], np.uint32), v=np.array([2], np.uint32),  # mean affinity > 1
                   qsum=np.array([(1 << 33)], np.uint64), count=np.array([1], np.uint32))
    with pytest.raises(RuntimeError):
        native_roots(too_big, 3, 0.9)
    empty = dict(u=np.zeros(0, np.uint32), v=np.zeros(0, np.uint32), qsum=np.zeros(0, np.uint64),
                 count=np.zeros(0, np.uint32))
    assert np.array_equal(native_roots(empty, 4, 0.9), np.arange(5))


def test_merge_queue_equals_oracle_on_random_graphs():
    """Few distinct affinity values, small counts, dense and sparse graphs, thresholds inside and
    outside the score range."""
    from oracle.watershed_ref import agglomerate

    rng = np.random.default_rng(7)
    for trial in range(80):
        n = int(rng.integers(2, 60))
        m = int(rng.integers(1, min(n * (n - 1) // 2, 4 * n) + 1))
        graph = random_graph(rng, n, m, int(rng.integers(2, 9)))
        threshold = float(rng.choice([0.0, 0.25, 0.5, 0.75, 0.9, 1.0, 1.5]))
        ref = agglomerate(n, graph, threshold)
        got = native_roots(graph, n, threshold)
        assert np.array_equal(got.astype(np.int64), ref), (trial, n, m, threshold)
