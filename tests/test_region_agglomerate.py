"""The host merge queue of K7 (exa_region_agglomerate, csrc/ws_agglomerate.h) against the oracle's
agglomerate() on region graphs built by the oracle -- no GPU needed."""

import ctypes

import numpy as np
import pytest
from scipy.ndimage import gaussian_filter


def region_graph_arrays(aff):
    from oracle.watershed_ref import region_graph, watershed_fragments

    frag, n = watershed_fragments(aff)
    stats = region_graph(aff, frag)
    keys = np.array([(a << 32) | b for a, b in stats], dtype=np.uint64)
    sums = np.array([v[0] for v in stats.values()], dtype=np.float64)
    cnts = np.array([v[1] for v in stats.values()], dtype=np.int32)
    return n, stats, keys, sums, cnts


def native_roots(n, keys, sums, cnts, threshold):
    from aind_exaspim_neuron_segmentation_b200 import _native

    root = np.zeros(n + 1, dtype=np.uint32)
    code = _native.lib().exa_region_agglomerate(
        n, keys.size, keys.ctypes.data_as(ctypes.c_void_p), sums.ctypes.data_as(ctypes.c_void_p),
        cnts.ctypes.data_as(ctypes.c_void_p), float(threshold), root.ctypes.data_as(ctypes.c_void_p))
    _native.check(code, None, "exa_region_agglomerate")
    return root


def same_partition(a, b):
    """Root labels are arbitrary representatives: compare the partitions they induce."""
    _, ia = np.unique(a, return_inverse=True)
    _, ib = np.unique(b, return_inverse=True)
    pairs = np.unique(np.stack([ia, ib]), axis=1).shape[1]
    return pairs == ia.max() + 1 == ib.max() + 1


@pytest.mark.parametrize("shape,seed,quant,threshold", [
    ((20, 24, 28), 1, None, 0.9), ((32, 32, 40), 2, None, 0.6), ((24, 24, 24), 3, 20, 0.9),
    ((24, 24, 24), 4, 5, 0.8), ((16, 40, 24), 5, None, 0.3), ((16, 16, 16), 6, None, 2.0)])
def test_merge_queue_equals_oracle(shape, seed, quant, threshold):
    from oracle.watershed_ref import agglomerate

    rng = np.random.default_rng(seed)
    f = np.stack([gaussian_filter(rng.normal(size=shape), 1.5) for _ in range(3)])
    aff = (1.0 / (1.0 + np.exp(-5.0 * f / f.std()))).astype(np.float32)
    if quant:
        aff = (np.round(aff * quant) / quant).astype(np.float32)
    n, stats, keys, sums, cnts = region_graph_arrays(aff)
    ref = agglomerate(n, stats, threshold)
    got = native_roots(n, keys, sums, cnts, threshold)
    assert got[0] == 0 and same_partition(got[1:], ref[1:])
    # the surviving representative is chosen by the same rule, so even the labels agree
    assert np.array_equal(got.astype(np.int64), ref)


def test_merge_queue_rejects_bad_edges_and_handles_empty_graph():
    keys = np.array([(2 << 32) | 1], dtype=np.uint64)   # a must be < b
    with pytest.raises(RuntimeError):
        native_roots(3, keys, np.ones(1), np.ones(1, np.int32), 0.9)
    root = native_roots(4, np.zeros(0, np.uint64), np.zeros(0), np.zeros(0, np.int32), 0.9)
    assert np.array_equal(root, np.arange(5))


def test_merge_queue_equals_oracle_on_random_graphs():
    """Arbitrary (non-grid) region graphs with many exact score ties: few distinct affinity values,
    small counts, dense and sparse graphs, thresholds inside and outside the score range."""
    from oracle.watershed_ref import agglomerate

    rng = np.random.default_rng(7)
    for trial in range(60):
        n = int(rng.integers(2, 60))
        m = int(rng.integers(1, min(n * (n - 1) // 2, 4 * n) + 1))
        pairs = set()
        while len(pairs) < m:
            a, b = (int(v) for v in rng.integers(1, n + 1, 2))
            if a != b:
                pairs.add((min(a, b), max(a, b)))
        levels = rng.integers(2, 9)
        stats = {}
        for a, b in sorted(pairs):
            c = int(rng.integers(1, 6))
            # sums of float32 values that are multiples of 1/levels: exact in float64, many ties
            vals = (rng.integers(0, levels + 1, c) / levels).astype(np.float32).astype(np.float64)
            stats[(a, b)] = [float(vals.sum()), c]
        keys = np.array([(a << 32) | b for a, b in stats], dtype=np.uint64)
        sums = np.array([v[0] for v in stats.values()], dtype=np.float64)
        cnts = np.array([v[1] for v in stats.values()], dtype=np.int32)
        threshold = float(rng.choice([0.0, 0.25, 0.5, 0.75, 0.9, 1.0, 1.5]))
        ref = agglomerate(n, {k: list(v) for k, v in stats.items()}, threshold)
        got = native_roots(n, keys, sums, cnts, threshold)
        assert np.array_equal(got.astype(np.int64), ref), (trial, n, m, threshold)
