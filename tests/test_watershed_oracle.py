"""Known-answer tests of the CPU restatement of affinities_to_segmentation (oracle/watershed_ref.py,
oracle/ws_ref.cpp).

Parity unpinned: waterz is not available (see the module docstring), so these tests pin the
restatement's own contract: hand-constructed affinity volumes with known answers, the compiled
watershed against its literal pure-Python transcription, and the two accumulation modes against
each other.
"""

import numpy as np
import pytest
from scipy.ndimage import gaussian_filter

from oracle.watershed_ref import (adapted_rand_agreement, affinities_to_segmentation_ref, agglomerate,
                                  region_graph, remove_small_segments, watershed_fragments,
                                  watershed_fragments_py)


def _two_blocks(shape=(12, 16, 16), wall=8, inside=0.95, across=0.02, seed=0):
    """Two bright blocks separated by a wall of weak z-edges between planes wall-1 and wall."""
    rng = np.random.default_rng(seed)
    aff = (inside + 0.04 * rng.random((3,) + shape)).astype(np.float32)
    aff[0, wall] = across          # aff[0][z] is the edge between plane z and plane z-1 (waterz)
    return aff


def test_two_blocks_are_two_segments():
    aff = _two_blocks()
    seg = affinities_to_segmentation_ref(aff, min_segment_size=10)
    assert seg.shape == aff.shape[1:] and seg.dtype.kind == "i"
    assert set(np.unique(seg)) == {1, 2}
    assert np.all(seg[:8] == seg[0, 0, 0]) and np.all(seg[8:] == seg[-1, 0, 0])
    assert seg[0, 0, 0] == 1        # renumbered in order of first appearance


def test_threshold_merges_medium_boundaries():
    # boundary affinity 0.3 -> score 0.7: merged by the 0.9 threshold, kept apart at 0.6
    aff = _two_blocks(across=0.3)
    merged = affinities_to_segmentation_ref(aff, agglomeration_thresholds=(0.6, 0.8, 0.9), min_segment_size=10)
    apart = affinities_to_segmentation_ref(aff, agglomeration_thresholds=(0.6,), min_segment_size=10)
    assert len(np.unique(merged)) == 1 and len(np.unique(apart)) == 2


def test_low_affinity_voxels_are_background_and_small_segments_removed():
    aff = _two_blocks()
    aff[:, :, :, 12:] = 0.01            # nothing links these voxels (x >= 12) to anything: background
    seg = affinities_to_segmentation_ref(aff, min_segment_size=10)
    assert np.all(seg[:, :, 12:] == 0) and seg[:, :, :12].min() >= 1
    frag, n = watershed_fragments(aff)
    assert n >= 2 and frag[:, :, 12:].max() == 0
    # size filter: ids with <= min_size voxels vanish, the rest are renumbered from 1
    lab = np.zeros((4, 4, 4), np.int64)
    lab[:2] = 7
    lab[3, 3, 3] = 9
    out = remove_small_segments(lab, 5)
    assert set(np.unique(out)) == {0, 1} and out[0, 0, 0] == 1 and out[3, 3, 3] == 0


def test_edge_convention_and_strict_low():
    """aff[c][i] joins voxel i and voxel i-1 along c; index 0 has no edge; m must be > low."""
    aff = np.zeros((3, 1, 1, 6), np.float32)
    # x-edges: (0-1) 0.5, (1-2) 0.1, (2-3) 0.05, (3-4) 0.8, (4-5) 0.1; the 0.9 at index 0 joins nothing
    aff[2, 0, 0] = [0.9, 0.5, 0.1, 0.05, 0.8, 0.1]
    frag, n = watershed_fragments(aff)
    # voxel 2's best edge is exactly low (0.1): not > low, so it is background, and so is voxel 5
    assert frag.ravel().tolist() == [1, 1, 0, 2, 2, 0] and n == 2


def test_plateau_is_divided_breadth_first_from_its_exits():
    """A 1-D plateau of equal edges with a stronger exit at each end is split in the middle, as the
    literal transcription of the algorithm says."""
    w = 9
    aff = np.zeros((3, 1, 1, w), np.float32)
    aff[2, 0, 0, 1:] = 0.5
    aff[2, 0, 0, 1] = 0.9        # edge 0-1
    aff[2, 0, 0, w - 1] = 0.8    # edge (w-2)-(w-1)
    frag, n = watershed_fragments(aff)
    ref, n_py = watershed_fragments_py(aff)
    assert n == n_py == 2 and np.array_equal(frag, ref)
    f = frag.ravel()
    assert f[0] == f[1] == 1 and f[-1] == f[-2] == 2
    assert np.all(np.diff(f) >= 0)                      # one cut
    assert 3 <= int((f == 1).sum()) <= 6                # somewhere in the middle


@pytest.mark.parametrize("shape,seed,quant", [((6, 7, 8), 1, 4), ((5, 9, 6), 2, 3), ((8, 8, 8), 3, 8),
                                              ((4, 12, 10), 4, 2), ((7, 7, 7), 5, None)])
def test_compiled_watershed_equals_python_transcription(shape, seed, quant):
    rng = np.random.default_rng(seed)
    f = np.stack([gaussian_filter(rng.normal(size=shape), 1.0) for _ in range(3)])
    aff = (1.0 / (1.0 + np.exp(-3.0 * f / f.std()))).astype(np.float32)
    if quant:
        aff = (np.round(aff * quant) / quant).astype(np.float32)   # plateaus, 0.0 and 1.0 (>= high)
    got, n = watershed_fragments(aff)
    ref, n_py = watershed_fragments_py(aff)
    assert n == n_py and np.array_equal(got, ref)
    # ids in order of first appearance
    first = [int(np.argmax(got.ravel() == i)) for i in range(1, n + 1)]
    assert first == sorted(first)


def test_region_graph_statistics():
    aff = _two_blocks(across=0.25)
    frag, n = watershed_fragments(aff)
    g = region_graph(aff, frag)
    assert np.all(g["u"] < g["v"]) and g["u"].min() >= 1 and g["v"].max() <= n
    assert np.array_equal(np.lexsort((g["v"], g["u"])), np.arange(g["u"].size))
    # all faces across the wall carry 0.25 exactly
    top = set(np.unique(frag[7]).tolist()) - {0}
    bottom = set(np.unique(frag[8]).tolist()) - {0}
    across = [i for i in range(g["u"].size)
              if (int(g["u"][i]) in top and int(g["v"][i]) in bottom)
              or (int(g["v"][i]) in top and int(g["u"][i]) in bottom)]
    assert across and sum(int(g["count"][i]) for i in across) == 16 * 16
    for i in across:
        assert int(g["qsum"][i]) == int(g["count"][i]) << 30
        assert g["fsum"][i] == np.float32(0.25) * g["count"][i]


def test_float32_and_exact_accumulation_agree():
    """waterz sums float32 affinities in scan/merge order; the exact mode (what the GPU implements)
    sums fixed-point values.  The two differ by float32 rounding only."""
    rng = np.random.default_rng(11)
    shape = (32, 40, 36)
    f = np.stack([gaussian_filter(rng.normal(size=shape), 2.0) for _ in range(3)])
    aff = (1.0 / (1.0 + np.exp(-6.0 * f / f.std()))).astype(np.float32)
    frag, n = watershed_fragments(aff)
    g = region_graph(aff, frag)
    for thr in (0.6, 0.9):
        a = agglomerate(n, g, thr, "float32")[frag]
        b = agglomerate(n, g, thr, "exact")[frag]
        assert adapted_rand_agreement(b, a) >= 0.99


def test_adapted_rand_properties():
    rng = np.random.default_rng(1)
    ref = rng.integers(0, 6, (10, 10, 10))
    assert adapted_rand_agreement(ref, ref) == 1.0
    perm = np.array([0, 5, 4, 3, 2, 1])[ref]            # relabelling does not matter
    assert adapted_rand_agreement(perm, ref) == 1.0
    split = ref.copy()
    split[:5][ref[:5] == 3] = 9                           # splitting a segment lowers the score
    assert 0 < adapted_rand_agreement(split, ref) < 1.0
    assert adapted_rand_agreement(np.ones_like(ref), ref) < 0.5   # everything merged
