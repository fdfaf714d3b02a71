"""Known-answer tests of the CPU restatement of affinities_to_segmentation (oracle/watershed_ref.py).

Parity unpinned: waterz is not available (see the module docstring), so these tests pin the
restatement's own contract on hand-constructed affinity volumes.
"""

import numpy as np

from oracle.watershed_ref import (adapted_rand_agreement, affinities_to_segmentation_ref,
                                  remove_small_segments, watershed_fragments)


def _two_blocks(shape=(12, 16, 16), wall=8, inside=0.95, across=0.02, seed=0):
    """Two bright blocks separated by a wall of weak z-edges between planes wall-1 and wall."""
    rng = np.random.default_rng(seed)
    aff = (inside + 0.04 * rng.random((3,) + shape)).astype(np.float32)
    aff[0, wall - 1] = across          # edge from plane wall-1 to plane wall (edge-to-next convention)
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
    aff[:, :, :, 12:] = 0.01            # nothing links these voxels: background
    aff[2, :, :, 11] = 0.01             # and no x-edge reaches into them
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
