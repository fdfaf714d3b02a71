"""GPU affinities_to_segmentation (csrc/watershed.cu) against the CPU restatement
oracle/watershed_ref.py.  Integer work: the bar is exact equality of the label volumes."""

import numpy as np
import pytest
import torch
from scipy.ndimage import gaussian_filter

pytestmark = pytest.mark.gpu


def smooth_affinities(shape, seed, sigma=2.0, gain=6.0, quant=None):
    rng = np.random.default_rng(seed)
    f = np.stack([gaussian_filter(rng.normal(size=shape), sigma) for _ in range(3)])
    f = f / f.std()
    aff = (1.0 / (1.0 + np.exp(-gain * f))).astype(np.float32)
    if quant:
        aff = (np.round(aff * quant) / quant).astype(np.float32)
    return aff


def _oracle(aff, thresholds, min_size):
    from oracle.watershed_ref import affinities_to_segmentation_ref

    return affinities_to_segmentation_ref(aff, thresholds, min_size)


@pytest.mark.parametrize("shape,seed,quant", [((24, 28, 36), 1, None), ((40, 48, 56), 2, None),
                                              ((17, 33, 9), 3, None), ((32, 32, 32), 4, 10),
                                              ((32, 40, 24), 5, 4), ((20, 20, 60), 6, 3),
                                              ((48, 48, 48), 7, 2), ((1, 64, 64), 8, 5),
                                              ((64, 1, 37), 9, 6), ((3, 3, 3), 10, 2)])
def test_fragments_equal_oracle(shape, seed, quant):
    """threshold 0 stops the merging at once and min size 0 keeps everything: the output is the
    watershed fragments, numbered in order of first appearance like the oracle's.  Coarsely
    quantised affinities give large plateaus (several breadth-first levels of the division),
    exact 0.0 / 1.0 values (>= high) and undivided plateaus without an exit."""
    from aind_exaspim_neuron_segmentation_b200 import affinities_to_segmentation
    from oracle.watershed_ref import watershed_fragments

    aff = smooth_affinities(shape, seed, quant=quant)
    got = affinities_to_segmentation(aff, [0.0], 0)
    ref, n = watershed_fragments(aff)
    assert got.dtype == np.uint64 and got.shape == shape
    assert int(got.max()) == n
    assert np.array_equal(got.astype(np.int64), ref)


@pytest.mark.parametrize("shape,seed,thr,min_size,quant", [
    ((24, 28, 36), 5, [0.6, 0.8, 0.9], 100, None),
    ((40, 48, 56), 6, [0.6, 0.8, 0.9], 100, None),
    ((40, 48, 56), 7, [0.3], 20, None),
    ((32, 40, 24), 8, [0.5, 0.7], 0, None),
    ((32, 32, 32), 9, [0.6, 0.8, 0.9], 10, 20),   # plateaus and exact score ties
])
def test_segmentation_equals_oracle(shape, seed, thr, min_size, quant):
    from aind_exaspim_neuron_segmentation_b200 import affinities_to_segmentation

    aff = smooth_affinities(shape, seed, quant=quant)
    got = affinities_to_segmentation(aff, thr, min_size)
    ref = _oracle(aff, thr, min_size)
    assert np.array_equal(got.astype(np.int64), ref)
    # CUDA tensor in -> CUDA tensor out, same labels
    dev = affinities_to_segmentation(torch.from_numpy(aff).cuda(), thr, min_size)
    assert dev.is_cuda and dev.dtype == torch.int64
    assert np.array_equal(dev.cpu().numpy(), ref)


def test_plateau_division_known_answer():
    """1-D plateau with an exit at each end (tests/test_watershed_oracle.py pins the oracle on it)."""
    from aind_exaspim_neuron_segmentation_b200 import affinities_to_segmentation
    from oracle.watershed_ref import watershed_fragments

    for w in (9, 10, 33, 200):
        aff = np.zeros((3, 1, 1, w), np.float32)
        aff[2, 0, 0, 1:] = 0.5
        aff[2, 0, 0, 1] = 0.9
        aff[2, 0, 0, w - 1] = 0.8
        got = affinities_to_segmentation(aff, [0.0], 0)
        ref, n = watershed_fragments(aff)
        assert n == 2 and np.array_equal(got.astype(np.int64), ref)


@pytest.mark.parametrize("rounds,tail", [(None, None), (0, 1), (1 << 30, 0), (3, 1)])
def test_agglomeration_rounds_equal_sequential_queue(monkeypatch, rounds, tail):
    """exa_region_agglomerate on the GPU -- default hand-over, host queue only, parallel rounds to
    the end, three rounds then the host queue -- against the oracle's sequential queue, on grid
    graphs and on arbitrary graphs with many exact score ties."""
    from oracle.watershed_ref import agglomerate
    from test_region_agglomerate import GRID_CASES, native_roots, random_graph, smooth_graph

    if rounds is not None:
        monkeypatch.setenv("EXA_WS_GPU_ROUNDS", str(rounds))
        monkeypatch.setenv("EXA_WS_HOST_TAIL", str(tail))
    for shape, seed, quant, threshold in GRID_CASES:
        frag, n, graph = smooth_graph(shape, seed, quant)
        ref = agglomerate(n, graph, threshold)
        got = native_roots(graph, n, threshold, device=0)
        assert np.array_equal(got.astype(np.int64), ref), (shape, seed, quant, threshold)
    rng = np.random.default_rng(17)
    for trial in range(60):
        n = int(rng.integers(2, 80))
        m = int(rng.integers(1, min(n * (n - 1) // 2, 4 * n) + 1))
        graph = random_graph(rng, n, m, int(rng.integers(2, 9)))
        threshold = float(rng.choice([0.0, 0.25, 0.5, 0.75, 0.9, 1.0, 1.5]))
        ref = agglomerate(n, graph, threshold)
        got = native_roots(graph, n, threshold, device=0)
        assert np.array_equal(got.astype(np.int64), ref), (trial, n, m, threshold)


def test_larger_volume_equals_oracle_all_paths(monkeypatch):
    """128^3 smooth field (2.7e5 fragments, 1.2e6 region edges): labels identical to the oracle
    with the default hand-over and with parallel rounds to the end."""
    from aind_exaspim_neuron_segmentation_b200 import affinities_to_segmentation

    aff = smooth_affinities((128, 128, 128), 21)
    ref = _oracle(aff, [0.6, 0.8, 0.9], 100)
    got = affinities_to_segmentation(aff, [0.6, 0.8, 0.9], 100)
    assert np.array_equal(got.astype(np.int64), ref)
    monkeypatch.setenv("EXA_WS_HOST_TAIL", "0")
    got = affinities_to_segmentation(aff, [0.6, 0.8, 0.9], 100)
    assert np.array_equal(got.astype(np.int64), ref)


def test_degenerate_inputs():
    from aind_exaspim_neuron_segmentation_b200 import affinities_to_segmentation

    shape = (12, 16, 20)
    zeros = np.zeros((3,) + shape, np.float32)
    assert affinities_to_segmentation(zeros).max() == 0           # no edge survives `low`
    ones = np.ones((3,) + shape, np.float32)
    seg = affinities_to_segmentation(ones)                         # every edge >= high: one segment
    assert seg.min() == 1 and seg.max() == 1
    small = affinities_to_segmentation(ones, min_segment_size=int(np.prod(shape)))
    assert small.max() == 0                                        # kept only if size > min size
    with pytest.raises(ValueError):
        affinities_to_segmentation(np.zeros((2,) + shape, np.float32))
    with pytest.raises(ValueError):
        affinities_to_segmentation(ones, [])


def test_predict_then_segment_matches_oracle_pipeline():
    """predict() -> affinities_to_segmentation(), the reference's README sequence, on the product;
    the same affinities through the oracle give identical labels."""
    from aind_exaspim_neuron_segmentation_b200 import UNet3D, affinities_to_segmentation, predict
    from helpers import lightsheet_volume, state_dict_for

    model = UNet3D(output_channels=3)
    model.load_state_dict(state_dict_for("rescaled", 51), strict=True)
    model = model.cuda().eval()
    vol = lightsheet_volume((64, 64, 64), 52)
    aff = predict(vol, model, verbose=False, patch_shape=(32, 32, 32), overlap=(8, 8, 8), trim=4)
    got = affinities_to_segmentation(aff, [0.6, 0.8, 0.9], 50)
    ref = _oracle(aff, [0.6, 0.8, 0.9], 50)
    assert np.array_equal(got.astype(np.int64), ref)
