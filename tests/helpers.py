"""Shared test helpers: seeded inputs/weights identical to tests/golden/make_golden.py."""

import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def float_image(vol, kind):
    """Floating-point images for the float-input cases: integer data stored as float32, quarter
    steps in float32, thirds in float64 (few thousand distinct values, not all exact in binary)."""
    if kind == "f32":
        return vol.astype(np.float32)
    if kind == "f32_quarter":
        return (vol.astype(np.float32) * np.float32(0.25)).astype(np.float32)
    if kind == "f64_third":
        return vol.astype(np.float64) / 3.0
    raise ValueError(kind)


def make_volume(shape, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 2000, tuple(shape), dtype=np.uint16)


def lightsheet_volume(shape, seed, n_paths=None):
    """Lightsheet-like uint16 volume: Poisson background + blurred bright random-walk neurites."""
    rng = np.random.default_rng(seed)
    shape = tuple(shape)
    vol = rng.poisson(40, shape).astype(np.float32) + rng.normal(0, 4, shape).astype(np.float32)
    n_paths = n_paths or max(4, int(np.prod(shape) / 60000))
    sig = np.zeros(shape, np.float32)
    for _ in range(n_paths):
        pos = np.array([rng.uniform(0, s) for s in shape])
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        amp = float(np.exp(rng.uniform(np.log(150), np.log(4000))))
        for _ in range(int(rng.integers(40, 200))):
            d = 0.9 * d + 0.1 * rng.normal(size=3)
            d /= np.linalg.norm(d)
            pos = pos + d
            idx = np.round(pos).astype(int)
            if np.any(idx < 1) or np.any(idx >= np.array(shape) - 1):
                break
            sig[idx[0] - 1:idx[0] + 2, idx[1] - 1:idx[1] + 2, idx[2] - 1:idx[2] + 2] = np.maximum(
                sig[idx[0] - 1:idx[0] + 2, idx[1] - 1:idx[1] + 2, idx[2] - 1:idx[2] + 2], amp)
    vol += sig
    return np.clip(vol, 0, 65535).astype(np.uint16)


def state_dict_for(kind, seed, out_channels=3):
    from oracle.unet_ref import rescaled_state_dict

    if kind == "rescaled":
        return rescaled_state_dict(seed, out_channels)
    if kind.startswith("rescaled/"):   # "rescaled/<trilinear>/<width_multiplier>"
        tri, width = (int(v) for v in kind.split("/")[1:])
        return rescaled_state_dict(seed, out_channels, bool(tri), width)
    from aind_exaspim_neuron_segmentation_b200.machine_learning.unet3d import UNet3D

    torch.manual_seed(seed)
    return UNet3D(output_channels=out_channels).state_dict()


def reduce_output(out):
    nz = np.nonzero(out[0])
    bbox = [[int(a.min()), int(a.max()) + 1] for a in nz] if nz[0].size else [[0, 0]] * 3
    return dict(
        sub=out[:, ::5, ::7, ::3].copy(),
        sum=out.astype(np.float64).sum(axis=(1, 2, 3)),
        sumsq=(out.astype(np.float64) ** 2).sum(axis=(1, 2, 3)),
        bbox=np.array(bbox),
        row=out[:, out.shape[1] // 2, out.shape[2] // 2, :].copy(),
        col=out[:, :, out.shape[2] // 3, out.shape[3] // 3].copy(),
    )


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))


def compare_with_golden(out, golden, atol):
    """max-abs comparison of the reduced views; returns the worst error."""
    red = reduce_output(out)
    assert red["bbox"].tolist() == golden["bbox"].tolist(), "non-zero bounding box differs"
    worst = 0.0
    for key in ("sub", "row", "col"):
        err = float(np.abs(red[key] - golden[key]).max())
        worst = max(worst, err)
        assert err <= atol, f"{key}: max abs err {err} > {atol}"
    n = out[0].size
    mean_err = np.abs(red["sum"] - golden["sum"]).max() / n
    assert mean_err <= atol, f"mean drift {mean_err}"
    return worst
