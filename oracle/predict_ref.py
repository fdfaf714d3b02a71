"""CPU restatement of the reference's sliding-window prediction driver.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function cites the
reference lines it restates; ``REF`` = src/aind_exaspim_neuron_segmentation.

The model forward is injected (``forward_fn``) so that the same driver can be
run with the torch restatement of the U-Net (oracle/unet_ref.py), with the
bf16-emulating variant, or with a stub.
"""

import itertools

import numpy as np


# --- preprocessing ---------------------------------------------------------
def clip_and_normalize(vol, brightness_clip=1000, percentiles=(1, 99.9)):
    """REF/inference.py:79-80 + REF/utils/img_util.py:504-533.

    ``np.minimum`` keeps the integer dtype; ``np.percentile`` is global over
    the volume (linear interpolation); the affine map runs in float64 and is
    clipped to [0, 1].  Returns (float64 volume, mn, mx).
    """
    clipped = np.minimum(vol, brightness_clip)
    mn, mx = np.percentile(clipped, percentiles)
    out = (clipped - mn) / (mx - mn + 1e-8)
    return np.clip(out, 0, 1), float(mn), float(mx)


# --- tiling ----------------------------------------------------------------
def axis_starts(dim, patch, overlap):
    """Per-axis window starts: REF/inference.py:389-393."""
    stride = patch - overlap
    return list(range(0, dim - patch + stride, stride))


def patch_starts(shape3, patch_shape, overlap):
    """All (z, y, x) starts, z-major: REF/inference.py:368-397."""
    per_axis = [axis_starts(d, p, o) for d, p, o in zip(shape3, patch_shape, overlap)]
    return list(itertools.product(*per_axis))


def n_patches(shape3, patch_shape, overlap):
    """REF/inference.py:340-365."""
    n = 1
    for d, p, o in zip(shape3, patch_shape, overlap):
        n *= len(axis_starts(d, p, o))
    return n


def extract_patch(norm_vol, start, patch_shape):
    """Box clipped to the volume, reflect-padded at the far end only.

    REF/utils/img_util.py:424-428 (get_patch_slices) and :378-379 (add_padding,
    np.pad mode='reflect' with pad=(0, P-len)); cast to float32 happens on
    assignment into the batch array, REF/inference.py:188-191.
    """
    sl = tuple(slice(s, min(s + p, d)) for s, p, d in zip(start, patch_shape, norm_vol.shape))
    box = norm_vol[sl]
    pad = [(0, p - n) for p, n in zip(patch_shape, box.shape)]
    return np.pad(box, pad, mode="reflect").astype(np.float32)


# --- stitching -------------------------------------------------------------
def coverage_count_axis(dim, patch, overlap, trim):
    """How many windows keep voxel p along one axis (REF/inference.py:101-116)."""
    cnt = np.zeros(dim, dtype=np.int64)
    keep = patch - 2 * trim if trim > 0 else patch
    for s in axis_starts(dim, patch, overlap):
        lo = s + trim
        hi = min(lo + keep, dim)
        if hi > lo:
            cnt[lo:hi] += 1
    return cnt


def predict_ref(
    vol,
    forward_fn,
    n_channels=3,
    batch_size=16,
    brightness_clip=1000,
    normalization_percentiles=(1, 99.9),
    patch_shape=(96, 96, 96),
    overlap=(32, 32, 32),
    trim=8,
    apply_sigmoid=True,
    norm_range=None,
    only_starts=None,
):
    """Whole driver: REF/inference.py:29-126 for a 3-D ``vol``.

    ``forward_fn`` maps float32 (B, 1, P, P, P) -> float32 logits
    (B, C, P, P, P).  Returns float32 (C, D, H, W).

    Two test-only extensions for checking sub-blocks of volumes the CPU cannot finish:
    ``norm_range=(mn, mx)`` replaces the percentiles of ``vol`` (pass those of the WHOLE volume
    when ``vol`` is a corner of it), and ``only_starts`` restricts the patch loop to the given
    window starts (a voxel's result is then the reference's wherever every window covering it is
    in the list; elsewhere it is a partial sum divided by a partial count).
    """
    vol = np.asarray(vol)
    while vol.ndim > 3:
        assert vol.shape[0] == 1
        vol = vol[0]
    if norm_range is None:
        norm, _, _ = clip_and_normalize(vol, brightness_clip, normalization_percentiles)
    else:
        mn, mx = norm_range   # REF/utils/img_util.py:527-531 with the given scalars
        norm = np.clip((np.minimum(vol, brightness_clip) - mn) / (mx - mn + 1e-8), 0, 1)
    starts = patch_starts(norm.shape, patch_shape, overlap)
    if only_starts is not None:
        wanted = {tuple(int(v) for v in s) for s in only_starts}
        assert wanted <= set(starts), "only_starts must be window starts of this volume"
        starts = [s for s in starts if s in wanted]
    acc = np.zeros((n_channels,) + norm.shape, dtype=np.float32)
    wgt = np.zeros(norm.shape, dtype=np.float16)
    for i in range(0, len(starts), batch_size):
        chunk = starts[i:i + batch_size]
        x = np.stack([extract_patch(norm, s, patch_shape) for s in chunk])[:, None]
        y = forward_fn(x)
        if apply_sigmoid:
            y = 1.0 / (1.0 + np.exp(-y.astype(np.float32)))
            y = y.astype(np.float32)
        if trim > 0:
            y = y[..., trim:-trim, trim:-trim, trim:-trim]
        for patch, st in zip(y, chunk):
            lo = [s + trim for s in st]
            hi = [min(a + n, d) for a, n, d in zip(lo, patch.shape[1:], norm.shape)]
            dst = tuple(slice(a, b) for a, b in zip(lo, hi))
            src = tuple(slice(0, b - a) for a, b in zip(lo, hi))
            acc[(slice(None),) + dst] += patch[(slice(None),) + src]
            wgt[dst] += 1
    np.divide(acc, wgt, out=acc, where=wgt != 0)
    return acc
