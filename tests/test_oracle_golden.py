"""Pin the CPU oracle against outputs of the unmodified reference (tests/golden/*.npz)."""

import numpy as np
import pytest

from oracle import predict_ref as pr
from oracle.unet_ref import make_forward_fn

from helpers import compare_with_golden, float_image, load_golden, make_volume, state_dict_for


def _run_oracle(meta):
    vol = make_volume(meta["shape"], meta["vol_seed"])
    if "float_image" in meta:
        vol = float_image(vol, meta["float_image"])
    sd = state_dict_for(*meta["weights"])
    kw = dict(meta["kwargs"])
    for key in ("patch_shape", "overlap", "normalization_percentiles"):
        if key in kw:
            kw[key] = tuple(kw[key])
    return pr.predict_ref(vol, make_forward_fn(sd), **kw)


@pytest.mark.parametrize("name", ["small_p32", "small_p48_trim0ish", "c1_rescaled_96", "multireflect_p32",
                                  "p128_single", "variant_convT", "variant_w2", "variant_convT_w2",
                                  "float32_integers", "float32_quarters", "float64_thirds"])
def test_oracle_matches_reference_golden(golden_meta, name):
    out = _run_oracle(golden_meta["cases"][name])
    # same fp32 arithmetic (torch CPU conv) -> only thread-count dependent reduction order differs
    worst = compare_with_golden(out, load_golden(name), atol=2e-6)
    assert worst <= 2e-6


def test_oracle_default_init_case(golden_meta):
    out = _run_oracle(golden_meta["cases"]["c1_default_96"])
    compare_with_golden(out, load_golden("c1_default_96"), atol=2e-6)
    assert out[:, :8].max() == 0.0 and out[:, 88:].max() == 0.0


def test_tiling_matches_reference(golden_meta):
    for t in golden_meta["tiling"]:
        starts = pr.patch_starts(t["dims"], t["patch"], t["overlap"])
        assert len(starts) == t["n"] == pr.n_patches(t["dims"], t["patch"], t["overlap"])
        assert [list(s) for s in starts[:3]] == t["first"]
        assert [list(s) for s in starts[-3:]] == t["last"]
        checksum = sum((i + 1) * (z * 1000003 + y * 1009 + x) for i, (z, y, x) in enumerate(starts))
        assert checksum == t["checksum"]


def test_normalisation_matches_reference(golden_meta):
    for n in golden_meta["norms"]:
        vol = make_volume(n["shape"], n["seed"])
        normed, mn, mx = pr.clip_and_normalize(vol, n["clip"], tuple(n["pct"]))
        assert mn == n["mn"] and mx == n["mx"]
        assert float(normed.mean()) == pytest.approx(n["mean"], rel=0, abs=1e-15)
        sample = normed.ravel()[::max(1, normed.size // 16)][:16]
        assert [float(v) for v in sample] == n["sample"]


def test_reflect_padding_excludes_edge():
    vol = np.arange(64, dtype=np.float64).reshape(1, 1, 64) * np.ones((2, 2, 1))
    patch = pr.extract_patch(vol, (0, 0, 0), (2, 2, 96))
    assert patch.shape == (2, 2, 96)
    assert patch[0, 0, 64:].tolist() == [float(v) for v in range(62, 30, -1)]


def test_sub_block_extensions_reproduce_the_full_run():
    """norm_range / only_starts (used to check 512^3 sub-blocks on the GPU box): a corner block
    run with the whole volume's percentiles equals the full run wherever all covering windows
    are among the selected ones."""
    from oracle.unet_ref import rescaled_state_dict

    sd = rescaled_state_dict(2)
    vol = make_volume((72, 56, 40), 4)
    kw = dict(patch_shape=(32, 32, 32), overlap=(8, 8, 8), trim=4)
    full = pr.predict_ref(vol, make_forward_fn(sd), **kw)
    _, mn, mx = pr.clip_and_normalize(vol, 1000, (1, 99.9))
    # corner that holds the windows starting at z in {0, 24}, y in {0, 24}, x = 0 completely
    corner = vol[:56, :, :32]
    sel = [(z, y, 0) for z in (0, 24) for y in (0, 24)]
    part = pr.predict_ref(corner, make_forward_fn(sd), norm_range=(mn, mx), only_starts=sel, **kw)
    # voxels covered by the selected windows only: z < 48 + 4 (next z window starts at 48), x < 24 + 4
    assert np.array_equal(part[:, :52, :, :28], full[:, :52, :, :28])


def test_variant_state_dict_layout_matches_module():
    """trilinear=False / width_multiplier variants: the product module, the oracle's weight recipe
    and (checked when the goldens were made: tests/golden/make_golden.py variants) the reference
    module agree on keys and shapes."""
    import pytest

    from aind_exaspim_neuron_segmentation_b200 import UNet3D
    from oracle.unet_ref import rescaled_state_dict

    for trilinear, width in ((True, 1), (False, 1), (True, 2), (False, 2), (True, 4)):
        sd = rescaled_state_dict(1, 3, trilinear, width)
        model = UNet3D(output_channels=3, trilinear=trilinear, width_multiplier=width)
        own = model.state_dict()
        assert set(own) == set(sd) and len(sd) == (128 if trilinear else 136)
        assert all(tuple(own[k].shape) == tuple(sd[k].shape) for k in sd)
        model.load_state_dict(sd, strict=True)
    with pytest.raises(NotImplementedError):
        UNet3D(output_channels=3, width_multiplier=0.5)
