"""CPU tests of the host-side logic behind the C ABI (no compute calls, no GPU)."""

import ctypes
import os
import re

import numpy as np
import pytest

from aind_exaspim_neuron_segmentation_b200 import _native, inference
from aind_exaspim_neuron_segmentation_b200.engine import percentiles_from_hist, plan_slab
from oracle import predict_ref as pr

from helpers import ROOT, make_volume


def test_library_loads_and_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "exaspim_b200.h")).read()
    declared = set(re.findall(r"\b(exa_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 18
    handle = ctypes.CDLL(_native.build())
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
    assert declared == set(_native.EXPORTED_SYMBOLS)
    assert b"sm_100a" in _native.lib().exa_version()


def test_create_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    code = _native.lib().exa_create(0, 0, ctypes.byref(h))
    assert code < 0 and not h.value
    assert b"no CPU fallback" in _native.lib().exa_last_error(None)
    with pytest.raises(RuntimeError):
        from aind_exaspim_neuron_segmentation_b200 import UNet3D

        UNet3D(3).eval()(torch.zeros(1, 1, 16, 16, 16))


def test_count_and_starts_match_oracle_and_c_abi(golden_meta):
    lib = _native.lib()
    for t in golden_meta["tiling"]:
        dims, patch, ov = t["dims"], t["patch"], t["overlap"]
        shape5 = (1, 1) + tuple(dims)
        assert inference.count_patches(shape5, patch, ov) == t["n"]
        starts = list(inference.generate_patch_starts(shape5, patch, ov))
        assert starts == [tuple(s) for s in pr.patch_starts(dims, patch, ov)]
        p3 = (ctypes.c_int32 * 3)(*patch)
        o3 = (ctypes.c_int32 * 3)(*ov)
        n = lib.exa_count_patches(*dims, p3, o3)
        assert n == t["n"]
        buf = (ctypes.c_int32 * (3 * n))()
        assert lib.exa_patch_starts(*dims, p3, o3, buf, n) == n
        assert np.array(buf[:]).reshape(-1, 3).tolist() == [list(s) for s in starts]
    with pytest.raises(AssertionError):
        inference.count_patches((96, 96, 96), (96, 96, 96), (32, 32, 32))


def test_percentiles_from_histogram_are_bit_exact():
    rng = np.random.default_rng(0)
    cases = [((96, 96, 96), 1000, (1, 99.9)), ((33, 17, 29), 700, (5, 99)), ((20, 20, 20), 300, (50, 50.5)),
             ((7, 5, 3), 1000, (0, 100)), ((64, 64, 64), 1000, (1, 99.9)), ((11, 13, 17), 50, (2.5, 97.5))]
    for shape, clip, pct in cases:
        for hi in (2000, 1200, 40):
            vol = rng.integers(0, hi, shape, dtype=np.uint16)
            clipped = np.minimum(vol, clip)
            hist = np.bincount(clipped.ravel(), minlength=clip + 1)
            mn, mx = percentiles_from_hist(hist, *pct)
            rmn, rmx = np.percentile(clipped, pct)
            assert mn == float(rmn) and mx == float(rmx), (shape, clip, pct, hi)
    # many random percentiles on one skewed volume
    vol = (rng.gamma(2.0, 60.0, (40, 40, 40))).astype(np.uint16)
    clipped = np.minimum(vol, 1000)
    hist = np.bincount(clipped.ravel(), minlength=1001)
    for q in rng.uniform(0, 100, 200):
        a, b = percentiles_from_hist(hist, float(q), float(100 - q))
        ra, rb = np.percentile(clipped, (q, 100 - q))
        assert a == float(ra) and b == float(rb)


def test_golden_norm_scalars(golden_meta):
    for n in golden_meta["norms"]:
        clipped = np.minimum(make_volume(n["shape"], n["seed"]), n["clip"])
        hist = np.bincount(clipped.ravel(), minlength=n["clip"] + 1)
        mn, mx = percentiles_from_hist(hist, *n["pct"])
        assert mn == n["mn"] and mx == n["mx"]


def _params(patch=(96, 96, 96), overlap=(32, 32, 32), trim=8):
    return _native.make_params(patch, overlap, trim, 1000, (1, 99.9))


def test_slab_plans_tile_the_volume():
    for dims, patch, ov, trim in [((1024, 64, 64), (96,) * 3, (32,) * 3, 8), ((512, 40, 40), (96,) * 3, (32,) * 3, 8),
                                  ((512, 40, 40), (128,) * 3, (32,) * 3, 8), ((200, 40, 40), (32,) * 3, (8,) * 3, 4),
                                  ((96, 96, 96), (96,) * 3, (32,) * 3, 8)]:
        p = _params(patch, ov, trim)
        nz = plan_slab(dims, p, 0, 0)["nz"]
        cover = pr.coverage_count_axis(dims[0], patch[0], ov[0], trim)
        for world in (1, 2, 3, 4, 8):
            rows = inference.split_rows(nz, world)
            assert rows[0][0] == 0 and rows[-1][1] == nz
            assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
            owned = np.zeros(dims[0], int)
            prev = None
            for r in rows:
                if r[1] == r[0]:
                    continue
                pl = plan_slab(dims, p, *r)
                owned[pl["out_z0"]:pl["out_z1"]] += 1
                stride = patch[0] - ov[0]
                assert pl["in_z0"] == stride * r[0]
                assert pl["in_z1"] == min(stride * (r[1] - 1) + patch[0], dims[0])
                if prev is not None:
                    assert (prev["halo_z0"], prev["halo_z1"]) == (pl["seed_z0"], pl["seed_z1"])
                    # the shared planes are exactly those of this slab covered twice at its start
                    assert all(cover[z] == 2 for z in range(pl["seed_z0"], pl["seed_z1"]))
                    assert pl["seed_z0"] == pl["out_z0"]
                # every owned plane is only covered by own rows or the previous slab's last row
                prev = pl
            assert (owned == 1).all()
            assert prev["halo_z1"] - prev["halo_z0"] == 0


def test_slab_plan_rejects_triple_overlap():
    p = _params((64, 64, 64), (48, 32, 32), 0)
    with pytest.raises(RuntimeError):
        plan_slab((256, 64, 64), p, 0, 2)


def test_input_handling():
    vol = make_volume((8, 9, 10), 1)
    out = inference._as_volume_u16(vol[None, None], 1000)
    assert out.dtype == np.uint16 and out.shape == (8, 9, 10) and np.array_equal(out, vol)
    big = vol.astype(np.int64) * 100
    out = inference._as_volume_u16(big, 1000)
    assert np.array_equal(out, np.minimum(big, 1000))
    with pytest.raises(TypeError):
        inference._as_volume_u16(vol.astype(np.float32), 1000)
    with pytest.raises(ValueError):
        inference._as_volume_u16(np.zeros((2, 1, 4, 4, 4), np.uint16), 1000)


def test_state_dict_layout_matches_reference(golden_meta):
    import torch

    from aind_exaspim_neuron_segmentation_b200 import UNet3D

    sd = UNet3D(output_channels=3).state_dict()
    ref = golden_meta["state_dict"]
    assert list(sd) == list(ref) and len(sd) == 128
    for k, v in sd.items():
        assert [list(v.shape), str(v.dtype)] == ref[k]
    assert sum(v.numel() for v in sd.values()) == 12951285
    # strict round trip through a file, as load_model does (inference.py:420-421)
    from oracle.unet_ref import rescaled_state_dict

    m = UNet3D(output_channels=3)
    m.load_state_dict(rescaled_state_dict(0), strict=True)
    with pytest.raises(RuntimeError):
        UNet3D(output_channels=3).load_state_dict(
            {k: v for k, v in sd.items() if k != "outc.conv.bias"}, strict=True)
    assert torch.equal(m.state_dict()["inc.double_conv.0.weight"],
                       rescaled_state_dict(0)["inc.double_conv.0.weight"])


def test_input_checks_raise_instead_of_altering_the_image():
    """Inputs the reference would treat differently from the uint16 kernels raise."""
    vol = make_volume((8, 8, 8), 1)
    assert inference._as_volume_u16(vol, 1000) is not None
    assert inference._as_volume_u16(vol.astype(np.int64), 1000).dtype == np.uint16
    assert inference._check_clip(vol, 1000.0) == 1000
    with pytest.raises(ValueError):
        inference._check_clip(vol, 1000.5)          # np.minimum would make the image float
    with pytest.raises(ValueError):
        inference._check_clip(vol, -1)
    big = vol.astype(np.int64) + 70000
    with pytest.raises(TypeError):
        inference._as_volume_u16(big, 100000)       # values > 65535 survive the clip
    assert inference._as_volume_u16(big, 60000).max() == 60000   # removed by the clip: exact
    with pytest.raises(TypeError):
        inference._as_volume_u16(-vol.astype(np.int32) - 1, 1000)
    with pytest.raises(TypeError):
        inference._as_volume_u16(vol.astype(np.float64), 1000)


def test_trainer_fails_loudly_without_gpu():
    """The training step has no CPU fallback either: exa_train_create fails without a device, and a
    train()-mode forward on CPU tensors raises instead of running torch modules."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _native.lib()
    h = ctypes.c_void_p()
    code = lib.exa_train_create(0, 0, ctypes.byref(h))
    assert code < 0 and not h.value
    assert b"no CPU fallback" in lib.exa_train_last_error(None)
    assert lib.exa_train_forward(None, None, 1, (ctypes.c_int32 * 3)(16, 16, 16), None, None) < 0
    from aind_exaspim_neuron_segmentation_b200 import UNet3D

    m = UNet3D(3).train()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(2, 1, 16, 16, 16))
    # the variants the Trainer never builds are refused in train() mode before any kernel runs
    for kw in (dict(trilinear=False), dict(width_multiplier=2)):
        wide = UNet3D(3, **kw).train()
        with pytest.raises((NotImplementedError, RuntimeError)):
            wide(torch.zeros(2, 1, 16, 16, 16))


def test_model_copies_and_training_mode():
    """Engines live outside the module: models deep-copy / pickle, and a model put back into
    training mode never silently runs the folded eval-mode engine."""
    import copy
    import io

    import torch

    from aind_exaspim_neuron_segmentation_b200 import UNet3D

    m = UNet3D(3).eval()
    m2 = copy.deepcopy(m)
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
    buf = io.BytesIO()
    torch.save(m, buf)
    assert not hasattr(m, "_engines")
    m.train()
    with pytest.raises(RuntimeError, match="eval"):
        m.engine()
    with pytest.raises(RuntimeError, match="eval"):
        inference._engine_for(m)


def test_sharded_rejects_triple_z_overlap(monkeypatch):
    """patch - 2*trim > 2*stride along z: a plane is covered by three z rows; predict() handles
    it (one slab), predict_sharded on more than one rank raises a clear error."""
    vol = make_volume((64, 32, 32), 2)
    monkeypatch.setattr(inference, "world_size_of", lambda group=None: 2)
    with pytest.raises(ValueError, match="three or more z rows"):
        inference.predict_sharded(vol, None, patch_shape=(32, 32, 32), overlap=(24, 8, 8), trim=0,
                                  backend=object())


def test_percentiles_from_histogram_and_value_table_are_bit_exact():
    """Float images: np.percentile from the histogram of ranks + the table of distinct values
    (exa_percentiles_from_hist_values), float32 (numpy takes the neighbour difference in float32)
    and float64."""
    from aind_exaspim_neuron_segmentation_b200.engine import percentiles_from_hist_values

    rng = np.random.default_rng(3)
    for dt in (np.float32, np.float64):
        for _ in range(12):
            n = int(rng.integers(5, 4000))
            pool = np.unique((rng.random(int(rng.integers(2, 300))) * 1000 / 3).astype(dt))
            a = rng.choice(pool, n).astype(dt)
            table = np.unique(a)
            hist = np.array([(a == t).sum() for t in table], np.uint64)
            for pct in ((1, 99.9), (5, 99), (0, 100), (50, 50.5), (33.3, 66.6)):
                ref = np.percentile(a, pct)
                got = percentiles_from_hist_values(hist, table.astype(np.float64), dt == np.float32, *pct)
                assert (float(ref[0]), float(ref[1])) == got, (dt, pct)


def test_npy_stack_and_sink_round_trip(tmp_path):
    """utils/img_util.py: a directory of .npy z-chunks as a sliceable source, one file per assigned
    plane range as a sink (what predict_streamed reads from / writes to), read() dispatch."""
    from aind_exaspim_neuron_segmentation_b200.utils import img_util

    rng = np.random.default_rng(0)
    vol = rng.integers(0, 2000, (37, 6, 5), dtype=np.uint16)
    src = tmp_path / "vol"
    src.mkdir()
    for i, (a, b) in enumerate(((0, 10), (10, 11), (11, 30), (30, 37))):
        np.save(src / f"chunk{i}.npy", vol[a:b])
    stack = img_util.read(str(src))
    assert isinstance(stack, img_util.NpyStack) and stack.shape == vol.shape and stack.dtype == vol.dtype
    for z0, z1 in ((0, 37), (3, 9), (9, 12), (10, 11), (29, 37), (12, 12)):
        assert np.array_equal(stack[z0:z1], vol[z0:z1])
    assert np.array_equal(stack[5], vol[5]) and np.array_equal(stack[-1], vol[-1])
    assert np.array_equal(stack[4:20, 1:3], vol[4:20, 1:3]) and np.array_equal(np.asarray(stack), vol)
    np.save(tmp_path / "single.npy", vol)
    assert np.array_equal(img_util.read(str(tmp_path / "single.npy")), vol)
    with pytest.raises(ValueError):
        img_util.read(str(tmp_path / "volume.xyz"))
    sink = img_util.NpySink(str(tmp_path / "out"), channels=3)
    res = rng.random((3, 37, 6, 5)).astype(np.float32)
    for z0, z1 in ((0, 16), (16, 17), (17, 37)):
        sink[:, z0:z1] = res[:, z0:z1]
    assert np.array_equal(img_util.NpySink.open(str(tmp_path / "out")), res)
