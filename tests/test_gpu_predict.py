"""GPU parity of the whole path (C ABI exa_predict, host buffers) vs the oracle / goldens."""

import numpy as np
import pytest
import torch

from helpers import (compare_with_golden, float_image, lightsheet_volume, load_golden, make_volume,
                     state_dict_for)

pytestmark = pytest.mark.gpu

BF16_TOL = 1e-2   # north_star: max-abs affinity error in BF16
FP32_TOL = 1e-4   # north_star: FP32 validation mode


def _model(kind, seed, precision="bf16", out_channels=3):
    from aind_exaspim_neuron_segmentation_b200 import UNet3D

    tri, width = True, 1
    if kind.startswith("rescaled/"):   # "rescaled/<trilinear>/<width_multiplier>" (make_golden.py)
        tri, width = (int(v) for v in kind.split("/")[1:])
    m = UNet3D(output_channels=out_channels, trilinear=bool(tri), width_multiplier=width,
               precision=precision)
    m.load_state_dict(state_dict_for(kind, seed, out_channels), strict=True)
    return m.cuda().eval()


def _kwargs(meta):
    kw = dict(meta["kwargs"])
    for key in ("patch_shape", "overlap", "normalization_percentiles"):
        if key in kw:
            kw[key] = tuple(kw[key])
    return kw


@pytest.mark.parametrize("name", ["c1_default_96", "c1_rescaled_96", "mixed_160x160x100", "small_p32",
                                  "small_p48_trim0ish", "multireflect_p32", "p128_single",
                                  "variant_convT", "variant_w2", "variant_convT_w2",
                                  "float32_integers", "float32_quarters", "float64_thirds"])
@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_predict_matches_reference_golden(golden_meta, name, precision):
    from aind_exaspim_neuron_segmentation_b200 import predict

    meta = golden_meta["cases"][name]
    if precision == "fp32" and name == "mixed_160x160x100":
        pytest.skip("fp32 validation mode is the slow SIMT path; covered by the other cases")
    vol = make_volume(meta["shape"], meta["vol_seed"])
    if "float_image" in meta:
        vol = float_image(vol, meta["float_image"])
    model = _model(*meta["weights"], precision=precision)
    before = vol.copy()
    out = predict(vol, model, verbose=False, **_kwargs(meta))
    assert out.dtype == np.float32 and out.shape == (3,) + tuple(meta["shape"])
    assert out.flags["C_CONTIGUOUS"] and np.array_equal(vol, before)
    worst = compare_with_golden(out, load_golden(name), BF16_TOL if precision == "bf16" else FP32_TOL)
    print(name, precision, "max abs err vs reference golden:", worst)


def test_predict_full_volume_vs_oracle_fp32():
    """Whole-array comparison (not only the golden sub-sample) on a ragged small case."""
    from aind_exaspim_neuron_segmentation_b200 import predict
    from oracle.predict_ref import predict_ref
    from oracle.unet_ref import make_forward_fn

    sd = state_dict_for("rescaled", 11)
    vol = lightsheet_volume((70, 45, 83), 12)
    kw = dict(patch_shape=(32, 32, 48), overlap=(8, 16, 16), trim=4)
    ref = predict_ref(vol, make_forward_fn(sd), **kw)
    for precision, tol in (("fp32", FP32_TOL), ("bf16", BF16_TOL)):
        out = predict(vol, _model("rescaled", 11, precision), verbose=False, **kw)
        err = np.abs(out - ref).max()
        assert err <= tol, (precision, err)
        assert np.array_equal(out == 0, ref == 0)   # uncovered shell is exactly zero


def test_foreground_mode_and_reference_module_interop():
    from aind_exaspim_neuron_segmentation_b200 import predict
    from oracle.predict_ref import predict_ref
    from oracle.unet_ref import make_forward_fn

    sd = state_dict_for("rescaled", 13, out_channels=1)
    vol = make_volume((40, 40, 40), 14)
    kw = dict(patch_shape=(32, 32, 32), overlap=(16, 16, 16), trim=4)
    out = predict(vol, _model("rescaled", 13, "bf16", 1), affinity_mode=False, verbose=False, **kw)
    ref = predict_ref(vol, make_forward_fn(sd), n_channels=1, **kw)[0]
    assert out.shape == (40, 40, 40)
    assert np.abs(out - ref).max() <= BF16_TOL
    with pytest.raises(ValueError):
        predict(vol, _model("rescaled", 13, "bf16", 1), affinity_mode=True, verbose=False, **kw)


def test_load_model_round_trip(tmp_path):
    from aind_exaspim_neuron_segmentation_b200 import load_model, predict

    sd = state_dict_for("rescaled", 15)
    path = tmp_path / "UNet3d-test.pth"
    torch.save(sd, path)
    model = load_model(str(path), affinity_mode=True, device="cuda")
    assert not model.training and next(model.parameters()).is_cuda
    vol = make_volume((32, 32, 32), 16)
    out = predict(vol, model, verbose=False, patch_shape=(32, 32, 32), overlap=(8, 8, 8), trim=4)
    assert out.shape == (3, 32, 32, 32) and out[:, 4:28, 4:28, 4:28].min() > 0
    assert out[:, :4].max() == 0 and out[:, 28:].max() == 0


def test_device_path_equals_host_path_and_slabs_equal_single():
    """exa_predict_device == exa_predict; z-row slabs with partial-sum hand-over == single slab."""
    from aind_exaspim_neuron_segmentation_b200 import _native
    from aind_exaspim_neuron_segmentation_b200.engine import percentiles_from_hist, plan_slab

    model = _model("rescaled", 17)
    eng = model.engine()
    shape = (150, 40, 56)
    vol = lightsheet_volume(shape, 18)
    params = _native.make_params((32, 32, 32), (8, 8, 8), 4, 1000, (1, 99.9), batch=7)
    host = eng.predict_host(vol, params)
    vdev = torch.from_numpy(vol).cuda()
    dev = eng.predict_device(vdev, params).cpu().numpy()
    assert np.array_equal(host, dev)

    hist = eng.histogram(vdev, 1000).cpu().numpy()
    assert np.array_equal(hist, np.bincount(np.minimum(vol, 1000).ravel(), minlength=1001))
    mn, mx = percentiles_from_hist(hist, 1, 99.9)
    nz = plan_slab(shape, params, 0, 0)["nz"]
    for cuts in ([0, 2, nz], [0, 1, 3, nz], [0, nz]):
        out = np.zeros_like(host)
        seed = None
        for r0, r1 in zip(cuts, cuts[1:]):
            pl = plan_slab(shape, params, r0, r1)
            slab = vdev[pl["in_z0"]:pl["in_z1"]].contiguous()
            eng.set_normalization(mn, mx, 1000)
            eng.slab_run(slab, shape, params, r0, r1)
            own = torch.empty((3, pl["out_z1"] - pl["out_z0"], shape[1], shape[2]), device="cuda")
            eng.slab_stitch(seed, own)
            out[:, pl["out_z0"]:pl["out_z1"]] = own.cpu().numpy()
            nh = pl["halo_z1"] - pl["halo_z0"]
            if nh > 0:
                seed = torch.empty((3, nh, shape[1], shape[2]), device="cuda")
                eng.slab_partial(seed)
            else:
                seed = None
        assert np.array_equal(out, host), cuts   # bit-identical: same summation order

    # the pipelined form (exa_slab_predict + exa_slab_finish): row groups inside each rank's rows,
    # D2H overlapped, the planes shared with the previous rank finished last
    # batch=2 with 4 patches per row: one-row groups whose finished y-bands are stitched and copied
    # while the row still runs
    params_band = _native.make_params((32, 32, 32), (8, 8, 8), 4, 1000, (1, 99.9), batch=2)
    assert np.array_equal(eng.predict_host(vol, params_band), host)
    for cuts, prm in [(c, q) for q in (params, params_band)
                      for c in ([0, 2, nz], [0, 1, 3, nz], [0, nz], [0, 3, 4, nz])]:
        out = np.zeros_like(host)
        pending = None   # (halo of the previous "rank")
        for r0, r1 in zip(cuts, cuts[1:]):
            pl = plan_slab(shape, params, r0, r1)
            slab = vdev[pl["in_z0"]:pl["in_z1"]].contiguous()
            n_own = pl["out_z1"] - pl["out_z0"]
            own = torch.full((3, n_own, shape[1], shape[2]), -1.0, device="cuda")
            own_host = torch.full((3, n_own, shape[1], shape[2]), -2.0).pin_memory()
            nh = pl["halo_z1"] - pl["halo_z0"]
            halo = torch.empty((3, nh, shape[1], shape[2]), device="cuda") if nh > 0 else None
            eng.set_normalization(mn, mx, 1000)
            eng.slab_predict(slab, shape, prm, r0, r1, own, own_host, halo)
            eng.slab_finish(pending, own, own_host)
            assert np.array_equal(own_host.numpy(), own.cpu().numpy())
            out[:, pl["out_z0"]:pl["out_z1"]] = own_host.numpy()
            pending = halo
        assert np.array_equal(out, host), cuts


def test_predict_streamed_equals_predict(tmp_path):
    """Out-of-core form: memory-mapped input and output, a few z patch-rows resident at a time,
    bit-identical to predict() (same summation order through the halo hand-over)."""
    from aind_exaspim_neuron_segmentation_b200 import predict, predict_streamed

    model = _model("rescaled", 23)
    shape = (150, 40, 56)
    vol = lightsheet_volume(shape, 24)
    kw = dict(patch_shape=(32, 32, 32), overlap=(8, 8, 8), trim=4)
    ref = predict(vol, model, verbose=False, **kw)
    src = np.memmap(tmp_path / "vol.u16", dtype=np.uint16, mode="w+", shape=shape)
    src[:] = vol
    for rows in (1, 2, 100):
        dst = np.memmap(tmp_path / f"aff{rows}.f32", dtype=np.float32, mode="w+", shape=(3,) + shape)
        dst[:] = -1
        assert predict_streamed(src, model, dst, rows_per_chunk=rows, **kw) is dst
        assert np.array_equal(np.asarray(dst), ref), rows
    # foreground mode writes (D, H, W); a volume below one patch stride gives zeros like predict()
    fg = _model("rescaled", 13, "bf16", 1)
    out1 = np.full(shape, -1, np.float32)
    predict_streamed(src, fg, out1, affinity_mode=False, rows_per_chunk=2, **kw)
    assert np.array_equal(out1, predict(vol, fg, affinity_mode=False, verbose=False, **kw))
    with pytest.raises(ValueError):
        predict_streamed(vol[None], model, np.empty((3,) + shape, np.float32), **kw)
    # file-backed source and sink (utils/img_util.py): a directory of .npy z-chunks read through
    # img_util.read(), one .npy file per finished plane range written by NpySink
    from aind_exaspim_neuron_segmentation_b200.utils import img_util

    chunks = tmp_path / "chunks"
    chunks.mkdir()
    for i, (a, b) in enumerate(((0, 40), (40, 41), (41, 120), (120, 150))):
        np.save(chunks / f"z{i:03d}.npy", vol[a:b])
    sink = img_util.NpySink(str(tmp_path / "aff_chunks"))
    predict_streamed(img_util.read(str(chunks)), model, sink, rows_per_chunk=2, **kw)
    assert np.array_equal(img_util.NpySink.open(str(tmp_path / "aff_chunks")), ref)


def test_config2_512_matches_oracle_on_sampled_blocks():
    """BASELINE config 2 at full size, on the bench's own volume and weights.

    The CPU oracle cannot finish 512 patches, so four 2x2x2 groups of windows (volume corner, far
    overhang corner with reflect padding, a face, the interior) are run through it with the WHOLE
    volume's percentiles; wherever only the group's windows cover a voxel the result must be the
    reference's (<= 1e-2 in bf16).  The fp32 validation mode is checked the same way on the first
    z row (<= 1e-4).  The sub-sampled checksum bench.py prints is tied to this checked output
    through tests/golden/bench_checksum.json."""
    import json
    import os

    import bench
    from aind_exaspim_neuron_segmentation_b200 import _native, predict
    from helpers import GOLDEN
    from oracle.predict_ref import predict_ref
    from oracle.unet_ref import make_forward_fn, rescaled_state_dict

    shape = (512, 512, 512)
    vol = bench.synth_planes(shape, 0, 512)
    sd = rescaled_state_dict(0)
    model = _model("rescaled", 0)
    out = predict(vol, model, verbose=False)
    assert out.shape == (3,) + shape and out.dtype == np.float32
    assert np.isfinite(out).all() and out.min() >= 0 and out.max() <= 1
    # uncovered shell: first 8 planes on every axis; overhang axes are covered to the end
    assert out[:, :8].max() == 0 and out[:, :, :8].max() == 0 and out[:, :, :, :8].max() == 0
    assert out[:, 8:, 8:, 8:].min() > 0
    assert np.array_equal(out, predict(vol, model, verbose=False, batch_size=64))   # deterministic

    mn, mx = (float(v) for v in np.percentile(np.minimum(vol, 1000), (1, 99.9)))
    fwd = make_forward_fn(sd)

    def group(k):
        """Axis range of the sub-volume holding windows k, k+1 and, in its coordinates, the range
        covered by no other window."""
        a = 64 * k
        b = min(a + 160, 512)
        lo = 0 if k == 0 else 24
        hi = b - a if k + 1 == 7 else 64 + 72
        return a, b, lo, hi

    worst = {}
    refs = {}
    for name, ks in dict(corner=(0, 0, 0), far_corner=(6, 6, 6), face=(3, 0, 6), interior=(2, 4, 1)).items():
        g = [group(k) for k in ks]
        sub = vol[g[0][0]:g[0][1], g[1][0]:g[1][1], g[2][0]:g[2][1]]
        ref = predict_ref(sub, fwd, norm_range=(mn, mx))
        loc = tuple(slice(a[2], a[3]) for a in g)
        glo = tuple(slice(a[0] + a[2], a[0] + a[3]) for a in g)
        got = out[(slice(None),) + glo]
        want = ref[(slice(None),) + loc]
        assert np.array_equal(got == 0, want == 0), name
        worst[name] = float(np.abs(got - want).max())
        refs[name] = (glo, want)
    print("512^3 bf16 vs oracle, max abs per block:", worst)
    assert max(worst.values()) <= BF16_TOL, worst

    # fp32 validation mode, first z row only (64 patches), same global normalisation
    eng = _model("rescaled", 0, "fp32").engine()
    params = _native.make_params((96, 96, 96), (32, 32, 32), 8, 1000, (1, 99.9), batch=32)
    eng.set_normalization(mn, mx, 1000)
    from aind_exaspim_neuron_segmentation_b200.engine import plan_slab

    pl = plan_slab(shape, params, 0, 1)
    eng.slab_run(torch.from_numpy(vol[pl["in_z0"]:pl["in_z1"]]).cuda(), shape, params, 0, 1)
    own = torch.empty((3, pl["out_z1"] - pl["out_z0"], 512, 512), device="cuda")
    eng.slab_stitch(None, own)
    glo, want = refs["corner"]
    z_hi = min(glo[0].stop, pl["out_z1"], 72)      # planes covered by row 0 only
    got32 = own[:, :z_hi, glo[1], glo[2]].cpu().numpy()
    err32 = float(np.abs(got32 - want[:, :z_hi]).max())
    print("512^3 fp32 mode vs oracle (row 0, corner block): max abs", err32)
    assert err32 <= FP32_TOL, err32

    checksum = float(out[:, ::37, ::41, ::43].astype(np.float64).sum())
    with open(os.path.join(GOLDEN, "bench_checksum.json")) as f:
        want_sum = json.load(f)["n1_512"]
    print("bench checksum", checksum, "recorded", want_sum["value"])
    assert abs(checksum - want_sum["value"]) <= want_sum["tol"]


def test_float_images_too_many_values_or_nan_raise():
    """predict() takes float images whose clipped values form a table of at most 65536 entries
    (goldens float32_integers / float32_quarters / float64_thirds above); anything else is refused,
    never approximated."""
    from aind_exaspim_neuron_segmentation_b200 import predict

    model = _model("rescaled", 7)
    rng = np.random.default_rng(5)
    noisy = (rng.random((48, 48, 48)) * 900).astype(np.float32)      # ~110 000 distinct values
    with pytest.raises(RuntimeError, match="65536"):
        predict(noisy, model, verbose=False, patch_shape=(32, 32, 32), overlap=(8, 8, 8), trim=4)
    bad = make_volume((48, 48, 48), 1).astype(np.float32)
    bad[3, 4, 5] = np.nan
    with pytest.raises(RuntimeError, match="NaN"):
        predict(bad, model, verbose=False, patch_shape=(32, 32, 32), overlap=(8, 8, 8), trim=4)
    # a clip below every noisy value leaves ONE distinct value: allowed (and all-constant input)
    flat = predict(noisy + 1000, model, verbose=False, brightness_clip=500, patch_shape=(32, 32, 32),
                   overlap=(8, 8, 8), trim=4)
    assert flat.shape == (3, 48, 48, 48) and np.isfinite(flat).all()
