"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference's module-level imports of I/O / post-processing packages that are not
installed here (waterz, kimimaro, fastremap, zarr, tifffile, gcsfs, s3fs, matplotlib,
google-cloud-storage) are satisfied with empty stub modules; none of them is touched by
``predict``.  Outputs are reduced to small fixtures: a strided sub-sample of the affinity
volume, per-channel float64 moments, the non-zero bounding box and a few full rows.
"""

import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.unet_ref import rescaled_state_dict  # noqa: E402


def import_reference():
    for name in ["kimimaro", "waterz", "fastremap", "gcsfs", "s3fs", "tifffile", "zarr",
                 "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "google", "google.cloud",
                 "google.cloud.storage"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib.colors"].ListedColormap = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].colors = sys.modules["matplotlib.colors"]
    sys.modules["google"].cloud = sys.modules["google.cloud"]
    sys.modules["google.cloud"].storage = sys.modules["google.cloud.storage"]
    sys.modules["google.cloud.storage"].Client = object
    for attr in ("mask_except", "renumber", "unique"):
        setattr(sys.modules["fastremap"], attr, None)
    sys.path.insert(0, "/root/reference/src")
    from aind_exaspim_neuron_segmentation import inference
    from aind_exaspim_neuron_segmentation.machine_learning.unet3d import UNet3D
    return inference, UNet3D


CASES = {
    # name: (volume shape, vol seed, weights, predict kwargs)
    "c1_default_96": ((96, 96, 96), 0, ("default", 0), {}),
    "c1_rescaled_96": ((96, 96, 96), 0, ("rescaled", 0), {}),
    "mixed_160x160x100": ((100, 160, 160), 3, ("rescaled", 1), {}),
    "small_p32": ((72, 56, 40), 4, ("rescaled", 2),
                  dict(patch_shape=(32, 32, 32), overlap=(8, 8, 8), trim=4)),
    "small_p48_trim0ish": ((64, 80, 48), 5, ("rescaled", 3),
                           dict(patch_shape=(48, 32, 48), overlap=(16, 8, 24), trim=2,
                                brightness_clip=700, normalization_percentiles=(5, 99))),
    # round 2: axes shorter than half a patch -> np.pad(mode="reflect") wraps more than once
    # (z: 10 voxels padded to 32; second y window: 16 voxels padded to 32)
    "multireflect_p32": ((10, 40, 21), 6, ("rescaled", 4),
                         dict(patch_shape=(32, 32, 32), overlap=(8, 8, 8), trim=4)),
    # BASELINE config 4: one full 128^3 patch through the reference
    "p128_single": ((128, 128, 128), 7, ("rescaled", 5),
                    dict(patch_shape=(128, 128, 128), overlap=(32, 32, 32), trim=8)),
    # round 2: the constructor's other arguments (unet3d.py:37): weights kind
    # "rescaled/<trilinear>/<width_multiplier>" builds UNet3D(3, trilinear, width_multiplier)
    "variant_convT": ((48, 56, 40), 8, ("rescaled/0/1", 6),
                      dict(patch_shape=(32, 32, 32), overlap=(8, 8, 8), trim=4)),
    "variant_w2": ((48, 56, 40), 8, ("rescaled/1/2", 6),
                   dict(patch_shape=(32, 32, 32), overlap=(8, 8, 8), trim=4)),
    "variant_convT_w2": ((48, 56, 40), 8, ("rescaled/0/2", 6),
                         dict(patch_shape=(32, 32, 32), overlap=(8, 8, 8), trim=4)),
    # round 2: floating-point images (inference.py:79-80 takes any dtype); optional 5th entry =
    # float_image() kind applied to the uint16 volume
    "float32_integers": ((48, 56, 40), 9, ("rescaled", 7),
                         dict(patch_shape=(32, 32, 32), overlap=(8, 8, 8), trim=4), "f32"),
    "float32_quarters": ((40, 40, 72), 10, ("rescaled", 7),
                         dict(patch_shape=(32, 32, 32), overlap=(8, 8, 8), trim=4, brightness_clip=300.1,
                              normalization_percentiles=(2, 99.5)), "f32_quarter"),
    "float64_thirds": ((40, 64, 40), 11, ("rescaled", 7),
                       dict(patch_shape=(32, 32, 32), overlap=(8, 8, 8), trim=4, brightness_clip=500),
                       "f64_third"),
}


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
    """Synthetic uint16 volume: uniform noise with a bright blob so that the clip bites."""
    rng = np.random.default_rng(seed)
    vol = rng.integers(0, 2000, shape, dtype=np.uint16)
    return vol


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


def main():
    """``--only name[,name]``: (re)generate just those predict cases and merge them into the
    existing golden_meta.json (np.savez stamps the archive with the wall clock, so untouched
    fixtures are left alone to keep the history quiet)."""
    only = None
    if "--only" in sys.argv:
        only = set(sys.argv[sys.argv.index("--only") + 1].split(","))
    inference, UNet3D = import_reference()
    torch.set_num_threads(os.cpu_count())
    meta = {}
    meta_path = os.path.join(HERE, "golden_meta.json")
    if only is not None:
        with open(meta_path) as f:
            old = json.load(f)
        meta = old["cases"]
    for name, case in CASES.items():
        shape, vseed, (wkind, wseed), kwargs = case[:4]
        if only is not None and name not in only:
            continue
        vol = make_volume(shape, vseed)
        if len(case) > 4:
            vol = float_image(vol, case[4])
        if wkind == "default":
            torch.manual_seed(wseed)
            model = UNet3D(output_channels=3).eval()
        elif wkind.startswith("rescaled/"):
            tri, width = (int(v) for v in wkind.split("/")[1:])
            model = UNet3D(output_channels=3, trilinear=bool(tri), width_multiplier=width)
            model.load_state_dict(rescaled_state_dict(wseed, 3, bool(tri), width), strict=True)
            model.eval()
        else:
            model = UNet3D(output_channels=3)
            model.load_state_dict(rescaled_state_dict(wseed), strict=True)
            model.eval()
        out = inference.predict(vol, model, verbose=False, **kwargs)
        assert out.dtype == np.float32 and out.shape == (3,) + shape
        red = reduce_output(out)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **red)
        meta[name] = dict(shape=shape, vol_seed=vseed, weights=[wkind, wseed], kwargs=kwargs,
                          min=float(out.min()), max=float(out.max()))
        if len(case) > 4:
            meta[name]["float_image"] = case[4]
        print(name, out.shape, float(out.min()), float(out.max()), red["bbox"].tolist())
    if only is not None:
        old["cases"] = meta
        with open(meta_path, "w") as f:
            json.dump(old, f, indent=1)
        print("updated golden_meta.json:", sorted(only))
        return

    # tiling helpers (inference.py:340-397)
    tiling = []
    for dims in [(96, 96, 96), (100, 160, 160), (512, 512, 512), (64, 200, 97), (1024, 96, 130),
                 (31, 64, 65), (128, 128, 128)]:
        for patch, ov in [((96, 96, 96), (32, 32, 32)), ((128, 128, 128), (32, 32, 32)),
                          ((32, 48, 64), (8, 16, 0)), ((64, 64, 64), (48, 32, 16))]:
            shape5 = (1, 1) + dims
            n = inference.count_patches(shape5, patch, ov)
            starts = list(inference.generate_patch_starts(shape5, patch, ov))
            assert n == len(starts)
            tiling.append(dict(dims=dims, patch=patch, overlap=ov, n=n,
                               first=starts[:3], last=starts[-3:],
                               checksum=int(sum((i + 1) * (z * 1000003 + y * 1009 + x)
                                                for i, (z, y, x) in enumerate(starts)))))
    # normalisation scalars (img_util.py:526) on clipped volumes
    norms = []
    from aind_exaspim_neuron_segmentation.utils import img_util
    for seed, shape, clip, pct in [(0, (96, 96, 96), 1000, (1, 99.9)), (7, (40, 50, 60), 1000, (1, 99.9)),
                                   (8, (33, 17, 29), 700, (5, 99)), (9, (64, 64, 64), 1000, (0, 100)),
                                   (10, (20, 20, 20), 300, (50, 50.5))]:
        vol = make_volume(shape, seed)
        clipped = np.minimum(vol, clip)
        mn, mx = np.percentile(clipped, pct)
        normed = img_util.normalize(clipped, percentiles=pct)
        norms.append(dict(seed=seed, shape=shape, clip=clip, pct=pct, mn=float(mn), mx=float(mx),
                          mean=float(normed.mean()),
                          sample=[float(v) for v in normed.ravel()[::max(1, normed.size // 16)][:16]]))
    # default init of the product module must reproduce the reference's (same RNG consumption)
    from aind_exaspim_neuron_segmentation_b200.machine_learning.unet3d import UNet3D as Mine
    torch.manual_seed(0)
    ref_sd = UNet3D(output_channels=3).state_dict()
    torch.manual_seed(0)
    my_sd = Mine(output_channels=3).state_dict()
    assert list(ref_sd) == list(my_sd)
    assert all(torch.equal(ref_sd[k], my_sd[k]) for k in ref_sd)
    keys = {k: [list(v.shape), str(v.dtype)] for k, v in ref_sd.items()}
    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(dict(cases=meta, tiling=tiling, norms=norms, state_dict=keys,
                       torch=torch.__version__, numpy=np.__version__), f, indent=1)
    print("wrote golden_meta.json")


if __name__ == "__main__":
    main()
