"""GPU tests of the remaining BASELINE.json configs (4: 128^3 patches, bf16 vs fp32 validation mode;
5: predict followed by the affinities_to_segmentation watershed, adapted-Rand agreement)."""

import numpy as np
import pytest

from helpers import lightsheet_volume, state_dict_for

pytestmark = pytest.mark.gpu

BF16_TOL = 1e-2   # north_star: max-abs affinity error in BF16
FP32_TOL = 1e-4   # north_star: FP32 validation mode


def _model(seed, precision):
    from aind_exaspim_neuron_segmentation_b200 import UNet3D

    m = UNet3D(output_channels=3, precision=precision)
    m.load_state_dict(state_dict_for("rescaled", seed), strict=True)
    return m.cuda().eval()


def test_config4_patch128_bf16_vs_fp32_validation_mode():
    """patch_shape=(128,128,128): the tcgen05 path (bf16) against the SIMT fp32 validation mode of
    the same library; the fp32 mode itself is pinned to the oracle (<= 1e-4) on smaller patches
    in test_gpu_forward / test_gpu_predict and here on a 128-wide sub-problem."""
    from aind_exaspim_neuron_segmentation_b200 import predict
    from oracle.predict_ref import predict_ref
    from oracle.unet_ref import make_forward_fn

    vol = lightsheet_volume((128, 128, 128), 31)
    kw = dict(patch_shape=(128, 128, 128), overlap=(32, 32, 32), trim=8)
    out16 = predict(vol, _model(30, "bf16"), verbose=False, **kw)
    out32 = predict(vol, _model(30, "fp32"), verbose=False, **kw)
    assert out16.shape == (3, 128, 128, 128) and out16.dtype == np.float32
    # exact-fit patch: non-zero only in [8, 120)^3
    assert out16[:, :8].max() == 0 and out16[:, 120:].max() == 0 and out16[:, 8:120, 8:120, 8:120].min() > 0
    err = float(np.abs(out16 - out32).max())
    print("P=128 bf16 vs fp32 validation mode: max abs", err)
    assert err <= BF16_TOL, err
    assert np.array_equal(out16 == 0, out32 == 0)
    # the same full 128^3 patch through the CPU oracle (one patch, a few seconds): every voxel
    ref128 = predict_ref(vol, make_forward_fn(state_dict_for("rescaled", 30)), **kw)
    e16, e32 = float(np.abs(out16 - ref128).max()), float(np.abs(out32 - ref128).max())
    print("P=128 vs CPU oracle: bf16", e16, "fp32 mode", e32)
    assert e16 <= BF16_TOL and e32 <= FP32_TOL, (e16, e32)
    assert np.array_equal(out16 == 0, ref128 == 0)
    # the fp32 mode against the CPU oracle on the same weights with a (128, 32, 32) patch
    small = lightsheet_volume((128, 32, 32), 32)
    kw2 = dict(patch_shape=(128, 32, 32), overlap=(32, 8, 8), trim=4)
    ref = predict_ref(small, make_forward_fn(state_dict_for("rescaled", 30)), **kw2)
    got = predict(small, _model(30, "fp32"), verbose=False, **kw2)
    assert float(np.abs(got - ref).max()) <= FP32_TOL
    got16 = predict(small, _model(30, "bf16"), verbose=False, **kw2)
    assert float(np.abs(got16 - ref).max()) <= BF16_TOL


def test_config5_watershed_adapted_rand_agreement():
    """predict -> affinities_to_segmentation, the product end to end, against the oracle end to end
    (oracle predict -> waterz restated in oracle/watershed_ref.py, parity unpinned): adapted-Rand
    agreement >= 0.99 with the reference's default thresholds (north_star); the product's
    segmentation of its own affinities equals the oracle's segmentation of the same array exactly.
    With random-init weights the default thresholds merge almost everything (SURVEY.md 8c caveat),
    so the fragment-level agreement (before agglomeration, where bf16 noise does move voxels) is
    checked as well."""
    from aind_exaspim_neuron_segmentation_b200 import affinities_to_segmentation, predict
    from oracle.predict_ref import predict_ref
    from oracle.unet_ref import make_forward_fn
    from oracle.watershed_ref import (adapted_rand_agreement, affinities_to_segmentation_ref,
                                      watershed_fragments)

    sd = state_dict_for("rescaled", 41)
    vol = lightsheet_volume((96, 160, 160), 42)
    out = predict(vol, _model(41, "bf16"), verbose=False)          # default 96^3 patches, 1x2x2
    ref = predict_ref(vol, make_forward_fn(sd))
    assert float(np.abs(out - ref).max()) <= BF16_TOL
    crop = (slice(None), slice(8, 88), slice(40, 120), slice(40, 120))
    seg_out = affinities_to_segmentation(out[crop]).astype(np.int64)
    assert np.array_equal(seg_out, affinities_to_segmentation_ref(out[crop]))
    seg_ref = affinities_to_segmentation_ref(ref[crop])
    score = adapted_rand_agreement(seg_out, seg_ref)
    frag_out, _ = watershed_fragments(out[crop])
    frag_ref, _ = watershed_fragments(ref[crop])
    frag_score = adapted_rand_agreement(frag_out, frag_ref)
    print("adapted-Rand agreement: segmentation", score, "fragments", frag_score,
          "segments", int(seg_ref.max()))
    assert score >= 0.99, score
    assert frag_score >= 0.9, frag_score


@pytest.mark.parametrize("patch,overlap,trim,shape", [
    ((32, 48, 80), (8, 16, 16), 4, (70, 100, 150)),     # anisotropic: mixes K1z2 / K1 paths per level
    ((64, 32, 112), (16, 8, 32), 8, (64, 60, 200)),     # levels with H % 16 != 0 and W % 8 != 0
])
def test_anisotropic_patch_shapes_match_oracle(patch, overlap, trim, shape):
    """Every multiple-of-16 patch shape must work: layers that do not meet a fast kernel's shape
    rules fall back to the generic tcgen05 kernel, never to wrong results."""
    from aind_exaspim_neuron_segmentation_b200 import predict
    from oracle.predict_ref import predict_ref
    from oracle.unet_ref import make_forward_fn

    sd = state_dict_for("rescaled", 51)
    vol = lightsheet_volume(shape, 52)
    kw = dict(patch_shape=patch, overlap=overlap, trim=trim)
    ref = predict_ref(vol, make_forward_fn(sd), **kw)
    out = predict(vol, _model(51, "bf16"), verbose=False, **kw)
    err = float(np.abs(out - ref).max())
    print(patch, "max abs err", err)
    assert err <= BF16_TOL, err
    assert np.array_equal(out == 0, ref == 0)


VARIANT_SCRIPT = """
import sys
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
from helpers import lightsheet_volume, state_dict_for
from aind_exaspim_neuron_segmentation_b200 import UNet3D, predict
from oracle.predict_ref import predict_ref
from oracle.unet_ref import make_forward_fn
sd = state_dict_for("rescaled", 61)
model = UNet3D(output_channels=3); model.load_state_dict(sd, strict=True); model = model.cuda().eval()
vol = lightsheet_volume((96, 112, 128), 62)
out = predict(vol, model, verbose=False)
ref = predict_ref(vol, make_forward_fn(sd))
print("MAXERR", float(np.abs(out - ref).max()), bool(np.array_equal(out == 0, ref == 0)))
"""


@pytest.mark.parametrize("env", [{"EXA_NO_PAIR": "1"}, {"EXA_NO_ZFOLD": "1"}, {"EXA_NO_TC_STEM": "1"},
                                 {"EXA_NO_MT2": "1"}, {"EXA_UP_CPT": "4"}, {"EXA_STITCH_OVERLAP": "1"}])
def test_alternative_kernel_variants_match_oracle(env):
    """The kernel variants behind the A/B knobs (one CTA per MMA, per-tap conv everywhere, SIMT
    stem, single M tile, 4-channel upsample threads, stitch on a second stream) stay correct: a
    96^3-patch predict() under each knob, in a fresh process, against the oracle."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = VARIANT_SCRIPT.format(root=root, tests=os.path.join(root, "tests"))
    res = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=600,
                         env={**os.environ, **env})
    assert res.returncode == 0, res.stderr[-2000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("MAXERR")][-1].split()
    assert float(line[1]) <= BF16_TOL and line[2] == "True", (env, line)
