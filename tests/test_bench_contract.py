"""bench.py contract pieces that need no GPU: the FLOP accounting behind the roofline numbers and
the JSON line of the reference arm (the CPU oracle port timed on a bounded sample)."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_flop_accounting_matches_survey_figures():
    import bench

    assert bench.patch_flops_algorithmic(96) == 370_145_230_848
    assert bench.patch_flops_algorithmic(128) == 877_381_287_936
    assert bench.conv_flops_executed(96) == bench.F_CONV_96 < bench.F_PATCH_96
    # trimmed layers: only the kept box (and that box grown by one voxel) is executed
    assert bench.layer_flops(17, 96) == 2 * 80 ** 3 * 32 * 27 * 32
    assert bench.layer_flops(16, 96) == 2 * 82 ** 3 * 32 * 27 * 64
    assert bench.layer_flops(16, 128) == 2 * 114 ** 3 * 32 * 27 * 64
    old = bench.PATCH
    try:
        bench.PATCH = (96,) * 3
        assert bench.n_patches_of((512,) * 3) == 512 and bench.n_patches_of((1024,) * 3) == 4096
        bench.PATCH = (128,) * 3
        assert bench.n_patches_of((512,) * 3) == 125 and bench.n_patches_of((1024,) * 3) == 1331
    finally:
        bench.PATCH = old
    assert bench.volume_shape(1) == (512, 512, 512) and bench.volume_shape(8) == (1024, 1024, 1024)


def test_training_step_flop_accounting():
    """`bench.py --workload train`: forward convs on full patches (nothing trimmed), data gradients
    for all layers but the stem, weight gradients for all."""
    import bench

    fl = bench.train_flops(16, 96)
    head = 2 * 96 ** 3 * 3 * 32
    assert fl["fprop"] == 16 * (bench.F_PATCH_96 - head) == fl["wgrad"] + fl["wgrad_stem"]
    assert fl["dgrad"] == fl["wgrad"] == fl["fprop"] - 16 * bench.F_STEM_96


def test_reference_arm_prints_one_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout[-2000:]          # nothing but the JSON line on stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "affinity voxels/sec"
    assert line["unit"] == "voxels/s" and line["higher_is_better"] is True
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["value"] > 0 and line["ms_per_step"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "voxels/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    base = line["cpu_baseline"]
    assert base["kind"] == "port" and base["cores"] >= 1 and base["value"] == line["value"]
    assert "sample" in base and line["config"]["workload"].startswith("predict() on a synthetic 512x512x512")


def test_train_reference_arm_prints_one_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "train",
                          "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout[-2000:]
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "training patch voxels/sec"
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["gpu_launches"] == 0
    assert line["config"]["batch"] == 16 and line["config"]["patch"] == [96, 96, 96]


def test_adapted_rand_on_device_equals_the_oracle_form():
    """bench.py --workload segment scores 1024^3 label volumes with a torch implementation of the
    oracle's adapted-Rand agreement; both must agree (here on CPU tensors)."""
    import numpy as np
    import torch

    import bench
    from oracle.watershed_ref import adapted_rand_agreement

    rng = np.random.default_rng(0)
    for _ in range(5):
        a = rng.integers(0, 9, (12, 14, 10))
        b = rng.integers(0, 6, (12, 14, 10))
        got = bench.adapted_rand_on_device(torch.from_numpy(a), torch.from_numpy(b))
        assert abs(got - adapted_rand_agreement(a, b)) < 1e-12
    z = np.zeros((4, 4, 4), np.int64)
    assert bench.adapted_rand_on_device(torch.from_numpy(z + 3), torch.from_numpy(z)) == 1.0
