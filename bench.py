#!/usr/bin/env python
"""Benchmark of the affinity-prediction hot path (BASELINE.json metric: affinity voxels/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the whole path over one synthetic uint16 volume:
histogram -> percentiles -> (gather+normalise, tensor-core stem, 17 tcgen05 convs, pools, upsamples,
fused head+sigmoid+trim) per wave of patches -> overlap stitch -> (N>1: halo exchange + all-gather).
N=1 runs BASELINE config 2 (512^3, patch 96^3 -> 512 patches); N GPUs run N x 512^3 voxels
(1024x512x512, 1024x1024x512, 1024^3 = BASELINE config 3), sharded by z patch-rows: weak scaling.
`--volume 1024` fixes the volume at 1024^3 for every N instead (the STRONG scaling curve of
SURVEY.md 8d-3; "scaling": "strong").

`value`   : voxels/s with the rank's uint16 slab already resident in HBM, device-timed (CUDA
            events), max over ranks.
`e2e`     : same metric through the public API with HOST buffers: pinned H2D of the slab and
            D2H of the rank's output planes inside the timed region.
`roofline`: the dominant kernel, conv3x3_zfold2_kernel (z-folded tcgen05 conv on CTA pairs; 8 of the
            17 conv layers, ~2/3 of the step): executed FLOPs of its launches / their summed CUDA
            event time inside the timed steps, against the measured dense bf16 peak
            (MEASURED_PEAKS.json, sustained figure: timed inside a long step).  `all_conv_kernels`
            and `per_layer` give the same for every conv launch.
`cpu_baseline` / `--impl reference`: the CPU oracle port of the reference path (oracle/),
            all host threads, on a bounded sample (4 patches) of the same workload.
`parity_check`: the bench's own output checked inside the run.  N=1: the corner block of the timed
            output against the oracle run of `cpu_baseline` (same 4 windows, the whole volume's
            percentiles).  N>1: a 512x160x160 volume predicted single-rank and sharded (fused
            peer-store gather AND the NCCL send/recv gather) on every rank -> bit_identical, and its
            corner against the oracle on rank 0 -> max_abs_vs_oracle.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PATCH = (96, 96, 96)
OVERLAP = (32, 32, 32)
TRIM = 8
F_PATCH_96 = 370_145_230_848          # 2*MACs of the 19 convs per 96^3 patch
F_STEM_96 = 2 * 96 ** 3 * 32 * 27     # inc.double_conv.0 (SIMT stem kernel)
F_HEAD_96 = 2 * 96 ** 3 * 3 * 32      # outc (fused into the last conv's epilogue)
# Work skipped because its results are trimmed away (inference.py:161-162): the last conv is
# evaluated on the kept 80^3 box only and up4.0 on that box grown by one voxel (82^3).  Only
# EXECUTED MACs are counted (no tile padding).
F_UP40_FULL = 2 * 96 ** 3 * 32 * 27 * 64
F_UP43_FULL = 2 * 96 ** 3 * 32 * 27 * 32
F_TRIM_SAVED = (F_UP40_FULL - 2 * 82 ** 3 * 32 * 27 * 64) + (F_UP43_FULL - 2 * 80 ** 3 * 32 * 27 * 32)
F_CONV_96 = F_PATCH_96 - F_STEM_96 - F_HEAD_96 - F_TRIM_SAVED  # executed by the tcgen05 convs

# (level, Cin, Cout) of conv layers 1..17 in the order of unet3d.py:64-74 (0 = the Cin=1 stem)
LAYER_SHAPES = {1: (0, 32, 32), 2: (1, 32, 64), 3: (1, 64, 64), 4: (2, 64, 128), 5: (2, 128, 128),
                6: (3, 128, 256), 7: (3, 256, 256), 8: (4, 256, 256), 9: (4, 256, 256),
                10: (3, 512, 256), 11: (3, 256, 128), 12: (2, 256, 128), 13: (2, 128, 64),
                14: (1, 128, 64), 15: (1, 64, 32), 16: (0, 64, 32), 17: (0, 32, 32)}


def layer_flops(i, p=96):
    """Executed FLOPs (2*MACs) of conv layer i per p^3 patch; the last two layers only compute the
    kept box ((p-16)^3) and that box grown by one voxel."""
    lvl, cin, cout = LAYER_SHAPES[i]
    side = {16: p - 2 * TRIM + 2, 17: p - 2 * TRIM}.get(i, p >> lvl)
    return 2 * side ** 3 * cout * 27 * cin


def conv_flops_executed(p=96):
    return sum(layer_flops(i, p) for i in LAYER_SHAPES)


def patch_flops_algorithmic(p=96):
    """2*MACs of all 19 convs on the full p^3 patch (370 145 230 848 at 96,
    877 381 287 936 at 128)."""
    full = sum(2 * (p >> lvl) ** 3 * cout * 27 * cin for lvl, cin, cout in LAYER_SHAPES.values())
    return full + 2 * p ** 3 * 32 * 27 + 2 * p ** 3 * 3 * 32


assert conv_flops_executed(96) == F_CONV_96 and patch_flops_algorithmic(96) == F_PATCH_96
assert patch_flops_algorithmic(128) == 877_381_287_936


def n_patches_of(shape):
    n = 1
    for d, pp, ov in zip(shape, PATCH, OVERLAP):
        n *= len(range(0, d - pp + (pp - ov), pp - ov))
    return n


STRONG_VOLUME = 0   # --volume: edge of the fixed cube for strong scaling (0 = weak scaling)


def volume_shape(n_gpus):
    if STRONG_VOLUME:
        return (STRONG_VOLUME,) * 3
    shape = [512, 512, 512]
    k, axis = n_gpus, 0
    while k > 1:
        shape[axis] *= 2
        k //= 2
        axis += 1
        if axis == 3:
            axis = 0
    return tuple(shape)


def synth_planes(shape, z0, z1, seed=1):
    """Planes [z0, z1) of a deterministic lightsheet-like uint16 volume of `shape`.

    Poisson(40) background plus sparse bright straight 'neurite' segments (peak intensity
    log-uniform 150..4000, 3-voxel cross-section) so that both the 1000-clip and the 99.9th
    percentile bite.  Generated in 64-plane chunks seeded by (seed, chunk) so that any rank can
    produce any plane range without building the whole volume.
    """
    d, h, w = shape
    out = np.empty((z1 - z0, h, w), np.uint16)
    for c in range(z0 // 64, (z1 + 63) // 64):
        rng = np.random.default_rng([seed, c])
        a, b = c * 64, min(c * 64 + 64, d)
        chunk = rng.poisson(40, (b - a, h, w)).astype(np.uint16)
        n_seg = max(8, (b - a) * h * w // 200_000)
        p0 = np.stack([rng.uniform(0, b - a, n_seg), rng.uniform(0, h, n_seg),
                       rng.uniform(0, w, n_seg)], 1)
        dirs = rng.normal(size=(n_seg, 3))
        dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        amp = np.exp(rng.uniform(np.log(150), np.log(4000), n_seg))
        t = np.arange(0, 60, 0.5)[None, :, None]
        pts = np.rint(p0[:, None, :] + dirs[:, None, :] * t).astype(np.int64)  # (n_seg, T, 3)
        val = np.broadcast_to(amp[:, None], pts.shape[:2]).ravel()
        pts = pts.reshape(-1, 3)
        for dz in (0, 1):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    q = pts + (dz, dy, dx)
                    ok = ((q[:, 0] >= 0) & (q[:, 0] < b - a) & (q[:, 1] >= 0) & (q[:, 1] < h)
                          & (q[:, 2] >= 0) & (q[:, 2] < w))
                    qq = q[ok]
                    np.maximum.at(chunk, (qq[:, 0], qq[:, 1], qq[:, 2]),
                                  np.minimum(val[ok] + 40, 65535).astype(np.uint16))
        lo, hi = max(a, z0), min(b, z1)
        out[lo - z0:hi - z0] = chunk[lo - a:hi - a]
    return out


# --- clocks -------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "200", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_tflops": p.get("bf16_tflops_sustained", p.get("bf16_tflops")),
                "hbm_gbs": p.get("hbm_gbs"), "source": "measured (MEASURED_PEAKS.json, sustained)"}
    return {"bf16_tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def hbm_kernels(prof, n_patches, voxels, steps, peak_gbs):
    """Achieved GB/s of the memory-bound kernels from their ALGORITHMIC bytes
    and their summed CUDA-event time, against the measured copy bandwidth."""
    up_bytes = 0
    for lvl_out, c in ((3, 256), (2, 128), (1, 64)):     # up1..up3: full output, input at half res
        n = 96 >> lvl_out
        up_bytes += n ** 3 * c * 2 + (n // 2) ** 3 * c * 2
    up_bytes += 84 ** 3 * 32 * 2 + 42 ** 3 * 32 * 2       # up4: the 84^3 box the last convs need
    algo = {
        # 2 B of uint16 in + 64 B of bf16 activations out per patch voxel (the bf16
        # (hi, lo) intermediate between the gather and the tensor-core stem is NOT algorithmic)
        "stem": n_patches * 96 ** 3 * (2 + 64),
        "upsample": n_patches * up_bytes,
        # stitch: every trimmed patch voxel read once (3 x 4 B), every output voxel written once
        "stitch": n_patches * 3 * 80 ** 3 * 4 + voxels * 12,
        "histogram": voxels * 2,
    }
    out = {}
    for k, b in algo.items():
        ms = prof[k][0] / steps
        if ms > 0:
            gbs = b / (ms * 1e-3) / 1e9
            out[k] = {"algorithmic_bytes": b, "ms_per_step": ms, "achieved_gbs": gbs,
                      "frac_of_measured_hbm": gbs / peak_gbs}
    return out


def ncu_traffic():
    """Per-launch DRAM bytes of the dominant conv kernel from the committed ncu capture."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("traffic_bytes_per_launch")
    return None


# --- CPU oracle leg ----------------------------------------------------------------------
CPU_SAMPLE_SHAPE = (96, 160, 160)  # 1 x 2 x 2 patches of the same synthetic volume


# voxels of the corner covered by its own 1 x 2 x 2 windows only (the next windows' kept boxes
# start at 64 + 8 along z and 128 + 8 along y, x): there the corner run equals the whole-volume run
CPU_SAMPLE_VALID = (slice(None), slice(0, 72), slice(0, 136), slice(0, 136))


def cpu_oracle_step(sd, vol, norm_range=None):
    from oracle.predict_ref import predict_ref
    from oracle.unet_ref import make_forward_fn

    t0 = time.perf_counter()
    out = predict_ref(vol, make_forward_fn(sd), patch_shape=PATCH, overlap=OVERLAP, trim=TRIM,
                      norm_range=norm_range)
    return time.perf_counter() - t0, out


def cpu_baseline(steps=1, warmup=0, norm_range=None, shape=None, want_output=False):
    """The oracle port timed on a 4-patch corner of the workload volume.  With norm_range (the
    whole volume's percentiles) its output is the reference's result for that corner of the whole
    volume (CPU_SAMPLE_VALID) and is returned for the parity check."""
    import torch

    from oracle.unet_ref import rescaled_state_dict

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = rescaled_state_dict(0)
    shape = shape or volume_shape(1)
    vol = synth_planes(shape, 0, 96)[:, :160, :160].copy()
    for _ in range(warmup):
        cpu_oracle_step(sd, vol, norm_range)
    times, out = [], None
    for _ in range(max(steps, 1)):
        sec, out = cpu_oracle_step(sd, vol, norm_range)
        times.append(sec)
    sec = float(np.mean(times))
    n_patch = 4
    vox_per_patch = 64 ** 3  # stitched output voxels per patch at stride 64 (512^3 / 512 patches)
    base = {"value": n_patch * vox_per_patch / sec, "unit": "voxels/s", "cores": cores,
            "kind": "port",
            "sample": (f"oracle/ CPU port of reference predict (torch fp32 convs, {cores} threads) on "
                       f"a {CPU_SAMPLE_SHAPE} corner of the same volume = 4 patches, "
                       f"{sec:.2f} s/step; scaled by the workload's 262144 output voxels per patch"),
            "sec_per_patch": sec / n_patch}
    return (base, sec, out) if want_output else (base, sec)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, sec = cpu_baseline(steps=args.steps, warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "affinity voxels/sec", "value": base["value"],
        "unit": "voxels/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong" if STRONG_VOLUME else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "voxels/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(n):
    shape = volume_shape(n)
    return {"workload": f"predict() on a synthetic {shape[0]}x{shape[1]}x{shape[2]} uint16 volume, "
                        f"patch {PATCH[0]}^3 overlap 32 trim 8 ({n_patches_of(shape)} "
                        "patches), random-init UNet3D, affinity_mode=True",
            "volume": list(shape), "patch_shape": list(PATCH), "overlap": list(OVERLAP), "trim": TRIM,
            "sharding": f"z patch-rows over {n} GPU(s)" if n > 1 else "single GPU",
            "l2": "per-wave activations (~8 GB) and the volume are far larger than the 126 MB L2; "
                  "no flush needed between steps"}


def recorded_checksum(n, shape):
    """Checksums of outputs that tests/test_gpu_predict.py checks against the oracle."""
    path = os.path.join(ROOT, "tests", "golden", "bench_checksum.json")
    if not os.path.exists(path) or PATCH != (96, 96, 96):
        return None
    with open(path) as f:
        rec = json.load(f)
    if tuple(shape) == (512, 512, 512):
        return rec.get("n1_512")
    if tuple(shape) == (1024, 1024, 1024):
        return rec.get("n8_1024")
    return None


PARITY_SHAPE = (512, 160, 160)   # 8 x 2 x 2 windows: every rank of up to 8 owns at least one z row


def sharded_parity_check(model, dev, rank, world):
    """N-GPU == 1-GPU, bit for bit, on every rank and for both gather paths; rank 0 also checks the
    corner of that volume against the CPU oracle.  Runs before the timed region."""
    import torch
    import torch.distributed as dist

    from aind_exaspim_neuron_segmentation_b200 import _native, predict
    from aind_exaspim_neuron_segmentation_b200.inference import SlabJob, _EngineSlabBackend

    vol = synth_planes(PARITY_SHAPE, 0, PARITY_SHAPE[0], seed=7)
    single = predict(vol, model, verbose=False, patch_shape=PATCH, overlap=OVERLAP, trim=TRIM)
    params = _native.make_params(PATCH, OVERLAP, TRIM, 1000, (1, 99.9), batch=32)
    backend = _EngineSlabBackend(model.engine("bf16"))
    same, modes = True, []
    saved = os.environ.get("EXA_GATHER")
    for gather_env in ("", "store", "nccl"):
        os.environ["EXA_GATHER"] = gather_env
        job = SlabJob(PARITY_SHAPE, params, 3, backend)
        for _ in range(2):   # the second run overwrites the peers' previous result in place
            full = job.run(job.upload(vol), gather=True)
            same = same and bool(np.array_equal(full.cpu().numpy(), single))
        modes.append(("fused peer-store" if gather_env == "store" else "copy-engine peer copies")
                     if job._fused else "nccl send/recv")
        z0, z1 = job.own_bounds()
        host_own = torch.empty((3, z1 - z0) + PARITY_SHAPE[1:], dtype=torch.float32).pin_memory()
        job.run_pipelined(job.upload(vol), host_own)
        same = same and bool(np.array_equal(host_own.numpy(), single[:, z0:z1]))
        del job
    if saved is None:
        os.environ.pop("EXA_GATHER", None)
    else:
        os.environ["EXA_GATHER"] = saved
    flag = torch.tensor([1 if same else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out = {"bit_identical": bool(flag.item() == 1), "gather_modes": modes, "volume": list(PARITY_SHAPE),
           "what": "predict() on one GPU vs SlabJob.run(gather) and run_pipelined on every rank"}
    if rank == 0:
        from oracle.predict_ref import predict_ref
        from oracle.unet_ref import make_forward_fn, rescaled_state_dict

        torch.set_num_threads(os.cpu_count() or 1)
        mn, mx = (float(v) for v in np.percentile(np.minimum(vol, 1000), (1, 99.9)))
        ref = predict_ref(vol[:96], make_forward_fn(rescaled_state_dict(0)), patch_shape=PATCH,
                          overlap=OVERLAP, trim=TRIM, norm_range=(mn, mx))[:, :72]
        out["max_abs_vs_oracle"] = float(np.abs(single[:, :72] - ref).max())
        out["tolerance"] = 1e-2
        out["ok"] = bool(out["bit_identical"] and out["max_abs_vs_oracle"] <= 1e-2)
        print(f"SHARDED_OK world={world} {json.dumps(out)}" if out["ok"] else
              f"SHARDED_MISMATCH world={world} {json.dumps(out)}", file=sys.stderr, flush=True)
    dist.barrier()
    return out



# --- config 5: predict -> affinities_to_segmentation -----------------------
SEG_THRESHOLDS = [0.6, 0.8, 0.9]      # reference defaults (inference.py:198-199)
SEG_MIN_SIZE = 100


def ws_profile():
    import ctypes

    from aind_exaspim_neuron_segmentation_b200 import _native

    buf = (ctypes.c_double * 8)()
    _native.check(_native.lib().exa_ws_last_profile(buf, 8), None, "exa_ws_last_profile")
    v = list(buf)
    return {"fragments_ms": v[0], "region_graph_ms": v[1], "parallel_rounds_ms": v[2],
            "host_queue_ms": v[3], "sizes_relabel_ms": v[4], "parallel_rounds": int(v[5]),
            "region_edges": int(v[6]), "edges_to_host_queue": int(v[7])}


def adapted_rand_on_device(seg, ref):
    """oracle.watershed_ref.adapted_rand_agreement (SURVEY.md 8c iv) on CUDA int64 tensors, for the
    1024^3 labels that numpy would take minutes to sort; small volumes use the oracle's numpy form."""
    import torch

    seg, ref = seg.reshape(-1), ref.reshape(-1)
    m = ref != 0
    seg, ref = seg[m], ref[m]
    if seg.numel() == 0:
        return 1.0
    pair = ref * (int(seg.max().item()) + 1) + seg
    _, pij = torch.unique(pair, return_counts=True)
    del pair
    _, ai = torch.unique(ref, return_counts=True)
    _, bj = torch.unique(seg, return_counts=True)
    num = 2.0 * (pij.double() ** 2).sum()
    return float((num / ((ai.double() ** 2).sum() + (bj.double() ** 2).sum())).item())


def run_segment(args):
    """`--workload segment`: BASELINE config 5 on one GPU.  predict() on the synthetic volume
    (`--volume` edge, default 512), then a step = affinities_to_segmentation on the affinities.
    `value`: voxels/s with the float32 affinities resident in HBM (label volume left in HBM);
    `e2e`: host float32 affinities in -> host uint64 labels out through the public call;
    `cpu_baseline`: the compiled restatement of waterz (oracle/ws_ref.cpp, one core) on a 128^3
    corner of the same affinities, whose labels the product must reproduce exactly
    (`parity_check`); `adapted_rand`: the product's segmentation of its bf16 affinities against
    its segmentation of the FP32-validation-mode affinities (north_star: >= 0.99)."""
    import torch

    from aind_exaspim_neuron_segmentation_b200 import UNet3D, affinities_to_segmentation, predict
    from oracle.unet_ref import rescaled_state_dict  # weights recipe only (seeded random init)

    torch.cuda.set_device(0)
    edge = args.volume or 512
    shape = (edge,) * 3
    model = UNet3D(output_channels=3)
    model.load_state_dict(rescaled_state_dict(0), strict=True)
    model = model.cuda().eval()
    vol = synth_planes(shape, 0, edge)
    t0 = time.perf_counter()
    aff = predict(vol, model, verbose=False, patch_shape=PATCH, overlap=OVERLAP, trim=TRIM)
    predict_s = time.perf_counter() - t0
    dev = torch.from_numpy(aff).cuda()
    voxels = float(edge) ** 3

    def step():
        return affinities_to_segmentation(dev, SEG_THRESHOLDS, SEG_MIN_SIZE)

    # at 1024^3 the sort buffers of this noise-like input take most of the 180 GB: one cold step,
    # no second copy of the affinities on the device (no e2e leg), no separate fragment count
    lean = edge >= 1024
    warm = 0 if lean else max(args.warmup, 1)
    steps = 1 if lean else args.steps
    for _ in range(warm):
        seg = step()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    phases = []
    t0 = time.perf_counter()
    for _ in range(steps):
        seg = step()
        phases.append(ws_profile())
    torch.cuda.synchronize()
    sec = (time.perf_counter() - t0) / steps
    clocks = sampler.stop()
    n_segments = int(seg.max().item())
    n_fragments = None
    if not lean:
        frags = affinities_to_segmentation(dev, [0.0], 0)
        n_fragments = int(frags.max().item())
        del frags

    # e2e: numpy in, numpy out
    e2e_sec, same = None, None
    if not lean:
        affinities_to_segmentation(aff, SEG_THRESHOLDS, SEG_MIN_SIZE)
        t0 = time.perf_counter()
        for _ in range(steps):
            seg_host = affinities_to_segmentation(aff, SEG_THRESHOLDS, SEG_MIN_SIZE)
        e2e_sec = (time.perf_counter() - t0) / steps
        same = bool(np.array_equal(seg_host.astype(np.int64), seg.cpu().numpy()))
        del seg_host
    seg_np = seg.cpu().numpy()
    del seg, dev
    torch.cuda.empty_cache()
    from aind_exaspim_neuron_segmentation_b200 import _native
    _native.lib().exa_ws_release_memory()   # the FP32-validation predict below wants the memory

    line = {
        "metric": "segmentation voxels/sec", "value": voxels / sec, "unit": "voxels/s", "n_gpus": 1,
        "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 fixed point",
        "data": "synthetic",
        "config": {"workload": f"affinities_to_segmentation(predict(synthetic {edge}^3 uint16 volume)), "
                               f"thresholds {SEG_THRESHOLDS}, min_segment_size {SEG_MIN_SIZE}",
                   "volume": list(shape), "n_fragments": n_fragments, "n_segments": n_segments,
                   "predict_s": predict_s,
                   "l2": "12 B of affinities per voxel (1.6 GB at 512^3) are far larger than the L2"},
        "e2e": None if e2e_sec is None else {
            "value": voxels / e2e_sec, "unit": "voxels/s", "ms_per_step": e2e_sec * 1e3,
            "h2d_bytes_per_step": int(voxels) * 12, "d2h_bytes_per_step": int(voxels) * 8,
            "equal_to_device_path": same,
            "note": "numpy float32 (3,D,H,W) in -> numpy uint64 (D,H,W) out (pageable buffers)"},
        "gpu_launches": None,
        "phases_last_step": phases[-1],
        "clocks": clocks,
    }
    # wall share of the host queue: what bounds this row
    line["roofline"] = {"bound": "host queue / hbm", "achieved": None, "peak": None, "unit": "GB/s",
                        "frac": None, "traffic": None,
                        "note": "the voxel-sized kernels are HBM-class work; the step is bounded by "
                                "the sequential tail of the merge queue (host_queue_ms)"}
    if not args.no_cpu:
        from oracle.watershed_ref import adapted_rand_agreement, affinities_to_segmentation_ref

        c = min(128, edge)
        corner = np.ascontiguousarray(aff[:, :c, :c, :c])
        t0 = time.perf_counter()
        ref = affinities_to_segmentation_ref(corner, SEG_THRESHOLDS, SEG_MIN_SIZE)
        cpu_s = time.perf_counter() - t0
        got = affinities_to_segmentation(corner, SEG_THRESHOLDS, SEG_MIN_SIZE).astype(np.int64)
        line["cpu_baseline"] = {"value": c ** 3 / cpu_s, "unit": "voxels/s", "cores": 1, "kind": "port",
                                "sample": f"oracle/ws_ref.cpp (compiled restatement of waterz) on the "
                                          f"{c}^3 corner of the same affinities, {cpu_s:.2f} s"}
        line["parity_check"] = {"labels_identical_to_oracle": bool(np.array_equal(got, ref)),
                                "volume": [c, c, c], "n_segments": int(ref.max())}
        # config 5's agreement: bf16 vs FP32-validation affinities through the same segmentation
        aff32 = predict(vol, model, verbose=False, patch_shape=PATCH, overlap=OVERLAP, trim=TRIM,
                        precision="fp32")
        seg32 = affinities_to_segmentation(torch.from_numpy(aff32).cuda(), SEG_THRESHOLDS, SEG_MIN_SIZE)
        if lean:
            torch.cuda.empty_cache()
            agreement = adapted_rand_on_device(torch.from_numpy(seg_np).cuda(), seg32)
        else:
            agreement = adapted_rand_agreement(seg_np, seg32.cpu().numpy())
            assert abs(agreement - adapted_rand_on_device(torch.from_numpy(seg_np).cuda(), seg32)) < 1e-12
        line["adapted_rand"] = {
            "agreement": agreement,
            "max_abs_affinity_diff": max(float(np.abs(aff[c] - aff32[c]).max()) for c in range(3)),
            "what": "segmentation of the bf16 affinities vs segmentation of the FP32-validation "
                    "affinities (both by the product), ignoring voxels that are background in the latter"}
        if not lean:
            # the default thresholds merge everything on random-init weights (one segment, SURVEY.md
            # 8c): the same comparison where the segmentation is NOT degenerate -- the watershed
            # fragments themselves and two lower thresholds
            del seg32
            dev16, dev32 = torch.from_numpy(aff).cuda(), torch.from_numpy(aff32).cuda()
            extra = []
            for thr, min_size in (([0.0], 0), ([0.45], 0), ([0.5], 0), ([0.55], 0)):
                a = affinities_to_segmentation(dev16, thr, min_size)
                b = affinities_to_segmentation(dev32, thr, min_size)
                extra.append({"thresholds": thr, "min_segment_size": min_size,
                              "segments_bf16": int(a.max().item()), "segments_fp32": int(b.max().item()),
                              "agreement": adapted_rand_on_device(a, b)})
                del a, b
            line["adapted_rand"]["non_degenerate"] = extra
    emit(line)


# --- training step -------------------------------------------------------------
def train_flops(batch, p=96):
    """Executed FLOPs of one step: forward convs (all 18, full patches -- nothing is trimmed in
    training), data gradients (17: the stem's input needs none) and weight gradients (18)."""
    stem = 2 * p ** 3 * 32 * 27
    convs = sum(2 * (p >> lvl) ** 3 * cout * 27 * cin for lvl, cin, cout in LAYER_SHAPES.values())
    # "wgrad" = the 17 tensor-core weight gradients (the Cin = 1 stem's runs on its own SIMT kernel
    # and is timed under its own category)
    return {"fprop": batch * (convs + stem), "dgrad": batch * convs, "wgrad": batch * convs,
            "wgrad_stem": batch * stem}


def run_train(args):
    """`--workload train`: one training step of reference train.py:136-140 (forward_pass + backward)
    on one GPU, B = 16 patches of 96^3 (Trainer / dataset defaults, train.py:39, data_handling.py:33).
    `value`: patch voxels/s with x and y resident in HBM; `e2e`: host x, y in -> loss on the host,
    every parameter's .grad on the device, through model(x) / criterion / loss.backward();
    `cpu_baseline`: the oracle's torch-autograd step (oracle/train_ref.py) on the host cores, B = 2."""
    import torch

    from aind_exaspim_neuron_segmentation_b200 import UNet3D
    from aind_exaspim_neuron_segmentation_b200.machine_learning.training import trainer_for_module
    from oracle.train_ref import train_inputs_structured  # seeded inputs only
    from oracle.unet_ref import rescaled_state_dict      # weights recipe only

    torch.cuda.set_device(0)
    B, P = args.train_batch, args.patch
    model = UNet3D(output_channels=3)
    model.load_state_dict(rescaled_state_dict(0), strict=True)
    model = model.cuda().train()
    crit = torch.nn.BCEWithLogitsLoss()
    x_host, y_host = train_inputs_structured(1, B, (P, P, P))
    x_host, y_host = x_host.pin_memory(), y_host.pin_memory()
    x_dev, y_dev = x_host.cuda(), y_host.cuda()
    voxels = float(B) * P ** 3

    def step(x, y):
        model.zero_grad(set_to_none=True)
        hat_y = model(x)
        loss = crit(hat_y, y)
        loss.backward()
        return loss

    def step_e2e():
        x = x_host.cuda(non_blocking=True)
        y = y_host.cuda(non_blocking=True)
        return float(step(x, y).detach().cpu())

    for _ in range(max(args.warmup, 1)):
        loss0 = step(x_dev, y_dev)
    torch.cuda.synchronize()
    eng = trainer_for_module(model, model.precision)
    launches0 = eng.launch_count
    sampler = ClockSampler(0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(x_dev, y_dev)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop()
    launches = (eng.launch_count - launches0) // args.steps
    # per-category pass, separately (events around every launch)
    eng.profile_begin()
    step(x_dev, y_dev)
    prof = eng.profile_end()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss_e2e = step_e2e()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) / args.steps * 1e3
    fl = train_flops(B, P)
    peaks = measured_peaks()
    cat_ms = {k: v[0] for k, v in prof.items()}
    tflops = {k: fl[k] / (cat_ms[k] * 1e-3) / 1e12 for k in ("fprop", "dgrad", "wgrad") if cat_ms[k] > 0}
    line = {
        "metric": "training patch voxels/sec", "value": voxels / (ms * 1e-3), "unit": "voxels/s",
        "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": train_workload_text(B, P), "batch": B, "patch": [P, P, P],
                   "l2": f"activations of one step ({eng.workspace_bytes / 2**30:.1f} GiB workspace) "
                         "are far larger than the L2",
                   "workspace_gib": eng.workspace_bytes / 2 ** 30},
        "e2e": {"value": voxels / (e2e_ms * 1e-3), "unit": "voxels/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(voxels) * 4 * 4, "d2h_bytes_per_step": 4,
                "note": "pinned float32 x (B,1,P,P,P) and y (B,3,P,P,P) in -> float loss out; "
                        "gradients stay on the device for the optimizer, as in train.py:139-147"},
        "gpu_launches": int(launches),
        "loss": float(loss0.detach()), "loss_e2e": loss_e2e,
        "per_category_ms": cat_ms, "per_category_launches": {k: v[1] for k, v in prof.items()},
        "tflops": tflops,
        "roofline": {"bound": "tensor", "achieved": tflops.get("wgrad"), "peak": peaks["bf16_tflops"],
                     "unit": "TFLOP/s",
                     "frac": (tflops.get("wgrad") or 0) / peaks["bf16_tflops"], "traffic": None,
                     "kernel": "wgrad_tc_kernel (17 weight gradients on tcgen05, MN-major operands): "
                               "executed FLOPs = batch x 368.4 GFLOP (all 3x3x3 convs but the stem) "
                               "/ summed event time of its 17 launches", "peak_source": peaks["source"]},
        "clocks": clocks,
    }
    if not args.no_cpu:
        from oracle.train_ref import train_step_ref

        torch.set_num_threads(os.cpu_count())
        xs, ys = x_host[:2].clone(), y_host[:2].clone()
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        t0 = time.perf_counter()
        train_step_ref(xs, ys, sd)
        cpu_s = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": 2 * P ** 3 / cpu_s, "unit": "voxels/s", "cores": os.cpu_count(),
                                "kind": "port",
                                "sample": f"oracle/train_ref.py (torch autograd, fp32) on 2 patches of "
                                          f"{P}^3, {cpu_s:.2f} s"}
    emit(line)


def train_workload_text(B, P):
    return (f"one training step (forward with batch-statistics BatchNorm, BCEWithLogitsLoss, "
            f"backward to all 74 parameter gradients) on {B} patches of {P}^3, "
            "UNet3D(output_channels=3)")


def run_train_reference(args):
    """`--workload train --impl reference`: the oracle's torch-autograd step (oracle/train_ref.py,
    the reference's own CPU arithmetic) on the host cores, each step a bounded sample of the
    workload (2 of the B patches)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from oracle.train_ref import train_inputs_structured, train_step_ref
    from oracle.unet_ref import rescaled_state_dict

    B, P = args.train_batch, args.patch
    torch.set_num_threads(os.cpu_count())
    sd = rescaled_state_dict(0)
    x, y = train_inputs_structured(1, 2, (P, P, P))
    for _ in range(min(args.warmup, 1)):
        train_step_ref(x, y, sd)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        train_step_ref(x, y, sd)
    sec = (time.perf_counter() - t0) / args.steps
    value = 2 * P ** 3 / sec
    base = {"value": value, "unit": "voxels/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"oracle/train_ref.py (torch autograd, fp32) on 2 of the {B} patches, "
                      f"{sec:.2f} s/step"}
    emit({"impl": "reference", "metric": "training patch voxels/sec", "value": value,
          "unit": "voxels/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
          "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
          "dtype": "f32", "data": "synthetic",
          "config": {"workload": train_workload_text(B, P), "batch": B, "patch": [P, P, P]},
          "cpu_baseline": base,
          "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0,
                  "d2h_bytes_per_step": 0},
          "gpu_launches": 0})


def run_train_library(args):
    """The reference's own GPU path for the same step: torch autograd over cuDNN with fp16
    autocast and GradScaler-style loss scaling (train.py:79-84,140), on the oracle's functional
    restatement of the model.  Not part of the bench contract; "impl": "library"."""
    import contextlib

    import torch

    from oracle.train_ref import is_parameter, train_inputs_structured, unet_forward_train
    from oracle.unet_ref import rescaled_state_dict

    dev = torch.device("cuda", 0)
    B, P = args.train_batch, args.patch
    p = {}
    for k, v in rescaled_state_dict(0).items():
        if v.dtype == torch.int64:
            continue
        t = v.to(dev)
        if is_parameter(k):
            t.requires_grad_(True)
        p[k] = t
    x, y = train_inputs_structured(1, B, (P, P, P))
    x, y = x.to(dev), y.to(dev)
    torch.backends.cudnn.benchmark = True
    out = {}
    for name, dtype in (("fp16_autocast", torch.float16), ("bf16_autocast", torch.bfloat16),
                        ("tf32", None)):
        def step():
            for t in p.values():
                t.grad = None
            ctx = torch.autocast("cuda", dtype=dtype) if dtype else contextlib.nullcontext()
            with ctx:
                logits = unet_forward_train(x, p, {})
                loss = torch.nn.functional.binary_cross_entropy_with_logits(logits.float(), y)
            (loss * 1024.0).backward()
            return loss
        try:
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            out[name] = {"ms_per_step": ms, "voxels_per_s": B * P ** 3 / (ms * 1e-3),
                         "peak_mem_gib": torch.cuda.max_memory_allocated() / 2 ** 30}
        except Exception as exc:  # noqa: BLE001
            out[name] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
        torch.cuda.empty_cache()
    emit({"impl": "library", "what": f"PyTorch/cuDNN training step (forward, BCE, backward) of the "
          f"same UNet3D, B={B} x {P}^3", "torch": torch.__version__,
          "cudnn": torch.backends.cudnn.version(), "settings": out})

# --- B200 arm ------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    from aind_exaspim_neuron_segmentation_b200 import UNet3D, _native
    from aind_exaspim_neuron_segmentation_b200.inference import SlabJob, _EngineSlabBackend
    from oracle.unet_ref import rescaled_state_dict  # weights recipe only (seeded random init)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n = max(world, 1)
    shape = volume_shape(n)
    model = UNet3D(output_channels=3)
    model.load_state_dict(rescaled_state_dict(0), strict=True)
    model = model.to(dev).eval()
    engine = model.engine("bf16")
    params = _native.make_params(PATCH, OVERLAP, TRIM, 1000, (1, 99.9), batch=args.batch)
    job = SlabJob(shape, params, 3, _EngineSlabBackend(engine), None)
    z0, z1 = job.slab_bounds()
    host = torch.from_numpy(synth_planes(shape, z0, z1)).pin_memory()
    slab = host.to(dev, non_blocking=True)
    torch.cuda.synchronize()
    oz0, oz1 = job.own_bounds()
    host_out = torch.empty((3, oz1 - oz0, shape[1], shape[2]), dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return job.run(slab, gather=True)

    from aind_exaspim_neuron_segmentation_b200 import predict

    def step_e2e():
        if n == 1:
            # the public call a user makes: host uint16 volume in, host float32 affinities out
            # (pinned buffers; H2D, all kernels and the row-pipelined D2H inside the call)
            return predict(host.numpy(), model, verbose=False, patch_shape=PATCH, overlap=OVERLAP,
                           trim=TRIM, batch_size=args.batch, out=host_out.numpy())
        # sharded form of the same call (predict_sharded(gather=False, out=...)): every rank
        # uploads its slab and gets its owned planes in pinned host memory, D2H row-pipelined
        return job.run_pipelined(host.to(dev, non_blocking=True), host_out)

    # ---- N>1: correctness of the sharded forms, inside the run the driver scales (SCALE_rNN) ----
    parity = sharded_parity_check(model, dev, rank, world) if world > 1 else None

    for _ in range(max(args.warmup, 3)):
        out = step_device()
    del out
    barrier()

    # ---- device-timed region: K steps, no per-launch instrumentation ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = engine.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        out = step_device()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = engine.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    # ---- per-kernel pass: the same K steps again with every launch bracketed by CUDA events
    # (roofline, per_layer, kernel_ms_per_step); not part of `value` ----
    engine.profile_begin()
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for _ in range(args.steps):
        out = step_device()
    pe1.record()
    barrier()
    prof_ms_step = pe0.elapsed_time(pe1) / args.steps
    prof = engine.profile_end()
    layers = engine.profile_layers()
    checksum = float(out[:, ::37, ::41, ::43].double().sum().item())
    corner = out[CPU_SAMPLE_VALID].cpu().numpy() if n == 1 else None   # for the parity check below
    del out
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    voxels = float(np.prod(shape))

    # ---- end-to-end (host buffers, pinned H2D + D2H inside the timed region) ----
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    barrier()
    wall = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([max(e0.elapsed_time(e1), wall)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = t.item() / args.steps
    h2d = torch.tensor([host.numel() * 2], dtype=torch.float64, device=dev)
    d2h = torch.tensor([host_out.numel() * 4], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(h2d)
        dist.all_reduce(d2h)

    if rank == 0:
        peaks = measured_peaks()
        n_patches_rank = (job.rows[1] - job.rows[0]) * n_patches_of((PATCH[0],) + tuple(shape[1:]))
        conv_ms, conv_launches = prof["conv"]
        conv_flops = n_patches_rank * conv_flops_executed(PATCH[0]) * args.steps
        achieved_all = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        # the dominant kernel: the z-folded CTA-pair conv (conv_zfold2.cuh)
        per_layer, dom_ms, dom_flops, dom_launches = {}, 0.0, 0.0, 0
        dom_name = "conv3x3_zfold2_kernel"
        for i, (lms, ln, lname) in enumerate(layers):
            if i == 0 or ln == 0:
                continue
            fl = n_patches_rank * layer_flops(i, PATCH[0]) * args.steps
            per_layer[str(i)] = {"kernel": lname, "cin": LAYER_SHAPES[i][1], "cout": LAYER_SHAPES[i][2],
                                 "ms_per_launch": lms / ln, "tflops": fl / (lms * 1e-3) / 1e12}
            if lname == dom_name:
                dom_ms += lms
                dom_flops += fl
                dom_launches += ln
        if dom_launches == 0:   # pair kernel disabled (EXA_NO_PAIR / EXA_NO_ZFOLD): fall back to all convs
            dom_name, dom_ms, dom_flops, dom_launches = "conv3x3 kernels (all)", conv_ms, conv_flops, conv_launches
        achieved = dom_flops / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
        line = {
            "metric": "affinity voxels/sec", "value": voxels / (ms_step * 1e-3), "unit": "voxels/s",
            "n_gpus": n, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if STRONG_VOLUME else "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(n),
            "e2e": {"value": voxels / (e2e_ms * 1e-3), "unit": "voxels/s",
                    "h2d_bytes_per_step": int(h2d.item()), "d2h_bytes_per_step": int(d2h.item()),
                    "ms_per_step": e2e_ms,
                    "note": ("predict(host uint16, out=pinned float32): H2D + kernels + row-pipelined D2H"
                             if n == 1 else
                             "pinned uint16 slab H2D + owned fp32 planes D2H per step, every rank")},
            "gpu_launches": int(launches),
            "roofline": {
                "kernel": f"{dom_name} ({dom_launches // max(args.steps, 1)} launches per step on rank 0)",
                "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"],
                "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops"],
                "peak_source": peaks["source"], "traffic": ncu_traffic(),
                "launches": dom_launches, "avg_launch_ms": dom_ms / max(dom_launches, 1),
                "flops_per_launch": dom_flops / max(dom_launches, 1),
                "share_of_step": dom_ms / (prof_ms_step * args.steps),
                "timing": ("per-launch CUDA events in a separate pass of the same K steps "
                           f"({prof_ms_step:.2f} ms/step instrumented vs {ms_step:.2f} timed)"),
                "all_conv_kernels": {"achieved": achieved_all, "frac": achieved_all / peaks["bf16_tflops"],
                                     "launches": conv_launches, "share_of_step": conv_ms / (prof_ms_step * args.steps),
                                     "flops_per_patch": conv_flops_executed(PATCH[0]),
                                     "flops_per_patch_untrimmed": patch_flops_algorithmic(PATCH[0])},
                "per_layer": per_layer,
            },
            "kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()},
            "hbm_kernels": hbm_kernels(prof, n_patches_rank, voxels / n, args.steps, peaks["hbm_gbs"]),
            "clocks": clocks,
            "checksum": checksum,
        }
        want = recorded_checksum(n, shape)
        if want is not None:
            line["checksum_ok"] = bool(abs(checksum - want["value"]) <= want["tol"])
        if parity is not None:
            line["parity_check"] = parity
        if n == 1 and not args.no_cpu:
            # the oracle run that is timed as the CPU baseline doubles as the parity check: same 4
            # windows, the WHOLE volume's percentiles -> the reference's result for that corner
            mn, mx = (float(v) for v in np.percentile(np.minimum(host.numpy(), 1000), (1, 99.9)))
            line["cpu_baseline"], _, ref = cpu_baseline(steps=1, norm_range=(mn, mx), shape=shape,
                                                        want_output=True)
            ref = ref[CPU_SAMPLE_VALID]
            line["parity_check"] = {
                "max_abs_vs_oracle": float(np.abs(corner - ref).max()), "tolerance": 1e-2,
                "zero_shell_identical": bool(np.array_equal(corner == 0, ref == 0)),
                "region": "corner [0:72, 0:136, 0:136] of the timed output (1x2x2 windows) vs the oracle "
                          "run of cpu_baseline with the whole volume's percentiles"}
            line["parity_check"]["ok"] = bool(line["parity_check"]["max_abs_vs_oracle"] <= 1e-2 and
                                              line["parity_check"]["zero_shell_identical"])
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_JSON_FD = None


def capture_stdout():
    """Keep the real stdout for the ONE JSON line and send everything else written to fd 1 (e.g.
    NCCL's "NCCL version ..." banner, which goes to stdout) to stderr."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


# --- "library on the same box" bar -----------------------------------------
def run_library_bar(args):
    """The reference's own GPU path is PyTorch/cuDNN: time the oracle's torch restatement of the
    network (forward + sigmoid, B=16 patches of 96^3 as inference.py:33) on cuda:0 in the three
    settings BASELINE.md names.  Not part of the bench contract: prints one JSON line with
    "impl": "library" for DESIGN.md; the product never takes this path."""
    import contextlib

    import torch

    from oracle.unet_ref import rescaled_state_dict, unet_forward

    dev = torch.device("cuda", 0)
    sd = {k: v.to(dev) for k, v in rescaled_state_dict(0).items()}
    torch.backends.cudnn.benchmark = True
    x = torch.rand(16, 1, 96, 96, 96, device=dev)
    out = {}
    for name in ("fp32_ieee", "tf32", "bf16_autocast", "bf16_autocast_channels_last"):
        tf32 = name != "fp32_ieee"
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        ctx = torch.autocast("cuda", dtype=torch.bfloat16) if name.startswith("bf16") \
            else contextlib.nullcontext()
        xin, w = x, sd
        if name.endswith("channels_last"):
            xin = x.contiguous(memory_format=torch.channels_last_3d)
            w = {k: (v.contiguous(memory_format=torch.channels_last_3d) if v.dim() == 5 else v)
                 for k, v in sd.items()}
        try:
            with ctx:
                for _ in range(2):
                    torch.sigmoid(unet_forward(xin, w))
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.steps):
                    torch.sigmoid(unet_forward(xin, w))
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            # one patch adds 64^3 stitched voxels at stride 64
            out[name] = {"ms_per_batch16": ms, "patches_per_s": 16e3 / ms,
                         "voxels_per_s": 16e3 / ms * 64 ** 3,
                         "conv_tflops": 16 * F_PATCH_96 / (ms * 1e-3) / 1e12}
        except Exception as exc:  # noqa: BLE001 -- a setting cuDNN cannot run is reported, not fatal
            out[name] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
        torch.cuda.empty_cache()
    emit({"impl": "library", "what": "PyTorch/cuDNN forward+sigmoid of the same UNet3D, B=16 x 96^3, "
          "forward only (no normalisation, patch extraction, D2H or stitching)",
          "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(), "settings": out})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "library"])
    ap.add_argument("--batch", type=int, default=32, help="patches per wave")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--volume", type=int, default=0, choices=[0, 256, 512, 1024],
                    help="fix the volume at N^3 for every --gpus (strong scaling); 0 = N x 512^3 (weak)")
    ap.add_argument("--train-batch", type=int, default=16,
                    help="--workload train: patches per step (Trainer default 16, train.py:39)")
    ap.add_argument("--workload", default="predict", choices=["predict", "segment", "train"],
                    help="predict = the headline hot path; segment = BASELINE config 5 (predict, then "
                         "affinities_to_segmentation timed; one GPU)")
    ap.add_argument("--patch", type=int, default=96, choices=[96, 128],
                    help="patch edge: 96 = the headline workload; 128 = BASELINE config 4 (not a bench line)")
    args = ap.parse_args()
    if args.workload == "train" and args.gpus != 1 and args.impl == "b200":
        # the reference trains on one GPU (train.py:77); data-parallel replicas would need a
        # gradient all-reduce that is not built, and N unsynchronised copies measure nothing
        raise SystemExit("bench.py --workload train runs on one GPU (the reference's Trainer is "
                         "single-GPU, train.py:77); use --gpus 1")
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: launched without torchrun -> start one process per GPU ourselves
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port",
               "29531", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    global PATCH, STRONG_VOLUME
    PATCH = (args.patch,) * 3
    STRONG_VOLUME = args.volume
    capture_stdout()
    if args.workload == "segment":
        run_segment(args)
    elif args.workload == "train":
        {"library": run_train_library, "reference": run_train_reference}.get(args.impl, run_train)(args)
    elif args.impl == "reference":
        run_reference(args)
    elif args.impl == "library":
        run_library_bar(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
