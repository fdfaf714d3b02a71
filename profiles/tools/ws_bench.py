"""Timing of the GPU affinities_to_segmentation next to the CPU restatement.

python profiles/tools/ws_bench.py [edge] [check]   (default 256 oracle) -> one JSON line.
Affinities: sigmoid of a Gaussian-filtered random field (sigma 3), device-resident for the GPU
timing.  check = oracle: the whole volume through oracle/watershed_ref.py (compiled restatement of
waterz, sequential queue), labels compared element for element; check = host: the product with
EXA_WS_GPU_ROUNDS=0 (its exact host queue only, itself tested against the oracle) as the
comparison, for sizes where the oracle takes too long; check = none.  Measurement infrastructure
only, never imported by the product.  EXA_WS_PROF=1 prints the phases on stderr."""

import json
import os
import sys
import time

import numpy as np
import torch
from scipy.ndimage import gaussian_filter

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    edge = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    check = sys.argv[2] if len(sys.argv) > 2 else "oracle"
    from aind_exaspim_neuron_segmentation_b200 import affinities_to_segmentation

    rng = np.random.default_rng(0)
    shape = (edge,) * 3
    f = np.stack([gaussian_filter(rng.normal(size=shape).astype(np.float32), 3.0) for _ in range(3)])
    aff = (1.0 / (1.0 + np.exp(-6.0 * f / f.std()))).astype(np.float32)
    del f
    dev = torch.from_numpy(aff).cuda()
    affinities_to_segmentation(dev[:, :64, :64, :64].contiguous())   # warm-up (CUB temp, context)
    torch.cuda.synchronize()
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        seg = affinities_to_segmentation(dev)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    frags = affinities_to_segmentation(dev, [0.0], 0)
    out = {
        "what": "affinities_to_segmentation, thresholds [0.6,0.8,0.9], min size 100, device-resident input",
        "volume": list(shape), "seconds": min(times), "all_seconds": times,
        "voxels_per_s": edge ** 3 / min(times),
        "n_fragments": int(frags.max()), "n_segments": int(seg.max())}
    del frags
    got = seg.cpu().numpy()
    if check == "oracle":
        from oracle.watershed_ref import affinities_to_segmentation_ref

        t0 = time.perf_counter()
        ref = affinities_to_segmentation_ref(aff)
        cpu_s = time.perf_counter() - t0
        out["cpu_oracle"] = {"volume": list(shape), "seconds": cpu_s, "voxels_per_s": edge ** 3 / cpu_s,
                             "kind": "port", "cores": 1, "equal_to_gpu": bool(np.array_equal(got, ref))}
    elif check == "host":
        os.environ["EXA_WS_GPU_ROUNDS"] = "0"
        t0 = time.perf_counter()
        ref = affinities_to_segmentation(dev).cpu().numpy()
        host_s = time.perf_counter() - t0
        out["host_queue_only"] = {"seconds": host_s, "equal_to_default": bool(np.array_equal(got, ref))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
