"""Timing of the GPU affinities_to_segmentation next to the CPU restatement.

python profiles/tools/ws_bench.py [edge]   (default 256) -> one JSON line.  Affinities: sigmoid of a
Gaussian-filtered random field (sigma 3), device-resident for the GPU timing; the CPU oracle
(oracle/watershed_ref.py, numpy/scipy + a Python merge queue) is timed on a 96^3 corner of the same
field -- measurement infrastructure only, never imported by the product."""

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
    from aind_exaspim_neuron_segmentation_b200 import affinities_to_segmentation
    from oracle.watershed_ref import affinities_to_segmentation_ref

    rng = np.random.default_rng(0)
    shape = (edge,) * 3
    f = np.stack([gaussian_filter(rng.normal(size=shape).astype(np.float32), 3.0) for _ in range(3)])
    aff = (1.0 / (1.0 + np.exp(-6.0 * f / f.std()))).astype(np.float32)
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
    sub = np.ascontiguousarray(aff[:, :96, :96, :96])
    t0 = time.perf_counter()
    ref = affinities_to_segmentation_ref(sub)
    cpu_s = time.perf_counter() - t0
    got = affinities_to_segmentation(sub)
    print(json.dumps({
        "what": "affinities_to_segmentation, thresholds [0.6,0.8,0.9], min size 100, device-resident input",
        "volume": list(shape), "seconds": min(times), "voxels_per_s": edge ** 3 / min(times),
        "n_fragments": int(frags.max()), "n_segments": int(seg.max()),
        "cpu_oracle": {"volume": [96, 96, 96], "seconds": cpu_s, "voxels_per_s": 96 ** 3 / cpu_s,
                       "equal_to_gpu": bool(np.array_equal(got.astype(np.int64), ref))}}))


if __name__ == "__main__":
    main()
