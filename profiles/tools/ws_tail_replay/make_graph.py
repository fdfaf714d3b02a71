"""Region graph of the smooth random field profiles/tools/ws_bench.py uses, made with the CPU
oracle (no GPU): python make_graph.py EDGE OUT.npz.  Measurement infrastructure only."""
import os
import sys

import numpy as np
from scipy.ndimage import gaussian_filter

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT)
from oracle.watershed_ref import region_graph, watershed_fragments  # noqa: E402

edge, out = int(sys.argv[1]), sys.argv[2]
rng = np.random.default_rng(0)
f = np.stack([gaussian_filter(rng.normal(size=(edge,) * 3).astype(np.float32), 3.0) for _ in range(3)])
aff = (1.0 / (1.0 + np.exp(-6.0 * f / f.std()))).astype(np.float32)
frag, n = watershed_fragments(aff)
g = region_graph(aff, frag)
print("fragments", n, "region edges", g["u"].size)
np.savez(out, n=n, u=g["u"], v=g["v"], q=g["qsum"], c=g["count"], k=np.arange(g["u"].size, dtype=np.uint32))
