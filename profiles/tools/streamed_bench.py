"""predict_streamed (memmap in -> memmap out) next to predict() on the same volume.

    python profiles/tools/streamed_bench.py [D H W]   (default 1024 512 512) -> one JSON line

Both calls start from a uint16 .npy memory map on local disk (page cache warm after the first pass)
and end with the float32 (3, D, H, W) result in a .npy memory map; predict() needs the whole volume
and result in host RAM, predict_streamed only rows_per_chunk z patch-rows.  Bit-identity is checked."""

import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import bench
    from aind_exaspim_neuron_segmentation_b200 import UNet3D, predict, predict_streamed
    from oracle.unet_ref import rescaled_state_dict   # seeded weight recipe only

    shape = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (1024, 512, 512)
    model = UNet3D(output_channels=3)
    model.load_state_dict(rescaled_state_dict(0), strict=True)
    model = model.cuda().eval()
    tmp = tempfile.mkdtemp(prefix="exa_streamed_")
    src = np.lib.format.open_memmap(os.path.join(tmp, "vol.npy"), mode="w+", dtype=np.uint16, shape=shape)
    src[:] = bench.synth_planes(shape, 0, shape[0])
    src.flush()
    src = np.load(os.path.join(tmp, "vol.npy"), mmap_mode="r")
    out_a = np.lib.format.open_memmap(os.path.join(tmp, "a.npy"), mode="w+", dtype=np.float32, shape=(3,) + shape)
    out_b = np.lib.format.open_memmap(os.path.join(tmp, "b.npy"), mode="w+", dtype=np.float32, shape=(3,) + shape)
    res = {}
    for name, fn in (("predict", lambda: out_a.__setitem__(slice(None), predict(np.asarray(src), model, verbose=False))),
                     ("predict_streamed", lambda: predict_streamed(src, model, out_b, rows_per_chunk=4))):
        fn()   # warm-up: page cache, workspaces
        torch.cuda.synchronize()
        times = []
        for _ in range(3):
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
        res[name] = {"seconds": min(times), "all_seconds": times,
                     "voxels_per_s": float(np.prod(shape)) / min(times)}
    same = all(np.array_equal(out_a[c], out_b[c]) for c in range(3))
    print(json.dumps({"what": "memmap uint16 volume -> memmap float32 affinities, wall clock",
                      "volume": list(shape), "bit_identical": bool(same),
                      "streamed_over_predict": res["predict_streamed"]["seconds"] / res["predict"]["seconds"],
                      **res}))
    for f in os.listdir(tmp):
        os.remove(os.path.join(tmp, f))
    os.rmdir(tmp)


if __name__ == "__main__":
    main()
