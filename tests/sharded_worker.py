"""torchrun worker of tests/test_gpu_sharded.py: one process per GPU, NCCL.

Every rank computes the single-GPU result itself (small volume) and checks that the sharded
forms are bit-identical to it: the fused gather over symmetric memory in both forms (copy-engine
transfers of finished bands, the default; stores from the stitch kernel, EXA_GATHER=store), the
grouped NCCL send/recv gather, and the row-pipelined
``run_pipelined`` / ``predict_sharded(gather=False)`` with host buffers.
"""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from helpers import lightsheet_volume, state_dict_for  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()

    from aind_exaspim_neuron_segmentation_b200 import UNet3D, _native, predict, predict_sharded
    from aind_exaspim_neuron_segmentation_b200.inference import SlabJob, _EngineSlabBackend

    model = UNet3D(output_channels=3)
    model.load_state_dict(state_dict_for("rescaled", 41), strict=True)
    model = model.to(dev).eval()
    shape = (200, 72, 88)
    vol = lightsheet_volume(shape, 42)
    kw = dict(patch_shape=(32, 32, 32), overlap=(8, 8, 8), trim=4)
    single = predict(vol, model, verbose=False, **kw)

    params = _native.make_params(kw["patch_shape"], kw["overlap"], kw["trim"], 1000, (1, 99.9), batch=7)
    backend = _EngineSlabBackend(model.engine("bf16"))
    modes = []
    for gather_env in ("", "store", "nccl"):
        os.environ["EXA_GATHER"] = gather_env
        job = SlabJob(shape, params, 3, backend)
        for step in range(2):   # twice: the second run overwrites the peers' previous result
            full = job.run(job.upload(vol), gather=True)
            assert np.array_equal(full.cpu().numpy(), single), (gather_env, step, rank)
        modes.append(("peer-store" if gather_env == "store" else "copy-engine") if job._fused else "nccl")
        z0, z1 = job.own_bounds()
        host = torch.empty((3, z1 - z0) + shape[1:], dtype=torch.float32).pin_memory()
        own = job.run_pipelined(job.upload(vol), host)
        assert np.array_equal(host.numpy(), single[:, z0:z1]) and np.array_equal(own.cpu().numpy(), host.numpy())
    os.environ["EXA_GATHER"] = ""
    a, b, planes = predict_sharded(vol, model, gather=False, **kw)
    assert np.array_equal(planes, single[:, a:b])
    assert np.array_equal(predict_sharded(vol, model, **kw), single)
    dist.barrier()
    if rank == 0:
        print(f"SHARDED_OK world={world} gather_modes={modes}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
