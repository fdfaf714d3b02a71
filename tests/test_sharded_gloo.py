"""world_size-2/3 gloo tests of the z-row slab sharding host logic (no GPU).

The slab compute is injected through SlabJob's ``backend`` seam: a CPU stand-in built from
the oracle (tests only) with the same four operations the native engine exposes.  What is
under test is the product's host logic: row partition, histogram partition + all-reduce,
partial-sum hand-over order and the gather -- the result must be bit-identical to the
single-rank oracle.
"""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import make_volume, state_dict_for


class OracleSlabBackend:
    device = torch.device("cpu")

    def __init__(self, sd, out_channels=3):
        from oracle.unet_ref import make_forward_fn

        self.fwd = make_forward_fn(sd)
        self.out_channels = out_channels

    def to_device(self, host_u16):
        return torch.from_numpy(np.ascontiguousarray(host_u16).astype(np.int32))

    def sync(self):
        pass

    def histogram(self, slab, clip):
        v = np.minimum(slab.numpy().ravel(), clip)
        return torch.from_numpy(np.bincount(v, minlength=clip + 1).astype(np.int64))

    def run(self, slab, shape, params, rows, mn, mx):
        from oracle import predict_ref as pr
        from aind_exaspim_neuron_segmentation_b200.engine import plan_slab

        self.shape, self.rows = shape, rows
        patch, ov, self.trim = tuple(params.patch), tuple(params.overlap), params.trim
        self.patch, self.ov = patch, ov
        plan = plan_slab(shape, params, *rows)
        z0 = plan["in_z0"]
        clipped = np.minimum(slab.numpy(), params.brightness_clip)
        norm = np.clip((clipped - mn) / (mx - mn + 1e-8), 0, 1)
        self.patches = []
        starts = [s for s in pr.patch_starts(shape, patch, ov)
                  if rows[0] <= s[0] // (patch[0] - ov[0]) < rows[1]]
        for s in starts:
            # extract from the slab with slab-local z, but clip against the GLOBAL volume end
            loc = (s[0] - z0, s[1], s[2])
            box = pr.extract_patch(norm[:min(s[0] + patch[0], shape[0]) - z0], loc, patch)
            y = 1.0 / (1.0 + np.exp(-self.fwd(box[None, None])[0].astype(np.float32)))
            t = self.trim
            self.patches.append((s, y.astype(np.float32)[:, t:patch[0] - t, t:patch[1] - t, t:patch[2] - t]))

    def _accumulate(self, z_lo, z_hi, seed):
        c, (d, h, w) = self.out_channels, self.shape
        acc = np.zeros((c, z_hi - z_lo, h, w), np.float32)
        if seed is not None:
            acc[:, :seed.shape[1]] = seed.numpy()
        for s, y in self.patches:
            lo = [a + self.trim for a in s]
            hi = [min(a + n, dim) for a, n, dim in zip(lo, y.shape[1:], self.shape)]
            za, zb = max(lo[0], z_lo), min(hi[0], z_hi)
            if zb <= za:
                continue
            acc[:, za - z_lo:zb - z_lo, lo[1]:hi[1], lo[2]:hi[2]] += \
                y[:, za - lo[0]:zb - lo[0], :hi[1] - lo[1], :hi[2] - lo[2]]
        return acc

    def partial(self, halo):
        from aind_exaspim_neuron_segmentation_b200.engine import plan_slab

        halo.copy_(torch.from_numpy(self._accumulate(self._plan["halo_z0"], self._plan["halo_z1"], None)))

    def stitch(self, seed, out):
        from oracle import predict_ref as pr

        z0, z1 = self._plan["out_z0"], self._plan["out_z1"]
        acc = self._accumulate(z0, z1, seed)
        cz = pr.coverage_count_axis(self.shape[0], self.patch[0], self.ov[0], self.trim)[z0:z1]
        cy = pr.coverage_count_axis(self.shape[1], self.patch[1], self.ov[1], self.trim)
        cx = pr.coverage_count_axis(self.shape[2], self.patch[2], self.ov[2], self.trim)
        wgt = (cz[:, None, None] * cy[None, :, None] * cx[None, None, :]).astype(np.float32)
        np.divide(acc, wgt[None], out=acc, where=wgt[None] != 0)
        out.copy_(torch.from_numpy(acc))


    # pipelined form (exa_slab_predict / exa_slab_finish): same pieces, the seed planes last
    def predict_rows(self, slab, shape, params, rows, mn, mx, own, out_host, halo):
        self.run(slab, shape, params, rows, mn, mx)
        if halo is not None:
            self.partial(halo)

    def finish_rows(self, seed, own, out_host):
        self.stitch(seed, own)
        if out_host is not None:
            out_host.copy_(own)


def _worker(rank, world, port, shape, kw, result_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    try:
        from aind_exaspim_neuron_segmentation_b200 import _native
        from aind_exaspim_neuron_segmentation_b200.inference import SlabJob

        sd = state_dict_for("rescaled", 21)
        vol = make_volume(shape, 22)
        backend = OracleSlabBackend(sd)
        params = _native.make_params(kw["patch_shape"], kw["overlap"], kw["trim"], 1000, (1, 99.9))
        job = SlabJob(shape, params, 3, backend)
        backend._plan = job.plan
        full = job.run(job.upload(vol), gather=True).numpy()
        own = job.run(job.upload(vol), gather=False).numpy()
        z0, z1 = job.own_bounds()
        assert np.array_equal(full[:, z0:z1], own)
        host = torch.empty(own.shape, dtype=torch.float32)
        piped = job.run_pipelined(job.upload(vol), host).numpy()
        assert np.array_equal(piped, own) and np.array_equal(host.numpy(), own)
        if rank == 0:
            np.save(result_path, full)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,shape", [(2, (100, 40, 36)), (3, (90, 36, 40))])
def test_sharded_equals_single_rank_oracle(tmp_path, world, shape):
    from oracle.predict_ref import predict_ref
    from oracle.unet_ref import make_forward_fn

    kw = dict(patch_shape=(32, 32, 32), overlap=(8, 8, 8), trim=4)
    path = str(tmp_path / "full.npy")
    mp.spawn(_worker, args=(world, _free_port(), shape, kw, path), nprocs=world, join=True)
    got = np.load(path)
    ref = predict_ref(make_volume(shape, 22), make_forward_fn(state_dict_for("rescaled", 21)), **kw)
    assert got.shape == ref.shape
    # identical summation order; conv thread counts differ between processes -> tiny fp noise
    assert np.abs(got - ref).max() <= 2e-6
    assert np.array_equal(got == 0, ref == 0)


def test_split_rows_and_hist_ranges_cover_volume():
    from aind_exaspim_neuron_segmentation_b200 import _native
    from aind_exaspim_neuron_segmentation_b200.inference import split_rows

    assert split_rows(16, 8) == [(2 * i, 2 * i + 2) for i in range(8)]
    assert split_rows(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]
    assert split_rows(8, 3) == [(0, 3), (3, 6), (6, 8)]
