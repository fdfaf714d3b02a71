"""Drop-in for the prediction half of reference ``inference.py``.

Same names, signatures, defaults and results as

* ``predict``                reference inference.py:29-126
* ``load_model``             reference inference.py:400-424
* ``count_patches``          reference inference.py:340-365
* ``generate_patch_starts``  reference inference.py:368-397

but the work -- clip, global percentile normalisation, halo patch extraction with
reflect padding, the U-Net forward, sigmoid, trim, overlap stitching and the
division by the coverage count -- runs in hand-written sm_100a CUDA behind the C
ABI of ``include/exaspim_b200.h``.  The host code here only validates arguments
and moves buffers.  There is no CPU fallback.

``predict_sharded`` is the multi-GPU form (one process per GPU, z-row slabs,
SURVEY.md 8e); with a world size of 1 it is the same computation as ``predict``.
"""

import itertools
import os
import warnings

import numpy as np
import torch

from . import _native
from .engine import Engine, percentiles_from_hist, percentiles_from_hist_values, plan_slab
from .machine_learning.unet3d import UNet3D, engine_for_module

__all__ = ["predict", "predict_sharded", "predict_streamed", "load_model", "count_patches",
           "generate_patch_starts", "affinities_to_segmentation"]


# --- tiling helpers (host integer logic) --------------------------------------------
def _axis_starts(dim, patch, overlap):
    stride = patch - overlap
    return range(0, dim - patch + stride, stride)


def count_patches(img_shape, patch_shape, overlap):
    """Number of sliding-window patches for a (1, 1, D, H, W) image shape."""
    assert len(img_shape) == 5, "Image must have shape (1, 1, D, H, W)"
    n = 1
    for dim, p, o in zip(img_shape[2:], patch_shape, overlap):
        n *= len(_axis_starts(dim, p, o))
    return n


def generate_patch_starts(img_shape, patch_shape, overlap):
    """Yield (z, y, x) window starts, z-major, for a (1, 1, D, H, W) image shape."""
    assert len(img_shape) == 5, "Image must have shape (1, 1, D, H, W)"
    per_axis = [_axis_starts(dim, p, o) for dim, p, o in zip(img_shape[2:], patch_shape, overlap)]
    yield from itertools.product(*per_axis)


# --- model loading --------------------------------------------------------------------
def load_model(path, affinity_mode=True, device="cuda", precision="bf16"):
    """Load a reference checkpoint (``torch.save(model.state_dict())``) strictly."""
    model = UNet3D(output_channels=3 if affinity_mode else 1, precision=precision)
    model.load_state_dict(torch.load(path, map_location=device))
    model.to(device)
    model.eval()
    return model


def _engine_for(model, precision=None):
    """Native engine for ``model``: ours directly, or any module with the same state_dict."""
    if isinstance(model, UNet3D):
        return model.engine(precision)
    if isinstance(model, Engine):
        return model
    if isinstance(model, torch.nn.Module):
        # e.g. the reference's own UNet3D instance: identical state_dict layout
        return engine_for_module(model, precision or "bf16")
    raise TypeError("model must be a torch.nn.Module with the reference UNet3D state_dict")


# --- input handling ---------------------------------------------------------------------
def _as_volume_u16(img, brightness_clip):
    """(…,D,H,W) array-like -> C-contiguous uint16 (D,H,W) holding min(img, clip) exactly.

    The reference computes ``np.minimum(img, brightness_clip)`` in the image's own
    dtype (inference.py:79).  For integer images that is representable in uint16
    whenever the clipped values are, which is what the kernels consume.  Inputs whose clipped
    values would NOT survive the conversion unchanged raise instead of being altered silently
    (values above 65535 that the clip does not remove, negative values, floating-point images).
    """
    arr = np.asarray(img)
    if arr.ndim > 5 or arr.ndim < 3:
        raise ValueError("image must have between 3 and 5 dimensions")
    if any(s != 1 for s in arr.shape[:-3]):
        raise ValueError("leading (batch/channel) dimensions must be 1")
    arr = arr.reshape(arr.shape[-3:])
    if arr.dtype == np.uint16:
        return np.ascontiguousarray(arr)
    if arr.dtype.kind == "b":
        return arr.astype(np.uint16)
    if arr.dtype.kind in "ui":
        if arr.dtype.kind == "i" and arr.size and arr.min() < 0:
            raise TypeError("negative intensities are not supported by the uint16 kernels")
        clip = _check_clip(arr, brightness_clip)
        if clip > 65535 and arr.size and arr.max() > 65535:
            raise TypeError(
                f"{arr.dtype} image with values above 65535 and brightness_clip={brightness_clip}: "
                "the clipped volume does not fit the uint16 kernels (lower the clip to <= 65535)")
        return np.minimum(arr, min(clip, 65535)).astype(np.uint16)
    raise TypeError(
        f"unsupported image dtype {arr.dtype}: predict_sharded / predict_streamed consume integer "
        "(ExaSPIM uint16) volumes; floating-point images go through predict()"
    )


def _check_clip(arr, brightness_clip):
    """brightness_clip as the integer the kernels use; anything np.minimum would treat
    differently (negative, non-integral: the clipped image would become float) raises."""
    if brightness_clip < 0:
        raise ValueError("brightness_clip must be >= 0")
    if float(brightness_clip) != int(brightness_clip):
        raise ValueError(f"brightness_clip={brightness_clip} is not an integer: np.minimum would turn "
                         "the image into floats, which the uint16 kernels do not reproduce")
    return int(brightness_clip)


# --- the hot path ---------------------------------------------------------------------
def predict(
    img,
    model,
    affinity_mode=True,
    batch_size=16,
    brightness_clip=1000,
    normalization_percentiles=(1, 99.9),
    patch_shape=(96, 96, 96),
    overlap=(32, 32, 32),
    trim=8,
    verbose=True,
    precision=None,
    out=None,
):
    """Affinity (or foreground) prediction for a 3-D volume; see reference inference.py:29-126.

    ``precision`` ("bf16" | "fp32") and ``out`` (a preallocated C-contiguous float32
    ``(C, D, H, W)`` array, e.g. pinned memory, filled and returned instead of a new array)
    are extensions; defaults keep the reference behaviour.

    Returns a new C-contiguous float32 array ``(3, D, H, W)`` (``(D, H, W)`` when
    ``affinity_mode=False``).  ``batch_size`` is accepted for compatibility and used as
    a lower bound on the number of patches per wave -- it has no numerical effect
    (eval-mode BatchNorm, no cross-sample op).
    """
    engine = _engine_for(model, precision)
    n_channels = 3 if affinity_mode else 1
    if engine.out_channels != n_channels:
        raise ValueError(
            f"model has {engine.out_channels} output channels but affinity_mode={affinity_mode} "
            f"needs {n_channels}"
        )
    if np.asarray(img).dtype.kind == "f":
        res = _predict_float(np.asarray(img), engine, n_channels, batch_size, brightness_clip,
                             normalization_percentiles, patch_shape, overlap, trim, out)
        return res if affinity_mode else res[0]
    _check_clip(img, brightness_clip)
    vol = _as_volume_u16(img, brightness_clip)
    shape5 = (1, 1) + vol.shape
    n_patches = count_patches(shape5, patch_shape, overlap)
    params = _native.make_params(patch_shape, overlap, trim, brightness_clip,
                                 normalization_percentiles, batch=max(int(batch_size), 32))
    if not verbose:
        out = engine.predict_host(vol, params, out=out)
        return out if affinity_mode else out[0]
    from tqdm import tqdm

    # the bar advances as waves of patches FINISH on the device (exa_set_progress_callback), like
    # the reference's per-batch update (inference.py:118-120)
    with tqdm(total=n_patches, desc="Predict") as pbar:
        state = {"done": 0}

        def advance(_user, done, _total):
            pbar.update(int(done) - state["done"])
            state["done"] = int(done)

        engine.set_progress(advance)
        try:
            out = engine.predict_host(vol, params, out=out)
        finally:
            engine.set_progress(None)
        pbar.update(n_patches - state["done"])
    return out if affinity_mode else out[0]


def _predict_float(arr, engine, n_channels, batch_size, brightness_clip, normalization_percentiles,
                   patch_shape, overlap, trim, out):
    """Floating-point images (the reference takes any dtype: inference.py:79-80).  The volume is
    rank-compressed on the GPU -- uint16 index of min(x, clip) among its distinct values, at most
    65536 of them, otherwise this raises -- and the normalisation table is evaluated on the exact
    values in float64, so the uint16 kernels reproduce the reference bit for bit
    (csrc/float_volume.cu).  float16 is widened to float32 like numpy's percentile/normalise do."""
    if arr.ndim > 5 or arr.ndim < 3 or any(s != 1 for s in arr.shape[:-3]):
        raise ValueError("image must have between 3 and 5 dimensions with leading dimensions of 1")
    if brightness_clip < 0:
        raise ValueError("brightness_clip must be >= 0")
    arr = arr.reshape(arr.shape[-3:])
    if arr.dtype not in (np.float32, np.float64):
        raise TypeError(f"unsupported floating-point image dtype {arr.dtype}: float32 and float64 "
                        "images are handled")
    shape = tuple(int(v) for v in arr.shape)
    c = n_channels
    dev = engine.device
    vol_dev = torch.from_numpy(np.ascontiguousarray(arr)).to(dev)
    idx, table = engine.compress_float_volume(vol_dev, brightness_clip)
    del vol_dev
    clip = int(table.size) - 1
    params = _native.make_params(patch_shape, overlap, trim, clip, normalization_percentiles,
                                 batch=max(int(batch_size), 32))
    hist = engine.histogram(idx, clip)
    mn, mx = percentiles_from_hist_values(hist.cpu().numpy().astype(np.uint64), table,
                                          arr.dtype == np.float32, params.pct_lo, params.pct_hi)
    engine.set_normalization_table(table, mn, mx)
    res = out if out is not None else np.empty((c,) + shape, dtype=np.float32)
    if tuple(res.shape) != (c,) + shape or res.dtype != np.float32 or not res.flags["C_CONTIGUOUS"]:
        raise ValueError(f"out must be a C-contiguous float32 array of shape {(c,) + shape}")
    nz = plan_slab(shape, params, 0, 0)["nz"]
    if nz == 0 or count_patches((1, 1) + shape, patch_shape, overlap) == 0:
        res[...] = 0.0
        return res
    own = torch.empty((c,) + shape, dtype=torch.float32, device=dev)
    engine.slab_predict(idx, shape, params, 0, nz, own, None, None)
    engine.slab_finish(None, own, None)
    res[...] = own.cpu().numpy()
    return res


# --- volumes larger than memory: chunks of z patch-rows, one after the other ------------------
def predict_streamed(
    img,
    model,
    out,
    affinity_mode=True,
    batch_size=16,
    brightness_clip=1000,
    normalization_percentiles=(1, 99.9),
    patch_shape=(96, 96, 96),
    overlap=(32, 32, 32),
    trim=8,
    rows_per_chunk=4,
    precision=None,
):
    """``predict`` for volumes that fit neither in host nor in device memory.

    ``img`` is anything with a ``(D, H, W)`` ``.shape`` that returns planes for ``img[z0:z1]``
    (``np.memmap``, a zarr / N5 / HDF5 array, ...); ``out`` anything that accepts
    ``out[:, z0:z1] = planes`` (``out[z0:z1]`` when ``affinity_mode=False``).  The reference holds
    the whole volume plus 16 B/voxel of accumulators in RAM (inference.py:80,91-92); here only
    ``rows_per_chunk`` z patch-rows are resident at a time.  Two passes over the input: the global
    percentiles need every voxel (exact 1001-bin histogram, accumulated per chunk), then the
    chunks run one after the other exactly like the ranks of ``predict_sharded`` -- each hands the
    partial sums of the planes it shares with the next one forward -- so the result is bit-identical
    to ``predict``.  Reading chunk i+1 from the source, computing chunk i (upload, kernels,
    row-pipelined download) and writing chunk i-1 to the sink overlap (two pinned staging buffers
    on either side, one reader and one writer thread).  Returns ``out``.
    """
    _check_clip(img, brightness_clip)
    shape = tuple(int(v) for v in img.shape)
    if len(shape) != 3:
        raise ValueError("predict_streamed expects a (D, H, W) array-like")
    engine = _engine_for(model, precision)
    c = 3 if affinity_mode else 1
    if engine.out_channels != c:
        raise ValueError(f"model has {engine.out_channels} output channels but affinity_mode={affinity_mode} "
                         f"needs {c}")
    params = _native.make_params(patch_shape, overlap, trim, brightness_clip,
                                 normalization_percentiles, batch=max(int(batch_size), 32))
    clip = params.brightness_clip
    dev = engine.device
    d, h, w = shape
    nz = plan_slab(shape, params, 0, 0)["nz"]
    rows_per_chunk = max(1, int(rows_per_chunk))
    if patch_shape[0] - 2 * trim > 2 * (patch_shape[0] - overlap[0]):
        rows_per_chunk = nz   # planes covered by three rows cannot be handed over pairwise
    stride_z = patch_shape[0] - overlap[0]

    # Three stages run concurrently (they were sequential in the first version): a reader thread
    # slices the source into one of two pinned staging buffers, this thread uploads and computes,
    # and a writer thread stores the previous chunk's planes into the sink.  numpy copies, file
    # I/O and the engine calls all release the GIL.
    from concurrent.futures import ThreadPoolExecutor

    step = max(rows_per_chunk * stride_z, 1)
    plans = [plan_slab(shape, params, r0, min(r0 + rows_per_chunk, nz))
             for r0 in range(0, nz, rows_per_chunk)]
    max_in = max([step] + [pl["in_z1"] - pl["in_z0"] for pl in plans])
    in_host = [torch.empty((max_in, h, w), dtype=torch.uint16).pin_memory() for _ in range(2)]

    def load(z0, z1, k):
        view = in_host[k][:z1 - z0]
        np.copyto(view.numpy(), _as_volume_u16(img[z0:z1], clip))
        return view

    def store(z0, z1, host):
        if affinity_mode:
            out[:, z0:z1] = host
        else:
            out[z0:z1] = host[0]

    with ThreadPoolExecutor(1) as reader, ThreadPoolExecutor(1) as writer:
        # pass 1: exact global percentiles from per-chunk histograms
        hist = torch.zeros(clip + 1, dtype=torch.int64, device=dev)
        spans = [(z0, min(z0 + step, d)) for z0 in range(0, d, step)]
        fut = reader.submit(load, spans[0][0], spans[0][1], 0)
        for i in range(len(spans)):
            host_in = fut.result()
            if i + 1 < len(spans):
                fut = reader.submit(load, spans[i + 1][0], spans[i + 1][1], (i + 1) % 2)
            hist += engine.histogram(host_in.to(dev, non_blocking=True), clip)
            torch.cuda.current_stream(dev).synchronize()   # the staging buffer is free again
        mn, mx = percentiles_from_hist(hist.cpu().numpy().astype(np.uint64), params.pct_lo,
                                       params.pct_hi)
        if nz == 0 or count_patches((1, 1) + shape, patch_shape, overlap) == 0:
            zero = np.zeros((c, min(step, d), h, w), dtype=np.float32)
            for z0, z1 in spans:
                store(z0, z1, zero[:, :z1 - z0])
            return out
        # pass 2: row chunks; the halo of one chunk seeds the next
        engine.set_normalization(mn, mx, clip)
        max_own = max(pl["out_z1"] - pl["out_z0"] for pl in plans)
        own_dev = torch.empty((c, max_own, h, w), dtype=torch.float32, device=dev)
        own_host = [torch.empty((c, max_own, h, w), dtype=torch.float32).pin_memory() for _ in range(2)]
        writes = [None, None]
        seed = None
        fut = reader.submit(load, plans[0]["in_z0"], plans[0]["in_z1"], 0)
        for i, pl in enumerate(plans):
            r0 = i * rows_per_chunk
            r1 = min(r0 + rows_per_chunk, nz)
            n_own = pl["out_z1"] - pl["out_z0"]
            n_halo = pl["halo_z1"] - pl["halo_z0"]
            halo = (torch.empty((c, n_halo, h, w), dtype=torch.float32, device=dev)
                    if n_halo > 0 else None)
            host_in = fut.result()
            if i + 1 < len(plans):
                fut = reader.submit(load, plans[i + 1]["in_z0"], plans[i + 1]["in_z1"], (i + 1) % 2)
            slab = host_in.to(dev, non_blocking=True)
            k = i % 2
            if writes[k] is not None:
                writes[k].result()        # the writer is done with this host buffer (chunk i-2)
            # dense (C, n_own, H, W) views of the re-used buffers
            od = own_dev.view(-1)[:c * n_own * h * w].view(c, n_own, h, w)
            oh = own_host[k].view(-1)[:c * n_own * h * w].view(c, n_own, h, w)
            engine.slab_predict(slab, shape, params, r0, r1, od, oh, halo)
            engine.slab_finish(seed, od, oh)
            torch.cuda.current_stream(dev).synchronize()
            writes[k] = writer.submit(store, pl["out_z0"], pl["out_z1"], oh.numpy())
            seed = halo
        for wfut in writes:
            if wfut is not None:
                wfut.result()
    return out


# --- downstream of the hot path: affinities -> segmentation ---------------------------------
def affinities_to_segmentation(affinities, agglomeration_thresholds=[0.6, 0.8, 0.9],  # noqa: B006
                               min_segment_size=100):
    """Affinity maps ``(3, D, H, W)`` -> segmentation; reference inference.py:196-237.

    Same signature and defaults (``waterz.agglomerate`` with ``aff_threshold_low=0.1``,
    ``aff_threshold_high=0.9999``, the segmentation of the LAST threshold, then
    ``remove_small_segments``) and waterz's reading of the array: ``affinities[c][z,y,x]`` is the
    edge between voxel (z,y,x) and its previous neighbour along axis c.  Watershed fragments (with
    the breadth-first plateau division), the region graph, the agglomeration rounds and the
    relabelling run on the GPU (``csrc/watershed.cu``); only the tail of the merge queue runs on
    the host.  waterz itself is not available offline, so the behaviour is checked against a
    restatement of its published algorithm (``oracle/ws_ref.cpp``; parity unpinned): edge
    statistics are summed exactly in fixed point where waterz sums float32 values, and ties
    between equal scores are broken by the rank of the region-graph edge.
    numpy in -> new ``uint64 (D, H, W)`` array out, the dtype waterz returns; a CUDA tensor in -> an
    ``int64`` CUDA tensor out (no host round trip, e.g. straight from ``predict_sharded``).
    """
    import ctypes

    thr = [float(t) for t in agglomeration_thresholds]
    if not thr:
        raise ValueError("agglomeration_thresholds must not be empty")
    thr_c = (ctypes.c_double * len(thr))(*thr)
    n_frag, n_seg = ctypes.c_int64(0), ctypes.c_int64(0)
    lib = _native.lib()
    if isinstance(affinities, torch.Tensor) and affinities.is_cuda:
        aff = affinities.to(torch.float32).contiguous()
        if aff.dim() != 4 or aff.shape[0] != 3:
            raise ValueError("affinities must have shape (3, D, H, W)")
        d, h, w = (int(v) for v in aff.shape[1:])
        seg = torch.empty((d, h, w), dtype=torch.int64, device=aff.device)
        with torch.cuda.device(aff.device):
            code = lib.exa_affinities_to_segmentation_device(
                ctypes.c_void_p(aff.data_ptr()), d, h, w, thr_c, len(thr), 0.1, 0.9999,
                int(min_segment_size), ctypes.c_void_p(seg.data_ptr()), ctypes.byref(n_frag),
                ctypes.byref(n_seg), ctypes.c_void_p(torch.cuda.current_stream(aff.device).cuda_stream))
        _native.check(code, None, "exa_affinities_to_segmentation_device")
        return seg
    aff = np.ascontiguousarray(np.asarray(affinities).astype(np.float32, copy=False))
    if aff.ndim != 4 or aff.shape[0] != 3:
        raise ValueError("affinities must have shape (3, D, H, W)")
    d, h, w = aff.shape[1:]
    seg = np.empty((d, h, w), dtype=np.uint64)
    code = lib.exa_affinities_to_segmentation(
        torch.cuda.current_device(), aff.ctypes.data_as(ctypes.c_void_p), d, h, w, thr_c, len(thr),
        0.1, 0.9999, int(min_segment_size), seg.ctypes.data_as(ctypes.c_void_p),
        ctypes.byref(n_frag), ctypes.byref(n_seg))
    _native.check(code, None, "exa_affinities_to_segmentation")
    return seg


# --- multi-GPU: z-row slabs ---------------------------------------------------------------
def world_size_of(group=None):
    import torch.distributed as dist

    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def split_rows(n_rows, world_size):
    """Contiguous, balanced [begin, end) row ranges, one per rank (earlier ranks get the extras)."""
    base, extra = divmod(n_rows, world_size)
    out, r = [], 0
    for g in range(world_size):
        n = base + (1 if g < extra else 0)
        out.append((r, r + n))
        r += n
    return out


_SYMMETRIC_OUTPUTS = {}


class _EngineSlabBackend:
    """Slab compute on the native engine (device tensors)."""

    stream_ordered = True   # everything is queued on torch's current stream

    def __init__(self, engine):
        self.engine = engine
        self.device = engine.device
        self.out_channels = engine.out_channels

    def histogram(self, slab_u16, clip):
        return self.engine.histogram(slab_u16, clip)

    def run(self, slab_u16, shape, params, rows, mn, mx):
        self.engine.set_normalization(mn, mx, params.brightness_clip)
        self.engine.slab_run(slab_u16, shape, params, rows[0], rows[1])

    def partial(self, halo):
        self.engine.slab_partial(halo)

    def stitch(self, seed, out):
        self.engine.slab_stitch(seed, out)

    def stitch_into(self, seed, full, z0):
        """Stitch the owned planes straight into full[:, z0:...] of the (C, D, H, W) output."""
        self.engine.slab_stitch(seed, full[:, z0:], channel_stride=full.stride(0))

    def predict_rows(self, slab_u16, shape, params, rows, mn, mx, own, out_host, halo):
        self.engine.set_normalization(mn, mx, params.brightness_clip)
        self.engine.slab_predict(slab_u16, shape, params, rows[0], rows[1], own, out_host, halo)

    def finish_rows(self, seed, own, out_host):
        self.engine.slab_finish(seed, own, out_host)

    def symmetric_full(self, shape, group):
        """The gathered (C, D, H, W) output in torch symmetric memory: every rank's copy is mapped
        into every other rank's address space over NVLink.  -> (full, handle, peer pointers)."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        # ONE buffer per process, re-used by every job of the same (shape, device, group) --
        # symmetric allocations are expensive to set up -- and replaced (the old one freed) when
        # a job of another shape comes along, so a stream of ragged block shapes cannot grow GPU
        # memory without bound.  Consequence, documented on SlabJob.run: the gathered DEVICE array
        # a job returns is only valid until the next gather (peers store into it remotely).
        pg = group if group is not None else dist.group.WORLD
        key = (tuple(shape), str(self.device), pg.group_name)
        if key not in _SYMMETRIC_OUTPUTS:
            _SYMMETRIC_OUTPUTS.clear()   # every rank takes this branch for the same job: collective
            full = symm.empty(*shape, dtype=torch.float32, device=self.device)
            hdl = symm.rendezvous(full, pg)
            # peers in ring order starting after this rank: when every rank walks its list in step
            # (copy-engine gather), no two ranks send to the same destination at the same time
            n = len(hdl.buffer_ptrs)
            ptrs = [int(hdl.buffer_ptrs[(hdl.rank + k) % n]) for k in range(1, n)]
            _SYMMETRIC_OUTPUTS[key] = (full, hdl, ptrs)
        return _SYMMETRIC_OUTPUTS[key]

    def set_peers(self, full, ptrs):
        self.engine.set_peer_outputs(full, ptrs)

    def to_device(self, host_u16):
        return torch.from_numpy(host_u16).to(self.device, non_blocking=True)

    def sync(self):
        torch.cuda.current_stream(self.device).synchronize()


class SlabJob:
    """One rank's share of a volume sharded by z patch-rows.

    Exchange steps, all small next to the convolutions:
    C1 all-reduce of the (clip+1)-bin histogram -> global percentiles,
    C2 send of raw partial sums for the planes shared with the next rank's first row
       (summed there in the reference's patch order, so the result is bit-identical to
       the single-GPU one),
    C3 all-gather of the owned output planes.
    ``backend`` is a test seam for the slab compute; the default is the native engine.
    """

    def __init__(self, shape, params, n_channels, backend, group=None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.shape = tuple(int(v) for v in shape)
        self.params = params
        self.backend = backend
        self.n_channels = n_channels
        if backend.out_channels != n_channels:
            raise ValueError("model output channels do not match affinity_mode")
        nz = plan_slab(self.shape, params, 0, 0)["nz"]
        self.all_rows = split_rows(nz, self.world)
        self.all_plans = [plan_slab(self.shape, params, *r) for r in self.all_rows]
        self.rows = self.all_rows[self.rank]
        self.plan = self.all_plans[self.rank]
        self.has_rows = self.rows[1] > self.rows[0]
        # plane range this rank histograms: from its slab start to the next slab's start
        nxt = [p["in_z0"] for p, r in zip(self.all_plans, self.all_rows) if r[1] > r[0]]
        nxt.append(self.shape[0])
        order = sum(1 for r in self.all_rows[:self.rank] if r[1] > r[0])
        self.hist_range = (nxt[order], nxt[order + 1]) if self.has_rows else (0, 0)
        self.own_planes = [max(p["out_z1"] - p["out_z0"], 0) if r[1] > r[0] else 0
                           for p, r in zip(self.all_plans, self.all_rows)]
        self._full = None  # gathered output, allocated once and re-used by every run()
        self._own = None   # owned planes of run_pipelined(), likewise
        self._fused = None  # tri-state: not tried / fused peer-store gather / NCCL gather
        self._hdl = None
        self._peer_ptrs = []

    def slab_bounds(self):
        """Input planes [z0, z1) this rank needs resident."""
        return (self.plan["in_z0"], self.plan["in_z1"]) if self.has_rows else (0, 0)

    def upload(self, vol_host):
        z0, z1 = self.slab_bounds()
        return self.backend.to_device(vol_host[z0:z1])

    def _peer(self, group_rank):
        if self.group is None:
            return group_rank
        return self.dist.get_global_rank(self.group, group_rank)

    def _normalization(self, slab):
        """C1: global histogram -> exact percentiles (mn, mx)."""
        be, clip = self.backend, self.params.brightness_clip
        if self.has_rows and self.hist_range[1] > self.hist_range[0]:
            z0 = self.plan["in_z0"]
            hist = be.histogram(slab[self.hist_range[0] - z0:self.hist_range[1] - z0], clip)
        else:
            hist = torch.zeros(clip + 1, dtype=torch.int64, device=be.device)
        if self.world > 1:
            self.dist.all_reduce(hist, op=self.dist.ReduceOp.SUM, group=self.group)
        return percentiles_from_hist(hist.cpu().numpy().astype(np.uint64), self.params.pct_lo,
                                     self.params.pct_hi)

    def _halo_buffer(self):
        n_halo = self.plan["halo_z1"] - self.plan["halo_z0"] if self.has_rows else 0
        if self.world == 1 or n_halo <= 0:
            return None
        return torch.empty((self.n_channels, n_halo) + self.shape[1:], dtype=torch.float32,
                           device=self.backend.device)

    def _exchange_halo(self, halo):
        """C2: partial sums of the shared planes go to their owner (the next rank).  -> seed"""
        dist = self.dist
        if self.world == 1 or not self.has_rows:
            return None
        ops, seed = [], None
        n_seed = self.plan["seed_z1"] - self.plan["seed_z0"]
        if n_seed > 0:
            seed = torch.empty((self.n_channels, n_seed) + self.shape[1:], dtype=torch.float32,
                               device=self.backend.device)
            ops.append(dist.P2POp(dist.irecv, seed, self._peer(self.rank - 1), self.group))
        if halo is not None:
            ops.append(dist.P2POp(dist.isend, halo, self._peer(self.rank + 1), self.group))
        if ops:
            # the partial sums are produced on the current stream (the engine launches there); NCCL
            # orders its streams after it at enqueue and wait() orders the current stream after the
            # transfer -- no host synchronisation, the launch queue stays full
            if not getattr(self.backend, "stream_ordered", False):
                self.backend.sync()
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return seed

    def run_pipelined(self, slab, out_host=None):
        """``run(slab, gather=False)`` as a row-group pipeline: finished planes are stitched and
        copied to ``out_host`` (float32 host tensor ``(C, own planes, H, W)``, ideally pinned)
        while later rows still compute; only the planes shared with the previous rank wait for
        the C2 exchange.  -> device float32 (C, own, H, W); ``out_host`` is complete on return."""
        be = self.backend
        c, (d, h, w) = self.n_channels, self.shape
        nz_own = self.own_planes[self.rank]
        if out_host is not None and tuple(out_host.shape) != (c, nz_own, h, w):
            raise ValueError(f"out_host must have shape {(c, nz_own, h, w)}")
        mn, mx = self._normalization(slab)
        own = self._own
        if own is None or tuple(own.shape) != (c, nz_own, h, w) or own.device != be.device:
            own = self._own = torch.empty((c, nz_own, h, w), dtype=torch.float32, device=be.device)
        halo = self._halo_buffer()
        if self.has_rows:
            be.predict_rows(slab, self.shape, self.params, self.rows, mn, mx, own, out_host, halo)
        seed = self._exchange_halo(halo)
        if self.has_rows:
            be.finish_rows(seed, own, out_host)
        return own

    def _fused_gather_ready(self):
        """C3 without a gather pass: the output lives in symmetric memory and the stitch kernel
        sends every finished band of planes to all ranks' copies (exa_set_peer_outputs): by
        copy-engine transfers on side streams (default; no SM time, overlapped with the following
        waves) or, with ``EXA_GATHER=store``, by stores from the stitch kernel itself.  Falls back to the grouped NCCL
        send/recv gather when the backend has no peer mapping (CPU test backend),
        ``EXA_GATHER=nccl`` is set, or the symmetric-memory rendezvous is not possible here."""
        if self._fused is None:
            self._fused = False
            if hasattr(self.backend, "symmetric_full") and os.environ.get("EXA_GATHER", "") != "nccl":
                try:
                    self._full, self._hdl, self._peer_ptrs = self.backend.symmetric_full(
                        (self.n_channels,) + self.shape, self.group)
                    self._fused = True
                except Exception as exc:  # noqa: BLE001 -- e.g. no P2P between the GPUs of this box
                    warnings.warn(f"symmetric-memory gather unavailable ({exc}); using NCCL send/recv")
        return self._fused

    def _run_fused_gather(self, slab):
        be, full = self.backend, self._full
        mn, mx = self._normalization(slab)   # the all-reduce also orders this step's peer stores
        z0 = int(sum(self.own_planes[:self.rank]))  # after every rank's reads of the last result
        own = full[:, z0:z0 + self.own_planes[self.rank]]
        halo = self._halo_buffer()
        be.set_peers(full, self._peer_ptrs)
        try:
            if self.has_rows:
                be.predict_rows(slab, self.shape, self.params, self.rows, mn, mx, own, None, halo)
            seed = self._exchange_halo(halo)
            if self.has_rows:
                be.finish_rows(seed, own, None)
        finally:
            be.set_peers(full, [])
        self._hdl.barrier()   # every rank's stores have landed before anyone reads `full`
        return full

    def run(self, slab, gather=True):
        """slab: device uint16 planes [in_z0, in_z1).  -> device float32 (C, D|own, H, W).

        The returned tensor is a buffer the job (or, for the fused gather, the process-wide
        symmetric-memory allocation) re-uses: it is valid until the next ``run`` / gather --
        copy it (``predict_sharded`` returns a host copy) if it must outlive that."""
        dist, be, p = self.dist, self.backend, self.params
        dev = be.device
        c, (d, h, w) = self.n_channels, self.shape
        nz_own = self.own_planes[self.rank]
        if gather and self.world > 1 and self._fused_gather_ready():
            return self._run_fused_gather(slab)
        mn, mx = self._normalization(slab)
        if self.has_rows:
            be.run(slab, self.shape, p, self.rows, mn, mx)
        halo = self._halo_buffer()
        if halo is not None:
            be.partial(halo)
        seed = self._exchange_halo(halo)
        if not gather or self.world == 1:
            own = self._own
            if own is None or tuple(own.shape) != (c, nz_own, h, w) or own.device != dev:
                own = self._own = torch.empty((c, nz_own, h, w), dtype=torch.float32, device=dev)
            if self.has_rows and nz_own > 0:
                be.stitch(seed, own)   # writes every voxel of the owned planes (shell = 0.0)
            else:
                own.zero_()
            return own
        # C3: every rank ends up with the full (C, D, H, W) array.  The owned planes are stitched
        # straight into it and the other ranks' planes arrive in place: the output is channel-
        # major, so a rank's planes are one contiguous chunk per channel -- C sends and C receives
        # per peer in ONE grouped NCCL call, no staging copies, no concatenation.
        if self._full is None or self._full.device != dev:
            self._full = torch.empty((c, d, h, w), dtype=torch.float32, device=dev)
        full = self._full
        starts = np.concatenate([[0], np.cumsum(self.own_planes)]).astype(int)
        z0 = int(starts[self.rank])
        if self.has_rows and nz_own > 0:
            if hasattr(be, "stitch_into"):
                be.stitch_into(seed, full, z0)
            else:
                own = torch.zeros((c, nz_own, h, w), dtype=torch.float32, device=dev)
                be.stitch(seed, own)
                full[:, z0:z0 + nz_own].copy_(own)
        be.sync()
        ops = []
        for g in range(self.world):
            if g == self.rank:
                continue
            gz0, gn = int(starts[g]), self.own_planes[g]
            for ch in range(c):
                if gn > 0:
                    ops.append(dist.P2POp(dist.irecv, full[ch, gz0:gz0 + gn], self._peer(g), self.group))
                if nz_own > 0:
                    ops.append(dist.P2POp(dist.isend, full[ch, z0:z0 + nz_own], self._peer(g), self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return full

    def own_bounds(self):
        z0 = self.plan["out_z0"] if self.has_rows else 0
        return z0, z0 + self.own_planes[self.rank]


def predict_sharded(
    img,
    model,
    affinity_mode=True,
    batch_size=16,
    brightness_clip=1000,
    normalization_percentiles=(1, 99.9),
    patch_shape=(96, 96, 96),
    overlap=(32, 32, 32),
    trim=8,
    precision=None,
    group=None,
    gather=True,
    backend=None,
    out=None,
):
    """``predict`` sharded by z patch-rows over the ranks of a ``torch.distributed`` group.

    One process per GPU; every rank passes the same ``img`` (only its slab, including the
    input halo it shares with the next rank, is uploaded).  With ``gather`` every rank
    returns the full ``(C, D, H, W)`` array; otherwise ``(z0, z1, planes)`` with the planes
    it owns -- those are produced by the row-group pipeline (``SlabJob.run_pipelined``), their
    device->host copies overlapped with the remaining rows, into ``out`` when given (C-contiguous
    float32 ``(C, z1 - z0, H, W)``, e.g. pinned memory; ``SlabJob.own_bounds`` gives the range).
    With a world size of 1 this is the same computation as ``predict``.
    """
    _check_clip(img, brightness_clip)
    vol = _as_volume_u16(img, brightness_clip)
    if backend is None:
        backend = _EngineSlabBackend(_engine_for(model, precision))
    if world_size_of(group) > 1 and patch_shape[0] - 2 * trim > 2 * (patch_shape[0] - overlap[0]):
        raise ValueError(
            "predict_sharded needs patch - 2*trim <= 2*(patch - overlap) along z: with more overlap a "
            "plane is covered by three or more z rows and the pairwise partial-sum hand-over between "
            "neighbouring ranks no longer reproduces the reference's summation order; use predict() "
            "(single GPU) or a smaller z overlap")
    params = _native.make_params(patch_shape, overlap, trim, brightness_clip,
                                 normalization_percentiles, batch=max(int(batch_size), 32))
    job = SlabJob(vol.shape, params, 3 if affinity_mode else 1, backend, group)
    if gather:
        full = job.run(job.upload(vol), gather=True).cpu().numpy()
        return full if affinity_mode else full[0]
    z0, z1 = job.own_bounds()
    shape = (job.n_channels, z1 - z0) + vol.shape[1:]
    if out is None:
        out = np.empty(shape, dtype=np.float32)
    if out.shape != shape or out.dtype != np.float32 or not out.flags.c_contiguous:
        raise ValueError(f"out must be a C-contiguous float32 array of shape {shape}")
    job.run_pipelined(job.upload(vol), torch.from_numpy(out))
    return z0, z1, (out if affinity_mode else out[0])
