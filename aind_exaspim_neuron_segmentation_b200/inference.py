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

import numpy as np
import torch

from . import _native
from .engine import Engine, percentiles_from_hist, plan_slab
from .machine_learning.unet3d import UNet3D

__all__ = ["predict", "predict_sharded", "load_model", "count_patches", "generate_patch_starts"]


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
        cache = model.__dict__.setdefault("_exa_b200_engines", {})
        device = next(model.parameters()).device
        sd = model.state_dict()
        fp = tuple((k, v.data_ptr(), v._version) for k, v in sd.items())
        key = (precision or "bf16", str(device))
        if key not in cache or cache[key][0] != fp:
            if model.training:
                raise RuntimeError("model must be in eval mode (inference.py:423)")
            cache[key] = (fp, Engine(sd, device, precision or "bf16"))
        return cache[key][1]
    raise TypeError("model must be a torch.nn.Module with the reference UNet3D state_dict")


# --- input handling ---------------------------------------------------------------------
def _as_volume_u16(img, brightness_clip):
    """(…,D,H,W) array-like -> C-contiguous uint16 (D,H,W) holding min(img, clip) exactly.

    The reference computes ``np.minimum(img, brightness_clip)`` in the image's own
    dtype (inference.py:79).  For integer images that is representable in uint16
    whenever the clipped values are, which is what the kernels consume.
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
        clip = min(max(int(brightness_clip), 0), 65535)
        return np.minimum(arr, clip).astype(np.uint16)
    raise TypeError(
        f"unsupported image dtype {arr.dtype}: the B200 path consumes integer (ExaSPIM uint16) "
        "volumes; there is no CPU fallback for floating-point images"
    )


def _check_clip(arr, brightness_clip):
    if brightness_clip < 0:
        raise ValueError("brightness_clip must be >= 0")


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
):
    """Affinity (or foreground) prediction for a 3-D volume; see reference inference.py:29-126.

    Returns a new C-contiguous float32 array ``(3, D, H, W)`` (``(D, H, W)`` when
    ``affinity_mode=False``).  ``batch_size`` is accepted for compatibility and used as
    a lower bound on the number of patches per wave -- it has no numerical effect
    (eval-mode BatchNorm, no cross-sample op).
    """
    _check_clip(img, brightness_clip)
    vol = _as_volume_u16(img, brightness_clip)
    engine = _engine_for(model, precision)
    n_channels = 3 if affinity_mode else 1
    if engine.out_channels != n_channels:
        raise ValueError(
            f"model has {engine.out_channels} output channels but affinity_mode={affinity_mode} "
            f"needs {n_channels}"
        )
    shape5 = (1, 1) + vol.shape
    n_patches = count_patches(shape5, patch_shape, overlap)
    pbar = None
    if verbose:
        from tqdm import tqdm

        pbar = tqdm(total=n_patches, desc="Predict")
    params = _native.make_params(patch_shape, overlap, trim, brightness_clip,
                                 normalization_percentiles, batch=max(int(batch_size), 32))
    out = engine.predict_host(vol, params)
    if pbar is not None:
        pbar.update(n_patches)
        pbar.close()
    return out if affinity_mode else out[0]


# --- multi-GPU: z-row slabs ---------------------------------------------------------------
def split_rows(n_rows, world_size):
    """Contiguous, balanced [begin, end) row ranges, one per rank (earlier ranks get the extras)."""
    base, extra = divmod(n_rows, world_size)
    out, r = [], 0
    for g in range(world_size):
        n = base + (1 if g < extra else 0)
        out.append((r, r + n))
        r += n
    return out


class _EngineSlabBackend:
    """Slab compute on the native engine (device tensors)."""

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

    def to_device(self, host_u16):
        return torch.from_numpy(host_u16).to(self.device, non_blocking=True)


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
    device_out=False,
):
    """``predict`` sharded by z patch-rows over the ranks of a ``torch.distributed`` group.

    Every rank passes the same ``img`` (only its slab plus a 32-plane input halo is
    uploaded).  Exchange steps, all small next to the convolutions:
    C1 all-reduce of the 1001-bin histogram (global percentiles), C2 send of raw partial
    sums for the planes shared with the next rank's first row (summed there in the
    reference's order), C3 gather of the owned output slabs to every rank.

    Returns the full ``(C, D, H, W)`` array on every rank when ``gather`` is true, else
    ``(z0, z1, slab)`` with the rank's own planes.  ``backend`` is a test seam for the slab
    compute; the default is the native engine.
    """
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    vol = _as_volume_u16(img, brightness_clip)
    shape = vol.shape
    if backend is None:
        backend = _EngineSlabBackend(_engine_for(model, precision))
    n_channels = 3 if affinity_mode else 1
    if backend.out_channels != n_channels:
        raise ValueError("model output channels do not match affinity_mode")
    params = _native.make_params(patch_shape, overlap, trim, brightness_clip,
                                 normalization_percentiles, batch=max(int(batch_size), 32))
    clip = params.brightness_clip
    full = plan_slab(shape, params, 0, 0)  # grid sizes only
    rows = split_rows(full["nz"], world)[rank]
    plan = plan_slab(shape, params, rows[0], rows[1])
    dev = backend.device
    h, w = shape[1], shape[2]

    # C1: global histogram -> exact percentiles.  Ranks histogram disjoint plane ranges.
    cuts = [round(shape[0] * g / world) for g in range(world + 1)]
    part = backend.to_device(vol[cuts[rank]:cuts[rank + 1]])
    hist = backend.histogram(part, clip) if part.numel() else torch.zeros(
        clip + 1, dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    mn, mx = percentiles_from_hist(hist.cpu().numpy().astype(np.uint64),
                                   params.pct_lo, params.pct_hi)

    # slab upload + all patches of the rank's rows
    has_rows = rows[1] > rows[0]
    if has_rows:
        slab = backend.to_device(vol[plan["in_z0"]:plan["in_z1"]])
        backend.run(slab, shape, params, rows, mn, mx)

    # C2: partial sums of the shared planes go to the owner (the next rank)
    seed = None
    reqs = []
    n_halo = plan["halo_z1"] - plan["halo_z0"]
    n_seed = plan["seed_z1"] - plan["seed_z0"]
    if world > 1:
        if has_rows and n_seed > 0:
            seed = torch.empty((n_channels, n_seed, h, w), dtype=torch.float32, device=dev)
            reqs.append(dist.irecv(seed, src=_global_rank(group, rank - 1), group=group))
        if has_rows and n_halo > 0 and rank + 1 < world:
            halo = torch.empty((n_channels, n_halo, h, w), dtype=torch.float32, device=dev)
            backend.partial(halo)
            if dev.type == "cuda":
                torch.cuda.current_stream(dev).synchronize()
            reqs.append(dist.isend(halo, dst=_global_rank(group, rank + 1), group=group))
        for r in reqs:
            r.wait()

    # own planes
    nz_own = max(plan["out_z1"] - plan["out_z0"], 0) if has_rows else 0
    own = torch.zeros((n_channels, nz_own, h, w), dtype=torch.float32, device=dev)
    if has_rows and nz_own > 0:
        backend.stitch(seed, own)
    if not gather:
        return plan["out_z0"], plan["out_z0"] + nz_own, own

    # C3: gather of the owned slabs (channel-major output => one strided copy per rank)
    if world == 1:
        result = own
    else:
        all_plans = [plan_slab(shape, params, *r) for r in split_rows(full["nz"], world)]
        sizes = [max(p["out_z1"] - p["out_z0"], 0) if r[1] > r[0] else 0
                 for p, r in zip(all_plans, split_rows(full["nz"], world))]
        pieces = [torch.empty((n_channels, s, h, w), dtype=torch.float32, device=dev)
                  for s in sizes]
        dist.all_gather(pieces, own, group=group) if len(set(sizes)) == 1 else \
            _all_gather_ragged(pieces, own, rank, world, group)
        result = torch.cat(pieces, dim=1)
    if device_out:
        return result if affinity_mode else result[0]
    out = result.cpu().numpy()
    return out if affinity_mode else out[0]


def _global_rank(group, group_rank):
    import torch.distributed as dist

    if group is None:
        return group_rank
    return dist.get_global_rank(group, group_rank)


def _all_gather_ragged(pieces, own, rank, world, group):
    """all_gather for per-rank slabs of different plane counts: one broadcast per rank."""
    import torch.distributed as dist

    for g in range(world):
        if pieces[g].numel() == 0:
            continue
        if g == rank:
            pieces[g].copy_(own)
        dist.broadcast(pieces[g], src=_global_rank(group, g), group=group)
