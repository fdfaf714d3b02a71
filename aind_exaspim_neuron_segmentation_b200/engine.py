"""Python handle on one native engine (one CUDA device): packed weights, workspaces, job state.

Thin by design -- every method is one call through the C ABI declared in
``include/exaspim_b200.h``; tensors cross the boundary as raw device pointers.
"""

import ctypes

import numpy as np
import torch

from . import _native


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Engine:
    """Owns an ``exa_engine``; built from a ``state_dict`` with the reference's layout."""

    def __init__(self, state_dict, device, precision="bf16"):
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError(
                "exaspim_b200 runs on CUDA devices only (no CPU fallback); got device "
                f"'{device}'"
            )
        self.device = torch.device("cuda", device.index if device.index is not None
                                   else torch.cuda.current_device())
        self.precision = precision
        self._lib = _native.lib()
        handle = ctypes.c_void_p()
        code = self._lib.exa_create(
            self.device.index,
            _native.PRECISION_BF16 if precision == "bf16" else _native.PRECISION_FP32,
            ctypes.byref(handle),
        )
        _native.check(code, None, "exa_create")
        self._h = handle
        self._load(state_dict)

    # -- weights ---------------------------------------------------------------
    def _load(self, state_dict):
        for name, value in state_dict.items():
            t = value.detach().to("cpu")
            if t.dtype == torch.int64:
                dtype, arr = _native.DTYPE_I64, t.numpy().astype(np.int64, copy=False)
            else:
                dtype, arr = _native.DTYPE_F32, t.to(torch.float32).numpy()
            dims = tuple(arr.shape)  # ascontiguousarray would promote 0-d to 1-d
            arr = np.ascontiguousarray(arr)
            shape = (ctypes.c_int64 * max(len(dims), 1))(*dims)
            code = self._lib.exa_load_weight(
                self._h, name.encode(), arr.ctypes.data_as(ctypes.c_void_p), shape, len(dims), dtype
            )
            _native.check(code, self._h, f"exa_load_weight({name})")
        _native.check(self._lib.exa_finalize_weights(self._h), self._h, "exa_finalize_weights")
        self.out_channels = self._lib.exa_out_channels(self._h)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.exa_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_progress(self, fn):
        """fn(user, patches_done, patches_total) from a CUDA host-function thread whenever a wave
        of patches has finished on the device; None switches it off."""
        self._progress = _native.PROGRESS_FN(fn) if fn is not None else None   # keep it alive
        cb = ctypes.cast(self._progress, ctypes.c_void_p) if fn is not None else ctypes.c_void_p(0)
        _native.check(self._lib.exa_set_progress_callback(self._h, cb, ctypes.c_void_p(0)), self._h,
                      "exa_set_progress_callback")

    @property
    def launch_count(self):
        return int(self._lib.exa_launch_count(self._h))

    PROFILE_CATEGORIES = ("histogram", "stem", "conv", "pool", "upsample", "head", "stitch")

    def profile_begin(self):
        _native.check(self._lib.exa_profile_begin(self._h), self._h, "exa_profile_begin")

    def profile_end(self):
        """-> {category: (device ms, launches)} since profile_begin (synchronises)."""
        n = len(self.PROFILE_CATEGORIES)
        ms = (ctypes.c_double * n)()
        cnt = (ctypes.c_int64 * n)()
        _native.check(self._lib.exa_profile_end(self._h, ms, cnt, n), self._h, "exa_profile_end")
        return {k: (ms[i], int(cnt[i])) for i, k in enumerate(self.PROFILE_CATEGORIES)}

    LAYER_KERNELS = ("none", "conv3x3_umma_kernel", "conv3x3_zfold_kernel", "conv3x3_zfold2_kernel")

    def profile_layers(self):
        """-> [(device ms, launches, kernel name)] * 18 per conv layer for the last profile_end."""
        ms = (ctypes.c_double * 18)()
        cnt = (ctypes.c_int64 * 18)()
        kind = (ctypes.c_int32 * 18)()
        _native.check(self._lib.exa_profile_layers(self._h, ms, cnt, kind, 18), self._h,
                      "exa_profile_layers")
        return [(ms[i], int(cnt[i]), self.LAYER_KERNELS[kind[i]]) for i in range(18)]

    # -- operator level ----------------------------------------------------------
    def forward(self, x):
        """float32 (B,1,Pz,Py,Px) cuda tensor -> float32 logits (B,C,Pz,Py,Px)."""
        if x.device != self.device:
            raise RuntimeError(f"input is on {x.device}, engine on {self.device}")
        if x.dim() != 5 or x.shape[1] != 1:
            raise ValueError("expected input of shape (B, 1, D, H, W)")
        x = x.to(torch.float32).contiguous()
        b, _, pz, py, px = x.shape
        out = torch.empty((b, self.out_channels, pz, py, px), dtype=torch.float32,
                          device=self.device)
        patch = (ctypes.c_int32 * 3)(pz, py, px)
        code = self._lib.exa_forward(self._h, _ptr(x), _ptr(out), b, patch,
                                     _stream_ptr(self.device))
        _native.check(code, self._h, "exa_forward")
        return out

    # -- whole path ---------------------------------------------------------------
    def predict_host(self, vol_u16, params, out=None):
        """numpy uint16 (D,H,W) -> numpy float32 (C,D,H,W), host buffers end to end."""
        vol_u16 = np.ascontiguousarray(vol_u16, dtype=np.uint16)
        d, h, w = vol_u16.shape
        if out is None:
            out = np.empty((self.out_channels, d, h, w), dtype=np.float32)
        elif (out.dtype != np.float32 or out.shape != (self.out_channels, d, h, w)
              or not out.flags["C_CONTIGUOUS"]):
            raise ValueError("out must be a C-contiguous float32 array of shape (C, D, H, W)")
        code = self._lib.exa_predict(
            self._h, vol_u16.ctypes.data_as(ctypes.c_void_p), d, h, w, ctypes.byref(params),
            out.ctypes.data_as(ctypes.c_void_p),
        )
        _native.check(code, self._h, "exa_predict")
        return out

    def predict_device(self, vol_dev, params, out=None):
        """cuda uint16 (D,H,W) tensor -> cuda float32 (C,D,H,W) tensor (stream-ordered)."""
        assert vol_dev.dtype == torch.uint16 and vol_dev.is_contiguous()
        d, h, w = vol_dev.shape
        if out is None:
            out = torch.empty((self.out_channels, d, h, w), dtype=torch.float32,
                              device=self.device)
        code = self._lib.exa_predict_device(
            self._h, _ptr(vol_dev), d, h, w, ctypes.byref(params), _ptr(out),
            _stream_ptr(self.device),
        )
        _native.check(code, self._h, "exa_predict_device")
        return out

    # -- slab pieces (z-row sharding) -----------------------------------------------
    def histogram(self, vol_dev, clip):
        hist = torch.zeros(clip + 1, dtype=torch.int64, device=self.device)
        code = self._lib.exa_histogram(self._h, _ptr(vol_dev), vol_dev.numel(), clip, _ptr(hist),
                                       _stream_ptr(self.device))
        _native.check(code, self._h, "exa_histogram")
        return hist

    def set_normalization(self, mn, mx, clip):
        _native.check(self._lib.exa_set_normalization(self._h, float(mn), float(mx), int(clip)),
                      self._h, "exa_set_normalization")

    def set_normalization_table(self, values, mn, mx):
        """Normalisation for a rank-compressed float volume: ``values`` are its distinct clipped
        intensities in ascending order (``compress_float_volume``)."""
        values = np.ascontiguousarray(values, dtype=np.float64)
        _native.check(self._lib.exa_set_normalization_table(
            self._h, values.ctypes.data_as(ctypes.c_void_p), int(values.size), float(mn), float(mx)),
            self._h, "exa_set_normalization_table")

    def compress_float_volume(self, vol_dev, clip):
        """cuda float32 / float64 (D,H,W) tensor -> (cuda uint16 ranks of min(x, clip), float64 table
        of the distinct clipped values); raises when there are more than 65536 of them."""
        if vol_dev.dtype not in (torch.float32, torch.float64) or not vol_dev.is_contiguous():
            raise TypeError("expected a contiguous float32 / float64 CUDA tensor")
        idx = torch.empty(vol_dev.shape, dtype=torch.uint16, device=self.device)
        table = np.zeros(65536, dtype=np.float64)
        n = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            code = self._lib.exa_compress_float_volume(
                _ptr(vol_dev), 1 if vol_dev.dtype == torch.float64 else 0, vol_dev.numel(), float(clip),
                _ptr(idx), table.ctypes.data_as(ctypes.c_void_p), ctypes.byref(n),
                _stream_ptr(self.device))
        _native.check(code, None, "exa_compress_float_volume")
        return idx, table[:n.value].copy()

    def slab_run(self, slab_dev, shape, params, row_begin, row_end):
        d, h, w = shape
        code = self._lib.exa_slab_run(self._h, _ptr(slab_dev), d, h, w, ctypes.byref(params),
                                      row_begin, row_end, _stream_ptr(self.device))
        _native.check(code, self._h, "exa_slab_run")

    def slab_partial(self, halo_dev):
        _native.check(self._lib.exa_slab_partial(self._h, _ptr(halo_dev),
                                                 _stream_ptr(self.device)),
                      self._h, "exa_slab_partial")

    def slab_stitch(self, seed_dev, out_dev, channel_stride=0):
        """out_dev: dense (C, own planes, H, W), or -- with channel_stride -- a view whose first
        element is plane out_z0 of channel 0 of a larger channel-major array."""
        _native.check(self._lib.exa_slab_stitch_strided(self._h, _ptr(seed_dev), _ptr(out_dev),
                                                        int(channel_stride), _stream_ptr(self.device)),
                      self._h, "exa_slab_stitch_strided")

    def slab_predict(self, slab_dev, shape, params, row_begin, row_end, out_dev, out_host, halo_dev):
        """Rows [row_begin,row_end) as a row-group pipeline; out_dev / out_host: dense
        (C, own planes, H, W) float32 (out_host a host tensor, ideally pinned, or None).  Finish
        with :meth:`slab_finish` after the halo/seed exchange."""
        d, h, w = shape
        cs = out_dev.stride(0) if out_dev.dim() == 4 and out_dev.shape[0] > 1 else out_dev[0].numel()
        hs = 0 if out_host is None else (out_host.stride(0) if out_host.shape[0] > 1 else out_host[0].numel())
        code = self._lib.exa_slab_predict(self._h, _ptr(slab_dev), d, h, w, ctypes.byref(params),
                                          row_begin, row_end, _ptr(out_dev), int(cs),
                                          _ptr(out_host), int(hs), _ptr(halo_dev),
                                          _stream_ptr(self.device))
        _native.check(code, self._h, "exa_slab_predict")

    def slab_finish(self, seed_dev, out_dev, out_host):
        cs = out_dev.stride(0) if out_dev.shape[0] > 1 else out_dev[0].numel()
        hs = 0 if out_host is None else (out_host.stride(0) if out_host.shape[0] > 1 else out_host[0].numel())
        code = self._lib.exa_slab_finish(self._h, _ptr(seed_dev), _ptr(out_dev), int(cs),
                                         _ptr(out_host), int(hs), _stream_ptr(self.device))
        _native.check(code, self._h, "exa_slab_finish")

    def set_peer_outputs(self, full, peer_ptrs):
        """Fused gather: stitches into ``full`` (this rank's (C, D, H, W) array) also store to the
        peer-mapped copies at ``peer_ptrs`` (ints).  An empty list switches it off."""
        arr = (ctypes.c_void_p * max(len(peer_ptrs), 1))(*[int(p) for p in peer_ptrs])
        code = self._lib.exa_set_peer_outputs(self._h, _ptr(full) if peer_ptrs else ctypes.c_void_p(0),
                                              full.numel() if peer_ptrs else 0, arr, len(peer_ptrs))
        _native.check(code, self._h, "exa_set_peer_outputs")


# -- host helpers that need no GPU --------------------------------------------------
def plan_slab(shape, params, row_begin, row_end):
    plan = _native.SlabPlan()
    d, h, w = shape
    code = _native.lib().exa_plan_slab(d, h, w, ctypes.byref(params), row_begin, row_end,
                                       ctypes.byref(plan))
    _native.check(code, None, "exa_plan_slab")
    return plan.as_dict()


def percentiles_from_hist_values(hist, values, is_f32, q_lo, q_hi):
    """np.percentile(method="linear") from the histogram of ranks and the table of values."""
    hist = np.ascontiguousarray(hist, dtype=np.uint64)
    values = np.ascontiguousarray(values, dtype=np.float64)
    mn, mx = ctypes.c_double(), ctypes.c_double()
    code = _native.lib().exa_percentiles_from_hist_values(
        hist.ctypes.data_as(ctypes.c_void_p), values.ctypes.data_as(ctypes.c_void_p), hist.size,
        1 if is_f32 else 0, q_lo, q_hi, ctypes.byref(mn), ctypes.byref(mx))
    _native.check(code, None, "exa_percentiles_from_hist_values")
    return mn.value, mx.value


def percentiles_from_hist(hist, q_lo, q_hi):
    hist = np.ascontiguousarray(hist, dtype=np.uint64)
    mn, mx = ctypes.c_double(), ctypes.c_double()
    code = _native.lib().exa_percentiles_from_hist(
        hist.ctypes.data_as(ctypes.c_void_p), hist.size, q_lo, q_hi, ctypes.byref(mn),
        ctypes.byref(mx))
    _native.check(code, None, "exa_percentiles_from_hist")
    return mn.value, mx.value
