"""ctypes binding of the C-ABI library ``libexaspim_b200.so`` (include/exaspim_b200.h).

The library holds every CUDA kernel of the path; this module only loads it,
declares the signatures and turns negative return codes into ``RuntimeError``.
There is no CPU fallback: if the library is missing it is built with nvcc
(``build()``), and if that is impossible the import of a compute entry point
fails loudly.
"""

import ctypes
import os
import shutil
import subprocess

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG_DIR, "csrc")
LIB_PATH = os.path.join(_PKG_DIR, "libexaspim_b200.so")
SOURCES = ["capi.cu", "engine.cu", "kernels_mem.cu", "conv_umma.cu", "conv_zfold.cu", "conv_stem.cu",
           "watershed.cu", "float_volume.cu", "train_kernels.cu", "train_wgrad.cu", "trainer.cu"]
HEADERS = ["common.cuh", "conv_umma.cuh", "conv_zfold.cuh", "conv_zfold2.cuh", "conv_stem.cuh", "engine.h",
           "kernels.h", "tmap.h", "watershed.h", "ws_agglomerate.h", "float_volume.h", "train_kernels.h",
           "trainer.h"]

PRECISION_BF16 = 0
PRECISION_FP32 = 1
DTYPE_F32 = 0
DTYPE_I64 = 1


class PredictParams(ctypes.Structure):
    """``exa_predict_params`` -- keyword arguments of reference predict()."""

    _fields_ = [
        ("patch", ctypes.c_int32 * 3),
        ("overlap", ctypes.c_int32 * 3),
        ("trim", ctypes.c_int32),
        ("brightness_clip", ctypes.c_int32),
        ("pct_lo", ctypes.c_double),
        ("pct_hi", ctypes.c_double),
        ("batch", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


class SlabPlan(ctypes.Structure):
    """``exa_slab_plan``."""

    _fields_ = [
        (name, ctypes.c_int32)
        for name in (
            "nz", "ny", "nx", "n_patches", "in_z0", "in_z1", "out_z0", "out_z1",
            "halo_z0", "halo_z1", "seed_z0", "seed_z1",
        )
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


def _nvcc():
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return cand if os.path.exists(cand) else None


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(_CSRC, f) for f in SOURCES + HEADERS]
    deps.append(os.path.join(_PKG_DIR, "..", "include", "exaspim_b200.h"))
    return any(os.path.exists(d) and os.path.getmtime(d) > built for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into ``libexaspim_b200.so`` (in-tree)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = _nvcc()
    if nvcc is None:
        if os.path.exists(LIB_PATH):
            return LIB_PATH  # GPU box without a toolchain mismatch: use the prebuilt library
        raise RuntimeError("nvcc not found and libexaspim_b200.so is not built")
    cmd = [
        nvcc, "--threads", "0", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
        "-std=c++17",
        "-Xcompiler", "-fPIC", "-shared", "-o", LIB_PATH + ".tmp",
    ] + [os.path.join(_CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    if verbose:
        print(res.stderr)
    return LIB_PATH


PROGRESS_FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64)

_lib = None

_SIGNATURES = {
    # name: (restype, argtypes)
    "exa_version": (ctypes.c_char_p, []),
    "exa_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "exa_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "exa_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "exa_load_weight": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p,
                                       ctypes.POINTER(ctypes.c_int64), ctypes.c_int, ctypes.c_int]),
    "exa_finalize_weights": (ctypes.c_int, [ctypes.c_void_p]),
    "exa_out_channels": (ctypes.c_int, [ctypes.c_void_p]),
    "exa_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                   ctypes.POINTER(ctypes.c_int32), ctypes.c_void_p]),
    "exa_predict": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_int, ctypes.POINTER(PredictParams), ctypes.c_void_p]),
    "exa_set_progress_callback": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "exa_predict_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_int,
                                          ctypes.POINTER(PredictParams), ctypes.c_void_p,
                                          ctypes.c_void_p]),
    "exa_plan_slab": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.POINTER(PredictParams), ctypes.c_int, ctypes.c_int,
                                     ctypes.POINTER(SlabPlan)]),
    "exa_histogram": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                     ctypes.c_void_p, ctypes.c_void_p]),
    "exa_percentiles_from_hist": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_double,
                                                 ctypes.c_double, ctypes.POINTER(ctypes.c_double),
                                                 ctypes.POINTER(ctypes.c_double)]),
    "exa_set_normalization_table": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                                   ctypes.c_double, ctypes.c_double]),
    "exa_compress_float_volume": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64,
                                                 ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p,
                                                 ctypes.POINTER(ctypes.c_int), ctypes.c_void_p]),
    "exa_percentiles_from_hist_values": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                                        ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                                        ctypes.POINTER(ctypes.c_double),
                                                        ctypes.POINTER(ctypes.c_double)]),
    "exa_set_normalization": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, ctypes.c_double,
                                             ctypes.c_int]),
    "exa_slab_run": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, ctypes.POINTER(PredictParams), ctypes.c_int,
                                    ctypes.c_int, ctypes.c_void_p]),
    "exa_slab_partial": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "exa_slab_stitch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p]),
    "exa_slab_stitch_strided": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_int64, ctypes.c_void_p]),
    "exa_slab_predict": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.POINTER(PredictParams), ctypes.c_int,
                                        ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                        ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                        ctypes.c_void_p]),
    "exa_slab_finish": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                       ctypes.c_void_p]),
    "exa_set_peer_outputs": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                            ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]),
    "exa_affinities_to_segmentation": (ctypes.c_int, [
        ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
        ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_double, ctypes.c_double,
        ctypes.c_int64, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64),
        ctypes.POINTER(ctypes.c_int64)]),
    "exa_ws_release_memory": (ctypes.c_int, []),
    "exa_ws_last_profile": (ctypes.c_int, [ctypes.POINTER(ctypes.c_double), ctypes.c_int]),
    "exa_region_agglomerate": (ctypes.c_int, [ctypes.c_int, ctypes.c_uint32, ctypes.c_int64,
                                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p]),
    "exa_affinities_to_segmentation_device": (ctypes.c_int, [
        ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
        ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_double, ctypes.c_double,
        ctypes.c_int64, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64),
        ctypes.POINTER(ctypes.c_int64), ctypes.c_void_p]),
    "exa_count_patches": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.POINTER(ctypes.c_int32),
                                         ctypes.POINTER(ctypes.c_int32)]),
    "exa_patch_starts": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.POINTER(ctypes.c_int32),
                                        ctypes.POINTER(ctypes.c_int32),
                                        ctypes.POINTER(ctypes.c_int32), ctypes.c_int]),
    "exa_profile_begin": (ctypes.c_int, [ctypes.c_void_p]),
    "exa_profile_end": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double),
                                       ctypes.POINTER(ctypes.c_int64), ctypes.c_int]),
    "exa_profile_layers": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double),
                                          ctypes.POINTER(ctypes.c_int64),
                                          ctypes.POINTER(ctypes.c_int32), ctypes.c_int]),
    "exa_launch_count": (ctypes.c_int64, [ctypes.c_void_p]),
    # training step
    "exa_train_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "exa_train_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "exa_train_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "exa_train_bind": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p,
                                      ctypes.POINTER(ctypes.c_int64), ctypes.c_int]),
    "exa_train_grad_elems": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64)]),
    "exa_train_grad_slot": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p,
                                           ctypes.POINTER(ctypes.c_int64),
                                           ctypes.POINTER(ctypes.c_int64)]),
    "exa_train_out_channels": (ctypes.c_int, [ctypes.c_void_p]),
    "exa_train_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                         ctypes.POINTER(ctypes.c_int32), ctypes.c_void_p,
                                         ctypes.c_void_p]),
    "exa_train_backward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p]),
    "exa_train_profile_begin": (ctypes.c_int, [ctypes.c_void_p]),
    "exa_train_profile_end": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double),
                                             ctypes.POINTER(ctypes.c_int64), ctypes.c_int]),
    "exa_train_launch_count": (ctypes.c_int64, [ctypes.c_void_p]),
    "exa_train_workspace_bytes": (ctypes.c_int64, [ctypes.c_void_p]),
    "exa_conv3d_weight_grad": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                              ctypes.c_void_p] + [ctypes.c_int] * 6 +
                               [ctypes.c_void_p, ctypes.c_void_p]),
    "exa_conv3d_data_grad": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                            ctypes.c_void_p] + [ctypes.c_int] * 6 +
                             [ctypes.c_void_p, ctypes.c_void_p]),
    "exa_bce_with_logits": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                           ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib():
    """Load (building first if needed) the C-ABI library."""
    global _lib
    if _lib is None:
        path = build()
        handle = ctypes.CDLL(path)
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError here = header/library mismatch
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(code, engine=None, what=""):
    """Raise ``RuntimeError`` with the library's message for a negative return code."""
    if code >= 0:
        return code
    msg = lib().exa_last_error(engine)
    msg = msg.decode() if msg else "unknown error"
    raise RuntimeError(f"exaspim_b200 {what} failed ({code}): {msg}")


def make_params(patch_shape, overlap, trim, brightness_clip, percentiles, batch=0):
    p = PredictParams()
    for i in range(3):
        p.patch[i] = int(patch_shape[i])
        p.overlap[i] = int(overlap[i])
    p.trim = int(trim)
    p.brightness_clip = int(min(max(int(brightness_clip), 0), 65535))
    p.pct_lo = float(percentiles[0])
    p.pct_hi = float(percentiles[1])
    p.batch = int(batch)
    p.reserved = 0
    return p
