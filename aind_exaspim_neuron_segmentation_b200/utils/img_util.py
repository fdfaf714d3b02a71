"""Image sources and sinks for the out-of-core path (``inference.predict_streamed``).

Reference: utils/img_util.py:23-121 (``read``: zarr / N5 / TIFF from local disk, GCS or S3).  The
same entry point is kept -- a path in, something array-like out -- and the array-likes it returns
are what ``predict_streamed`` consumes: a ``(D, H, W)`` ``.shape`` and planes for ``img[z0:z1]``.
zarr, tifffile, gcsfs and s3fs are imported lazily (they are not needed for anything else in this
package and are absent from the offline image it is built in); ``.npy`` files and directories of
``.npy`` z-chunks are handled without any dependency, so that volumes larger than host memory can
be streamed from plain files.
"""

import os
import re

import numpy as np


def read(img_path):
    """Open an image volume by extension; reference img_util.py:23-48.

    ``.zarr`` / ``.n5`` -> the zarr group or array (lazy, sliceable), ``.tif`` / ``.tiff`` -> ndarray,
    ``.npy`` -> read-only memory map, a directory of ``*.npy`` z-chunks -> :class:`NpyStack`.
    """
    if ".zarr" in img_path or ".n5" in img_path:
        try:
            import zarr
        except ImportError as exc:  # pragma: no cover - depends on the environment
            raise ImportError("reading zarr / N5 volumes needs the 'zarr' package") from exc
        if ".n5" in img_path:
            return zarr.open(zarr.n5.N5Store(img_path), mode="r")
        return zarr.open(img_path, mode="r")
    if ".tif" in img_path:
        try:
            import tifffile
        except ImportError as exc:  # pragma: no cover - depends on the environment
            raise ImportError("reading TIFF volumes needs the 'tifffile' package") from exc
        return tifffile.imread(img_path)
    if img_path.endswith(".npy"):
        return np.load(img_path, mmap_mode="r")
    if os.path.isdir(img_path) and any(f.endswith(".npy") for f in os.listdir(img_path)):
        return NpyStack(img_path)
    raise ValueError(f"Unsupported image format: {img_path}")


def _natural_key(name):
    return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", name)]


class NpyStack:
    """A ``(D, H, W)`` volume stored as ``*.npy`` chunks of consecutive z planes in one directory
    (natural sort order of the file names = z order).  Chunks are memory-mapped; ``stack[z0:z1]``
    returns the planes as one array, touching only the chunks involved."""

    def __init__(self, directory):
        names = sorted((f for f in os.listdir(directory) if f.endswith(".npy")), key=_natural_key)
        if not names:
            raise ValueError(f"no .npy chunks in {directory}")
        self._chunks = [np.load(os.path.join(directory, n), mmap_mode="r") for n in names]
        first = self._chunks[0]
        if first.ndim != 3 or any(c.ndim != 3 or c.shape[1:] != first.shape[1:] or c.dtype != first.dtype
                                  for c in self._chunks):
            raise ValueError("chunks must be 3-D arrays with equal (H, W) and dtype")
        self._starts = np.concatenate([[0], np.cumsum([c.shape[0] for c in self._chunks])]).astype(int)
        self.shape = (int(self._starts[-1]),) + tuple(first.shape[1:])
        self.dtype = first.dtype
        self.ndim = 3

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, key):
        rest = ()
        if isinstance(key, tuple):
            key, rest = key[0], key[1:]
        if isinstance(key, (int, np.integer)):
            z = int(key) + (self.shape[0] if key < 0 else 0)
            return self[z:z + 1][(0,) + rest] if rest else self[z:z + 1][0]
        if not isinstance(key, slice) or key.step not in (None, 1):
            raise IndexError("NpyStack supports contiguous z slices (and integers) on the first axis")
        z0, z1, _ = key.indices(self.shape[0])
        parts = []
        first = int(np.searchsorted(self._starts, z0, side="right")) - 1
        for i in range(max(first, 0), len(self._chunks)):
            a, b = int(self._starts[i]), int(self._starts[i + 1])
            if a >= z1:
                break
            parts.append(self._chunks[i][max(z0, a) - a:min(z1, b) - a])
        if not parts:
            out = np.empty((0,) + self.shape[1:], self.dtype)
        else:
            out = np.concatenate(parts) if len(parts) != 1 else np.asarray(parts[0])
        return out[(slice(None),) + rest] if rest else out

    def __array__(self, dtype=None, copy=None):
        arr = self[0:self.shape[0]]
        return arr.astype(dtype) if dtype is not None else arr


class NpySink:
    """Slice-assignable sink for ``predict_streamed``: every ``sink[:, z0:z1] = planes`` (or
    ``sink[z0:z1] = planes`` for single-channel output) is written as one ``.npy`` file named after
    its plane range, so the result never has to exist in memory as a whole.  ``NpySink.open`` reads
    it back as an array (for tests and small volumes)."""

    def __init__(self, directory, channels=3):
        os.makedirs(directory, exist_ok=True)
        self.directory = directory
        self.channels = channels

    def __setitem__(self, key, value):
        zkey = key[1] if isinstance(key, tuple) else key
        if isinstance(key, tuple) and key[0] != slice(None):
            raise IndexError("NpySink takes sink[:, z0:z1] = planes or sink[z0:z1] = planes")
        if not isinstance(zkey, slice) or zkey.step not in (None, 1) or zkey.start is None or zkey.stop is None:
            raise IndexError("NpySink needs an explicit contiguous z range")
        np.save(os.path.join(self.directory, f"z{zkey.start:07d}_{zkey.stop:07d}.npy"), np.asarray(value))

    @staticmethod
    def open(directory):
        names = sorted((f for f in os.listdir(directory) if re.fullmatch(r"z\d{7}_\d{7}\.npy", f)))
        parts = [np.load(os.path.join(directory, n), mmap_mode="r") for n in names]
        axis = 1 if parts and parts[0].ndim == 4 else 0
        return np.concatenate(parts, axis=axis)
