"""Device-level operator API over ``libyamb200.so``.

``Backend`` owns one ``yam_ctx`` on one GPU.  Operators take and return
``torch`` CUDA tensors (used only as device-memory handles; all arithmetic is in
the hand-written sm_100a kernels) shaped ``(h, w)`` or ``(n, h, w)`` for a stack
of frames (``(..., 3)`` for BGR input of ``bgr2gray``).  Everything is enqueued
on torch's current CUDA stream, so operators chain without host round trips.

``to_device`` / ``to_host`` move NumPy arrays through pinned staging buffers.
There is no CPU implementation behind any of these calls.
"""
from __future__ import annotations

import ctypes as C
import math
import threading
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import (
    BORDER_REFLECT101,
    BORDER_REPLICATE,
    MORPH_CLOSE,
    MORPH_DILATE,
    MORPH_ERODE,
    MORPH_OPEN,
    PROPS_STRIDE,
    SHAPE_CROSS,
    SHAPE_ELLIPSE,
    SHAPE_RECT,
    YAM_F32,
    YAM_I32,
    YAM_U8,
    YAM_U16,
    BackendUnavailable,
    YamError,
)

_SHAPES = {"rectangular": SHAPE_RECT, "elliptical": SHAPE_ELLIPSE, "cross": SHAPE_CROSS}


def _torch():
    import torch  # deferred: importing the package must not require CUDA

    return torch


def _dtype_code(t) -> int:
    torch = _torch()
    table = {torch.uint8: YAM_U8, torch.uint16: YAM_U16, torch.float32: YAM_F32, torch.int32: YAM_I32}
    try:
        return table[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype}; expected uint8, uint16, float32 or int32") from None


def shape_code(kernel_shape: str) -> int:
    """core/segmentation.py:265-274: unknown names fall back to rectangular."""
    return _SHAPES.get(str(kernel_shape).lower(), SHAPE_RECT)


class Backend:
    """One libyamb200 context bound to one CUDA device."""

    def __init__(self, device: int = 0) -> None:
        torch = _torch()
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise BackendUnavailable(
                "no CUDA device visible; yamimageprocessor_b200 runs on B200 only and has no CPU fallback"
            )
        self.device_index = int(device)
        self.device = torch.device("cuda", self.device_index)
        handle = C.c_void_p()
        # one process per GPU: this process's share of the host cores (torchrun exports LOCAL_WORLD_SIZE)
        import os

        local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
        self.host_threads = int(self.lib.yam_set_host_threads(max(1, (os.cpu_count() or 4) // local_world)))
        _lib.check("yam_ctx_create", self.lib.yam_ctx_create(self.device_index, C.byref(handle)))
        self._ctx = handle
        self._pinned: Dict[Tuple[str, int], object] = {}
        self._lock = threading.RLock()
        self._last_stream = None

    # ------------------------------------------------------------------ plumbing
    def close(self) -> None:
        if getattr(self, "_ctx", None):
            self.lib.yam_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self) -> None:  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    def _call(self, name: str, *args) -> None:
        torch = _torch()
        with self._lock:
            cur = torch.cuda.current_stream(self.device)
            stream = cur.cuda_stream
            # The context's scratch arena and staging buffers are shared by every caller of this Backend and
            # are ordered by the stream the work is enqueued on.  Callers normally stay on one stream; when
            # the current stream CHANGES, the previous stream is drained first so that two streams can never
            # have kernels in flight on the same scratch (rare, so the common path pays nothing).
            prev = self._last_stream
            if prev is not None and prev.cuda_stream != stream:
                prev.synchronize()
            self._last_stream = cur
            _lib.check("yam_ctx_set_stream", self.lib.yam_ctx_set_stream(self._ctx, C.c_void_p(stream)))
            _lib.check(name, getattr(self.lib, name)(self._ctx, *args))

    def launch_count(self, reset: bool = False) -> int:
        return int(self.lib.yam_ctx_launch_count(self._ctx, 1 if reset else 0))

    def synchronize(self) -> None:
        _torch().cuda.synchronize(self.device)

    def _check(self, t, *, ndim=(2, 3), dtypes: Optional[Sequence] = None, name="image"):
        torch = _torch()
        if not isinstance(t, torch.Tensor) or t.device != self.device:
            raise TypeError(f"{name} must be a CUDA tensor on {self.device}")
        if t.dim() not in ndim:
            raise ValueError(f"{name} must have {ndim} dimensions, got shape {tuple(t.shape)}")
        if dtypes is not None and t.dtype not in dtypes:
            raise TypeError(f"{name} dtype {t.dtype} not supported (expected one of {list(dtypes)})")
        if t.numel() == 0:
            raise ValueError(f"{name} is empty")
        return t.contiguous()

    @staticmethod
    def _nhw(t) -> Tuple[int, int, int]:
        if t.dim() == 2:
            return 1, int(t.shape[0]), int(t.shape[1])
        return int(t.shape[0]), int(t.shape[1]), int(t.shape[2])

    @staticmethod
    def _p(t) -> C.c_void_p:
        return C.c_void_p(t.data_ptr())

    # ------------------------------------------------------------------ host <-> device
    def _staging(self, kind: str, nbytes: int):
        torch = _torch()
        key = (kind, self.device_index)
        buf = self._pinned.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, pin_memory=True)
            self._pinned[key] = buf
        return buf

    _NP2T = None

    @classmethod
    def _np2t(cls):
        if cls._NP2T is None:
            torch = _torch()
            cls._NP2T = {np.dtype(np.uint8): torch.uint8, np.dtype(np.uint16): torch.uint16,
                         np.dtype(np.float32): torch.float32, np.dtype(np.int32): torch.int32,
                         np.dtype(np.int64): torch.int64}
        return cls._NP2T

    def pinned_empty(self, shape, dtype) -> np.ndarray:
        """A NumPy array in page-locked memory (inputs placed here upload without a staging copy)."""
        torch = _torch()
        t = torch.empty(tuple(shape), dtype=self._np2t()[np.dtype(dtype)], pin_memory=True)
        return t.numpy()

    def to_device(self, array: np.ndarray):
        """NumPy -> CUDA tensor, async on the current stream.

        Page-locked inputs (``pinned_empty``) are copied directly; pageable inputs go through a
        pinned staging buffer.
        """
        torch = _torch()
        a = np.ascontiguousarray(array)
        if a.dtype not in self._np2t() or a.dtype == np.int64:
            raise TypeError(f"unsupported dtype {a.dtype}; expected uint8, uint16, float32 or int32")
        out = torch.empty(a.shape, dtype=self._np2t()[a.dtype], device=self.device)
        if a.size == 0:
            return out
        src = torch.from_numpy(a)
        if src.is_pinned():
            out.copy_(src, non_blocking=True)
            # keep the host array alive until the copy has been consumed
            out._yam_host_ref = a  # type: ignore[attr-defined]
            return out
        nbytes = a.nbytes
        with self._lock:  # fill + enqueue must be atomic with respect to other host threads
            stage = self._staging("h2d", nbytes)
            torch.cuda.current_stream(self.device).synchronize()  # previous user of the staging buffer
            stage.numpy()[:nbytes] = a.reshape(-1).view(np.uint8)
            out.view(torch.uint8).reshape(-1).copy_(stage[:nbytes], non_blocking=True)
        return out

    def to_host(self, t) -> np.ndarray:
        """CUDA tensor -> fresh C-contiguous NumPy array backed by page-locked memory (synchronous)."""
        torch = _torch()
        t = t.contiguous()
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        host.copy_(t, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return host.numpy()

    # ------------------------------------------------------------------ K1
    def bgr2gray(self, img):
        """cv2.cvtColor(BGR2GRAY); 2-D (or stack of 2-D) input is returned unchanged like the reference."""
        torch = _torch()
        if img.dim() == 2 or (img.dim() == 3 and img.shape[-1] not in (3, 4)):
            return img
        img = self._check(img, ndim=(3, 4), dtypes=(torch.uint8, torch.uint16, torch.float32))
        if img.shape[-1] == 4:  # cv2.cvtColor(BGRA, COLOR_BGR2GRAY) ignores alpha (verified against cv2 4.13)
            img = img[..., :3].contiguous()
        lead = img.shape[:-1]
        n = 1 if img.dim() == 3 else int(img.shape[0])
        h, w = int(lead[-2]), int(lead[-1])
        out = torch.empty(lead, dtype=img.dtype, device=self.device)
        self._call("yam_bgr2gray", self._p(img), self._p(out), n, h, w, _dtype_code(img))
        return out

    # ------------------------------------------------------------------ K2
    def minmax(self, img) -> np.ndarray:
        torch = _torch()
        img = self._check(img, dtypes=(torch.uint8, torch.uint16, torch.float32))
        n, h, w = self._nhw(img)
        host = (C.c_double * (2 * n))()
        self._call("yam_minmax", self._p(img), n, h, w, _dtype_code(img), None, C.cast(host, C.c_void_p))
        return np.array(host, dtype=np.float64).reshape(n, 2)

    def normalize_minmax(self, img, alpha: float = 0.0, beta: float = 255.0):
        torch = _torch()
        img = self._check(img, dtypes=(torch.uint8, torch.uint16, torch.float32))
        n, h, w = self._nhw(img)
        out = torch.empty_like(img)
        self._call("yam_normalize_minmax", self._p(img), self._p(out), n, h, w, _dtype_code(img),
                   float(alpha), float(beta))
        return out

    def convert_scale_abs(self, img, alpha: float = 1.0, beta: float = 0.0):
        torch = _torch()
        img = self._check(img, ndim=(2, 3, 4), dtypes=(torch.uint8, torch.uint16, torch.float32))
        out = torch.empty(img.shape, dtype=torch.uint8, device=self.device)
        self._call("yam_convert_scale_abs", self._p(img), self._p(out), img.numel(), _dtype_code(img),
                   float(alpha), float(beta))
        return out

    def lut_u8(self, img, table: np.ndarray):
        torch = _torch()
        img = self._check(img, ndim=(2, 3, 4), dtypes=(torch.uint8,))
        tab = np.ascontiguousarray(table, dtype=np.uint8)
        if tab.size != 256:
            raise ValueError("LUT must have 256 entries")
        out = torch.empty_like(img)
        self._call("yam_lut_u8", self._p(img), self._p(out), img.numel(), tab.ctypes.data_as(C.c_void_p))
        return out

    def threshold(self, img, thresh: float, maxval: float = 255.0):
        torch = _torch()
        img = self._check(img, dtypes=(torch.uint8, torch.uint16, torch.float32))
        out = torch.empty_like(img)
        self._call("yam_threshold", self._p(img), self._p(out), img.numel(), _dtype_code(img),
                   float(thresh), float(maxval))
        return out

    # ------------------------------------------------------------------ K3 / K4 / K5
    def gaussian(self, img, ksize: int, sigma: float = 0.0, border: int = BORDER_REFLECT101):
        torch = _torch()
        img = self._check(img, dtypes=(torch.uint8, torch.uint16, torch.float32))
        n, h, w = self._nhw(img)
        out = torch.empty_like(img)
        self._call("yam_gaussian", self._p(img), self._p(out), n, h, w, _dtype_code(img), int(ksize),
                   float(sigma), int(border))
        return out

    def box(self, img, ksize: int):
        torch = _torch()
        img = self._check(img, dtypes=(torch.uint8, torch.uint16))
        n, h, w = self._nhw(img)
        out = torch.empty_like(img)
        self._call("yam_box", self._p(img), self._p(out), n, h, w, _dtype_code(img), int(ksize))
        return out

    def median(self, img, ksize: int):
        torch = _torch()
        img = self._check(img, dtypes=(torch.uint8, torch.uint16))
        n, h, w = self._nhw(img)
        out = torch.empty_like(img)
        self._call("yam_median", self._p(img), self._p(out), n, h, w, _dtype_code(img), int(ksize))
        return out

    def split_channels(self, img):
        """(h, w, c) interleaved -> (c, h, w) planes."""
        torch = _torch()
        img = self._check(img, ndim=(3,), dtypes=(torch.uint8, torch.uint16, torch.float32))
        h, w, c = (int(v) for v in img.shape)
        out = torch.empty((c, h, w), dtype=img.dtype, device=self.device)
        self._call("yam_split_channels", self._p(img), self._p(out), h * w, c, _dtype_code(img))
        return out

    def merge_channels(self, planes):
        """(c, h, w) planes -> (h, w, c) interleaved."""
        torch = _torch()
        planes = self._check(planes, ndim=(3,), dtypes=(torch.uint8, torch.uint16, torch.float32))
        c, h, w = (int(v) for v in planes.shape)
        out = torch.empty((h, w, c), dtype=planes.dtype, device=self.device)
        self._call("yam_merge_channels", self._p(planes), self._p(out), h * w, c, _dtype_code(planes))
        return out

    # ------------------------------------------------------------------ K9
    def adaptive_threshold(self, img, block_size: int = 11, C_: float = 2.0):
        torch = _torch()
        img = self._check(img, dtypes=(torch.uint8, torch.uint16))
        n, h, w = self._nhw(img)
        out = torch.empty(img.shape, dtype=torch.uint8, device=self.device)
        self._call("yam_adaptive_threshold", self._p(img), self._p(out), n, h, w, _dtype_code(img),
                   int(block_size), float(C_))
        return out

    # ------------------------------------------------------------------ K6
    def morph(self, img, op: int, kernel_shape: str = "Rectangular", kernel_size: int = 3, iterations: int = 1):
        torch = _torch()
        img = self._check(img, dtypes=(torch.uint8, torch.uint16))
        if int(iterations) == 0:      # cv2.erode / dilate / morphologyEx with iterations = 0 copy the input
            return img.clone()
        n, h, w = self._nhw(img)
        out = torch.empty_like(img)
        self._call("yam_morph", self._p(img), self._p(out), n, h, w, _dtype_code(img), int(op),
                   shape_code(kernel_shape), int(kernel_size), int(iterations))
        return out

    def erode(self, img, kernel_shape="Rectangular", kernel_size=3, iterations=1):
        return self.morph(img, MORPH_ERODE, kernel_shape, kernel_size, iterations)

    def dilate(self, img, kernel_shape="Rectangular", kernel_size=3, iterations=1):
        return self.morph(img, MORPH_DILATE, kernel_shape, kernel_size, iterations)

    def morph_open(self, img, kernel_shape="Rectangular", kernel_size=3, iterations=1):
        return self.morph(img, MORPH_OPEN, kernel_shape, kernel_size, iterations)

    def morph_close(self, img, kernel_shape="Rectangular", kernel_size=3, iterations=1):
        return self.morph(img, MORPH_CLOSE, kernel_shape, kernel_size, iterations)

    def morph_open_close(self, img, kernel_size: int = 5, iterations: int = 1):
        torch = _torch()
        img = self._check(img, dtypes=(torch.uint8, torch.uint16))
        if int(iterations) == 0:
            return img.clone()
        n, h, w = self._nhw(img)
        out = torch.empty_like(img)
        self._call("yam_morph_open_close", self._p(img), self._p(out), n, h, w, _dtype_code(img),
                   int(kernel_size), int(iterations))
        return out

    # ------------------------------------------------------------------ K7 / K8
    def histogram(self, img):
        torch = _torch()
        img = self._check(img, dtypes=(torch.uint8, torch.uint16))
        n, h, w = self._nhw(img)
        bins = 256 if img.dtype == torch.uint8 else 65536
        hist = torch.empty((n, bins), dtype=torch.int64, device=self.device)
        self._call("yam_histogram", self._p(img), n, h, w, _dtype_code(img), self._p(hist))
        return hist

    def otsu_threshold(self, img, maxval: float = 255.0, want_image: bool = True):
        """Returns (thresholds int32[n] on device, thresholded image or None)."""
        torch = _torch()
        img = self._check(img, dtypes=(torch.uint8, torch.uint16))
        n, h, w = self._nhw(img)
        t = torch.empty((n,), dtype=torch.int32, device=self.device)
        out = torch.empty_like(img) if want_image else None
        self._call("yam_otsu_threshold", self._p(img), self._p(out) if out is not None else None, n, h, w,
                   _dtype_code(img), float(maxval), self._p(t), None)
        return t, out

    def otsu_begin(self, img):
        """First half of ``otsu_threshold`` (histogram + scan, thresholds stay on the device).  The whole
        operator is asynchronous since the scan moved to the device, so this simply enqueues it; the
        two-step form is kept for schedules written against it.  Returns a handle for ``otsu_finish``."""
        torch = _torch()
        img = self._check(img, dtypes=(torch.uint8, torch.uint16))
        t, _ = self.otsu_threshold(img, want_image=False)
        return {"img": img, "t": t}

    def otsu_finish(self, handle, maxval: float = 255.0, want_image: bool = True):
        """Second half: (thresholds int32[n] on device, thresholded image or None)."""
        t = handle["t"]
        return t, (self.threshold_frames(handle["img"], t, maxval) if want_image else None)

    def threshold_frames(self, img, t_dev, maxval: float = 255.0):
        """dst = src > t_dev[frame] ? maxval : 0 per frame, thresholds int32[n] on the device."""
        torch = _torch()
        img = self._check(img, dtypes=(torch.uint8, torch.uint16))
        n, h, w = self._nhw(img)
        out = torch.empty_like(img)
        self._call("yam_threshold_frames", self._p(img), self._p(out), n, h, w, _dtype_code(img), self._p(t_dev),
                   float(maxval))
        return out

    def otsu_from_histogram_device(self, hist, want_certified: bool = False):
        """Otsu thresholds (int32[n], device) of histograms that are already on the device: int64 / uint64
        counts, shape (bins,) or (n, bins), bins = 256 | 65536 (e.g. an all-reduced mosaic histogram).
        ``want_certified``: also return int32[n] = 1 where the parallel certificate decided the frame."""
        torch = _torch()
        if not isinstance(hist, torch.Tensor) or hist.device != self.device or hist.dtype not in (torch.int64, torch.uint64):
            raise TypeError("otsu_from_histogram_device expects an int64 CUDA tensor on this backend's device")
        hist = hist.contiguous()
        bins = int(hist.shape[-1])
        n = int(hist.numel() // bins)
        t = torch.empty((n,), dtype=torch.int32, device=self.device)
        cert = torch.empty((n,), dtype=torch.int32, device=self.device) if want_certified else None
        self._call("yam_otsu_from_hist_dev", self._p(hist), bins, n, self._p(t), self._p(cert) if cert is not None else None)
        return (t, cert) if want_certified else t

    def equalize_hist(self, img):
        torch = _torch()
        img = self._check(img, dtypes=(torch.uint8,))
        n, h, w = self._nhw(img)
        out = torch.empty_like(img)
        self._call("yam_equalize_hist", self._p(img), self._p(out), n, h, w)
        return out

    def equalize_hist_bgr(self, img):
        """Preprocessor.histogram_equalization on a colour image (core/preprocessing.py:77-79): equalise the Y
        plane of cv2's YCrCb, convert back.  (h, w, 3|4) uint8 BGR[A] -> (h, w, 3) uint8 (cv2's
        BGR2YCrCb ignores alpha and the reference converts back to three channels)."""
        torch = _torch()
        img = self._check(img, ndim=(3,), dtypes=(torch.uint8,))
        if img.shape[-1] not in (3, 4):
            raise ValueError(f"equalize_hist_bgr expects (h, w, 3|4), got {tuple(img.shape)}")
        if img.shape[-1] == 4:
            img = img[..., :3].contiguous()
        h, w = int(img.shape[0]), int(img.shape[1])
        y = torch.empty((h, w), dtype=torch.uint8, device=self.device)
        self._call("yam_bgr_luma_ycrcb", self._p(img), self._p(y), h * w)
        y_eq = self.equalize_hist(y)
        out = torch.empty_like(img)
        self._call("yam_bgr_replace_luma_ycrcb", self._p(img), self._p(y_eq), self._p(out), h * w)
        return out

    def clahe(self, img, clip_limit: float = 2.0, tile_grid: Tuple[int, int] = (8, 8)):
        torch = _torch()
        img = self._check(img, dtypes=(torch.uint8, torch.uint16))
        n, h, w = self._nhw(img)
        out = torch.empty_like(img)
        self._call("yam_clahe", self._p(img), self._p(out), n, h, w, _dtype_code(img), float(clip_limit),
                   int(tile_grid[0]), int(tile_grid[1]))
        return out

    def clahe_luts(self, img, clip_limit: float = 2.0, tile_grid: Tuple[int, int] = (8, 8)):
        """Per-tile CLAHE LUTs of a single image: tensor (tiles_y, tiles_x, bins) in the image dtype."""
        torch = _torch()
        img = self._check(img, ndim=(2,), dtypes=(torch.uint8, torch.uint16))
        h, w = int(img.shape[0]), int(img.shape[1])
        bins = 256 if img.dtype == torch.uint8 else 65536
        luts = torch.empty((int(tile_grid[1]), int(tile_grid[0]), bins), dtype=img.dtype, device=self.device)
        self._call("yam_clahe_luts", self._p(img), h, w, _dtype_code(img), float(clip_limit), int(tile_grid[0]),
                   int(tile_grid[1]), self._p(luts))
        return luts

    def clahe_apply(self, rows, luts, tile_size: Tuple[int, int], y_offset: int = 0):
        """Blend precomputed LUTs over `rows` (a block of rows starting at global row y_offset);
        luts is (tiles_y, tiles_x, bins), tile_size = (tile_w, tile_h) of the WHOLE image."""
        torch = _torch()
        rows = self._check(rows, ndim=(2,), dtypes=(torch.uint8, torch.uint16))
        if luts.dtype != rows.dtype or luts.dim() != 3:
            raise TypeError("luts must be (tiles_y, tiles_x, bins) in the image dtype")
        luts = luts.contiguous()
        out = torch.empty_like(rows)
        self._call("yam_clahe_apply", self._p(rows), self._p(out), int(rows.shape[0]), int(rows.shape[1]),
                   _dtype_code(rows), self._p(luts), int(luts.shape[1]), int(luts.shape[0]), int(tile_size[0]),
                   int(tile_size[1]), int(y_offset))
        return out

    def relabel(self, labels, remap):
        """In place: labels = remap[labels] (remap int32 on device, remap[0] must be 0)."""
        torch = _torch()
        labels = self._check(labels, ndim=(2, 3), dtypes=(torch.int32,), name="labels")
        if remap.dtype != torch.int32 or remap.device != self.device:
            raise TypeError("remap must be an int32 CUDA tensor")
        remap = remap.contiguous()
        self._call("yam_relabel", self._p(labels), labels.numel(), self._p(remap), remap.numel())
        return labels

    def merge_strip_labels(self, edges, offsets, total: Optional[int] = None):
        """Roots of the cross-strip label union (see yam_merge_strip_labels): ``edges`` int32
        [world, 2, w] first/last label rows per strip, ``offsets`` int64 [world + 1] on the device
        (``total`` = offsets[-1] when the caller already has it on the host).
        Returns int32 [total + 1]: smallest global id of every id's merged component."""
        torch = _torch()
        edges = self._check(edges, ndim=(3,), dtypes=(torch.int32,), name="edges")
        offsets = offsets.to(device=self.device, dtype=torch.int64).contiguous()
        world, two, w = (int(v) for v in edges.shape)
        if two != 2 or offsets.numel() != world + 1:
            raise ValueError("edges must be [world, 2, w] and offsets [world + 1]")
        if total is None:
            total = int(offsets[-1].item())
        root = torch.empty((total + 1,), dtype=torch.int32, device=self.device)
        self._call("yam_merge_strip_labels", self._p(edges), self._p(offsets), world, w, total, self._p(root))
        return root

    def merge_strips_remap(self, packed, width: int, offsets: Sequence[int], rank: int, count: int = 1, counts_dev=None):
        """Cross-strip label merge + raster-first renumbering in one call (yam_merge_strips_remap).
        ``packed`` int32 [world, stride >= 2*width]: first and last label row of every strip;
        ``offsets`` host ints [world + 1].  Returns (remaps, total int32[1]): ``remaps`` is the int32
        table of strip ``rank`` (count_rank + 1 entries) or, for ``count`` > 1, the list of tables of
        strips rank .. rank + count - 1 (views of one buffer).

        ``counts_dev`` (a strided int32 device view with one real count per strip): ``offsets`` are then prefixes
        of UPPER BOUNDS and the host never needs the counts (yam_merge_strips_remap_bounded); the result is
        ``(remaps, total, overflow int32[1])`` -- overflow != 0 means a count exceeded its bound and the merge
        has to be repeated with exact offsets."""
        torch = _torch()
        packed = self._check(packed, ndim=(2,), dtypes=(torch.int32,), name="packed")
        world, stride = int(packed.shape[0]), int(packed.shape[1])
        offs = np.ascontiguousarray(offsets, dtype=np.int64)
        if offs.size != world + 1:
            raise ValueError("offsets must have world + 1 entries")
        nbytes = int(self.lib.yam_merge_strips_workspace_bytes(int(offs[-1])))
        if nbytes < 0:
            raise ValueError("too many components for int32 labels")
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=self.device)
        sizes = [int(offs[r + 1] - offs[r]) + 1 for r in range(rank, rank + count)]
        remap = torch.empty((sum(sizes),), dtype=torch.int32, device=self.device)
        total = torch.empty((1,), dtype=torch.int32, device=self.device)
        overflow = None
        if counts_dev is None:
            self._call("yam_merge_strips_remap", self._p(packed), stride, world, int(width),
                       offs.ctypes.data_as(C.c_void_p), int(rank), int(count), self._p(ws), self._p(remap), self._p(total))
        else:
            if counts_dev.dtype != torch.int32 or counts_dev.dim() != 1 or int(counts_dev.shape[0]) != world:
                raise TypeError("counts_dev must be an int32 device vector with one entry per strip")
            overflow = torch.empty((1,), dtype=torch.int32, device=self.device)
            self._call("yam_merge_strips_remap_bounded", self._p(packed), stride, world, int(width),
                       offs.ctypes.data_as(C.c_void_p), self._p(counts_dev), int(counts_dev.stride(0)), int(rank), int(count),
                       self._p(ws), self._p(remap), self._p(total), self._p(overflow))
        if count == 1:
            return (remap, total) if overflow is None else (remap, total, overflow)
        views, at = [], 0
        for n in sizes:
            views.append(remap[at:at + n])
            at += n
        return (views, total) if overflow is None else (views, total, overflow)

    # ------------------------------------------------------------------ watershed front half
    def threshold_inv(self, img, thresh: float = 0.0, maxval: float = 255.0, t_dev=None):
        """cv2.threshold(..., THRESH_BINARY_INV): dst = src > t ? 0 : maxval; ``t_dev`` (int32[n] on the
        device, e.g. from ``otsu_threshold``) takes precedence over ``thresh``."""
        torch = _torch()
        img = self._check(img, dtypes=(torch.uint8, torch.uint16))
        n, h, w = self._nhw(img)
        out = torch.empty_like(img)
        self._call("yam_threshold_inv", self._p(img), self._p(out), n, h, w, _dtype_code(img),
                   self._p(t_dev) if t_dev is not None else None, float(thresh), float(maxval))
        return out

    def distance_transform(self, mask):
        """cv2.distanceTransform(mask, DIST_L2, 5): float32 chamfer distance to the nearest zero pixel."""
        torch = _torch()
        mask = self._check(mask, dtypes=(torch.uint8,), name="mask")
        n, h, w = self._nhw(mask)
        dist = torch.empty(mask.shape, dtype=torch.float32, device=self.device)
        launches = C.c_int(0)
        self._call("yam_distance_transform", self._p(mask), self._p(dist), n, h, w, C.cast(C.byref(launches), C.c_void_p))
        self.last_distance_launches = int(launches.value)
        return dist

    def watershed_markers(self, gray, kernel_size: int = 3, opening_iterations: int = 2, dilation_iterations: int = 3,
                          distance_threshold_factor: float = 0.7):
        """The marker construction of ``Detector.watershed_segmentation`` (core/segmentation.py:97-111) for one
        uint8 frame, every stage on the device.  Returns a dict of CUDA tensors: ``thresh`` (what the reference
        step returns), ``opening``, ``sure_bg``, ``dist``, ``sure_fg``, ``markers`` (int32: 0 unknown,
        1 background, k + 1 for marker k in raster-first order) and ``n_markers`` (int32[1])."""
        torch = _torch()
        gray = self._check(gray, ndim=(2,), dtypes=(torch.uint8,), name="gray")
        t, _ = self.otsu_threshold(gray, want_image=False)
        thresh = self.threshold_inv(gray, t_dev=t)
        opening = self.morph_open(thresh, "Rectangular", kernel_size, opening_iterations)
        sure_bg = self.dilate(opening, "Rectangular", kernel_size, dilation_iterations)
        dist = self.distance_transform(opening)
        dmax = float(self.minmax(dist)[0, 1])
        # cv2.threshold on float32 compares against the threshold cast to float32
        fg_f = self.threshold(dist, float(np.float32(distance_threshold_factor * np.float64(np.float32(dmax)))), 255.0)
        sure_fg = self.convert_scale_abs(fg_f, 1.0, 0.0)         # np.uint8(sure_fg)
        labels, counts = self.ccl_label(sure_fg)
        markers = torch.empty_like(labels)
        self._call("yam_watershed_combine", self._p(labels), self._p(sure_bg), self._p(sure_fg), self._p(markers), int(labels.numel()))
        return {"thresh": thresh, "opening": opening, "sure_bg": sure_bg, "dist": dist, "sure_fg": sure_fg,
                "markers": markers, "n_markers": counts}

    def region_moments(self, labels, n_labels: int):
        """int64 [n_labels, 3] on device: per label sum r^2, sum c^2, sum r*c (second-order raw moments)."""
        torch = _torch()
        labels = self._check(labels, ndim=(2,), dtypes=(torch.int32,), name="labels")
        out = torch.zeros((int(n_labels), 3), dtype=torch.int64, device=self.device)
        if n_labels > 0:
            self._call("yam_region_moments", self._p(labels), int(labels.shape[0]), int(labels.shape[1]), int(n_labels), self._p(out))
        return out

    def region_perimeter_counts(self, labels, n_labels: int):
        """int64 [n_labels, 3] on device: border pixels per label in the three weight classes of
        skimage.measure.perimeter(neighborhood=4) (1, sqrt 2, (1 + sqrt 2) / 2); see ``contour_columns``."""
        torch = _torch()
        labels = self._check(labels, ndim=(2,), dtypes=(torch.int32,), name="labels")
        out = torch.zeros((int(n_labels), 3), dtype=torch.int64, device=self.device)
        if n_labels > 0:
            self._call("yam_region_perimeter", self._p(labels), int(labels.shape[0]), int(labels.shape[1]), int(n_labels), self._p(out))
        return out

    def region_convex_area(self, labels, n_labels: int, props=None):
        """int64 [n_labels] on device: skimage's ``area_convex`` per label (pixels of the convex hull image).
        ``props`` = the ``region_props`` table of the same label image (computed here when omitted)."""
        torch = _torch()
        labels = self._check(labels, ndim=(2,), dtypes=(torch.int32,), name="labels")
        if props is None:
            props = self.region_props(labels, None, int(n_labels))
        if not isinstance(props, torch.Tensor) or props.device != self.device or props.dtype != torch.int64 \
                or tuple(props.shape) != (int(n_labels), PROPS_STRIDE) or not props.is_contiguous():
            raise TypeError("region_convex_area: props must be the contiguous int64 [n_labels, 8] table of region_props on this device")
        out = torch.zeros((int(n_labels),), dtype=torch.int64, device=self.device)
        if n_labels > 0:
            self._call("yam_region_convex_area", self._p(labels), int(labels.shape[0]), int(labels.shape[1]), int(n_labels),
                       self._p(props), self._p(out))
        return out

    def checksum64(self, t, index_base: int = 0, accumulate=None):
        """Order-independent content checksum (yam_checksum64) of a uint8 / uint16 / int32 CUDA tensor;
        returns an int64 tensor [1] (bit pattern of the uint64 sum), adding into ``accumulate`` if given."""
        torch = _torch()
        if not isinstance(t, torch.Tensor) or t.device != self.device:
            raise TypeError("checksum64 expects a CUDA tensor on this backend's device")
        t = t.contiguous()
        out = accumulate if accumulate is not None else torch.zeros((1,), dtype=torch.int64, device=self.device)
        if t.numel():
            self._call("yam_checksum64", self._p(t), int(t.numel()), _dtype_code(t), int(index_base), self._p(out))
        return out

    def otsu_from_histogram(self, hist: np.ndarray) -> int:
        """cv2's Otsu recurrence on a host histogram (int64/uint64 counts); host-only helper."""
        h = np.ascontiguousarray(hist, dtype=np.uint64)
        t = C.c_int()
        _lib.check("yam_otsu_from_hist", self.lib.yam_otsu_from_hist(h.ctypes.data_as(C.c_void_p), int(h.size),
                                                                      C.cast(C.byref(t), C.c_void_p)))
        return int(t.value)

    # ------------------------------------------------------------------ remaining menu steps
    def add_weighted(self, a, alpha: float, b, beta: float, gamma: float = 0.0):
        torch = _torch()
        a = self._check(a, ndim=(2, 3), dtypes=(torch.uint8, torch.uint16), name="a")
        b = self._check(b, ndim=(2, 3), dtypes=(a.dtype,), name="b")
        if a.shape != b.shape:
            raise ValueError("add_weighted: operands must have the same shape")
        out = torch.empty_like(a)
        self._call("yam_add_weighted", self._p(a), self._p(b), self._p(out), int(a.numel()), _dtype_code(a),
                   float(alpha), float(beta), float(gamma))
        return out

    def sharpen(self, img, strength: float = 1.0):
        """Unsharp mask: addWeighted(img, 1+s, GaussianBlur(img, (0,0), sigma=3), -s)."""
        blurred = self.gaussian(img, 0, 3.0)
        return self.add_weighted(img, 1.0 + float(strength), blurred, -float(strength), 0.0)

    def select_channel(self, img, channel: str = "All"):
        """SelectChannelModule: img is (h, w, 3) BGR or a (h, w) plane (replicated to BGR first)."""
        torch = _torch()
        modes = {"B": 0, "G": 1, "R": 2, "RG": 3, "GB": 4, "BR": 5}
        img = self._check(img, ndim=(2, 3), dtypes=(torch.uint8, torch.uint16), name="image")
        if img.dim() == 3 and img.shape[-1] != 3:
            raise ValueError("select_channel expects an interleaved 3-channel image or a single plane")
        if channel not in modes:
            if img.dim() == 3:
                return img
            out = torch.empty(tuple(img.shape) + (3,), dtype=img.dtype, device=self.device)
            self._call("yam_gray2bgr", self._p(img), self._p(out), int(img.numel()), _dtype_code(img))
            return out
        if img.dim() == 2:
            if modes[channel] <= 2:
                return img.clone()
            if img.dtype != torch.uint8:
                raise TypeError("two-channel means are defined for uint8 only")
            return img.clone()  # (x + x) >> 1 == x
        if modes[channel] > 2 and img.dtype != torch.uint8:
            raise TypeError("two-channel means are defined for uint8 only")
        out = torch.empty(tuple(img.shape[:2]), dtype=img.dtype, device=self.device)
        self._call("yam_select_channel", self._p(img), self._p(out), int(out.numel()), _dtype_code(img), modes[channel])
        return out

    def border_clear(self, img, border_distance: int):
        """remove_border_regions: (h, w), (n, h, w) planes or (h, w, 3) colour."""
        torch = _torch()
        img = self._check(img, ndim=(2, 3), dtypes=(torch.uint8, torch.uint16), name="image")
        out = torch.empty_like(img)
        if img.dim() == 3 and img.shape[-1] in (3, 4):
            n, h, w, ch = 1, int(img.shape[0]), int(img.shape[1]), int(img.shape[2])
        else:
            n, h, w = self._nhw(img)
            ch = 1
        self._call("yam_border_clear", self._p(img), self._p(out), n, h, w, ch, _dtype_code(img), int(border_distance))
        return out

    def edge_filter(self, img, kind: str, ksize: int = 3):
        """Sobel / Prewitt / Laplacian magnitude image (uint8), kind in {'sobel','prewitt','laplacian'}."""
        torch = _torch()
        kinds = {"sobel": 0, "prewitt": 1, "laplacian": 2}
        img = self._check(img, dtypes=(torch.uint8, torch.uint16))
        n, h, w = self._nhw(img)
        out = torch.empty(img.shape, dtype=torch.uint8, device=self.device)
        self._call("yam_edge_filter", self._p(img), self._p(out), n, h, w, _dtype_code(img), kinds[kind], int(ksize))
        return out

    def mask_row_moments(self, mask):
        """int64 [h, 4] on device: per row (count, sum x, sum x^2, sum x^3) over the set pixels."""
        torch = _torch()
        mask = self._check(mask, ndim=(2,), dtypes=(torch.uint8, torch.uint16), name="mask")
        h, w = int(mask.shape[0]), int(mask.shape[1])
        out = torch.empty((h, 4), dtype=torch.int64, device=self.device)
        self._call("yam_mask_row_moments", self._p(mask), h, w, _dtype_code(mask), self._p(out))
        return out

    # ------------------------------------------------------------------ fused binary segmentation
    def adaptive_threshold_bits(self, img, block_size: int = 11, C_: float = 2.0, mask_thresh=None, maxval: float = 255.0):
        """Adaptive threshold with a 1-bit-per-pixel result: int32 tensor (..., h, ceil(w/32)).

        ``mask_thresh`` (int32[n] on the device, e.g. the Otsu thresholds): also return the global threshold
        mask ``img > t[frame] ? maxval : 0`` (= ``threshold_frames``), written in the same pass over ``img``;
        the result is then ``(bits, mask)``."""
        torch = _torch()
        img = self._check(img, dtypes=(torch.uint8, torch.uint16))
        n, h, w = self._nhw(img)
        wpr = (w + 31) // 32
        shape = (h, wpr) if img.dim() == 2 else (n, h, wpr)
        bits = torch.empty(shape, dtype=torch.int32, device=self.device)
        if mask_thresh is None:
            self._call("yam_adaptive_threshold_bits", self._p(img), self._p(bits), n, h, w, _dtype_code(img),
                       int(block_size), float(C_))
            return bits
        mask = torch.empty_like(img)
        self._call("yam_adaptive_threshold_bits_mask", self._p(img), self._p(bits), n, h, w, _dtype_code(img),
                   int(block_size), float(C_), self._p(mask_thresh), self._p(mask), float(maxval))
        return bits, mask

    def bits_morph(self, bits, width: int, op: int, kernel_size: int = 3, iterations: int = 1):
        """Rectangular erode / dilate / open / close / open+close (op = MORPH_* or 4) on packed bits."""
        torch = _torch()
        bits = self._check(bits, dtypes=(torch.int32,), name="bits")
        n, h, wpr = self._nhw(bits)
        if wpr != (int(width) + 31) // 32:
            raise ValueError("bits tensor does not match the image width")
        if int(iterations) == 0:
            return bits.clone()
        out = torch.empty_like(bits)
        self._call("yam_bits_morph", self._p(bits), self._p(out), n, h, int(width), int(op), int(kernel_size),
                   int(iterations))
        return out

    def bits_unpack(self, bits, width: int):
        torch = _torch()
        bits = self._check(bits, dtypes=(torch.int32,), name="bits")
        n, h, wpr = self._nhw(bits)
        shape = (h, int(width)) if bits.dim() == 2 else (n, h, int(width))
        mask = torch.empty(shape, dtype=torch.uint8, device=self.device)
        self._call("yam_bits_unpack", self._p(bits), self._p(mask), n, h, int(width))
        return mask

    def ccl_label_bits(self, bits, width: int):
        torch = _torch()
        bits = self._check(bits, dtypes=(torch.int32,), name="bits")
        n, h, wpr = self._nhw(bits)
        shape = (h, int(width)) if bits.dim() == 2 else (n, h, int(width))
        labels = torch.empty(shape, dtype=torch.int32, device=self.device)
        counts = torch.empty((n,), dtype=torch.int32, device=self.device)
        self._call("yam_ccl_label_bits", self._p(bits), self._p(labels), n, h, int(width), self._p(counts), None)
        return labels, counts

    def ccl_resolve_bits(self, bits, width: int):
        """First half of ``ccl_label_bits``: returns ``(workspace, counts)``; ``ccl_emit`` writes labels
        from the workspace, optionally renumbered through a remap table (cross-strip merge)."""
        torch = _torch()
        bits = self._check(bits, dtypes=(torch.int32,), name="bits")
        n, h, wpr = self._nhw(bits)
        if wpr != (int(width) + 31) // 32:
            raise ValueError("bits tensor does not match the image width")
        nbytes = int(self.lib.yam_ccl_workspace_bytes(self._ctx, n, h, int(width)))
        if nbytes < 0:
            _lib.check("yam_ccl_workspace_bytes", -1)
        workspace = torch.empty((nbytes,), dtype=torch.uint8, device=self.device)
        counts = torch.empty((n,), dtype=torch.int32, device=self.device)
        self._call("yam_ccl_resolve_bits", self._p(bits), n, h, int(width), self._p(workspace), self._p(counts))
        return workspace, counts

    def ccl_emit(self, bits, width: int, workspace, remap=None, rows: Optional[Tuple[int, int]] = None, out=None):
        """Labels of rows ``[rows[0], rows[1])`` (default: all) of a resolved mask; ``remap`` (int32,
        entry 0 = 0) maps strip-local to global labels while they are written.  ``out``: a contiguous
        int32 CUDA tensor with room for the rows (e.g. a slice of an all-gather send buffer)."""
        torch = _torch()
        bits = self._check(bits, dtypes=(torch.int32,), name="bits")
        n, h, wpr = self._nhw(bits)
        r0, r1 = (0, n * h) if rows is None else (int(rows[0]), int(rows[1]))
        if rows is None:
            shape = (h, int(width)) if bits.dim() == 2 else (n, h, int(width))
        else:
            shape = (r1 - r0, int(width))
        if out is not None:
            if out.dtype != torch.int32 or out.device != self.device or not out.is_contiguous() or \
                    out.numel() < (r1 - r0) * int(width):
                raise ValueError("ccl_emit: `out` must be a contiguous int32 CUDA tensor holding the rows")
            labels = out
        else:
            labels = torch.empty(shape, dtype=torch.int32, device=self.device)
        rptr = None
        if remap is not None:
            remap = self._check(remap, ndim=(1,), dtypes=(torch.int32,), name="remap")
            rptr = self._p(remap)
        if r1 > r0:
            # the table length travels with the table: a local label beyond it maps to 0 instead of reading past
            # the allocation (tables sized from bounds whose overflow is only detected afterwards)
            self._call("yam_ccl_emit_rows_bounded", self._p(bits), n, h, int(width), self._p(workspace), rptr,
                       int(remap.numel()) if remap is not None else 0, r0, r1, self._p(labels))
        return labels

    def segment_fused(self, img, block_size: int = 11, C_: float = 2.0, morph_ksize: int = 5, iterations: int = 1,
                      mask_thresh=None, maxval: float = 255.0):
        """adaptive threshold -> open -> close (rectangular) -> connected components, the mask never
        leaving its 1-bit-per-pixel form.  Same labels as the unfused chain.  ``mask_thresh`` (int32[n] on the
        device): the first kernel also writes the global threshold mask of ``img`` (see
        ``adaptive_threshold_bits``); the result is then ``(labels, counts, mask)``."""
        w = int(img.shape[-1])
        if mask_thresh is None:
            bits = self.adaptive_threshold_bits(img, block_size, C_)
            return self.ccl_label_bits(self.bits_morph(bits, w, 4, morph_ksize, iterations), w)
        bits, mask = self.adaptive_threshold_bits(img, block_size, C_, mask_thresh=mask_thresh, maxval=maxval)
        labels, counts = self.ccl_label_bits(self.bits_morph(bits, w, 4, morph_ksize, iterations), w)
        return labels, counts, mask

    # ------------------------------------------------------------------ K10 / K11
    def ccl_label(self, mask):
        """Returns (labels int32 like mask, counts int32[n] on device)."""
        torch = _torch()
        mask = self._check(mask, dtypes=(torch.uint8,), name="mask")
        n, h, w = self._nhw(mask)
        labels = torch.empty(mask.shape, dtype=torch.int32, device=self.device)
        counts = torch.empty((n,), dtype=torch.int32, device=self.device)
        self._call("yam_ccl_label", self._p(mask), self._p(labels), n, h, w, self._p(counts), None)
        return labels, counts

    def region_props(self, labels, intensity=None, n_labels: Optional[int] = None):
        """Per-label accumulators int64[n_labels, 8] on device (see include/yamb200.h)."""
        torch = _torch()
        labels = self._check(labels, ndim=(2,), dtypes=(torch.int32,), name="labels")
        h, w = int(labels.shape[0]), int(labels.shape[1])
        if n_labels is None:
            n_labels = int(labels.max().item())
        idt = 0
        iptr = None
        if intensity is not None:
            intensity = self._check(intensity, ndim=(2,), dtypes=(torch.uint8, torch.uint16), name="intensity")
            if tuple(intensity.shape) != (h, w):
                raise ValueError("intensity image must match the label image shape")
            idt = _dtype_code(intensity)
            iptr = self._p(intensity)
        props = torch.empty((int(n_labels), PROPS_STRIDE), dtype=torch.int64, device=self.device)
        if n_labels > 0:
            self._call("yam_region_props", self._p(labels), iptr, idt, h, w, int(n_labels), self._p(props))
        return props


    def region_props_stack(self, labels, intensity=None, counts=None):
        """Region tables of a labelled stack (n, h, w) in one launch.  Returns (props int64[total, 8] on
        device, offsets int64[n + 1] on host): frame f owns rows offsets[f]:offsets[f + 1]."""
        torch = _torch()
        labels = self._check(labels, ndim=(3,), dtypes=(torch.int32,), name="labels")
        n, h, w = self._nhw(labels)
        if counts is None:
            counts = labels.reshape(n, -1).amax(dim=1)
        cnt = counts.detach().cpu().numpy().astype(np.int64) if hasattr(counts, "detach") else np.asarray(counts, np.int64)
        offsets = np.zeros(n + 1, np.int64)
        np.cumsum(cnt, out=offsets[1:])
        total = int(offsets[-1])
        idt = 0
        iptr = None
        if intensity is not None:
            intensity = self._check(intensity, ndim=(3,), dtypes=(torch.uint8, torch.uint16), name="intensity")
            if tuple(intensity.shape) != (n, h, w):
                raise ValueError("intensity stack must match the label stack shape")
            idt = _dtype_code(intensity)
            iptr = self._p(intensity)
        props = torch.empty((total, PROPS_STRIDE), dtype=torch.int64, device=self.device)
        if total > 0:
            offs_dev = torch.from_numpy(offsets).to(self.device)
            self._call("yam_region_props_stack", self._p(labels), iptr, idt, n, h, w, self._p(offs_dev), total,
                       self._p(props))
        return props, offsets


_default: Dict[int, Backend] = {}
_default_lock = threading.Lock()


def get_backend(device: int = 0) -> Backend:
    """Process-wide Backend per device (created on first use; raises if no GPU / no library)."""
    with _default_lock:
        be = _default.get(device)
        if be is None:
            be = Backend(device)
            _default[device] = be
        return be


def props_table(props: np.ndarray) -> Dict[str, np.ndarray]:
    """Turn the int64[n,8] accumulators into skimage-style columns (core/extraction.py:70-87)."""
    area = props[:, 0]
    safe = np.maximum(area, 1).astype(np.float64)
    return {
        "area": area.copy(),
        "sum_y": props[:, 1].copy(),
        "sum_x": props[:, 2].copy(),
        "sum_intensity": props[:, 3].copy(),
        "centroid_row": props[:, 1] / safe,
        "centroid_col": props[:, 2] / safe,
        "mean_intensity": props[:, 3] / safe,
        "bbox": props[:, 4:8].copy(),
    }


def shape_columns(props: np.ndarray, moments: np.ndarray) -> Dict[str, np.ndarray]:
    """skimage.measure.regionprops columns that follow from the first- and second-order sums
    (core/extraction.py:81-85): ``extent`` = area / bbox area; inertia tensor of the central moments
    mu[p,q] = sum (r - rbar)^p (c - cbar)^q (skimage: [[mu02, -mu11], [-mu11, mu20]] / mu00), its
    eigenvalues l1 >= l2, ``eccentricity`` = sqrt(1 - l2 / l1) and ``orientation`` =
    0.5 * atan2(-2 b, c - a) with (a, b, b, c) the tensor (pi/4 conventions when a == c).  float64 on
    the host from exact integer sums.  (``perimeter`` and ``solidity`` need contour / hull geometry:
    ``contour_columns``.)"""
    area = props[:, 0].astype(np.float64)
    safe = np.maximum(area, 1.0)
    sr, sc = props[:, 1].astype(np.float64), props[:, 2].astype(np.float64)
    srr, scc, src_ = (moments[:, i].astype(np.float64) for i in range(3))
    mu20 = srr - sr * sr / safe          # rows
    mu02 = scc - sc * sc / safe          # columns
    mu11 = src_ - sr * sc / safe
    a, b, c = mu02 / safe, -mu11 / safe, mu20 / safe
    common = np.sqrt(np.maximum(((a - c) / 2.0) ** 2 + b * b, 0.0))
    l1, l2 = (a + c) / 2.0 + common, (a + c) / 2.0 - common
    with np.errstate(divide="ignore", invalid="ignore"):
        ecc = np.where(l1 > 0, np.sqrt(np.maximum(1.0 - l2 / np.where(l1 > 0, l1, 1.0), 0.0)), 0.0)
    orient = np.where(a - c == 0, np.where(b < 0, np.pi / 4.0, -np.pi / 4.0), 0.5 * np.arctan2(-2.0 * b, c - a))
    bbox_area = ((props[:, 6] - props[:, 4]) * (props[:, 7] - props[:, 5])).astype(np.float64)
    return {"extent": area / np.maximum(bbox_area, 1.0), "eccentricity": ecc, "orientation": orient,
            "inertia_tensor_eigvals": np.stack([l1, l2], axis=1)}


def contour_columns(props: np.ndarray, perimeter_counts: np.ndarray, convex_area: np.ndarray) -> Dict[str, np.ndarray]:
    """``perimeter`` and ``solidity`` (core/extraction.py:80,83) from the exact integer tables of
    ``yam_region_perimeter`` / ``yam_region_convex_area``: perimeter = n1 + n2 * sqrt 2 + n3 * (1 + sqrt 2) / 2
    (skimage.measure.perimeter's class weights), solidity = area / area_convex; float64 on the host."""
    c = perimeter_counts.astype(np.float64)
    perimeter = c[:, 0] + c[:, 1] * np.sqrt(2.0) + c[:, 2] * ((1.0 + np.sqrt(2.0)) / 2.0)
    area = props[:, 0].astype(np.float64)
    cva = convex_area.astype(np.int64)
    return {"perimeter": perimeter, "area_convex": cva.copy(), "solidity": area / np.maximum(cva, 1).astype(np.float64)}


__all__ = ["Backend", "get_backend", "props_table", "shape_columns", "contour_columns", "shape_code", "BackendUnavailable", "YamError"]
