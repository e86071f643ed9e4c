"""Lazy image handles and tile iteration (mirror of the reference's tiling types).

Reference: ``core/tiled_image.py`` (``TileBox:12``, ``_iter_tile_boxes:15``,
``TiledImageRecord:52``) and ``processing/tiled_records.py``
(``TiledPipelineImage:15``).  Boxes are ``(left, top, right, bottom)``, disjoint,
row-major, clipped at the right/bottom edges.  ``.npy`` files are memory-mapped;
rasters are opened with PIL on demand and converted RGB(A) -> BGR(A) except for
modes ``F`` and ``I;16`` (the whole application is BGR).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, Dict, Iterator, Optional, Tuple

import numpy as np

TileBox = Tuple[int, int, int, int]
TileSize = Tuple[int, int]


def iter_tile_boxes(width: int, height: int, tile_size: Optional[TileSize]) -> Iterator[TileBox]:
    """Row-major disjoint boxes covering ``width x height``; ``tile_size`` is ``(w, h)``."""
    if tile_size is None:
        yield (0, 0, width, height)
        return
    tw, th = tile_size
    if tw <= 0 or th <= 0:
        raise ValueError("tile_size must contain positive integers")
    for top in range(0, height, th):
        for left in range(0, width, tw):
            yield (left, top, min(left + tw, width), min(top + th, height))


def _checked(box: TileBox, width: int, height: int) -> TileBox:
    left, top, right, bottom = box
    if not (0 <= left < right <= width and 0 <= top < bottom <= height):
        raise ValueError("box coordinates must define a region within the image bounds")
    return left, top, right, bottom


def _to_bgr(a: np.ndarray) -> np.ndarray:
    if a.ndim == 3 and a.shape[2] == 3:
        return a[..., ::-1]
    if a.ndim == 3 and a.shape[2] == 4:
        out = a.copy()
        out[..., :3] = a[..., 2::-1]
        return out
    return a


_RAW_MODES = {"F", "I;16"}


@dataclass
class TiledImageRecord:
    """Handle around on-disk pixels (``.npy`` memmap or PIL raster)."""

    path: Path
    metadata: Dict[str, Any] = field(default_factory=dict)
    mode: Optional[str] = None
    size: Optional[Tuple[int, int]] = None  # (width, height)
    shape: Optional[Tuple[int, ...]] = None
    dtype: Optional[np.dtype] = None
    _cached_array: Optional[np.ndarray] = field(default=None, init=False, repr=False)
    _image_handle: Any = field(default=None, init=False, repr=False)
    _memmap: Optional[np.memmap] = field(default=None, init=False, repr=False)

    @classmethod
    def from_npy(cls, path: Path, *, metadata: Optional[Dict[str, Any]] = None,
                 memmap: Optional[np.memmap] = None) -> "TiledImageRecord":
        mm = memmap if memmap is not None else np.load(str(path), mmap_mode="r", allow_pickle=False)
        rec = cls(path=Path(path), metadata=dict(metadata or {}), shape=tuple(mm.shape), dtype=mm.dtype)
        rec._memmap = mm
        return rec

    @classmethod
    def from_raster(cls, path: Path, *, metadata: Dict[str, Any], image: Any) -> "TiledImageRecord":
        rec = cls(path=Path(path), metadata=dict(metadata), mode=image.mode, size=image.size)
        rec._image_handle = image
        return rec

    def close(self) -> None:
        if self._image_handle is not None:
            try:
                self._image_handle.close()
            finally:
                self._image_handle = None
        if self._memmap is not None:
            backing = getattr(self._memmap, "_mmap", None)
            if backing is not None:
                backing.close()
            self._memmap = None

    def _raster(self):
        if self._image_handle is None:
            from PIL import Image

            self._image_handle = Image.open(self.path)
        return self._image_handle

    def to_array(self) -> np.ndarray:
        if self._cached_array is not None:
            return self._cached_array
        if self._memmap is not None:
            arr = np.asarray(self._memmap)
        else:
            img = self._raster()
            arr = np.array(img)
            if img.mode not in _RAW_MODES:
                arr = _to_bgr(arr)
        self._cached_array = arr
        if self.shape is None:
            self.shape = tuple(arr.shape)
        if self.dtype is None:
            self.dtype = arr.dtype
        return arr

    def _dims(self) -> Tuple[int, int]:
        if self.size is not None:
            return self.size
        if self.shape is not None and len(self.shape) >= 2:
            return int(self.shape[1]), int(self.shape[0])
        arr = self.to_array()
        if arr.ndim < 2:
            raise ValueError("Cannot infer dimensions from a 1-D array")
        self.shape = tuple(arr.shape)
        return arr.shape[1], arr.shape[0]

    def read_region(self, box: TileBox) -> np.ndarray:
        if self._memmap is not None:
            shp = self.shape or tuple(self._memmap.shape)
            if len(shp) < 2:
                raise ValueError("np.ndarray images must be at least 2-D")
            left, top, right, bottom = _checked(box, shp[1], shp[0])
            return np.asarray(self._memmap[top:bottom, left:right, ...])
        img = self._raster()
        if img.size is None:
            raise ValueError("Image size unavailable for tiled reads")
        left, top, right, bottom = _checked(box, img.size[0], img.size[1])
        region = img.crop((left, top, right, bottom))
        arr = np.array(region)
        return arr if region.mode in _RAW_MODES else _to_bgr(arr)

    def iter_tiles(self, tile_size: Optional[TileSize] = None) -> Iterator[Tuple[TileBox, np.ndarray]]:
        width, height = self._dims()
        for box in iter_tile_boxes(width, height, tile_size):
            yield box, self.read_region(box)


@dataclass
class TiledPipelineImage:
    """Lazy handle + tiling hint handed to pipeline steps."""

    handle: Any  # TiledImageRecord or anything with the same read_region/iter_tiles/to_array surface
    tile_size: Optional[TileSize] = None
    shape: Optional[Tuple[int, ...]] = field(default=None, repr=False)

    def close(self) -> None:
        self.handle.close()

    def infer_shape(self) -> Tuple[int, ...]:
        if self.shape is None:
            hshape = getattr(self.handle, "shape", None)
            hsize = getattr(self.handle, "size", None)
            if hshape is not None:
                self.shape = tuple(hshape)
            elif hsize is not None:
                self.shape = (int(hsize[1]), int(hsize[0]))
            else:
                self.shape = tuple(self.handle.to_array().shape)
        return self.shape

    def iter_tiles(self, tile_size: Optional[TileSize] = None) -> Iterator[Tuple[TileBox, np.ndarray]]:
        yield from self.handle.iter_tiles(tile_size if tile_size is not None else self.tile_size)

    def read_region(self, box: TileBox) -> np.ndarray:
        return self.handle.read_region(box)

    def to_array(self) -> np.ndarray:
        arr = self.handle.to_array()
        self.shape = tuple(arr.shape)
        return arr

    @property
    def dtype(self) -> Optional[np.dtype]:
        d = getattr(self.handle, "dtype", None)
        if d is not None:
            return d
        arr = self.handle.to_array()
        self.shape = tuple(arr.shape)
        return arr.dtype


__all__ = ["TileBox", "TileSize", "TiledImageRecord", "TiledPipelineImage", "iter_tile_boxes"]
