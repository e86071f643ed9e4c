"""Row-strip sharded mosaic pipeline (BASELINE config 4: preprocess + segment a huge uint16 mosaic).

One process per GPU (``torch.distributed``).  Rank ``r`` owns a contiguous strip of whole CLAHE
tile rows and produces, for its rows, exactly the pixels the dense single-GPU run produces
(the reference's dense path, ``PipelineManager.apply(ndarray)``, is the parity target — its tiled
path has no halo, ``SURVEY.md`` §0 fact 5):

  source rows (+ halo)  --Gaussian k-->  g
  g core rows           --tile histograms -> LUTs-->   all-gather LUTs (tiles x 128 KiB)
  g (+ halo)            --CLAHE apply with GLOBAL geometry-->  c
  c core rows           --histogram--> all-reduce (65536 x int64)  --Otsu scan-->  t, Otsu mask
  c (+ halo)            --adaptive threshold -> open -> close-->  packed 1-bit mask (cropped to the core)
  mask core             --CCL resolve-->  per-strip roots; boundary label rows all-gathered,
                          equivalences united on the device, then the label image is written ONCE in
                          global raster-first numbering (CCL emit through the remap table)
  labels, c             --region props--> partial tables, all-reduced (sum / min / max)

Halo rows are over-fetched from the shared source once (Gaussian r + adaptive r + 4 morphology r
rows) and recomputed locally, so no mid-pipeline halo exchange is needed; the only collectives are
the LUT all-gather, the histogram all-reduce, the boundary-row all-gather and the table all-reduce
(all tiny: latency-, not bandwidth-bound).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Optional, Sequence, Tuple

import numpy as np

from ..backend import Backend
from . import ingest, sharding


_CCL_MAX_PX = 1 << 30   # pixels per labelling call (32-bit pixel indices inside yam_ccl_*)


@dataclass
class MosaicParams:
    gauss_ksize: int = 11
    clip_limit: float = 2.0
    tile_grid: Tuple[int, int] = (8, 8)
    block_size: int = 11
    C: float = 2.0
    morph_ksize: int = 5
    ccl_max_px: int = _CCL_MAX_PX  # pixels per labelling call; larger strips are labelled as merged sub-strips
    trace: Any = None              # optional callable(phase_name, rank): profiling hook, called at phase ends
    reuse_count_bounds: bool = True  # size the merge tables from the previous run over a same-shaped source (no host wait)
    count_bound_slack: Tuple[float, int] = (0.25, 1024)   # bound = count * (1 + slack[0]) + slack[1] per (sub-)strip


@dataclass
class StripResult:
    rows: Tuple[int, int]          # [start, stop) rows of the mosaic owned by this rank
    clahe: Any                     # device tensors for the core rows
    otsu_mask: Any
    labels: Any
    otsu_threshold: int
    n_components: int              # global count
    props: Optional[Any] = None    # int64 [n_components, 8] (global, reduced) when requested


class TorchComm:
    """Collectives over torch.distributed (NCCL on the GPU box)."""

    def __init__(self, group=None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def all_gather(self, tensor):
        """[world, *tensor.shape]: one ncclAllGather straight into the result (no per-rank list copies)."""
        import torch

        # integer payloads travel as bytes: NCCL does not take every unsigned dtype
        t = tensor.contiguous()
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out.view(torch.uint8).reshape(-1), t.view(torch.uint8).reshape(-1),
                                         group=self.group)
        return out

    def all_reduce(self, tensor, op: str = "sum"):
        ops = {"sum": self.dist.ReduceOp.SUM, "min": self.dist.ReduceOp.MIN, "max": self.dist.ReduceOp.MAX}
        self.dist.all_reduce(tensor, op=ops[op], group=self.group)
        return tensor

    def once(self, fn):
        """Host work whose inputs are identical on every strip of this process: run it once."""
        return fn()


# per (source shape, sharding, segmentation parameters): upper bounds of the component counts of every (sub-)strip,
# learnt from the previous run over such a source (MosaicParams.reuse_count_bounds)
_COUNT_BOUNDS: dict = {}


class LocalComm:
    """In-process stand-in: `world` threads, one per emulated rank, meet at a barrier.

    Lets the whole sharded pipeline (halo over-fetch, LUT gather, histogram reduce, label merge)
    run against the dense result on a single GPU; the collectives are plain tensor ops."""

    def __init__(self, world: int):
        import threading

        self.world = world
        self._slots = [None] * world
        self._barrier = threading.Barrier(world)

    def bind(self, rank: int) -> "_BoundLocalComm":
        return _BoundLocalComm(self, rank)


class _BoundLocalComm:
    def __init__(self, parent: LocalComm, rank: int):
        self.parent, self.rank, self.world = parent, rank, parent.world

    def all_gather(self, tensor):
        import torch

        p = self.parent
        p._slots[self.rank] = tensor
        p._barrier.wait()
        out = torch.stack(list(p._slots))
        p._barrier.wait()
        return out

    def all_reduce(self, tensor, op: str = "sum"):
        stacked = self.all_gather(tensor)
        red = stacked.sum(0) if op == "sum" else (stacked.min(0).values if op == "min" else stacked.max(0).values)
        tensor.copy_(red)
        return tensor

    def once(self, fn):
        p = self.parent
        if self.rank == 0:
            p._shared = fn()
        p._barrier.wait()
        out = p._shared
        p._barrier.wait()
        return out


class HybridComm:
    """`local` strip-threads in this process x `nproc` processes (torch.distributed): lets one GPU
    own several strips (a 65536^2 mosaic on 1, 2, 4 or 8 GPUs always runs as 8 strips of whole CLAHE
    tile rows; with fewer GPUs each GPU works through several strips, in lock-step threads).
    Global strip index = process_rank * local + local_index."""

    def __init__(self, local: int, use_dist: bool):
        import threading

        self.local = local
        self.dist = None
        self.nproc = 1
        self.prank = 0
        if use_dist:
            import torch.distributed as dist

            self.dist = dist
            self.nproc = dist.get_world_size()
            self.prank = dist.get_rank()
        self.world = self.local * self.nproc
        self._slots = [None] * local
        self._result = None
        self._barrier = threading.Barrier(local)

    def bind(self, local_index: int) -> "_BoundHybridComm":
        return _BoundHybridComm(self, local_index)


class _BoundHybridComm:
    def __init__(self, parent: HybridComm, li: int):
        self.parent, self.li = parent, li
        self.world = parent.world
        self.rank = parent.prank * parent.local + li

    def all_gather(self, tensor):
        import torch

        p = self.parent
        p._slots[self.li] = tensor.contiguous()
        p._barrier.wait()
        if self.li == 0:
            local = torch.stack(p._slots)  # (local, ...)
            if p.dist is not None:
                everything = torch.empty((p.nproc * p.local,) + tuple(local.shape[1:]), dtype=local.dtype,
                                         device=local.device)
                p.dist.all_gather_into_tensor(everything.view(torch.uint8).reshape(-1), local.view(torch.uint8).reshape(-1))
            else:
                everything = local
            p._result = everything
        p._barrier.wait()
        out = p._result
        p._barrier.wait()
        return out

    def all_reduce(self, tensor, op: str = "sum"):
        stacked = self.all_gather(tensor)
        red = stacked.sum(0) if op == "sum" else (stacked.min(0).values if op == "min" else stacked.max(0).values)
        tensor.copy_(red)
        return tensor

    def once(self, fn):
        p = self.parent
        if self.li == 0:
            p._shared = fn()
        p._barrier.wait()
        out = p._shared
        p._barrier.wait()
        return out


def strip_rows(height: int, tiles_y: int, rank: int, world: int) -> Tuple[int, int]:
    if height % tiles_y:
        raise ValueError("mosaic height must be divisible by the CLAHE grid for strip sharding")
    if tiles_y % world:
        raise ValueError(f"CLAHE tile rows ({tiles_y}) must be divisible by the number of ranks ({world})")
    th = height // tiles_y
    per = tiles_y // world
    return rank * per * th, (rank + 1) * per * th


def run_strip(be: Backend, source, rank: int = 0, world: int = 1, params: Optional[MosaicParams] = None,
              with_props: bool = False, device_source=None, comm=None) -> StripResult:
    """Process this rank's strip of ``source`` (an (H, W) uint16 array / memmap).

    ``device_source`` may pass the already-uploaded rows [r0, r1) (benchmarks keep the input
    resident); otherwise the rows are read from ``source`` and uploaded.
    """
    import torch

    p = params or MosaicParams()
    if world > 1 and comm is None:
        comm = TorchComm()
    if world == 1:
        comm = None
    mark = (lambda name: p.trace(name, rank)) if p.trace is not None else (lambda name: None)
    H, W = int(source.shape[0]), int(source.shape[1])
    tiles_x, tiles_y = p.tile_grid
    if W % tiles_x:
        raise ValueError("mosaic width must be divisible by the CLAHE grid for strip sharding")
    c0, c1 = strip_rows(H, tiles_y, rank, world)
    th, tw = H // tiles_y, W // tiles_x
    hg = p.gauss_ksize // 2
    halo_seg = p.block_size // 2 + 4 * (p.morph_ksize // 2)   # adaptive + erode,dilate,dilate,erode
    a0, a1 = max(0, c0 - halo_seg), min(H, c1 + halo_seg)      # rows of CLAHE output needed
    r0, r1 = max(0, a0 - hg), min(H, a1 + hg)                  # rows of input needed
    # The labeller indexes pixels with 32 bits, so a strip of 2^31 pixels or more (65536^2 on one or two
    # GPUs) is resolved as `k_sub` equal sub-strips that the cross-strip merge below stitches exactly like
    # strips of different ranks.
    rows_core = c1 - c0
    k_sub = max(1, -(-(rows_core * W) // int(p.ccl_max_px)))
    while rows_core % k_sub:
        k_sub += 1
    sub_rows = rows_core // k_sub
    # Learnt bounds of the component counts (see the merge below) are read HERE, before any collective: every
    # strip of this step -- other processes, or the lock-step threads of an emulated run, which share the table --
    # then sizes its tables from the same previous step, whichever strip finishes first and updates the table.
    hint_key = (H, W, world, k_sub, p.block_size, p.morph_ksize)
    bounds = _COUNT_BOUNDS.get(hint_key) if p.reuse_count_bounds else None

    # rows are streamed from the (memmap) source through a pinned ring, gather overlapping the DMA
    import os
    import time

    _t0 = time.perf_counter()
    x = device_source if device_source is not None else ingest.upload_rows(be, source, r0, r1)
    if os.environ.get("YAM_E2E_TRACE") and device_source is None:
        import sys

        torch.cuda.synchronize()
        sys.stderr.write(f"[e2e] rank {rank}: upload_rows {(r1 - r0) * W * 2 / 1e9:.2f} GB in {1e3 * (time.perf_counter() - _t0):.1f} ms\n")
    g = be.gaussian(x, p.gauss_ksize, 0.0)                     # exact on [a0, a1): artificial edges are hg rows away
    mark("gaussian")

    # CLAHE: LUTs of the tile rows this rank owns, gathered from every rank
    luts_local = be.clahe_luts(g[c0 - r0: c1 - r0], p.clip_limit, (tiles_x, tiles_y // world))
    mark("clahe_luts")
    if comm is not None:  # [world, tile rows per strip, tiles_x, bins] is the global LUT table, already in order
        luts = comm.all_gather(luts_local).reshape(tiles_y, tiles_x, -1)
        mark("lut_all_gather")
    else:
        luts = luts_local
    c = be.clahe_apply(g[a0 - r0: a1 - r0], luts, (tw, th), y_offset=a0)
    c_core = c[c0 - a0: c1 - a0]
    mark("clahe")

    # Otsu on the global histogram: all-reduce, then the certified parallel scan on the device (every rank
    # scans the same histogram and gets the same threshold); nothing is read back, the host does not wait.
    hist = be.histogram(c_core)[0]
    if comm is not None:
        comm.all_reduce(hist, "sum")
    t_dev = be.otsu_from_histogram_device(hist)
    mark("histogram_all_reduce_scan")

    # segmentation on the extended rows, cropped to the core
    # (the binary mask stays 1 bit/pixel from the threshold through open/close into the labelling;
    # cropping whole rows of the packed mask is a view, not a copy)
    # (the adaptive kernel has every pixel of c staged in shared memory: it writes the Otsu mask of the same rows in
    # the same pass -- the threshold is on the device by now -- instead of a separate DRAM-bound threshold pass)
    raw_bits, otsu_ext = be.adaptive_threshold_bits(c, p.block_size, p.C, mask_thresh=t_dev, maxval=255)
    otsu_mask = otsu_ext[c0 - a0: c1 - a0]
    bits = be.bits_morph(raw_bits, W, 4, p.morph_ksize, 1)
    bits_core = bits[c0 - a0: c1 - a0]
    # labelling in two steps: resolve now, write the label image once the global numbering is known.
    subs = []
    for i in range(k_sub):
        b = bits_core[i * sub_rows:(i + 1) * sub_rows]
        ws_i, cnt_i = be.ccl_resolve_bits(b, W)
        subs.append((b, ws_i, cnt_i))

    mark("segment")

    if comm is not None or k_sub > 1:
        # one all-gather carries every (sub-)strip's first / last label row and its component count
        stride = 2 * W + 8
        pack = torch.empty((k_sub, stride), dtype=torch.int32, device=be.device)
        for i, (b, ws_i, cnt_i) in enumerate(subs):
            be.ccl_emit(b, W, ws_i, rows=(0, 1), out=pack[i, :W])
            be.ccl_emit(b, W, ws_i, rows=(sub_rows - 1, sub_rows), out=pack[i, W:2 * W])
            pack[i, 2 * W:2 * W + 1].copy_(cnt_i)
        packed = comm.all_gather(pack).reshape(-1, stride) if comm is not None else pack
        mark("boundary_all_gather")
        labels = torch.empty((rows_core, W), dtype=torch.int32, device=be.device)
        counts_view = packed[:, 2 * W]                      # one real count per (sub-)strip, on the device
        cnt_host = torch.empty((int(packed.shape[0]),), dtype=torch.int32, pin_memory=True)
        cnt_host.copy_(counts_view, non_blocking=True)
        cnt_ready = torch.cuda.Event()
        cnt_ready.record()
        # The host needs the counts only to SIZE the merge tables.  A repeated run over a source of the same
        # shape (time series of mosaics, the bench's steps) sizes them from the previous run's counts plus
        # head room and leaves the real counts on the device (yam_merge_strips_remap_bounded): no host wait in
        # the middle of the pipeline.  Every rank derives the same bounds from the same all-gathered counts.
        overflow_dev = None

        def emit_all(remaps):
            for i, (b, ws_i, cnt_i) in enumerate(subs):
                be.ccl_emit(b, W, ws_i, remap=remaps[i], out=labels[i * sub_rows:(i + 1) * sub_rows])

        if bounds is not None:
            offs = np.concatenate([[0], np.cumsum(bounds)])
            remaps, total_dev, overflow_dev = be.merge_strips_remap(packed, W, offs, rank * k_sub, k_sub, counts_dev=counts_view)
            mark("merge_remap")
            emit_all([remaps] if k_sub == 1 else remaps)
        cnt_ready.synchronize()        # (with bounds: everything is enqueued by now, this wait costs nothing extra)
        counts = cnt_host.numpy().astype(np.int64)
        if bounds is None or int(overflow_dev[0].item()) != 0:
            # first run over this shape, or a count outgrew its bound: exact offsets.
            # union of the ids that touch across (sub-)strip boundaries + raster-first renumbering: one
            # library call, small kernels, all on the device (yam_merge_strips_remap);
            # the label image is written once, already in global numbering
            offs = np.concatenate([[0], np.cumsum(counts)])
            remaps, total_dev = be.merge_strips_remap(packed, W, offs, rank * k_sub, k_sub)
            mark("merge_remap")
            emit_all([remaps] if k_sub == 1 else remaps)
        if p.reuse_count_bounds:
            _COUNT_BOUNDS[hint_key] = (counts + np.ceil(counts * float(p.count_bound_slack[0])).astype(np.int64)
                                       + int(p.count_bound_slack[1])).astype(np.int64)
    else:
        b, ws_i, total_dev = subs[0]
        labels = be.ccl_emit(b, W, ws_i)
    mark("emit")
    total = int(total_dev[0].item())   # the only wait for the labelling: everything above is enqueued
    t = int(t_dev[0].item())
    mark("merge")

    props = None
    if with_props:
        props = be.region_props(labels, c_core, total)
        if total:
            # strip-local row coordinates -> mosaic coordinates, then reduce across ranks
            area = props[:, 0]
            seen = area > 0
            props[:, 1] += area * c0
            props[:, 4] = torch.where(seen, props[:, 4] + c0, props[:, 4])
            props[:, 6] = torch.where(seen, props[:, 6] + c0, props[:, 6])
            if comm is not None:
                sums = comm.all_reduce(props[:, 0:4].contiguous(), "sum")
                mins = comm.all_reduce(props[:, 4:6].contiguous(), "min")
                maxs = comm.all_reduce(props[:, 6:8].contiguous(), "max")
                props = torch.cat([sums, mins, maxs], dim=1)
    return StripResult((c0, c1), c_core, otsu_mask, labels, t, total, props)


def input_rows(height: int, rank: int, world: int, params: Optional[MosaicParams] = None) -> Tuple[int, int]:
    """Rows [r0, r1) of the source this rank reads (core + over-fetched halo)."""
    p = params or MosaicParams()
    c0, c1 = strip_rows(height, p.tile_grid[1], rank, world)
    halo = p.block_size // 2 + 4 * (p.morph_ksize // 2) + p.gauss_ksize // 2
    return max(0, c0 - halo), min(height, c1 + halo)


def run_emulated(be: Backend, source, world: int, params: Optional[MosaicParams] = None, with_props: bool = False):
    """Run all `world` strips on ONE GPU with in-process collectives (threads); returns the list of
    StripResult in rank order.  Used by the single-GPU parity tests of the sharded path."""
    import threading

    import torch

    comm = LocalComm(world)
    results = [None] * world
    errors = []

    def work(r: int):
        try:
            torch.cuda.set_device(be.device)
            results[r] = run_strip(be, source, r, world, params, with_props, comm=comm.bind(r))
        except BaseException as exc:  # pragma: no cover - surfaced below
            errors.append(exc)
            comm._barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results


def run_local_strips(be: Backend, source, local: int, use_dist: bool, params: Optional[MosaicParams] = None,
                     with_props: bool = False, device_sources: Optional[Sequence[Any]] = None):
    """Run this process's `local` strips (lock-step threads) of a mosaic split into
    local x world_size strips; returns their StripResults in strip order."""
    import threading

    import torch

    comm = HybridComm(local, use_dist)
    results = [None] * local
    errors = []

    def work(li: int):
        try:
            torch.cuda.set_device(be.device)
            bound = comm.bind(li)
            dev = device_sources[li] if device_sources is not None else None
            results[li] = run_strip(be, source, bound.rank, comm.world, params, with_props, device_source=dev,
                                    comm=bound if comm.world > 1 else None)
        except BaseException as exc:  # pragma: no cover
            errors.append(exc)
            comm._barrier.abort()

    threads = [threading.Thread(target=work, args=(li,)) for li in range(local)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results


class RowSource:
    """(H, W) row-sliceable view of a lazy image handle for ``ingest.upload_rows``.

    Accepts a ``TiledPipelineImage`` / ``TiledImageRecord`` (the reference's, ``processing/tiled_records.py:15``
    and ``core/tiled_image.py:52``, or this package's mirrors): rows come from
    ``read_region((0, r0, W, r1))``, which is a zero-copy memmap slice for ``.npy`` sources
    (``core/tiled_image.py:137-146``) and a PIL crop otherwise.  Nothing is densified."""

    ndim = 2

    def __init__(self, image):
        self._image = image
        shape = image.infer_shape() if hasattr(image, "infer_shape") else tuple(image.shape)
        if len(shape) != 2:
            raise ValueError(f"the mosaic pipeline takes a single-channel (H, W) image, got shape {tuple(shape)}")
        self.shape = (int(shape[0]), int(shape[1]))
        dt = getattr(image, "dtype", None)
        if dt is None:
            dt = image.read_region((0, 0, 1, 1)).dtype
        self.dtype = np.dtype(dt)

    def __getitem__(self, rows):
        if not isinstance(rows, slice) or rows.step not in (None, 1):
            raise TypeError("RowSource supports contiguous row slices only")
        r0, r1, _ = rows.indices(self.shape[0])
        return self._image.read_region((0, r0, self.shape[1], r1))


def run_source(be: Backend, image, params: Optional[MosaicParams] = None, strips_per_process: int = 1,
               with_props: bool = False):
    """The sharded pipeline over a lazy handle or a dense (H, W) array: this process's strips
    (``strips_per_process`` x world_size strips in total when ``torch.distributed`` is initialised,
    else ``strips_per_process`` strips on this GPU).  Rows stream host -> pinned ring -> HBM
    (``ingest.upload_rows``).  Returns the StripResults of this process in strip order."""
    import torch.distributed as dist

    source = image if isinstance(image, np.ndarray) else RowSource(image)
    use_dist = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    return run_local_strips(be, source, int(strips_per_process), use_dist, params, with_props)


__all__ = ["HybridComm", "RowSource", "run_source", "LocalComm", "MosaicParams", "run_local_strips", "StripResult", "TorchComm", "input_rows", "run_emulated", "run_strip",
           "strip_rows"]
