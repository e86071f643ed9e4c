"""Multi-GPU partitioning of the hot path (one process per GPU, ``torch.distributed``).

* Time-lapse batches / file batches: frames are independent units (the reference itself treats
  planes independently, ``processing/pipeline_manager.py:475-492``, and fans files out per process,
  ``ui/segmentation.py:2519-2536``) -> contiguous frame blocks per rank, NO data-path collective;
  only the per-frame region tables are gathered at the end.
* Mosaic row strips (65536^2): contiguous row strips per rank with a halo of ``halo`` rows read
  from the shared source (the memmap) so neighbourhood operators see the same pixels as the dense
  run; global statistics (Otsu histogram, min/max) are all-reduced.

The helpers are backend-agnostic (gloo on CPU in the tests, nccl on the GPU box).
"""
from __future__ import annotations

from typing import Any, List, Sequence, Tuple

import numpy as np


def frame_block(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [start, stop) of frames for ``rank`` (sizes differ by at most one)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def row_strip(height: int, rank: int, world: int, halo: int = 0, align: int = 1) -> Tuple[int, int, int, int]:
    """(core_start, core_stop, read_start, read_stop) rows of rank's strip; core rows are a multiple
    of ``align`` (e.g. the CLAHE tile height) except possibly the last strip."""
    units = (height + align - 1) // align
    u0, u1 = frame_block(units, rank, world)
    c0, c1 = min(u0 * align, height), min(u1 * align, height)
    return c0, c1, max(0, c0 - halo), min(height, c1 + halo)


def allreduce_histogram(hist: Any, group=None):
    """Sum per-rank histograms (int64 tensor) across ranks — the one collective Otsu needs."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist


def gather_tables(local: Sequence[Any], group=None) -> List[Any]:
    """Gather per-frame result objects from every rank in rank order (rank 0 gets the full list)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return list(local)
    world = dist.get_world_size(group)
    bucket: List[Any] = [None] * world
    dist.all_gather_object(bucket, list(local), group=group)
    out: List[Any] = []
    for part in bucket:
        out.extend(part)
    return out


def merge_label_strips(strips: Sequence[np.ndarray], counts: Sequence[int]) -> Tuple[np.ndarray, int]:
    """Host-side cross-strip label merge (CPU reference for the multi-GPU CCL path).

    ``strips`` are per-strip canonical label images (1..counts[i]) of vertically adjacent row
    strips.  Labels are offset to be globally unique, components touching across a strip boundary
    (8-connectivity) are united, and the result is renumbered in raster-first order.
    """
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    total = int(offs[-1])
    parent = np.arange(total + 1, dtype=np.int64)

    def find(x: int) -> int:
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    glob = [np.where(s > 0, s.astype(np.int64) + offs[i], 0) for i, s in enumerate(strips)]
    for i in range(len(glob) - 1):
        up, down = glob[i][-1], glob[i + 1][0]
        w = up.shape[0]
        for dx in (-1, 0, 1):
            a = up[max(0, -dx): w - max(0, dx)]
            b = down[max(0, dx): w - max(0, -dx)]
            both = (a > 0) & (b > 0)
            for la, lb in set(zip(a[both].tolist(), b[both].tolist())):
                ra, rb = find(la), find(lb)
                if ra != rb:
                    parent[max(ra, rb)] = min(ra, rb)
    full = np.concatenate(glob, axis=0)
    roots = np.array([find(i) for i in range(total + 1)], dtype=np.int64)
    rooted = roots[full]
    flat = rooted.ravel()
    nz = np.nonzero(flat)[0]
    if nz.size == 0:
        return np.zeros(full.shape, np.int32), 0
    uniq, first = np.unique(flat[nz], return_index=True)
    order = np.argsort(nz[first], kind="stable")
    remap = np.zeros(total + 1, np.int32)
    remap[uniq[order]] = np.arange(1, len(uniq) + 1, dtype=np.int32)
    return remap[rooted].astype(np.int32), int(len(uniq))




def boundary_roots(tops: Sequence[np.ndarray], bottoms: Sequence[np.ndarray], counts: Sequence[int]):
    """Union the components that touch across strip boundaries (8-connectivity).

    Works on the boundary rows only.  Labels are made globally unique as ``offs[r] + local``;
    returns ``(involved, roots, offs)``: the sorted global ids that take part in any cross-strip
    pair and, for each, the smallest global id of its merged set.  Every other label is its own root.
    Vectorised minimum-label propagation with pointer jumping over the (few) involved labels.
    """
    world = len(counts)
    offs = np.concatenate([[0], np.cumsum(np.asarray(counts, dtype=np.int64))])
    pair_list = []
    for r in range(world - 1):
        up = np.asarray(bottoms[r]).astype(np.int64)
        down = np.asarray(tops[r + 1]).astype(np.int64)
        w = up.shape[0]
        for dx in (-1, 0, 1):
            a = up[max(0, -dx): w - max(0, dx)]
            b = down[max(0, dx): w - max(0, -dx)]
            both = (a > 0) & (b > 0)
            if both.any():
                pair_list.append(np.stack([a[both] + offs[r], b[both] + offs[r + 1]], axis=1))
    if not pair_list:
        return np.zeros(0, np.int64), np.zeros(0, np.int64), offs
    pairs = np.unique(np.concatenate(pair_list, axis=0), axis=0)
    involved, inv = np.unique(pairs.ravel(), return_inverse=True)
    pa, pb = inv.reshape(-1, 2)[:, 0], inv.reshape(-1, 2)[:, 1]
    root = np.arange(involved.size, dtype=np.int64)  # compact ids are ordered like the global ids
    while True:
        m = np.minimum(root[pa], root[pb])
        new = root.copy()
        np.minimum.at(new, pa, m)
        np.minimum.at(new, pb, m)
        new = new[new]
        if np.array_equal(new, root):
            break
        root = new
    while True:
        nxt = root[root]
        if np.array_equal(nxt, root):
            break
        root = nxt
    return involved, involved[root], offs


def boundary_remaps(tops: Sequence[np.ndarray], bottoms: Sequence[np.ndarray], counts: Sequence[int]):
    """Host version of the cross-strip label merge: ``(remaps, total)`` with ``remaps[r]`` an int32
    table of size counts[r]+1 mapping local to global labels (entry 0 stays 0).  Global labels are
    assigned in raster-first order: a merged component keeps the position of its part in the
    earliest strip, whose per-strip label order already is raster order."""
    involved, roots, offs = boundary_roots(tops, bottoms, counts)
    total_local = int(offs[-1])
    root = np.arange(total_local + 1, dtype=np.int64)
    root[involved] = roots
    is_root = root == np.arange(total_local + 1)
    is_root[0] = False
    rank = np.cumsum(is_root)
    glob = rank[root].astype(np.int32)
    glob[0] = 0
    remaps = [np.concatenate([[0], glob[offs[r] + 1: offs[r + 1] + 1]]).astype(np.int32) for r in range(len(counts))]
    return remaps, int(is_root.sum())


__all__ = ["allreduce_histogram", "boundary_remaps", "boundary_roots", "frame_block", "gather_tables", "merge_label_strips", "row_strip"]
