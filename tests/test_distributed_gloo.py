"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: frame blocks, row strips with halos,
histogram all-reduce (Otsu on a sharded mosaic), result gathering and cross-strip label merge."""
from __future__ import annotations

import os
import socket

import numpy as np
import pytest

from oracle import np_oracle as O
from yamimageprocessor_b200.host import sharding


def test_frame_blocks_partition_exactly():
    for n in (0, 1, 7, 1024):
        for world in (1, 2, 3, 8):
            blocks = [sharding.frame_block(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.frame_block(4, 2, 2)


def test_row_strips_cover_rows_with_halo():
    h = 1000
    for world in (1, 2, 4, 8):
        strips = [sharding.row_strip(h, r, world, halo=8, align=125) for r in range(world)]
        assert strips[0][0] == 0 and strips[-1][1] == h
        for r, (c0, c1, r0, r1) in enumerate(strips):
            assert r0 == max(0, c0 - 8) and r1 == min(h, c1 + 8) and (c0 % 125 == 0)
            if r + 1 < world:
                assert c1 == strips[r + 1][0]


def test_merge_label_strips_equals_dense_labelling(rng):
    for dens in (0.2, 0.45, 0.6):
        m = (rng.random((90, 70)) < dens).astype(np.uint8) * 255
        n_want, want = O.ccl_label(m)
        for world in (2, 3, 5):
            parts, counts = [], []
            for r in range(world):
                c0, c1, _, _ = sharding.row_strip(90, r, world)
                n, lab = O.ccl_label(m[c0:c1])
                parts.append(lab)
                counts.append(n)
            got, n_got = sharding.merge_label_strips(parts, counts)
            assert n_got == n_want and np.array_equal(got, want)


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank: int, world: int, port: int, tmpdir: str):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        mosaic = rng.integers(0, 65536, (64, 48), dtype=np.uint16)
        mosaic[:, :24] //= 3
        # --- sharded Otsu: local histogram of the core rows, all-reduce, same threshold everywhere
        c0, c1, r0, r1 = sharding.row_strip(64, rank, world, halo=5)
        hist = torch.from_numpy(O.histogram(mosaic[c0:c1]))
        sharding.allreduce_histogram(hist)
        t = O.otsu_from_hist(hist.numpy())
        assert t == O.otsu_value(mosaic)
        # --- halo rows make a neighbourhood operator on the strip equal to the dense result
        dense = O.gaussian_fixed(mosaic, 11, 0.0)
        local = O.gaussian_fixed(mosaic[r0:r1], 11, 0.0)[c0 - r0: c0 - r0 + (c1 - c0)]
        interior = slice(5 if c0 == 0 else 0, None)  # rows whose window does not cross the mosaic border reflect
        assert np.array_equal(local[interior] if r0 > 0 else local, dense[c0:c1][interior] if r0 > 0 else dense[c0:c1])
        # --- frame-sharded batch: gather per-frame results in frame order
        f0, f1 = sharding.frame_block(7, rank, world)
        mine = [(i, int(i * i)) for i in range(f0, f1)]
        gathered = sharding.gather_tables(mine)
        assert gathered == [(i, i * i) for i in range(7)]
        # --- the mosaic's collectives (host/mosaic.py TorchComm): stacked all-gather of byte views
        # (uint16 LUT rows, int32 boundary rows + count) and the int64 histogram all-reduce
        from yamimageprocessor_b200.host.mosaic import TorchComm

        comm = TorchComm()
        assert (comm.world, comm.rank) == (world, rank)
        lut = torch.from_numpy((np.arange(12, dtype=np.uint16).reshape(1, 3, 4) + 100 * rank))
        got = comm.all_gather(lut)
        assert tuple(got.shape) == (world, 1, 3, 4) and got.dtype == torch.uint16
        for r in range(world):
            assert np.array_equal(got[r].numpy(), np.arange(12, dtype=np.uint16).reshape(1, 3, 4) + 100 * r)
        pack = torch.arange(9, dtype=torch.int32) * (rank + 1)
        assert np.array_equal(comm.all_gather(pack).numpy(), np.stack([np.arange(9) * (r + 1) for r in range(world)]))
        h2 = torch.full((5,), rank + 1, dtype=torch.int64)
        assert comm.all_reduce(h2, "sum").tolist() == [sum(range(1, world + 1))] * 5
        # --- bench.py's `check` block: per-strip content checksums (index base = first row x width) are added by
        # an int64 all-reduce that wraps mod 2^64, and equal the checksum of the dense array
        strip_sum = O.checksum64(mosaic[c0:c1], index_base=c0 * mosaic.shape[1])
        as_i64 = strip_sum - (1 << 64) if strip_sum >= (1 << 63) else strip_sum
        sums = torch.tensor([as_i64, (1 << 63) - 1 - rank], dtype=torch.int64)      # second entry: forces the wrap
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        assert int(sums[0].item()) & ((1 << 64) - 1) == O.checksum64(mosaic)
        assert int(sums[1].item()) & ((1 << 64) - 1) == sum((1 << 63) - 1 - r for r in range(world)) & ((1 << 64) - 1)
        np.save(os.path.join(tmpdir, f"ok{rank}.npy"), np.array([t]))
    finally:
        dist.destroy_process_group()


def test_gloo_world_size_2(tmp_path):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0.npy").exists() and (tmp_path / "ok1.npy").exists()
