"""Row-strip sharded mosaic pipeline vs the dense single-GPU run (bit-exact), with the ranks
emulated on one GPU (host/mosaic.py LocalComm): halo over-fetch, CLAHE LUT gather with global
geometry, global Otsu histogram, cross-strip label merge, reduced region tables."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest

from oracle import np_oracle as O
from yamimageprocessor_b200 import synth
from yamimageprocessor_b200.host import mosaic

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def dense(backend, frame, p):
    x = backend.to_device(frame)
    g = backend.gaussian(x, p.gauss_ksize, 0.0)
    c = backend.clahe(g, p.clip_limit, p.tile_grid)
    t, om = backend.otsu_threshold(c, 255)
    m = backend.morph_open_close(backend.adaptive_threshold(c, p.block_size, p.C), p.morph_ksize, 1)
    labels, counts = backend.ccl_label(m)
    n = int(backend.to_host(counts)[0])
    props = backend.region_props(labels, c, n)
    return (backend.to_host(c), int(backend.to_host(t)[0]), backend.to_host(om), backend.to_host(labels), n,
            backend.to_host(props))


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_sharded_mosaic_equals_dense(backend, world):
    p = mosaic.MosaicParams()
    frame = synth.nuclei(1024, 768, seed=11)
    # long vertical structures so that components cross several strip boundaries
    frame[:, 100:104] = 60000
    frame[300:800, 400:403] = 55000
    want_c, want_t, want_om, want_lab, want_n, want_props = dense(backend, frame, p)
    res = mosaic.run_emulated(backend, frame, world, p, with_props=True)
    assert [r.rows for r in res] == [mosaic.strip_rows(1024, 8, r, world) for r in range(world)]
    got_c = np.concatenate([backend.to_host(r.clahe) for r in res])
    got_om = np.concatenate([backend.to_host(r.otsu_mask) for r in res])
    got_lab = np.concatenate([backend.to_host(r.labels) for r in res])
    assert all(r.otsu_threshold == want_t for r in res)
    assert all(r.n_components == want_n for r in res)
    assert np.array_equal(got_c, want_c), "CLAHE output differs from the dense run"
    assert np.array_equal(got_om, want_om), "Otsu mask differs"
    assert np.array_equal(got_lab, want_lab), "labels differ"
    for r in res:
        assert np.array_equal(backend.to_host(r.props), want_props), "reduced region table differs"


@pytest.mark.parametrize("world", [1, 2])
def test_sharded_mosaic_with_labelling_sub_strips(backend, world):
    """Strips of >= 2^31 pixels are labelled as merged sub-strips (65536^2 on one or two GPUs); forced
    here with a tiny ccl_max_px: 1024 rows -> 4 (world 1) / 2 x 4 (world 2) sub-strips."""
    p = mosaic.MosaicParams(ccl_max_px=768 * 128)
    frame = synth.nuclei(1024, 768, seed=13)
    frame[:, 200:204] = 60000
    frame[100:900, 500:503] = 55000
    want_c, want_t, want_om, want_lab, want_n, want_props = dense(backend, frame, mosaic.MosaicParams())
    res = mosaic.run_emulated(backend, frame, world, p, with_props=True)
    assert all(r.n_components == want_n and r.otsu_threshold == want_t for r in res)
    assert np.array_equal(np.concatenate([backend.to_host(r.labels) for r in res]), want_lab)
    assert np.array_equal(backend.to_host(res[0].props), want_props)


def test_repeated_runs_size_the_merge_from_learnt_bounds(backend):
    """Second run over a same-shaped source: the merge tables are sized from the previous counts and the real
    counts stay on the device (yam_merge_strips_remap_bounded, no host wait).  Same labels as the first run and
    as the dense run, with tight and with very loose bounds; bounds that are too small overflow and the merge
    falls back to exact sizes."""
    p = mosaic.MosaicParams(count_bound_slack=(0.0, 1))
    frame = synth.nuclei(1024, 768, seed=21)
    frame[:, 300:303] = 60000
    want = dense(backend, frame, p)

    def check(res, what):
        assert all(r.n_components == want[4] for r in res), what
        assert np.array_equal(np.concatenate([backend.to_host(r.labels) for r in res]), want[3]), what

    for world in (4, 1):
        pp = mosaic.MosaicParams(ccl_max_px=768 * 256, count_bound_slack=(0.0, 1)) if world == 1 else p   # world 1: four sub-strips
        mosaic._COUNT_BOUNDS.clear()
        for attempt in range(3):                                   # exact, then bounded with bound = count + 1
            check(mosaic.run_emulated(backend, frame, world, pp), (world, "attempt", attempt))
            assert len(mosaic._COUNT_BOUNDS) == 1
        (key, learnt), = mosaic._COUNT_BOUNDS.items()
        assert learnt.shape == (4,) and int(learnt.sum()) >= want[4]
        mosaic._COUNT_BOUNDS[key] = learnt * 10 + 1000            # very loose bounds: large gaps in the id space
        check(mosaic.run_emulated(backend, frame, world, pp), (world, "loose"))
        mosaic._COUNT_BOUNDS[key] = np.maximum(learnt // 2, 1)     # too small: overflow -> exact merge
        check(mosaic.run_emulated(backend, frame, world, pp), (world, "overflow"))
        assert np.array_equal(mosaic._COUNT_BOUNDS[key], learnt)   # re-learnt from the real counts
    mosaic._COUNT_BOUNDS.clear()


def test_dense_chain_matches_oracle(backend):
    """anchor: the dense chain the sharded path is compared with equals the CPU oracle"""
    p = mosaic.MosaicParams()
    frame = synth.nuclei(512, 512, seed=12)
    c, t, om, lab, n, props = dense(backend, frame, p)
    oc = O.clahe(O.gaussian_fixed(frame, 11, 0.0), 2.0, (8, 8))
    ot, oom = O.otsu_threshold(oc, 255)
    olab = O.ccl_label(O.morph_close(O.morph_open(O.adaptive_threshold(oc, 11, 2), "Rectangular", 5, 1), "Rectangular", 5, 1))
    assert np.array_equal(c, oc) and t == ot and np.array_equal(om, oom)
    assert n == olab[0] and np.array_equal(lab, olab[1])


M64 = (1 << 64) - 1


def _u64(t):
    return int(t.item()) & M64


def test_checksum64_matches_oracle(backend, rng):
    """yam_checksum64 (bench.py's `check` block) against its NumPy statement: every dtype, an index base, and
    per-strip sums that add up to the checksum of the whole array."""
    for dt, hi in ((np.uint8, 1 << 8), (np.uint16, 1 << 16), (np.int32, 1 << 31)):
        a = rng.integers(0, hi, (37, 53)).astype(dt)
        if dt == np.int32:
            a[3, 5:9] = -1                      # bit patterns are zero-extended, not sign-extended
        t = backend.to_device(a)
        assert _u64(backend.checksum64(t)) == O.checksum64(a)
        assert _u64(backend.checksum64(t, index_base=123456789012)) == O.checksum64(a, 123456789012)
    a = rng.integers(0, 1 << 16, (64, 48)).astype(np.uint16)
    acc = backend.checksum64(backend.to_device(a[:20]))
    backend.checksum64(backend.to_device(a[20:]), index_base=20 * 48, accumulate=acc)
    assert _u64(acc) == O.checksum64(a)


@pytest.mark.parametrize("world", [1, 4])
def test_check_block_of_sharded_run_equals_dense_and_oracle(backend, world):
    """bench.py's parity evidence at reduced size: the per-strip checksums of the label image and of the Otsu
    mask (index base = first row x width) add up to the checksum of the dense single-strip run, which equals the
    oracle's checksum of the oracle's own labels.  (tools/check_bench_check_block.py does the same for whole bench
    lines against live cv2 on the CPU: profiles/r02_check_c4_*_vs_cv2.json.)"""
    p = mosaic.MosaicParams()
    frame = synth.nuclei(1024, 768, seed=17)
    frame[:, 500:503] = 60000
    W = frame.shape[1]
    _, want_t, want_om, want_lab, want_n, _ = dense(backend, frame, p)
    res = mosaic.run_emulated(backend, frame, world, p)
    sums = [0, 0]
    for r in res:
        base = r.rows[0] * W
        sums[0] = (sums[0] + _u64(backend.checksum64(r.labels, index_base=base))) & M64
        sums[1] = (sums[1] + _u64(backend.checksum64(r.otsu_mask, index_base=base))) & M64
    assert all(r.otsu_threshold == want_t and r.n_components == want_n for r in res)
    assert sums[0] == O.checksum64(want_lab) == _u64(backend.checksum64(backend.to_device(want_lab)))
    assert sums[1] == O.checksum64(want_om) == _u64(backend.checksum64(backend.to_device(want_om)))
    oc = O.clahe(O.gaussian_fixed(frame, 11, 0.0), 2.0, (8, 8))
    olab = O.ccl_label(O.morph_close(O.morph_open(O.adaptive_threshold(oc, 11, 2), "Rectangular", 5, 1), "Rectangular", 5, 1))[1]
    assert sums[0] == O.checksum64(olab) and sums[1] == O.checksum64(O.otsu_threshold(oc, 255)[1])


def test_strip_geometry_errors():
    with pytest.raises(ValueError):
        mosaic.strip_rows(1000, 8, 0, 3)
    with pytest.raises(ValueError):
        mosaic.strip_rows(1001, 8, 0, 2)
    assert mosaic.input_rows(1024, 1, 2) == (512 - 18, 1024)


def test_merge_strips_remap_matches_host_reference(backend, rng):
    """yam_merge_strips_remap (union + raster-first renumbering on the device) against the NumPy
    reference sharding.boundary_remaps and against the dense labelling, on masks whose components
    cross several strip boundaries (incl. U-shapes that join two components of an upper strip)."""
    import torch

    from yamimageprocessor_b200.host import sharding

    for dens, world, (h, w) in ((0.45, 4, (64, 70)), (0.58, 8, (96, 129)), (0.3, 2, (40, 33)), (0.62, 3, (90, 64))):
        m = (rng.random((h, w)) < dens).astype(np.uint8) * 255
        m[:, 5] = 255                                   # a column through every strip
        n_want, want = O.ccl_label(m)
        parts, counts, tops, bottoms = [], [], [], []
        for r in range(world):
            c0, c1, _, _ = sharding.row_strip(h, r, world)
            n, lab = O.ccl_label(m[c0:c1])
            parts.append(lab); counts.append(n); tops.append(lab[0]); bottoms.append(lab[-1])
        remaps_ref, total_ref = sharding.boundary_remaps(tops, bottoms, counts)
        stride = 2 * w + 8
        packed = np.zeros((world, stride), np.int32)
        for r in range(world):
            packed[r, :w], packed[r, w:2 * w], packed[r, 2 * w] = tops[r], bottoms[r], counts[r]
        offs = np.concatenate([[0], np.cumsum(counts)])
        dev = backend.to_device(packed)
        got_full = []
        for r in range(world):
            remap, total = backend.merge_strips_remap(dev, w, offs, r)
            remap = backend.to_host(remap)
            assert int(backend.to_host(total)[0]) == total_ref == n_want
            assert np.array_equal(remap, remaps_ref[r]), f"remap of strip {r} (world {world})"
            got_full.append(remap[parts[r]])
        assert np.array_equal(np.concatenate(got_full), want)


def test_mosaic_module_through_pipeline_manager_tiled_handle(backend, tmp_path):
    """BASELINE config 4's route: a .npy memmap behind a TiledPipelineImage goes through
    PipelineManager.apply -> _apply_tiled -> the supports_tiled_input step (reference
    processing/pipeline_manager.py:412-416); nothing is densified on the host, labels equal the
    dense chain, for 1 and 4 strips per process."""
    from yamimageprocessor_b200.host.pipeline import PipelineManager
    from yamimageprocessor_b200.host.tiles import TiledImageRecord, TiledPipelineImage
    from yamimageprocessor_b200.modules import b200_backend as plugin

    p = mosaic.MosaicParams()
    frame = synth.nuclei(1024, 768, seed=21)
    frame[:, 300:303] = 60000
    np.save(tmp_path / "mosaic.npy", frame)
    want = dense(backend, frame, p)

    class NoDensify(TiledImageRecord):
        def to_array(self):
            raise AssertionError("the tiled handle must not be densified")

    mm = np.load(tmp_path / "mosaic.npy", mmap_mode="r")
    rec = NoDensify.from_npy(tmp_path / "mosaic.npy", memmap=mm)
    handle = TiledPipelineImage(rec, tile_size=(256, 256))
    mod = plugin.MosaicModule()
    assert mod.supports_tiled_input() and not mod.pipeline_execution_metadata().requires_gpu
    for strips in (1, 4):
        step = mod.create_pipeline_step()
        step.enabled = True
        step.params["strips"] = strips
        assert step.supports_tiled_input
        out = PipelineManager([step]).apply(handle)
        assert out.dtype == np.int32 and out.shape == frame.shape
        assert np.array_equal(out, want[3]), f"labels differ ({strips} strips)"
        res = mod.last_results
        assert res[0].otsu_threshold == want[1] and res[0].n_components == want[4]
        assert np.array_equal(np.concatenate([backend.to_host(r.otsu_mask) for r in res]), want[2])
    # a dense ndarray takes the same step (executor route: DEVICE_STEPS["Mosaic"])
    from yamimageprocessor_b200.host.executor import B200Executor
    step = mod.create_pipeline_step(); step.enabled = True
    assert np.array_equal(B200Executor(backend).execute(step, frame), want[3])


def test_mosaic_real_nccl_two_ranks_equal_dense(backend, tmp_path):
    """The NCCL data path (TorchComm: LUT all-gather, histogram all-reduce, boundary all-gather) with
    two real ranks on two GPUs (torchrun) against the dense single-GPU run of this process."""
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    p = mosaic.MosaicParams()
    size = 2048
    frame = synth.nuclei(size, size, seed=31)
    frame[:, 700:704] = 60000
    np.save(tmp_path / "m.npy", frame)
    want = dense(backend, frame, p)
    script = tmp_path / "rank.py"
    script.write_text(
        "import os, sys, numpy as np, torch, torch.distributed as dist\n"
        f"sys.path.insert(0, {str(ROOT)!r})\n"
        "from yamimageprocessor_b200.backend import get_backend\n"
        "from yamimageprocessor_b200.host import mosaic\n"
        "lr = int(os.environ['LOCAL_RANK']); torch.cuda.set_device(lr)\n"
        "dist.init_process_group('nccl', device_id=torch.device('cuda', lr))\n"
        "be = get_backend(lr)\n"
        f"src = np.load({str(tmp_path / 'm.npy')!r}, mmap_mode='r')\n"
        "r = mosaic.run_source(be, src, mosaic.MosaicParams(), strips_per_process=1, with_props=True)[0]\n"
        f"np.savez({str(tmp_path)!r} + f'/out{{dist.get_rank()}}.npz', labels=be.to_host(r.labels), om=be.to_host(r.otsu_mask),\n"
        "         c=be.to_host(r.clahe), t=r.otsu_threshold, n=r.n_components, props=be.to_host(r.props), rows=np.array(r.rows))\n"
        "dist.barrier(); dist.destroy_process_group()\n")
    proc = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                           "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                          capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stderr[-3000:]
    outs = [np.load(tmp_path / f"out{r}.npz") for r in range(2)]
    assert [tuple(o["rows"]) for o in outs] == [(0, size // 2), (size // 2, size)]
    assert all(int(o["t"]) == want[1] and int(o["n"]) == want[4] for o in outs)
    assert np.array_equal(np.concatenate([o["c"] for o in outs]), want[0])
    assert np.array_equal(np.concatenate([o["om"] for o in outs]), want[2])
    assert np.array_equal(np.concatenate([o["labels"] for o in outs]), want[3])
    assert all(np.array_equal(o["props"], want[5]) for o in outs)
