"""bench.py contract on a box without a GPU: the reference arm prints one JSON line with the agreed keys (it times
the reference's CPU path -- the unmodified reference when /root/reference is importable, else the restated call
sites), and the GPU arm refuses to run instead of falling back to the CPU."""
from __future__ import annotations

import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_line():
    r = _run("--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "megapixels/s per pipeline" and d["unit"] == "megapixels/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "u16" and d["data"] == "synthetic"
    assert d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == pytest.approx(d["value"]) and cb["sample"]
    e = d["e2e"]
    assert e["value"] == pytest.approx(d["value"]) and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_gpu_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the GPU arm runs")
    r = _run("--workload", "c1", "--steps", "1", "--warmup", "1", timeout=300)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout) or "BackendUnavailable" in (r.stderr + r.stdout)


def test_committed_c4_lines_share_the_cv2_verified_check_block():
    """The committed 65536^2 lines (1, 2, 4, 8 GPUs) carry one and the same `check` block, and it is the one the
    reference's cv2 call sites produce over the same mosaic on the CPU (tools/check_bench_check_block.py)."""
    import json
    from pathlib import Path

    prof = Path(__file__).resolve().parents[1] / "profiles"
    cpu = json.loads((prof / "r02_check_c4_65536_vs_cv2.json").read_text())
    assert cpu["equal"] is True and cpu["mosaic"] == [65536, 65536]
    want = cpu["gpu bench line"]
    cpu_side = next(v for k, v in cpu.items() if k.startswith("cpu (cv2"))
    assert cpu_side == want
    for n in (1, 2, 4, 8):
        text = (prof / f"r02_bench_c4_n{n}.json").read_text()      # torchrun output: an NCCL banner may precede the line
        line = json.loads(next(ln for ln in text.splitlines() if ln.startswith("{")))
        assert line["n_gpus"] == n and line["config"]["mosaic"] == [65536, 65536]
        assert {k: line["check"][k] for k in want} == want, f"N = {n}"


def test_bench_check_block_cites_the_cv2_cpu_chain():
    """bench.cv2_reference_check: the committed CPU / cv2 values are attached to a c4 line's `check` block with a
    verdict; sizes without a committed result (or unreadable files) only drop the key."""
    import json
    from pathlib import Path

    import bench

    prof = Path(__file__).resolve().parents[1] / "profiles"
    line = json.loads((prof / "r02_bench_c4_n1.json").read_text())
    good = bench.cv2_reference_check(65536, line["check"])
    assert good["equal"] is True and good["values"]["components"] == 6673681 and "r02_check_c4_65536_vs_cv2.json" in good["source"]
    bad = dict(line["check"], components=1)
    assert bench.cv2_reference_check(65536, bad)["equal"] is False
    assert bench.cv2_reference_check(16384, json.loads((prof / "r02_bench_c4_16k_n1.json").read_text())["check"])["equal"] is True
    assert bench.cv2_reference_check(4096, line["check"]) is None
