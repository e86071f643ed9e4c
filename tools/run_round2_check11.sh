set -x
python -m pytest tests/test_gpu_mosaic.py -m gpu -x -q > gpurun_out/s14_pytest.log 2>&1; tail -5 gpurun_out/s14_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/s14_bench_c4.json 2> gpurun_out/s14_bench_c4.err; python - <<'P'
import json
d=json.loads([l for l in open("gpurun_out/s14_bench_c4.json") if l.startswith("{")][-1])
print("c4", d["ms_per_step"], d["check"]["otsu_threshold"], d["check"]["components"], d["check"]["labels_checksum64"], d["check"]["otsu_mask_checksum64"], d["gpu_launches"], d["e2e"]["value"])
P
tail -3 gpurun_out/s14_bench_c4.err
