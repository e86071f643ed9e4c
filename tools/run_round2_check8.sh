set -x
python -m pytest tests/test_gpu_ops.py tests/test_gpu_pipeline.py tests/test_gpu_mosaic.py -m gpu -x -q -k "adaptive or gauss or mosaic or segment or fused" > gpurun_out/s11_pytest.log 2>&1; tail -3 gpurun_out/s11_pytest.log
python tools/time_adaptive.py 8192 2>&1 | grep block
python tools/time_c4_ops.py 2>&1 | grep -E "gauss|adaptive|checks"
python bench.py --steps 5 --warmup 3 > gpurun_out/s11_bench_c4.json 2> gpurun_out/s11_bench_c4.err; python - <<'P'
import json
d=json.loads([l for l in open("gpurun_out/s11_bench_c4.json") if l.startswith("{")][-1])
print(d["ms_per_step"], d["check"])
P
