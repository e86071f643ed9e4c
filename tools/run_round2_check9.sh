set -x
python -m pytest tests/test_gpu_ops.py tests/test_gpu_pipeline.py -m gpu -x -q -k "clahe or hist or otsu or Otsu or CLAHE" > gpurun_out/s12_pytest.log 2>&1; tail -3 gpurun_out/s12_pytest.log
python tools/time_c5_parts.py 2>&1 | grep -E "clahe|hist"
for w in c1 c5; do python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/s12_bench_$w.json 2>/dev/null; python - <<P
import json
d=json.loads([l for l in open("gpurun_out/s12_bench_$w.json") if l.startswith("{")][-1])
print("$w", d["ms_per_step"], [(o["op"][:12], round(o["ms"],4)) for o in d["roofline"]["ops"]])
P
done
