python bench.py > gpurun_out/h1_bench_c4.json 2> gpurun_out/h1_bench_c4.err; echo "rc=$?"
python - <<P
import json
d=json.loads([l for l in open("gpurun_out/h1_bench_c4.json") if l.startswith("{")][-1])
print("c4 N=1", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], d["check"]["otsu_threshold"], d["check"]["components"], d["check"]["labels_checksum64"], d["check"]["otsu_mask_checksum64"], d["clocks"]["reasons"], d["n1_schedules"])
P
