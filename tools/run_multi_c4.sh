set -x
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/h${N}_bench_c4.json 2> gpurun_out/h${N}_bench_c4.err; echo "c4 N=$N rc=$?"
python - <<P
import json
d=json.loads([l for l in open("gpurun_out/h${N}_bench_c4.json") if l.startswith("{")][-1])
print("c4 N=$N", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], d["check"]["otsu_threshold"], d["check"]["components"], d["check"]["labels_checksum64"], d["check"]["otsu_mask_checksum64"], d["clocks"]["reasons"])
P
