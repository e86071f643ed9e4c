set -x
N=${1:-2}
YAM_MOSAIC_TRACE=events python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/t${N}_bench_c4.json 2> gpurun_out/t${N}_bench_c4.err; echo "rc=$?"
grep "\[trace\]" gpurun_out/t${N}_bench_c4.err | tail -30
