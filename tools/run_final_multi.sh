set -x
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/g${N}_bench_c4.json 2> gpurun_out/g${N}_bench_c4.err; echo "c4 N=$N rc=$?"
if [ "$N" = "8" ]; then python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --workload c5 --gpus $N --steps 5 --warmup 3 > gpurun_out/g${N}_bench_c5.json 2> gpurun_out/g${N}_bench_c5.err; echo "c5 N=$N rc=$?"; fi
if [ "$N" = "2" ]; then python -m pytest tests/test_gpu_mosaic.py -m gpu -x -q > gpurun_out/g2_pytest_mosaic.log 2>&1; tail -2 gpurun_out/g2_pytest_mosaic.log; fi
