set -x
N=${1:-2}
if [ "$N" = "2" ]; then python -m pytest tests/test_gpu_mosaic.py -m gpu -x -q > gpurun_out/f2_pytest_mosaic.log 2>&1; tail -3 gpurun_out/f2_pytest_mosaic.log; fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/f${N}_bench_c4.json 2> gpurun_out/f${N}_bench_c4.err; echo "c4 N=$N rc=$?"
tail -c 300 gpurun_out/f${N}_bench_c4.err
