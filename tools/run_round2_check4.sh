set -x
python -m pytest tests -m gpu -x -q > gpurun_out/s7_pytest.log 2>&1; tail -4 gpurun_out/s7_pytest.log
python tools/time_c5_parts.py > gpurun_out/s7_time_c5_parts.log 2>&1; cat gpurun_out/s7_time_c5_parts.log
python tools/time_c4_ops.py > gpurun_out/s7_time_c4_ops.log 2>&1; cat gpurun_out/s7_time_c4_ops.log
for w in c1 c5; do
  python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/s7_bench_$w.json 2> gpurun_out/s7_bench_$w.err; echo "bench $w rc=$?"
done
