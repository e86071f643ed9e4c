# GPU-box check: parity suite, then the single-GPU bench lines (no profiler)
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest.log
tail -5 gpurun_out/s3_pytest.log
for w in c1 c5 c4; do
  python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/s3_bench_$w.json 2> gpurun_out/s3_bench_$w.err; echo "bench $w rc=$?"
done
