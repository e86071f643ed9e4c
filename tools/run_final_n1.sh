set -x
python -m pytest tests -m gpu -x -q > gpurun_out/g1_pytest.log 2>&1; tail -3 gpurun_out/g1_pytest.log
for w in c1 c2 c3 c5; do
  python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/g1_bench_$w.json 2> gpurun_out/g1_bench_$w.err; echo "bench $w rc=$?"
done
python bench.py > gpurun_out/g1_bench_c4.json 2> gpurun_out/g1_bench_c4.err; echo "bench default rc=$?"
python bench.py --impl reference > gpurun_out/g1_bench_ref.json 2> gpurun_out/g1_bench_ref.err; echo "bench reference rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/g1_launches_bench_c4.csv python bench.py --steps 2 --warmup 1 > gpurun_out/g1_ncu_bench.log 2>&1; tail -2 gpurun_out/g1_ncu_bench.log
