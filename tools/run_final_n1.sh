# round-2 final single-GPU run: parity suite, every bench line, the reference arm, launch list of the default bench
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/f1_pytest.log 2>&1; tail -4 gpurun_out/f1_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/f1_smoke.log 2>&1; tail -2 gpurun_out/f1_smoke.log
for w in c1 c2 c3 c5; do
  python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/f1_bench_$w.json 2> gpurun_out/f1_bench_$w.err; echo "bench $w rc=$?"
done
python bench.py > gpurun_out/f1_bench_c4.json 2> gpurun_out/f1_bench_c4.err; echo "bench default rc=$?"
python bench.py --impl reference > gpurun_out/f1_bench_ref.json 2> gpurun_out/f1_bench_ref.err; echo "bench reference rc=$?"
python tools/time_c4_ops.py > gpurun_out/f1_time_c4_ops.log 2>&1; cat gpurun_out/f1_time_c4_ops.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/f1_launches_bench_c4.csv python bench.py --steps 2 --warmup 1 > gpurun_out/f1_ncu_bench.log 2>&1; tail -2 gpurun_out/f1_ncu_bench.log
ls -la gpurun_out | head -30
