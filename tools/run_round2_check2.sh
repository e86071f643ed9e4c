set -x
python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k otsu > gpurun_out/s4_pytest_otsu.log 2>&1; tail -3 gpurun_out/s4_pytest_otsu.log
python tools/time_c5_parts.py > gpurun_out/s4_time_c5_parts.log 2>&1; cat gpurun_out/s4_time_c5_parts.log
for w in c1 c5; do
  python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/s4_bench_$w.json 2> gpurun_out/s4_bench_$w.err; echo "bench $w rc=$?"
done
ncu --set full --clock-control none --import-source on -k regex:"clahe|hist16|otsu|threshold" -o gpurun_out/s4_c1_c5_hist -f python tools/profile_c1_otsu.py > gpurun_out/s4_ncu_c1.log 2>&1; tail -2 gpurun_out/s4_ncu_c1.log
ROWS=65536 ncu --set full --clock-control none -o gpurun_out/s4_c4_full -f python tools/profile_c4_strip.py > gpurun_out/s4_ncu_c4_full.log 2>&1; tail -2 gpurun_out/s4_ncu_c4_full.log
