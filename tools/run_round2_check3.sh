set -x
python -m pytest tests/test_gpu_ops.py tests/test_gpu_pipeline.py -m gpu -x -q -k "otsu or hist or Otsu" > gpurun_out/s5_pytest.log 2>&1; tail -3 gpurun_out/s5_pytest.log
python tools/time_c5_parts.py > gpurun_out/s5_time_c5_parts.log 2>&1; cat gpurun_out/s5_time_c5_parts.log
YAM_HIST_NO_CLUSTER=1 python tools/time_c5_parts.py > gpurun_out/s5_time_c5_parts_nocluster.log 2>&1; grep -i hist gpurun_out/s5_time_c5_parts_nocluster.log
ncu --set full --clock-control none -k regex:"clahe" -o /tmp/s5_clahe_stack -f python tools/profile_clahe_stack.py > gpurun_out/s5_ncu_clahe.log 2>&1; tail -2 gpurun_out/s5_ncu_clahe.log
ncu -i /tmp/s5_clahe_stack.ncu-rep --page raw --csv > gpurun_out/s5_clahe_stack_raw.csv
python tools/ncu_summary.py /tmp/s5_clahe_stack.ncu-rep > gpurun_out/s5_clahe_stack_summary.txt; cat gpurun_out/s5_clahe_stack_summary.txt
ROWS=65536 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/s5_c4_full_dram.csv python tools/profile_c4_strip.py > gpurun_out/s5_ncu_c4_full.log 2>&1; tail -2 gpurun_out/s5_ncu_c4_full.log
ls -la gpurun_out/
