set -x
YAM_HIST_TRACE=1 python tools/time_c5_parts.py 2>&1 | grep -i "hist" | sort | uniq -c | head
for mode in 0 1; do
  YAM_HIST_NO_CLUSTER=$mode ncu --set full --clock-control none --import-source on -k regex:hist16 -c 2 -o /tmp/s6_hist16_$mode -f python tools/profile_hist16.py > gpurun_out/s6_ncu_hist16_$mode.log 2>&1
  ncu -i /tmp/s6_hist16_$mode.ncu-rep --page source --csv > gpurun_out/s6_hist16_${mode}_source.csv
  python tools/ncu_summary.py /tmp/s6_hist16_$mode.ncu-rep
  python tools/ncu_source_summary.py gpurun_out/s6_hist16_${mode}_source.csv 1 > gpurun_out/s6_hist16_${mode}_source.txt; cat gpurun_out/s6_hist16_${mode}_source.txt
done
