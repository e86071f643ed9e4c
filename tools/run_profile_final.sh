set -x
ncu --set full --clock-control none -k regex:"hist16|otsu|adaptive|reduce" -o /tmp/z_final -f python tools/profile_c4_strip.py > gpurun_out/z_ncu_final.log 2>&1; tail -2 gpurun_out/z_ncu_final.log
python tools/ncu_summary.py /tmp/z_final.ncu-rep > gpurun_out/z_ncu_final_summary.txt; cat gpurun_out/z_ncu_final_summary.txt
ncu -i /tmp/z_final.ncu-rep --page raw --csv > gpurun_out/z_ncu_final_raw.csv
ls -la gpurun_out/z_*
