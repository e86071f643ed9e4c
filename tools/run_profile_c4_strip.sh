set -x
python tools/time_c4_ops.py > gpurun_out/s3_time_c4_ops.log 2>&1
ncu --set full --clock-control none --import-source on -o gpurun_out/s3_c4_strip -f python tools/profile_c4_strip.py > gpurun_out/s3_ncu_c4_strip.log 2>&1
tail -3 gpurun_out/s3_ncu_c4_strip.log
cat gpurun_out/s3_time_c4_ops.log
