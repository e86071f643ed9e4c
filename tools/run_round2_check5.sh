set -x
python -m pytest tests/test_gpu_ops.py tests/test_gpu_pipeline.py tests/test_gpu_mosaic.py -m gpu -x -q -k "otsu or hist or Otsu or clahe or mosaic or equalize" > gpurun_out/s8_pytest.log 2>&1; tail -3 gpurun_out/s8_pytest.log
python tools/time_c4_ops.py > gpurun_out/s8_time_c4_ops.log 2>&1; cat gpurun_out/s8_time_c4_ops.log
python tools/time_c5_parts.py > gpurun_out/s8_time_c5_parts.log 2>&1; cat gpurun_out/s8_time_c5_parts.log
