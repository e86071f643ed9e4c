set -x
python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "otsu or hist or clahe or equalize" > gpurun_out/s9_pytest.log 2>&1; tail -3 gpurun_out/s9_pytest.log
python tools/time_c4_ops.py 2>&1 | grep -E "clahe_luts|histogram"
python tools/time_c5_parts.py 2>&1 | grep -E "hist|clahe"
