python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "otsu" 2>&1 | tail -2
python tools/time_otsu_sparse.py 2>&1 | tail -4
