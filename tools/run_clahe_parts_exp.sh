python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "morph_zero or clahe" 2>&1 | tail -2
python tools/time_c5_parts.py 2>&1 | grep "c1: clahe"
YAM_CLAHE_PARTS_MIN_AREA=131072 python tools/time_c5_parts.py 2>&1 | grep "c1: clahe"
YAM_CLAHE_PARTS_MIN_AREA=131072 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "clahe" 2>&1 | tail -2
