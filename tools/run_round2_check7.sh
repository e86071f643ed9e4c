set -x
python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "adaptive or gauss" > gpurun_out/s10_pytest.log 2>&1; tail -3 gpurun_out/s10_pytest.log
python tools/time_adaptive.py 8192 2>&1 | grep block
YAM_ADAPTIVE_PROF=1 python tools/time_adaptive.py 8192 2>&1 | grep -E "prof" | tail -3
python tools/time_c4_ops.py 2>&1 | grep -E "gauss|adaptive"
