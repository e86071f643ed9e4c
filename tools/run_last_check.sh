# last GPU check of the round (3.5 GPU-minutes left): new tests first, then the rest of the suite, then smoke()
timeout 150 python -m pytest tests/test_gpu_regiongeom.py tests/test_gpu_n4.py tests/test_gpu_ops.py tests/test_gpu_pipeline.py tests/test_gpu_mosaic.py \
    -m gpu -v --tb=short -p no:cacheprovider > gpurun_out/last_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/last_pytest.log
tail -5 gpurun_out/last_pytest.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/last_smoke.log 2>&1; tail -2 gpurun_out/last_smoke.log
