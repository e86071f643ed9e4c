python -m pytest tests -m gpu -x -q > gpurun_out/z_pytest.log 2>&1; tail -3 gpurun_out/z_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
