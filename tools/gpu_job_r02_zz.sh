# round 2, last call (1 GPU): C++ adapter test and smoke() at HEAD
set -x
timeout 90 python -m pytest tests/test_cpp_adapter.py -m gpu -q -x > gpurun_out/r02_pytest_gpu_cpp_head.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_gpu_cpp_head.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
