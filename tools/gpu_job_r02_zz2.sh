# round 2 (1 GPU): C++ adapter test with SparseMatrixDevice::mmult
timeout 60 python -m pytest tests/test_cpp_adapter.py -m gpu -q -x 2>&1 | tail -5
