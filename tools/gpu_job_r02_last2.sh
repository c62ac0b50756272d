# round 2 (1 GPU): the loaded-flags variant of the strip sweep (MFMGB_MF_ARITH_FLAGS=0) against the cell kernel / oracle
MFMGB_MF_ARITH_FLAGS=0 timeout 26 python -m pytest tests/test_gpu_matrix_free.py -q -x -k "stencil_sweep or slab" 2>&1 | tail -4
