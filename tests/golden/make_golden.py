"""Regenerates tests/golden/kat.npz: the inputs of the reference's known-answer tests that depend on
libstdc++'s std::default_random_engine, plus the gold values those tests assert.

Run from the repo root:  python tests/golden/make_golden.py
Needs oracle/libstdrand.so (built from oracle/stdrand.cpp with the g++ in this image); the
reference itself is not needed (its tests only define the inputs, which are restated here with
file:line citations).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

out = {}

# tests/test_sparse_matrix_device.cu:36-57 (serial_mv): 10x10, row i draws 5 columns from
# default_random_engine(seed = i), uniform_int_distribution(0, 9); set(i, col, i + j) (last write wins)
size = 10
dense = np.zeros((size, size))
pattern = np.zeros((size, size), dtype=np.uint8)
for i in range(size):
    cols = oracle.std_uniform_int(5, 0, size - 1, seed=i)
    for j, c in enumerate(cols):
        dense[i, c] = float(i + j)
        pattern[i, c] = 1
out["serial_mv_dense"] = dense
out["serial_mv_pattern"] = pattern
out["serial_mv_x"] = np.arange(size, dtype=np.float64)
out["serial_mv_y"] = dense @ np.arange(size, dtype=np.float64)  # integers: exact

# tests/test_direct_solver_device.cu:49-53: x_ref ~ N(10, 2), default_random_engine default seed
out["direct_solver_xref"] = oracle.std_normal(30, 10.0, 2.0)

# tests/hierarchy_driver.cc:153-164 / tests/test_hierarchy_device.cu:293-297: U(0,1) stream
out["uniform01_first32"] = oracle.std_uniform01(32)

# tests/test_hierarchy_device.cu:365-371 two-grid gold (cube, no distortion), asserted to 1e-6 %
out["gold_rate_device_cube"] = np.array(0.14933479171507894)
# tests/test_smoother_device.cu:36-114: tridiag(-1,4,-1), b = 1, x0 = 0 -> x = 0.25
out["smoother_expected"] = np.full(30, 0.25)

# tests/test_agglomerate.cc:120-286 (simple_agglomerate_3d, one rank): agglomerate id of every cell of the 8 x 8 x 8
# mesh in deal.II's traversal order for the block partitioner nx = 2, ny = 3, nz = 4.  512 integers: transcribed from
# the reference test file when it is available (this container), kept from the previous fixture otherwise.
ref_test = "/root/reference/tests/test_agglomerate.cc"
old = os.path.join(ROOT, "tests", "golden", "kat.npz")
if os.path.exists(ref_test):
    import re

    text = open(ref_test).read()
    body = text[text.index("simple_agglomerate_3d"):]
    block = body[body.index("ref_agglomerates = {"):]
    block = block[:block.index("};")]
    ids = [int(t) for t in re.findall(r"\d+", block)]
    assert len(ids) == 512, len(ids)
    out["agglomerates_3d_ids"] = np.array(ids, dtype=np.int32)
elif os.path.exists(old) and "agglomerates_3d_ids" in np.load(old).files:
    out["agglomerates_3d_ids"] = np.load(old)["agglomerates_3d_ids"]

np.savez(os.path.join(ROOT, "tests", "golden", "kat.npz"), **out)
print("wrote tests/golden/kat.npz:", {k: v.shape for k, v in out.items()})
