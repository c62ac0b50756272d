"""GPU probe: SpMV / fused-kernel bandwidth vs lanes-per-row on the cfg1 operators (A, R, P) and the dense solve.
Writes a JSON summary to stdout.  Usage: python tools/probe_spmv.py [cells] [block] [neig]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mfmg_b200 import device as d  # noqa: E402
from mfmg_b200 import hostsetup as hs  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 128
block = int(sys.argv[2]) if len(sys.argv) > 2 else 8
neig = int(sys.argv[3]) if len(sys.argv) > 3 else 1
degree = int(sys.argv[4]) if len(sys.argv) > 4 else 1

P = hs.LaplaceProblem.create(3, degree, cells)
R = hs.build_restrictor(P, (block,) * 3, neig)
Ac = hs.galerkin(P.A, R)
stream = torch.cuda.Stream()
h = d.CudaHandle(0, stream=stream.cuda_stream)
Ad = d.SparseMatrixDevice.from_host(h, P.A)
Rd = d.SparseMatrixDevice.from_host(h, R)
Pd = Rd.transpose()
n, nc = P.n, R.n_rows
rng = np.random.default_rng(0)
x = d.DeviceVector.from_host(h, rng.standard_normal(n))
y = d.DeviceVector(h, n)
xc = d.DeviceVector.from_host(h, rng.standard_normal(nc))
yc = d.DeviceVector(h, nc)


def timeit(fn, reps=20):
    with torch.cuda.stream(stream):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {"n": n, "nnz": P.A.nnz, "nc": nc, "nnzR": R.nnz}
bA = 12 * P.A.nnz + 4 * (n + 1) + 16 * n
bR = 12 * R.nnz + 4 * (nc + 1) + 8 * n + 8 * nc
bP = 12 * R.nnz + 4 * (n + 1) + 8 * nc + 16 * n
for lanes in (1, 2, 4, 8, 16, 32):
    Ad.set_lanes_per_row(lanes)
    ms = timeit(lambda: Ad.vmult(y, x))
    out[f"A_lanes{lanes}"] = {"ms": ms, "gbs": bA / ms / 1e6}
    Rd.set_lanes_per_row(lanes)
    ms = timeit(lambda: Rd.vmult(yc, x))
    out[f"R_lanes{lanes}"] = {"ms": ms, "gbs": bR / ms / 1e6}
    Pd.set_lanes_per_row(lanes)
    ms = timeit(lambda: Pd.vmult(y, xc))
    out[f"P_lanes{lanes}"] = {"ms": ms, "gbs": bP / ms / 1e6}
op = d.CudaMatrixOperator(d.SparseMatrixDevice.from_host(h, Ac))
t0 = time.time()
s = d.CudaSolver(h, op, {})
h.synchronize()
out["dense_factor_s"] = time.time() - t0
ms = timeit(lambda: s.apply(xc, yc))
out["dense_solve"] = {"ms": ms, "gbs": (8 * nc * nc + 16 * nc) / ms / 1e6}
# copy bandwidth reference on this box
a = torch.empty(1 << 28, dtype=torch.float64, device="cuda")
bb = torch.empty_like(a)
with torch.cuda.stream(stream):
    ms = timeit(lambda: bb.copy_(a), 10)
out["copy_gbs"] = 2 * a.numel() * 8 / ms / 1e6
print(json.dumps(out, indent=1))
