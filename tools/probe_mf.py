"""GPU probe: matrix-free apply bandwidth (cfg4 fine level: 256^3 cells Q1) and stage variants."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mfmg_b200 import device as d  # noqa: E402
from mfmg_b200 import hostsetup as hs  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 256
degree = int(sys.argv[2]) if len(sys.argv) > 2 else 1
mat = sys.argv[3] if len(sys.argv) > 3 else "discontinuous"
P = hs.LaplaceProblem.create(3, degree, cells, mat, assemble_matrix=False)
stream = torch.cuda.Stream()
h = d.CudaHandle(0, stream=stream.cuda_stream)
M = d.MatrixFreeLaplaceDevice(h, 3, degree, P.cells, P.h, P.coef_per_q(), P.constrained)
n = P.n
ncells = int(np.prod(P.cells))
nq = (degree + 1) ** 3
rng = np.random.default_rng(0)
x = d.DeviceVector.from_host(h, rng.standard_normal(n))
y = d.DeviceVector(h, n)


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


ms = timeit(lambda: M.apply(x, y))
alg = 16 * n + (0 if "stencil" in M.kernel else 8 * ncells * (1 if "per-cell" in M.kernel else nq)) + n
flops = ncells * 2 * (16 * (degree + 1) ** 4 + 40)
print(json.dumps({"cells": cells, "degree": degree, "n": n, "ms": ms, "alg_bytes": alg, "gbs": alg / ms / 1e6,
                  "approx_tflops": flops / ms / 1e9}))

