"""GPU probe: first-cycle difference device vs oracle over operator / smoother / mode / size variants."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from helpers import two_level_problem  # noqa: E402
from mfmg_b200 import device as d  # noqa: E402

handle = d.CudaHandle(0)
for cells, block in ((4, 2), (6, 2), (8, 2), (12, 3)):
    P, R, Ac = two_level_problem(3, 1, cells, block, 2, "constant")
    for mf in (False, True):
        for smoother in ("Jacobi", "Chebyshev"):
            for precond in (True, False):
                for nu in (1, 2):
                    fine_d = d.MatrixFreeLaplaceDevice(handle, 3, 1, P.cells, P.h, P.coef_per_q(), P.constrained) if mf \
                        else d.SparseMatrixDevice.from_host(handle, P.A)
                    fine_o = oracle.MatrixFreeLaplace(3, 1, P.cells, P.h, P.coef_per_q(), P.constrained) if mf \
                        else (P.n, P.A.rowptr, P.A.col, P.A.val)
                    H = d.Hierarchy(handle, [fine_d, d.SparseMatrixDevice.from_host(handle, Ac)],
                                    [d.SparseMatrixDevice.from_host(handle, R)],
                                    {"is preconditioner": precond, "smoother": {"type": smoother, "n_smoothing_steps": nu}})
                    Ho = oracle.Hierarchy([fine_o, (Ac.n_rows, Ac.rowptr, Ac.col, Ac.val)],
                                          [(R.n_rows, R.n_cols, R.rowptr, R.col, R.val)], nu, precond,
                                          chebyshev={} if smoother == "Chebyshev" else None)
                    rng = np.random.default_rng(1)
                    x0 = oracle.std_uniform01(P.n, skip=P.constrained)
                    b_h = rng.standard_normal(P.n) * (P.constrained == 0)
                    errs = []
                    for bb in (np.zeros(P.n), b_h):
                        x = d.DeviceVector.from_host(handle, x0)
                        b = d.DeviceVector.from_host(handle, bb)
                        H.vmult(x, b)
                        xo = Ho.vmult(bb, x0)
                        errs.append(np.linalg.norm(x.to_host() - xo) / max(np.linalg.norm(xo), 1e-300))
                    print(f"cells={cells} mf={mf} {smoother:9s} precond={precond} nu={nu}: rel err b=0 {errs[0]:.2e}, b!=0 {errs[1]:.2e}",
                          flush=True)
