"""CPU ORACLE (test infrastructure) -- Python face of oracle/mfmg_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product (mfmg_b200/) never does.  Parity status: pinned against the
reference test-suite's known-answer tests (tests/test_oracle_kat.py, tests/golden/); PCG iteration
counts and residual histories are "parity unpinned" (no reference test asserts them).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_RAND = None

_i64p = ctypes.POINTER(ctypes.c_int64)
_vp = ctypes.c_void_p


def _build():
    import subprocess

    subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libmfmg_oracle.so")
        if not os.path.exists(path):
            _build()
        L = ctypes.CDLL(path)
        L.orc_spmv.argtypes = [ctypes.c_int64, _vp, _vp, _vp, _vp, _vp]
        L.orc_spmv_transpose.argtypes = [ctypes.c_int64, ctypes.c_int64, _vp, _vp, _vp, _vp, _vp]
        L.orc_csr_transpose.argtypes = [ctypes.c_int64, ctypes.c_int64, _vp, _vp, _vp, _vp, _vp, _vp]
        L.orc_inv_diag.restype = ctypes.c_int64
        L.orc_inv_diag.argtypes = [ctypes.c_int64, _vp, _vp, _vp, _vp]
        L.orc_residual_neg.argtypes = [ctypes.c_int64, _vp, _vp, _vp, _vp, _vp, _vp]
        L.orc_jacobi_apply.argtypes = [ctypes.c_int64, _vp, _vp, _vp, _vp, ctypes.c_double, _vp, _vp, _vp]
        L.orc_csr_to_dense.argtypes = [ctypes.c_int64, _vp, _vp, _vp, _vp]
        L.orc_lu_factor.restype = ctypes.c_int
        L.orc_lu_factor.argtypes = [ctypes.c_int64, _vp, _vp]
        L.orc_lu_solve.argtypes = [ctypes.c_int64, _vp, _vp, _vp, _vp]
        L.orc_hierarchy_new.restype = _vp
        L.orc_hierarchy_new.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double]
        L.orc_hierarchy_set_explicit_transpose.argtypes = [_vp, ctypes.c_int]
        L.orc_hierarchy_set_coarse_storage.argtypes = [_vp, ctypes.c_int]
        L.orc_hierarchy_set_chebyshev.argtypes = [_vp, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int]
        L.orc_hierarchy_chebyshev_info.argtypes = [_vp, ctypes.c_int, _vp]
        L.orc_hierarchy_set_operator.argtypes = [_vp, ctypes.c_int, ctypes.c_int64, _vp, _vp, _vp]
        L.orc_hierarchy_set_mf_operator.argtypes = [_vp, _vp]
        L.orc_hierarchy_set_restrictor.argtypes = [_vp, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, _vp, _vp, _vp]
        L.orc_hierarchy_finalize.restype = ctypes.c_int
        L.orc_hierarchy_finalize.argtypes = [_vp]
        L.orc_hierarchy_free.argtypes = [_vp]
        L.orc_hierarchy_apply.argtypes = [_vp, _vp, _vp, ctypes.c_int]
        L.orc_hierarchy_vmult.argtypes = [_vp, _vp, _vp]
        L.orc_dot.restype = ctypes.c_double
        L.orc_dot.argtypes = [ctypes.c_int64, _vp, _vp]
        L.orc_pcg.restype = ctypes.c_int
        L.orc_pcg.argtypes = [_vp, ctypes.c_int64, _vp, _vp, _vp, _vp, _vp, ctypes.c_double, ctypes.c_int, _vp]
        L.orc_mf_new.restype = _vp
        L.orc_mf_new.argtypes = [ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp]
        L.orc_mf_free.argtypes = [_vp]
        L.orc_mf_apply.argtypes = [_vp, _vp, _vp]
        L.orc_mf_diag.argtypes = [_vp, _vp]
        L.orc_num_threads.restype = ctypes.c_int
        L.orc_set_num_threads.argtypes = [ctypes.c_int]
        _LIB = L
    return _LIB


def randlib():
    global _RAND
    if _RAND is None:
        path = os.path.join(_HERE, "libstdrand.so")
        if not os.path.exists(path):
            _build()
        R = ctypes.CDLL(path)
        R.stdrand_uniform01.argtypes = [ctypes.c_uint, ctypes.c_int64, _vp]
        R.stdrand_uniform01_masked.argtypes = [ctypes.c_uint, ctypes.c_int64, _vp, _vp]
        R.stdrand_normal.argtypes = [ctypes.c_uint, ctypes.c_double, ctypes.c_double, ctypes.c_int64, _vp]
        R.stdrand_uniform_int.argtypes = [ctypes.c_uint, ctypes.c_int, ctypes.c_int, ctypes.c_int64, _vp]
        _RAND = R
    return _RAND


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _csr(rowptr, col, val):
    return (np.ascontiguousarray(rowptr, dtype=np.int64), np.ascontiguousarray(col, dtype=np.int32),
            _f64(val))


# -- libstdc++ random streams of the reference's tests/driver ---------------------------------
def std_uniform01(n: int, seed: int = 1, skip=None) -> np.ndarray:
    """std::default_random_engine + uniform_real_distribution<double>(0,1), in index order
    (tests/hierarchy_driver.cc:153-164).  skip[i] != 0: no draw, value 0 (constrained DoFs)."""
    out = np.empty(n, dtype=np.float64)
    if skip is None:
        randlib().stdrand_uniform01(seed, n, out.ctypes.data)
    else:
        skip = np.ascontiguousarray(skip, dtype=np.uint8)
        randlib().stdrand_uniform01_masked(seed, n, skip.ctypes.data, out.ctypes.data)
    return out


def std_normal(n: int, mean: float, stddev: float, seed: int = 1) -> np.ndarray:
    out = np.empty(n, dtype=np.float64)
    randlib().stdrand_normal(seed, mean, stddev, n, out.ctypes.data)
    return out


def std_uniform_int(n: int, lo: int, hi: int, seed: int) -> np.ndarray:
    out = np.empty(n, dtype=np.int32)
    randlib().stdrand_uniform_int(seed, lo, hi, n, out.ctypes.data)
    return out


def minstd_uniform01_py(n: int, seed: int = 1) -> np.ndarray:
    """Pure-Python restatement of the same stream: minstd_rand0 (x <- 16807 x mod 2^31-1) and
    libstdc++'s generate_canonical<double,53> (two draws per double; range 2^31-2)."""
    m, a = 2147483647, 16807
    x = seed % m or 1
    rng = float(m - 1 - 1 + 1)  # max - min + 1 = 2147483646
    out = np.empty(n)
    for i in range(n):
        s, mult = 0.0, 1.0
        for _ in range(2):
            x = (a * x) % m
            s += float(x - 1) * mult
            mult *= rng
        v = s / mult
        if v >= 1.0:
            v = np.nextafter(1.0, 0.0)
        out[i] = v
    return out


# -- kernels -------------------------------------------------------------------------------------
def spmv(n_rows, rowptr, col, val, x):
    rowptr, col, val = _csr(rowptr, col, val)
    x = _f64(x)
    y = np.empty(n_rows)
    lib().orc_spmv(n_rows, rowptr.ctypes.data, col.ctypes.data, val.ctypes.data, x.ctypes.data, y.ctypes.data)
    return y


def spmv_transpose(n_rows, n_cols, rowptr, col, val, x):
    rowptr, col, val = _csr(rowptr, col, val)
    x = _f64(x)
    y = np.empty(n_cols)
    lib().orc_spmv_transpose(n_rows, n_cols, rowptr.ctypes.data, col.ctypes.data, val.ctypes.data,
                             x.ctypes.data, y.ctypes.data)
    return y


def csr_transpose(n_rows, n_cols, rowptr, col, val):
    rowptr, col, val = _csr(rowptr, col, val)
    nnz = int(rowptr[-1])
    t_rowptr = np.empty(n_cols + 1, dtype=np.int64)
    t_col = np.empty(max(nnz, 1), dtype=np.int32)
    t_val = np.empty(max(nnz, 1), dtype=np.float64)
    lib().orc_csr_transpose(n_rows, n_cols, rowptr.ctypes.data, col.ctypes.data, val.ctypes.data,
                            t_rowptr.ctypes.data, t_col.ctypes.data, t_val.ctypes.data)
    return t_rowptr, t_col[:nnz], t_val[:nnz]


def inv_diag(n, rowptr, col, val):
    rowptr, col, val = _csr(rowptr, col, val)
    d = np.empty(n)
    lib().orc_inv_diag(n, rowptr.ctypes.data, col.ctypes.data, val.ctypes.data, d.ctypes.data)
    return d


def residual_neg(n, rowptr, col, val, x, b):
    rowptr, col, val = _csr(rowptr, col, val)
    x, b = _f64(x), _f64(b)
    r = np.empty(n)
    lib().orc_residual_neg(n, rowptr.ctypes.data, col.ctypes.data, val.ctypes.data, x.ctypes.data,
                           b.ctypes.data, r.ctypes.data)
    return r


def jacobi_apply(n, rowptr, col, val, b, x, omega=1.0, dinv=None):
    """One sweep; returns the new x (input not modified)."""
    rowptr, col, val = _csr(rowptr, col, val)
    if dinv is None:
        dinv = inv_diag(n, rowptr, col, val)
    dinv, b = _f64(dinv), _f64(b)
    x = _f64(x).copy()
    work = np.empty(n)
    lib().orc_jacobi_apply(n, rowptr.ctypes.data, col.ctypes.data, val.ctypes.data, dinv.ctypes.data,
                           omega, b.ctypes.data, x.ctypes.data, work.ctypes.data)
    return x


def lu_factor_csr(n, rowptr, col, val):
    rowptr, col, val = _csr(rowptr, col, val)
    lu = np.empty((n, n))
    piv = np.empty(max(n, 1), dtype=np.int32)
    lib().orc_csr_to_dense(n, rowptr.ctypes.data, col.ctypes.data, val.ctypes.data, lu.ctypes.data)
    info = lib().orc_lu_factor(n, lu.ctypes.data, piv.ctypes.data)
    return lu, piv, info


def lu_solve(lu, piv, b):
    n = lu.shape[0]
    b = _f64(b)
    x = np.empty(n)
    lib().orc_lu_solve(n, lu.ctypes.data, piv.ctypes.data, b.ctypes.data, x.ctypes.data)
    return x


class MatrixFreeLaplace:
    """orc_mf: the matrix-free operator of tests/laplace_matrix_free.hpp on a uniform grid."""

    def __init__(self, dim, degree, cells, h, coef_per_q, constrained):
        self.cells = np.ascontiguousarray(list(cells) + [1] * (3 - dim), dtype=np.int64)
        self.h = _f64(list(h) + [1.0] * (3 - dim))
        self.coef = _f64(coef_per_q)
        self.constr = np.ascontiguousarray(constrained, dtype=np.uint8)
        self.ptr = lib().orc_mf_new(dim, degree, self.cells.ctypes.data, self.h.ctypes.data,
                                    self.coef.ctypes.data, self.constr.ctypes.data)
        self.n = int(self.constr.shape[0])

    def apply(self, x):
        x = _f64(x)
        y = np.empty(self.n)
        lib().orc_mf_apply(self.ptr, x.ctypes.data, y.ctypes.data)
        return y

    def diag(self):
        d = np.empty(self.n)
        lib().orc_mf_diag(self.ptr, d.ctypes.data)
        return d

    def __del__(self):
        if getattr(self, "ptr", None):
            lib().orc_mf_free(self.ptr)
            self.ptr = None


class Hierarchy:
    """orc_hierarchy: mfmg::Hierarchy::apply/vmult restated (include/mfmg/common/hierarchy.hpp:238-309).

    operators: list of (n, rowptr, col, val) per level, finest first (level 0 may be a
    MatrixFreeLaplace); restrictors: list of (n_rows, n_cols, rowptr, col, val), restrictors[i]
    maps level i to level i+1."""

    def __init__(self, operators, restrictors, n_smoothing_steps=1, is_preconditioner=True, omega=1.0,
                 explicit_transpose=True, coarse_storage="auto", chebyshev=None):
        """coarse_storage: "dense" = the dense getrf/getrs restatement, "band" = the same factorisation on band
        storage (bit-identical results, orc_band_lu_factor), "auto" = band when the bandwidth is below n / 4."""
        L = lib()
        self._keep = []
        self.n_levels = len(operators)
        self.ptr = L.orc_hierarchy_new(self.n_levels, n_smoothing_steps, int(is_preconditioner), omega)
        L.orc_hierarchy_set_explicit_transpose(self.ptr, int(explicit_transpose))
        L.orc_hierarchy_set_coarse_storage(self.ptr, {"auto": 0, "dense": 1, "band": 2}[coarse_storage])
        if chebyshev is not None:
            # smoother.type Chebyshev (source/dealii/dealii_matrix_free_smoother.cc:34-60); keys = the reference's
            # parameter names, defaults = dealii::PreconditionChebyshev::AdditionalData at the pinned deal.II
            c = dict(chebyshev)
            L.orc_hierarchy_set_chebyshev(self.ptr, int(c.get("degree", 0)), float(c.get("smoothing_range", 0.0)),
                                          float(c.get("max_eigenvalue", 1.0)), int(c.get("eig_cg_n_iterations", 8)),
                                          {"mod11": 0, "constant": 1}[c.get("initial_guess", "mod11")])
        self.A0 = None
        for li, op in enumerate(operators):
            if isinstance(op, MatrixFreeLaplace):
                assert li == 0
                self._keep.append(op)
                L.orc_hierarchy_set_mf_operator(self.ptr, op.ptr)
                self.n = op.n
            else:
                n, rowptr, col, val = op
                rowptr, col, val = _csr(rowptr, col, val)
                self._keep += [rowptr, col, val]
                L.orc_hierarchy_set_operator(self.ptr, li, n, rowptr.ctypes.data, col.ctypes.data, val.ctypes.data)
                if li == 0:
                    self.n = n
                    self.A0 = (n, rowptr, col, val)
        for li, r in enumerate(restrictors):
            n_rows, n_cols, rowptr, col, val = r
            rowptr, col, val = _csr(rowptr, col, val)
            self._keep += [rowptr, col, val]
            L.orc_hierarchy_set_restrictor(self.ptr, li + 1, n_rows, n_cols, rowptr.ctypes.data,
                                           col.ctypes.data, val.ctypes.data)
        self.info = L.orc_hierarchy_finalize(self.ptr)

    def chebyshev_info(self, level=0):
        """(lambda_min, lambda_max incl. the 1.2 safety factor, theta, delta) of a level's Chebyshev smoother."""
        out = np.zeros(4)
        lib().orc_hierarchy_chebyshev_info(self.ptr, level, out.ctypes.data)
        return tuple(out)

    def vmult(self, b, x0=None):
        """x = Hierarchy::vmult(x, b); x0 only matters when is_preconditioner is false."""
        b = _f64(b)
        x = np.zeros(self.n) if x0 is None else _f64(x0).copy()
        lib().orc_hierarchy_vmult(self.ptr, x.ctypes.data, b.ctypes.data)
        return x

    def pcg(self, b, x0, tol, max_it, A=None):
        """deal.II SolverCG recurrence with this hierarchy as preconditioner.  Returns
        (x, iterations (negative = not converged), residual history)."""
        A = A or self.A0
        b = _f64(b)
        x = _f64(x0).copy()
        hist = np.zeros(max_it + 1)
        if A is None:
            it = lib().orc_pcg(self.ptr, self.n, None, None, None, b.ctypes.data, x.ctypes.data, tol,
                               max_it, hist.ctypes.data)
        else:
            n, rowptr, col, val = A
            it = lib().orc_pcg(self.ptr, n, rowptr.ctypes.data, col.ctypes.data, val.ctypes.data,
                               b.ctypes.data, x.ctypes.data, tol, max_it, hist.ctypes.data)
        nit = it if it >= 0 else -it - 1
        return x, it, hist[:nit + 1]

    def __del__(self):
        if getattr(self, "ptr", None):
            lib().orc_hierarchy_free(self.ptr)
            self.ptr = None


def cg_unpreconditioned(A, b, x0, tol, max_it):
    n, rowptr, col, val = A
    rowptr, col, val = _csr(rowptr, col, val)
    b = _f64(b)
    x = _f64(x0).copy()
    hist = np.zeros(max_it + 1)
    it = lib().orc_pcg(None, n, rowptr.ctypes.data, col.ctypes.data, val.ctypes.data, b.ctypes.data,
                       x.ctypes.data, tol, max_it, hist.ctypes.data)
    nit = it if it >= 0 else -it - 1
    return x, it, hist[:nit + 1]


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(n)


# -- independent numpy restatement of the assembly (cross-check of the host setup path) ---------
def assemble_laplace_py(dim, degree, cells, coef_per_q, constrained):
    """Cell-loop assembly with deal.II's distribute_local_to_global constraint handling
    (tests/laplace.hpp:172-199), pure Python/numpy, small cases only.  Returns a dense matrix.
    Shape functions/quadrature from orc_mf (oracle's own tables, independent of hostsetup)."""
    p, n1 = degree, degree + 1
    nodes = [cells[d] * p + 1 for d in range(dim)]
    n = int(np.prod(nodes))
    h = [1.0 / cells[d] for d in range(dim)]
    A = np.zeros((n, n))
    ndof = n1 ** dim
    # cell matrix by applying the oracle's matrix-free cell kernel to unit vectors on a 1-cell grid
    one_cell = [1] * dim
    for cz in range(cells[2] if dim == 3 else 1):
        for cy in range(cells[1]):
            for cx in range(cells[0]):
                cell = cx + cells[0] * (cy + cells[1] * cz)
                mf = MatrixFreeLaplace(dim, degree, one_cell, h, coef_per_q[cell:cell + 1],
                                       np.zeros(ndof, dtype=np.uint8))
                K = np.stack([mf.apply(np.eye(ndof)[i]) for i in range(ndof)], axis=1)
                idx = []
                for az in range(n1 if dim == 3 else 1):
                    for ay in range(n1):
                        for ax in range(n1):
                            gx, gy, gz = cx * p + ax, cy * p + ay, cz * p + az
                            idx.append(gx + nodes[0] * (gy + (nodes[1] * gz if dim == 3 else 0)))  # lexicographic
                for a in range(ndof):
                    for b in range(ndof):
                        ia, ib = idx[a], idx[b]
                        if constrained[ia] or constrained[ib]:
                            if ia == ib:
                                A[ia, ia] += abs(K[a, a])
                        else:
                            A[ia, ib] += K[a, b]
    return A
