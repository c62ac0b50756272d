"""Shared builders for the parity tests: the reference tests' matrices and small two-level problems."""
import functools
import os

import numpy as np
import scipy.sparse as sp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kat.npz")


def golden():
    return np.load(GOLDEN)


def csr_arrays(m):
    m = sp.csr_matrix(m)
    m.sort_indices()
    return m.shape[0], m.shape[1], m.indptr.astype(np.int64), m.indices.astype(np.int32), m.data.astype(np.float64)


def tridiag_matrix(size=30):
    """tests/test_smoother_device.cu:36-58 / tests/test_direct_solver_device.cu:27-47: tridiag(-1, 4, -1).
    (The end rows are written exactly like the reference: set(i, j_min, -1), set(i, j_max, -1), set(i, i, 4).)"""
    a = np.zeros((size, size))
    for i in range(size):
        j_max = min(size - 1, i + 1)
        j_min = 0 if i == 0 else i - 1
        a[i, j_min] = -1.0
        a[i, j_max] = -1.0
        a[i, i] = 4.0
    m = sp.lil_matrix((size, size))
    for i in range(size):
        for j in range(max(0, i - 1), min(size, i + 2)):
            m[i, j] = a[i, j]
    return sp.csr_matrix(m)


def banded_matrix(n_rows=30, nnz_per_row=10):
    """tests/test_sparse_matrix_device_operator.cu:31-50: entries (i, i+j) = i + (i+j), j < 10; 30 x 39."""
    n_cols = n_rows + nnz_per_row - 1
    rows, cols, vals = [], [], []
    for i in range(n_rows):
        for j in range(nnz_per_row):
            rows.append(i)
            cols.append(i + j)
            vals.append(float(i + (i + j)))
    # explicit zeros must stay in the pattern (entry (0,0) = 0)
    m = sp.csr_matrix((np.array(vals), (np.array(rows), np.array(cols))), shape=(n_rows, n_cols))
    return m


def serial_mv_matrix():
    """tests/test_sparse_matrix_device.cu:36-57 (pattern and values from the committed golden file)."""
    g = golden()
    dense, pattern = g["serial_mv_dense"], g["serial_mv_pattern"]
    rows, cols = np.nonzero(pattern)
    return sp.csr_matrix((dense[rows, cols], (rows, cols)), shape=dense.shape)


@functools.lru_cache(maxsize=None)
def two_level_problem(dim, degree, cells, block, n_eig, material="constant", eigensolver="free"):
    """(problem, R, A_c) through the host setup path."""
    from mfmg_b200 import hostsetup as hs

    P = hs.LaplaceProblem.create(dim, degree, cells, material)
    R = hs.build_restrictor(P, (block,) * dim, n_eig, eigensolver=eigensolver)
    Ac = hs.galerkin(P.A, R)
    return P, R, Ac


def oracle_hierarchy(P, R, Ac, nu=1, is_preconditioner=True, omega=1.0, explicit_transpose=True):
    import oracle

    return oracle.Hierarchy([(P.n, P.A.rowptr, P.A.col, P.A.val), (Ac.n_rows, Ac.rowptr, Ac.col, Ac.val)],
                            [(R.n_rows, R.n_cols, R.rowptr, R.col, R.val)], nu, is_preconditioner, omega,
                            explicit_transpose)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (d if d > 0 else 1.0)


def slab_parts(world, cells, h, block, n_eig, degree=1, material="constant"):
    """All ranks' hostsetup.build_slab_part results in one process (the setup-time gathers are emulated by two
    passes: the first records what every rank contributes)."""
    from mfmg_b200 import hostsetup as hs

    contributions = {}

    def run(rank, record):
        calls = [0]

        def gather(obj):
            k = calls[0]
            calls[0] += 1
            if record:
                contributions.setdefault(k, {})[rank] = obj
                return [obj] * world
            return [contributions[k][r] for r in range(world)]

        return hs.build_slab_part(degree, cells, h, material, block, n_eig, world, rank, gather)

    for r in range(world):
        run(r, True)
    return [run(r, False) for r in range(world)]


def slab_vector(part, v_global):
    """[owned | ghost] local copy of a global vector."""
    return np.concatenate([v_global[part.row_begin:part.row_end], v_global[part.ghost_global]])
