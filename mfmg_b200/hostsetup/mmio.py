"""Matrix-Market interchange with a reference build.

mfmg writes operators and vectors with EpetraExt (`matrix_market_output_file`, source/dealii/dealii_utils.cc:63-91):
matrices as `%%MatrixMarket matrix coordinate real general` (1-based, one `row col value` triple per line, values
`%22.16e`), vectors as `%%MatrixMarket matrix array real general` (n x 1).  These helpers read and write exactly those
two forms, so A / R / A_c produced here can be fed to a reference build elsewhere (and its residual histories or
operators compared with ours) without either side linking the other.  Host-side setup utility; not on the hot path.
"""
from __future__ import annotations

import numpy as np

from .problems import HostCSR


def write_matrix(path: str, A: HostCSR, comment: str = "") -> None:
    """Coordinate real general, rows in ascending order, columns in storage order, `%22.16e` values."""
    rows = np.repeat(np.arange(A.n_rows, dtype=np.int64), np.diff(A.rowptr)) + 1
    cols = A.col.astype(np.int64) + 1
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        for line in comment.splitlines():
            f.write("% " + line + "\n")
        f.write(f"{A.n_rows} {A.n_cols} {A.nnz}\n")
        np.savetxt(f, np.column_stack([rows, cols, A.val]), fmt=["%d", "%d", "%22.16e"])


def read_matrix(path: str) -> HostCSR:
    """Coordinate real general (or symmetric) -> CSR with ascending columns; duplicate entries are summed."""
    import scipy.sparse as sp

    with open(path) as f:
        header = f.readline().strip().lower().split()
        if header[:3] != ["%%matrixmarket", "matrix", "coordinate"] or header[3] not in ("real", "integer"):
            raise ValueError(f"{path}: not a real coordinate Matrix-Market file")
        symmetric = header[4] == "symmetric"
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        n_rows, n_cols, nnz = (int(t) for t in line.split())
        data = np.loadtxt(f, ndmin=2) if nnz else np.zeros((0, 3))
    if data.shape[0] != nnz:
        raise ValueError(f"{path}: expected {nnz} entries, found {data.shape[0]}")
    r, c, v = data[:, 0].astype(np.int64) - 1, data[:, 1].astype(np.int64) - 1, data[:, 2]
    if symmetric:
        off = r != c
        r, c, v = np.concatenate([r, c[off]]), np.concatenate([c, r[off]]), np.concatenate([v, v[off]])
    m = sp.coo_matrix((v, (r, c)), shape=(n_rows, n_cols)).tocsr()
    m.sum_duplicates()
    m.sort_indices()
    return HostCSR.from_scipy(m)


def write_vector(path: str, v: np.ndarray) -> None:
    v = np.asarray(v, dtype=np.float64).reshape(-1)
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix array real general\n")
        f.write(f"{len(v)} 1\n")
        np.savetxt(f, v, fmt="%22.16e")


def read_vector(path: str) -> np.ndarray:
    with open(path) as f:
        header = f.readline().strip().lower().split()
        if header[:3] != ["%%matrixmarket", "matrix", "array"]:
            raise ValueError(f"{path}: not an array Matrix-Market file")
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        n, m = (int(t) for t in line.split())
        v = np.loadtxt(f, ndmin=1)
    if v.size != n * m:
        raise ValueError(f"{path}: expected {n * m} values, found {v.size}")
    return v.reshape(-1)
