"""Host setup: the reference's Laplace/diffusion test problems on structured grids.

Restates (does not copy) tests/laplace.hpp:87-204 (mesh, FE_Q(p), Gauss (p+1)^d quadrature, cell
matrix, Dirichlet elimination while assembling) and the material functions of
tests/test_hierarchy_helpers.hpp:75-188 for the uniform `hyper_cube` meshes the benchmark
configs use, with lexicographic DoF numbering.  This is SETUP: it produces the
CSR operator that user code hands to mfmg through MeshEvaluator::evaluate_global; none of it is
on the V-cycle hot path.  The heavy loop lives in assemble.cpp (g++ -fopenmp).
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libmfmg_b200_host.so")
        if not os.path.exists(path):
            from .. import build

            build.build_host()
        lib = ctypes.CDLL(path)
        i64p = ctypes.POINTER(ctypes.c_int64)
        lib.hs_assemble_count.restype = ctypes.c_int64
        lib.hs_assemble_count.argtypes = [ctypes.c_int, ctypes.c_int, i64p, ctypes.c_int64,
                                          ctypes.c_int64, i64p]
        lib.hs_assemble_fill.restype = ctypes.c_int
        lib.hs_assemble_fill.argtypes = [ctypes.c_int, ctypes.c_int, i64p, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p]
        lib.hs_set_num_threads.argtypes = [ctypes.c_int]
        lib.hs_get_max_threads.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def set_num_threads(n: int) -> None:
    """OpenMP threads of the assembly loops (launchers like torchrun export OMP_NUM_THREADS=1)."""
    _lib().hs_set_num_threads(int(n))


def num_threads() -> int:
    return int(_lib().hs_get_max_threads())


# ----------------------------------------------------------------------------------------------
# reference element
# ----------------------------------------------------------------------------------------------
def gauss_unit(nq: int):
    """Gauss-Legendre points/weights on [0,1] (dealii::QGauss<1>(nq))."""
    x, w = np.polynomial.legendre.leggauss(nq)
    return 0.5 * (x + 1.0), 0.5 * w


def lagrange_1d(degree: int, pts: np.ndarray):
    """Values and derivatives of the equidistant Lagrange basis of FE_Q(degree) (degree <= 2 has
    equidistant support points) at pts in [0,1].  Returns (shape[q,a], dshape[q,a])."""
    nodes = np.arange(degree + 1) / degree
    nq = len(pts)
    shape = np.ones((nq, degree + 1))
    dshape = np.zeros((nq, degree + 1))
    for a in range(degree + 1):
        for c in range(degree + 1):
            if c != a:
                shape[:, a] *= (pts - nodes[c]) / (nodes[a] - nodes[c])
        for e in range(degree + 1):
            if e == a:
                continue
            t = np.full(nq, 1.0 / (nodes[a] - nodes[e]))
            for c in range(degree + 1):
                if c != a and c != e:
                    t *= (pts - nodes[c]) / (nodes[a] - nodes[c])
            dshape[:, a] += t
    return shape, dshape


def reference_matrices(dim: int, degree: int, h) -> np.ndarray:
    """G[q, a, b] = (grad phi_a . grad phi_b)(x_q) JxW_q on a cell of size h (lexicographic local
    DoFs and quadrature points, x fastest): the per-quadrature-point pieces of the cell matrix of
    tests/laplace.hpp:186-194."""
    n1 = degree + 1
    qp, qw = gauss_unit(n1)
    S, D = lagrange_1d(degree, qp)
    h = list(h) + [1.0] * (3 - len(h))
    # per-direction factor tables: val[d][q,a], grad[d][q,a]
    vals = [S if d < dim else np.ones((1, 1)) for d in range(3)]
    grads = [D / h[d] if d < dim else np.zeros((1, 1)) for d in range(3)]
    wts = [qw * h[d] if d < dim else np.ones(1) for d in range(3)]
    nqd = [v.shape[0] for v in vals]
    nad = [v.shape[1] for v in vals]
    nq = nqd[0] * nqd[1] * nqd[2]
    ndof = nad[0] * nad[1] * nad[2]
    # grad component g of phi_a at q: product over directions, derivative in direction g
    gradphi = np.zeros((3, nq, ndof))
    for g in range(dim):
        fx = grads[0] if g == 0 else vals[0]
        fy = grads[1] if g == 1 else vals[1]
        fz = grads[2] if g == 2 else vals[2]
        # index order: q = qx + nqx*(qy + nqy*qz); a = ax + nax*(ay + nay*az)
        t = np.einsum("zc,yb,xa->zyxcba", fz, fy, fx)
        gradphi[g] = t.reshape(nq, ndof)
    jxw = np.einsum("z,y,x->zyx", wts[2], wts[1], wts[0]).reshape(nq)
    G = np.einsum("gqa,gqb,q->qab", gradphi, gradphi, jxw)
    return np.ascontiguousarray(G)


def quadrature_points(dim: int, degree: int, cells, h, origin=None) -> np.ndarray:
    """Physical quadrature points, shape (n_cells, nq, dim), cells and points lexicographic."""
    n1 = degree + 1
    qp, _ = gauss_unit(n1)
    axes = []
    for d in range(dim):
        o = 0.0 if origin is None else origin[d]
        c = o + np.arange(cells[d])[:, None] * h[d] + qp[None, :] * h[d]  # (cells_d, n1)
        axes.append(c)
    if dim == 2:
        X = np.broadcast_to(axes[0][None, :, None, :], (cells[1], cells[0], n1, n1))
        Y = np.broadcast_to(axes[1][:, None, :, None], (cells[1], cells[0], n1, n1))
        pts = np.stack([X, Y], axis=-1).reshape(cells[0] * cells[1], n1 * n1, 2)
    else:
        shp = (cells[2], cells[1], cells[0], n1, n1, n1)
        X = np.broadcast_to(axes[0][None, None, :, None, None, :], shp)
        Y = np.broadcast_to(axes[1][None, :, None, None, :, None], shp)
        Z = np.broadcast_to(axes[2][:, None, None, :, None, None], shp)
        pts = np.stack([X, Y, Z], axis=-1).reshape(cells[0] * cells[1] * cells[2], n1 ** 3, 3)
    return pts


# ----------------------------------------------------------------------------------------------
# material properties, tests/test_hierarchy_helpers.hpp:75-188
# ----------------------------------------------------------------------------------------------
def material_value(kind: str, pts: np.ndarray) -> np.ndarray:
    """kappa at points pts[..., dim]."""
    dim = pts.shape[-1]
    if kind == "constant":
        return np.ones(pts.shape[:-1])
    if kind == "linear_x":
        return 1.0 + np.abs(pts[..., 0])
    if kind == "linear":
        val = np.ones(pts.shape[:-1])
        for d in range(dim):
            val = val + (1.0 + d) * np.abs(pts[..., d])
        return val
    if kind == "discontinuous":
        odd = np.zeros(pts.shape[:-1], dtype=np.int64)
        for d in range(dim):
            odd += np.floor(pts[..., d] * 100).astype(np.int64) % 2
        return np.where(odd == dim, 100.0, 10.0)
    raise NotImplementedError(kind)


def coefficient_table(kind: str, dim: int, degree: int, cells, h=None, origin=None):
    """coef[cell, q] for the whole grid.  Returns a (n_cells, 1) array when the coefficient is
    the same at every quadrature point of every cell (detected, not assumed)."""
    cells = [int(c) for c in cells]
    h = [1.0 / c for c in cells] if h is None else list(h)
    if kind == "constant":
        return np.ones((int(np.prod(cells)), 1))
    pts = quadrature_points(dim, degree, cells, h, origin)
    coef = material_value(kind, pts)
    if np.all(coef == coef[:, :1]):
        return np.ascontiguousarray(coef[:, :1])
    return np.ascontiguousarray(coef)


def boundary_mask(dim: int, degree: int, cells, faces=None) -> np.ndarray:
    """uint8[n_nodes]: 1 on nodes of the selected boundary faces.  faces[d] = (low, high) booleans;
    default = the whole boundary (all faces get boundary id 1, tests/laplace.hpp:99-108, and
    interpolate_boundary_values on id 1, :139-140)."""
    nodes = [cells[d] * degree + 1 for d in range(dim)]
    if faces is None:
        faces = [(True, True)] * dim
    shape = tuple(reversed(nodes))
    m = np.zeros(shape, dtype=np.uint8)
    for d in range(dim):
        ax = dim - 1 - d
        lo, hi = faces[d]
        sl = [slice(None)] * dim
        if lo:
            sl[ax] = 0
            m[tuple(sl)] = 1
        if hi:
            sl[ax] = nodes[d] - 1
            m[tuple(sl)] = 1
    return m.reshape(-1)


# ----------------------------------------------------------------------------------------------
# assembly
# ----------------------------------------------------------------------------------------------
@dataclass
class HostCSR:
    """Host CSR in the upload layout: int64 row offsets, int32 column indices, f64 values."""
    n_rows: int
    n_cols: int
    rowptr: np.ndarray
    col: np.ndarray
    val: np.ndarray

    @property
    def nnz(self) -> int:
        return int(self.rowptr[-1])

    def to_scipy(self):
        import scipy.sparse as sp

        return sp.csr_matrix((self.val, self.col, self.rowptr), shape=(self.n_rows, self.n_cols))

    @staticmethod
    def from_scipy(m) -> "HostCSR":
        m = m.tocsr()
        return HostCSR(m.shape[0], m.shape[1], np.ascontiguousarray(m.indptr, dtype=np.int64),
                       np.ascontiguousarray(m.indices, dtype=np.int32),
                       np.ascontiguousarray(m.data, dtype=np.float64))


def assemble(dim: int, degree: int, cells, G: np.ndarray, coef: np.ndarray,
             constrained: np.ndarray, row_begin: int = 0, row_end: int | None = None):
    """Assemble rows [row_begin,row_end) of the global matrix.  Returns (HostCSR, diag)."""
    lib = _lib()
    cells_a = (ctypes.c_int64 * 3)(*([int(c) for c in cells] + [1] * (3 - dim)))
    n = 1
    for d in range(dim):
        n *= cells[d] * degree + 1
    if row_end is None:
        row_end = n
    nloc = row_end - row_begin
    rowptr = np.empty(nloc + 1, dtype=np.int64)
    nnz = lib.hs_assemble_count(dim, degree, cells_a, row_begin, row_end,
                                rowptr.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    col = np.empty(nnz, dtype=np.int32)
    val = np.empty(nnz, dtype=np.float64)
    diag = np.empty(nloc, dtype=np.float64)
    G = np.ascontiguousarray(G, dtype=np.float64)
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    constrained = np.ascontiguousarray(constrained, dtype=np.uint8)
    assert constrained.shape[0] == n
    assert coef.shape[0] == int(np.prod(cells[:dim]))
    rc = lib.hs_assemble_fill(dim, degree, cells_a, G.shape[0], G.ctypes.data, coef.ctypes.data,
                              coef.shape[1], constrained.ctypes.data, row_begin, row_end,
                              rowptr.ctypes.data, col.ctypes.data, val.ctypes.data,
                              diag.ctypes.data)
    if rc != 0:
        raise RuntimeError("hs_assemble_fill failed")
    return HostCSR(nloc, n, rowptr, col, val), diag


@dataclass
class LaplaceProblem:
    """The `Laplace<dim, VectorType>` test problem (tests/laplace.hpp) on a uniform grid of the
    unit cube: `cells` per direction, FE_Q(degree), material `kind`, homogeneous Dirichlet on the
    whole boundary."""
    dim: int
    degree: int
    cells: tuple
    material: str = "constant"
    h: tuple = field(default=None)
    G: np.ndarray = field(default=None, repr=False)
    coef: np.ndarray = field(default=None, repr=False)
    constrained: np.ndarray = field(default=None, repr=False)
    A: HostCSR = field(default=None, repr=False)
    diag: np.ndarray = field(default=None, repr=False)

    @property
    def nodes(self):
        return tuple(c * self.degree + 1 for c in self.cells)

    @property
    def n(self) -> int:
        return int(np.prod(self.nodes))

    @staticmethod
    def create(dim: int, degree: int, cells_per_dim: int, material: str = "constant",
               assemble_matrix: bool = True) -> "LaplaceProblem":
        cells = (cells_per_dim,) * dim
        p = LaplaceProblem(dim, degree, cells, material)
        p.h = tuple(1.0 / c for c in cells)
        p.G = reference_matrices(dim, degree, p.h)
        p.coef = coefficient_table(material, dim, degree, cells, p.h)
        p.constrained = boundary_mask(dim, degree, cells)
        if assemble_matrix:
            p.A, p.diag = assemble(dim, degree, cells, p.G, p.coef, p.constrained)
        return p

    @staticmethod
    def create_box(dim: int, degree: int, cells, h, material: str = "constant",
                   assemble_matrix: bool = True, origin=None, faces=None) -> "LaplaceProblem":
        """Same problem on a box of `cells` (per direction) cells of size `h`: the weak-scaling domains
        (cubes stacked along z) and the slab-local sub-boxes of the multi-GPU setup (`origin`: physical position
        of the box corner, for the material; `faces[d] = (low, high)`: which faces are Dirichlet -- cut planes
        of a slab are not)."""
        cells = tuple(int(c) for c in cells)
        p = LaplaceProblem(dim, degree, cells, material)
        p.h = tuple(float(x) for x in h)
        p.G = reference_matrices(dim, degree, p.h)
        p.coef = coefficient_table(material, dim, degree, cells, p.h, origin)
        p.constrained = boundary_mask(dim, degree, cells, faces)
        if assemble_matrix:
            p.A, p.diag = assemble(dim, degree, cells, p.G, p.coef, p.constrained)
        return p

    def coef_per_q(self) -> np.ndarray:
        """coef[cell, q] expanded to one value per quadrature point (matrix-free table layout,
        tests/laplace_matrix_free.hpp:100-119)."""
        nq = (self.degree + 1) ** self.dim
        if self.coef.shape[1] == nq:
            return self.coef
        return np.ascontiguousarray(np.repeat(self.coef, nq, axis=1))


def csr_from_dealii_sparse_matrix(rowstart, colnums, values, n_cols: int) -> HostCSR:
    """Upload layout from deal.II's `SparseMatrix` storage, the host half of `convert_matrix`
    (source/cuda/utils.cu:39-89).  deal.II keeps the rows of a SQUARE matrix with the diagonal entry FIRST and the
    other columns ascending; the reference moves the diagonal to its sorted position (:64-81) so that the device sees
    a regular CSR with ascending columns.  Rectangular matrices are already ascending and pass through."""
    rowstart = np.asarray(rowstart, dtype=np.int64)
    col = np.array(colnums, dtype=np.int32)
    val = np.array(values, dtype=np.float64)
    n_rows = len(rowstart) - 1
    if n_rows == n_cols:
        for row in range(n_rows):
            k0, k1 = int(rowstart[row]), int(rowstart[row + 1])
            if k1 - k0 < 2:
                continue
            if col[k0] != row:
                raise ValueError(f"row {row}: deal.II stores the diagonal first, found column {col[k0]}")
            d_val = val[k0]
            pos = 1
            while pos < k1 - k0 and col[k0 + pos] < row:       # utils.cu:71-76
                col[k0 + pos - 1] = col[k0 + pos]
                val[k0 + pos - 1] = val[k0 + pos]
                pos += 1
            col[k0 + pos - 1] = row
            val[k0 + pos - 1] = d_val
    return HostCSR(n_rows, n_cols, rowstart.copy(), col, val)
