"""GPU parity tests of the matrix-free fine-level operator through the C ABI:
against the oracle's cell-loop restatement of tests/laplace_matrix_free.hpp and against the assembled matrix
(tests/test_hierarchy.cc:644-695), and of the V-cycle / PCG with a matrix-free level 0 + assembled coarse level."""
import numpy as np
import pytest

import oracle
from helpers import rel_err, two_level_problem

pytestmark = pytest.mark.gpu

MF_CASES = [(2, 1, 40, "constant"), (2, 1, 33, "discontinuous"), (2, 2, 20, "linear"), (2, 2, 17, "discontinuous"),
            (3, 1, 40, "linear"), (3, 1, 12, "discontinuous"), (3, 1, 35, "constant"), (3, 2, 8, "linear_x"),
            (3, 2, 18, "discontinuous"), (3, 1, 3, "constant"), (3, 1, 70, "constant")]


def _mf(handle, P):
    from mfmg_b200 import device as d

    return d.MatrixFreeLaplaceDevice(handle, P.dim, P.degree, P.cells, P.h, P.coef_per_q(), P.constrained)


@pytest.mark.parametrize("dim,degree,cells,mat", MF_CASES)
def test_mf_apply_and_diagonal_vs_oracle(handle, dim, degree, cells, mat):
    from mfmg_b200 import device as d
    from mfmg_b200 import hostsetup as hs

    P = hs.LaplaceProblem.create(dim, degree, cells, mat)
    M = _mf(handle, P)
    assert M.size == P.n
    if dim == 3 and degree == 1:   # one coefficient for the whole grid: the factorised-stencil z-sweep serves it
        assert ("stencil" in M.kernel) == (mat == "constant"), M.kernel
    Mo = oracle.MatrixFreeLaplace(dim, degree, P.cells, P.h, P.coef_per_q(), P.constrained)
    rng = np.random.default_rng(3)
    x_h = rng.standard_normal(P.n)          # non-zero on constrained DoFs too: y_i = x_i there
    x, y = d.DeviceVector.from_host(handle, x_h), d.DeviceVector(handle, P.n)
    M.apply(x, y)
    y_ref = Mo.apply(x_h)
    assert rel_err(y.to_host(), y_ref) < 1e-12
    assert np.array_equal(y.to_host()[P.constrained != 0], x_h[P.constrained != 0])
    # == assembled matrix on vectors that vanish on constrained DoFs (tests/test_hierarchy.cc:684-694)
    x0 = x_h.copy()
    x0[P.constrained != 0] = 0.0
    x.upload(x0)
    M.apply(x, y)
    y_mat = oracle.spmv(P.n, P.A.rowptr, P.A.col, P.A.val, x0)
    free = P.constrained == 0
    assert np.linalg.norm((y.to_host() - y_mat)[free]) < 1e-9
    # diagonal: compute_diagonal semantics (constrained := 1)
    dg = M.diagonal().to_host()
    assert rel_err(dg, Mo.diag()) < 1e-12
    assert np.all(dg[~free] == 1.0)
    # deterministic: two applies are bitwise equal
    y2 = d.DeviceVector(handle, P.n)
    M.apply(x, y2)
    assert np.array_equal(y.to_host(), y2.to_host())


@pytest.mark.parametrize("dim,degree,cells,block,ne,mat", [(2, 1, 32, 4, 2, "constant"), (3, 1, 16, 4, 2, "linear"),
                                                           (3, 2, 6, 3, 1, "discontinuous"), (3, 1, 40, 8, 1, "constant")])
@pytest.mark.parametrize("nu,precond", [(1, True), (2, False)])
def test_vcycle_with_matrix_free_fine_level(handle, dim, degree, cells, block, ne, mat, nu, precond):
    from mfmg_b200 import device as d

    P, R, Ac = two_level_problem(dim, degree, cells, block, ne, mat)
    M = _mf(handle, P)
    ops = [M, d.SparseMatrixDevice.from_host(handle, Ac)]
    H = d.Hierarchy(handle, ops, [d.SparseMatrixDevice.from_host(handle, R)],
                    {"is preconditioner": precond, "smoother": {"n_smoothing_steps": nu}})
    Mo = oracle.MatrixFreeLaplace(dim, degree, P.cells, P.h, P.coef_per_q(), P.constrained)
    Ho = oracle.Hierarchy([Mo, (Ac.n_rows, Ac.rowptr, Ac.col, Ac.val)], [(R.n_rows, R.n_cols, R.rowptr, R.col, R.val)],
                          nu, precond)
    rng = np.random.default_rng(9)
    b_h, x_h = rng.standard_normal(P.n), rng.standard_normal(P.n)
    for graph in (False, True):
        H.use_graph(graph)
        b, x = d.DeviceVector.from_host(handle, b_h), d.DeviceVector.from_host(handle, x_h)
        H.vmult(x, b)
        assert rel_err(x.to_host(), Ho.vmult(b_h, x_h)) < 1e-12


def test_pcg_with_matrix_free_operator(handle):
    from mfmg_b200 import device as d

    P, R, Ac = two_level_problem(3, 1, 16, 4, 2, "discontinuous")
    M = _mf(handle, P)
    H = d.Hierarchy(handle, [M, d.SparseMatrixDevice.from_host(handle, Ac)],
                    [d.SparseMatrixDevice.from_host(handle, R)], {"is preconditioner": True})
    Mo = oracle.MatrixFreeLaplace(3, 1, P.cells, P.h, P.coef_per_q(), P.constrained)
    Ho = oracle.Hierarchy([Mo, (Ac.n_rows, Ac.rowptr, Ac.col, Ac.val)], [(R.n_rows, R.n_cols, R.rowptr, R.col, R.val)],
                          1, True)
    x0 = oracle.std_uniform01(P.n, skip=P.constrained)
    x_ref, it_ref, hist_ref = Ho.pcg(np.zeros(P.n), x0, 1e-8, 500)
    x, b = d.DeviceVector.from_host(handle, x0), d.DeviceVector.from_host(handle, np.zeros(P.n))
    it, hist = d.solver_cg(handle, None, x, b, H, 1e-8, 500)     # A == the hierarchy's matrix-free level-0 operator
    assert it == it_ref
    assert np.max(np.abs(hist - hist_ref) / hist_ref) < 1e-10


def test_mf_unsupported_degree_raises(handle):
    from mfmg_b200 import device as d

    with pytest.raises(d.NotImplementedExc):
        d.MatrixFreeLaplaceDevice(handle, 3, 3, (2, 2, 2), (0.5, 0.5, 0.5), np.ones((8, 64)), np.zeros(343, dtype=np.uint8))


@pytest.mark.parametrize("world,cells,block,mat", [(3, (33, 9, 12), (3, 3, 2), "linear"),
                                                   (2, (40, 20, 16), (4, 4, 4), "discontinuous"),
                                                   (3, (33, 9, 12), (3, 3, 2), "constant"),
                                                   (2, (70, 40, 16), (5, 4, 4), "constant")])
def test_mf_slab_operator_on_one_gpu(handle, world, cells, block, mat):
    """The z-slab form of the 3D Q1 operator (row-partitioned hierarchy): every rank's slab operator, fed with its
    [owned | ghost] copy of a global vector, reproduces the owned rows of the global operator and of its diagonal."""
    from helpers import slab_parts, slab_vector
    from mfmg_b200 import device as d
    from mfmg_b200 import hostsetup as hs

    h = (0.05, 0.04, 0.03)
    P = hs.LaplaceProblem.create_box(3, 1, cells, h, mat)
    Mo = oracle.MatrixFreeLaplace(3, 1, P.cells, P.h, P.coef_per_q(), P.constrained)
    rng = np.random.default_rng(1)
    x_h = rng.standard_normal(P.n)
    y_ref, d_ref = Mo.apply(x_h), Mo.diag()
    for part in slab_parts(world, cells, h, block, 1, 1, mat):
        mf = part.mf
        M = d.MatrixFreeLaplaceDevice(handle, 3, 1, mf["cells"], mf["h"], mf["coef"], mf["constrained"],
                                      own_planes=mf["own_planes"])
        assert M.size == part.n_owned and M.vector_size == part.n_owned + part.n_ghost
        assert ("stencil" in M.kernel) == (mat == "constant"), M.kernel
        x = d.DeviceVector.from_host(handle, slab_vector(part, x_h))
        y = d.DeviceVector(handle, part.n_owned)
        M.apply(x, y)
        sl = slice(part.row_begin, part.row_end)
        assert np.max(np.abs(y.to_host() - y_ref[sl])) <= 1e-12 * np.abs(y_ref).max()
        assert np.max(np.abs(M.diagonal().to_host() - d_ref[sl])) <= 1e-12 * np.abs(d_ref).max()


def test_generic_kernel_still_serves_3d_q1(handle, monkeypatch):
    """MFMGB_MF_GENERIC=1 keeps the colour-phase cell kernel reachable for 3D Q1 (it serves 2D and Q2 by default)."""
    from mfmg_b200 import device as d
    from mfmg_b200 import hostsetup as hs

    P = hs.LaplaceProblem.create(3, 1, 20, "discontinuous")
    fast = _mf(handle, P)
    monkeypatch.setenv("MFMGB_MF_GENERIC", "1")
    slow = _mf(handle, P)
    assert fast.kernel.startswith("Q1 node-owner") and slow.kernel.startswith("generic")
    rng = np.random.default_rng(0)
    x = d.DeviceVector.from_host(handle, rng.standard_normal(P.n))
    y1, y2 = d.DeviceVector(handle, P.n), d.DeviceVector(handle, P.n)
    fast.apply(x, y1)
    slow.apply(x, y2)
    assert rel_err(y1.to_host(), y2.to_host()) < 1e-13


@pytest.mark.parametrize("arith_flags", ["1", "0"])
@pytest.mark.parametrize("mat", ["constant"])
def test_stencil_sweep_equals_cell_kernel_and_fused_epilogues(handle, monkeypatch, mat, arith_flags):
    """The constant-coefficient stencil z-sweep against the per-cell kernel of the same operator (MFMGB_MF_STENCIL=0),
    and its fused residual / Jacobi epilogues against their unfused definitions; with the constraint flags computed
    from the box geometry (default when the constrained set is exactly the box faces) and loaded from the flag array
    (MFMGB_MF_ARITH_FLAGS=0: the form every other constrained set gets)."""
    import ctypes

    from mfmg_b200 import device as d
    from mfmg_b200 import hostsetup as hs

    monkeypatch.setenv("MFMGB_MF_ARITH_FLAGS", arith_flags)
    P = hs.LaplaceProblem.create_box(3, 1, (45, 31, 23), (0.02, 0.03, 0.05), mat)
    fast = _mf(handle, P)
    monkeypatch.setenv("MFMGB_MF_STENCIL", "0")
    cell = _mf(handle, P)
    assert "stencil" in fast.kernel and "per-cell" in cell.kernel
    rng = np.random.default_rng(4)
    x_h, b_h = rng.standard_normal(P.n), rng.standard_normal(P.n)
    x, b = d.DeviceVector.from_host(handle, x_h), d.DeviceVector.from_host(handle, b_h)
    y1, y2 = d.DeviceVector(handle, P.n), d.DeviceVector(handle, P.n)
    fast.apply(x, y1)
    cell.apply(x, y2)
    assert rel_err(y1.to_host(), y2.to_host()) < 1e-13
    # fused epilogues through a hierarchy's smoother path: one V-cycle of each operator agrees to 1e-12
    R = hs.build_restrictor(P, (5, 4, 4), 1)
    Ac = hs.galerkin(P.A, R)
    outs = []
    for M in (fast, cell):
        H = d.Hierarchy(handle, [M, d.SparseMatrixDevice.from_host(handle, Ac)],
                        [d.SparseMatrixDevice.from_host(handle, R)], {"is preconditioner": False,
                                                                      "smoother": {"n_smoothing_steps": 2}})
        xx = d.DeviceVector.from_host(handle, x_h)
        H.vmult(xx, b)
        outs.append(xx.to_host())
    assert rel_err(outs[0], outs[1]) < 1e-12
