"""CPU-only checks of the host setup path (problem generation, AMGe restrictor, Galerkin product)
against the oracle's independent restatement."""
import numpy as np
import pytest

import oracle
from mfmg_b200 import hostsetup as hs
from mfmg_b200.hostsetup import amge


@pytest.mark.parametrize("dim,degree,cells,mat", [(2, 1, 4, "constant"), (2, 2, 3, "linear"),
                                                  (3, 1, 3, "discontinuous"), (3, 2, 2, "linear_x")])
def test_assembly_matches_independent_cell_loop(dim, degree, cells, mat):
    P = hs.LaplaceProblem.create(dim, degree, cells, mat)
    dense = P.A.to_scipy().toarray()
    ref = oracle.assemble_laplace_py(dim, degree, list(P.cells), P.coef_per_q(), P.constrained)
    assert np.abs(dense - ref).max() <= 1e-14 * np.abs(ref).max()
    assert np.abs(dense - dense.T).max() == 0.0


def test_pattern_sizes_match_survey_appendix_a():
    # nnz = (3N-2)^d for Q1, (8c+1)^d for Q2: constrained rows/columns stay in the pattern
    P = hs.LaplaceProblem.create(2, 1, 32)
    assert (P.n, P.A.nnz) == (1089, 9409)
    P = hs.LaplaceProblem.create(2, 1, 64)
    assert (P.n, P.A.nnz) == (4225, 37249)
    P = hs.LaplaceProblem.create(3, 2, 3)
    assert P.A.nnz == (8 * 3 + 1) ** 3
    P = hs.LaplaceProblem.create(3, 1, 16)
    assert P.A.nnz == (3 * 17 - 2) ** 3


def test_row_range_assembly_equals_slices_of_global():
    P = hs.LaplaceProblem.create(3, 1, 6, "linear")
    n = P.n
    a, _ = hs.assemble(3, 1, P.cells, P.G, P.coef, P.constrained, 100, 250)
    full = P.A.to_scipy()
    assert np.array_equal(a.to_scipy().toarray(), full[100:250].toarray())
    assert a.n_cols == n


def test_block_agglomerates_cover_all_cells_once():
    aggs = amge.block_agglomerates(3, (6, 6, 4), (2, 3, 4))
    seen = np.zeros((4, 6, 6), dtype=int)
    for (ox, oy, oz), (sx, sy, sz) in aggs:
        seen[oz:oz + sz, oy:oy + sy, ox:ox + sx] += 1
    assert np.all(seen == 1) and len(aggs) == 3 * 2 * 1


def test_restrictor_entries_follow_the_formula():
    # tests/test_restriction_matrix.cc:157-167: R entries = diag_agg / diag_glob * eigvec
    P = hs.LaplaceProblem.create(2, 1, 8)
    R = hs.build_restrictor(P, (4, 4), 2)
    assert R.n_rows == 4 * 2 and R.n_cols == P.n and R.nnz == 8 * 25
    aggs = amge.block_agglomerates(2, P.cells, (4, 4))
    origin, size = aggs[3]
    g = amge._local_global_nodes(P, origin, size)
    cl = amge._local_cells(P, origin, size)
    vecs, d = amge.local_eigenvectors(P, P.coef[cl], P.constrained[g], size, 2)
    Rd = R.to_scipy().toarray()
    for k in range(2):
        assert np.allclose(Rd[3 * 2 + k, g], d / P.diag[g] * vecs[k], rtol=1e-14, atol=0)


def test_galerkin_operator_is_spd():
    P = hs.LaplaceProblem.create(3, 1, 8)
    R = hs.build_restrictor(P, (4, 4, 4), 2)
    Ac = hs.galerkin(P.A, R).to_scipy().toarray()
    assert np.allclose(Ac, Ac.T, atol=1e-13)
    assert np.linalg.eigvalsh(Ac).min() > 0


def test_partial_blocks_and_eigensolver_modes():
    P = hs.LaplaceProblem.create(2, 1, 10)
    for mode in ("free", "host_lapack", "device_lapack"):
        R = hs.build_restrictor(P, (4, 4), 1, eigensolver=mode)
        assert R.n_rows == 9 and np.all(np.isfinite(R.val))


def test_chunked_galerkin_is_bitwise_the_plain_product():
    """The blocked R (A R^T) used past scipy's int32 range (cfg3) equals the two-product form bit for bit."""
    P = hs.LaplaceProblem.create(3, 1, 16, "discontinuous")
    R = hs.build_restrictor(P, (4, 4, 4), 2)
    a = hs.galerkin(P.A, R)
    b = hs.galerkin(P.A, R, max_chunk_nnz=20000)
    assert np.array_equal(a.rowptr, b.rowptr) and np.array_equal(a.col, b.col) and np.array_equal(a.val, b.val)


def test_matrix_market_round_trip(tmp_path):
    """hostsetup.mmio writes / reads the two Matrix-Market forms mfmg produces through EpetraExt
    (source/dealii/dealii_utils.cc:63-91): bit-exact round trip, and scipy reads the same files."""
    import scipy.io

    P = hs.LaplaceProblem.create(2, 1, 6, "linear")
    R = hs.build_restrictor(P, (2, 2), 2)
    for name, M in (("A", P.A), ("R", R)):
        path = str(tmp_path / f"{name}.mtx")
        hs.mmio.write_matrix(path, M, comment="mfmg_b200 test")
        back = hs.mmio.read_matrix(path)
        assert (back.n_rows, back.n_cols, back.nnz) == (M.n_rows, M.n_cols, M.nnz)
        assert np.array_equal(back.rowptr, M.rowptr) and np.array_equal(back.col, M.col)
        assert np.array_equal(back.val, M.val)          # %22.16e round-trips doubles exactly
        assert abs(scipy.io.mmread(path).tocsr() - M.to_scipy()).max() == 0.0
    v = np.random.default_rng(0).standard_normal(17)
    hs.mmio.write_vector(str(tmp_path / "v.mtx"), v)
    assert np.array_equal(hs.mmio.read_vector(str(tmp_path / "v.mtx")), v)


@pytest.mark.parametrize("dense_limit", [1500, 10])
def test_local_eigenvectors_diagonal_kats(dense_limit):
    """tests/test_eigenvectors.cc:74-127 (diagonal) and :178-232 (diagonal_constraint) for the eigensolve core of the
    host setup: local matrix diag(1, 2, ..., 81) on the 9 x 9 Q1 nodes (hyper_cube refined 3 times); 5 lowest pairs =
    (i + 1, e_i); with DoF 0 constrained = (i + 2, e_{i+1}); eigenvector entries compared in absolute value, 1e-12.
    Both the dense branch and the shift-invert Lanczos branch (dense_limit = 10)."""
    import scipy.sparse as sp

    n, ne = 81, 5
    K = sp.diags(np.arange(1, n + 1, dtype=float)).tocsr()
    w, v = hs.lowest_eigenpairs(K, np.zeros(n, dtype=np.uint8), ne, shift=float(np.mean(K.diagonal())),
                                dense_limit=dense_limit)
    assert np.allclose(w, np.arange(1, ne + 1), rtol=0, atol=1e-10)
    ref = np.zeros((ne, n))
    ref[np.arange(ne), np.arange(ne)] = 1.0
    assert np.max(np.abs(np.abs(v) - ref)) < 1e-10
    c = np.zeros(n, dtype=np.uint8)
    c[0] = 1
    w, v = hs.lowest_eigenpairs(K, c, ne, shift=float(np.mean(K.diagonal())), dense_limit=dense_limit)
    assert np.allclose(w, np.arange(2, ne + 2), rtol=0, atol=1e-10)
    ref = np.zeros((ne, n))
    ref[np.arange(ne), np.arange(1, ne + 1)] = 1.0
    assert np.max(np.abs(np.abs(v) - ref)) < 1e-10


def test_block_agglomerates_match_reference_partition():
    """tests/test_agglomerate.cc:69-117 (simple_agglomerate_2d, one rank): 8 x 8 cells (hyper_cube refined 3 times),
    block partitioner nx = 2, ny = 3.  The reference lists the agglomerate id of every cell in deal.II's traversal
    order (Morton / z-order), ids numbered by first visit.  hostsetup.block_agglomerates numbers its blocks
    lexicographically (a permutation); the PARTITION -- which cells share an agglomerate, including
    the 3 + 3 + 2 split of the 8 cell rows -- must be the reference's."""
    ref = [1, 1, 1, 1, 2, 2, 2, 2, 1, 1, 3, 3, 2, 2, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6, 5, 5, 7, 7, 6, 6, 8, 8,
           3, 3, 3, 3, 4, 4, 4, 4, 9, 9, 9, 9, 10, 10, 10, 10, 7, 7, 7, 7, 8, 8, 8, 8, 11, 11, 11, 11, 12, 12, 12, 12]
    aggs = hs.block_agglomerates(2, (8, 8), (2, 3))
    assert len(aggs) == 12
    owner = {}
    for ia, ((ox, oy), (sx, sy)) in enumerate(aggs):
        for y in range(oy, oy + sy):
            for x in range(ox, ox + sx):
                owner[(x, y)] = ia

    def morton(k):   # deal.II child order on a refined hyper_cube: bit pairs (y, x) from the coarsest level down
        x = y = 0
        for level in range(3):
            d = (k >> (2 * (2 - level))) & 3
            x = 2 * x + (d & 1)
            y = 2 * y + (d >> 1)
        return x, y

    first_visit, got = {}, []
    for k in range(64):
        a = owner[morton(k)]
        first_visit.setdefault(a, len(first_visit) + 1)
        got.append(first_visit[a])
    assert got == ref


def test_block_agglomerates_3d_match_reference_partition():
    """tests/test_agglomerate.cc:120-286 (simple_agglomerate_3d, one rank): 8^3 cells, blocks 2 x 3 x 4; the reference's
    512 per-cell ids (tests/golden/kat.npz, transcribed by tests/golden/make_golden.py) in z-order, first-visit ids."""
    from helpers import golden

    ref = golden()["agglomerates_3d_ids"].tolist()
    aggs = hs.block_agglomerates(3, (8, 8, 8), (2, 3, 4))
    owner = {}
    for ia, ((ox, oy, oz), (sx, sy, sz)) in enumerate(aggs):
        for z in range(oz, oz + sz):
            for y in range(oy, oy + sy):
                for x in range(ox, ox + sx):
                    owner[(x, y, z)] = ia

    def morton(k):
        x = y = z = 0
        for level in range(3):
            d = (k >> (3 * (2 - level))) & 7
            x, y, z = 2 * x + (d & 1), 2 * y + ((d >> 1) & 1), 2 * z + (d >> 2)
        return x, y, z

    first_visit, got = {}, []
    for k in range(512):
        a = owner[morton(k)]
        first_visit.setdefault(a, len(first_visit) + 1)
        got.append(first_visit[a])
    assert got == ref


def test_multilevel_setup_and_oracle_convergence():
    """hostsetup.build_multilevel: Galerkin operators stay symmetric, the aggregation restrictor partitions the coarse
    DoFs, and the oracle's PCG iteration count grows by at most 2 from 2 to 4 levels (own convergence study for
    SURVEY 8f-2; unpinned by the reference)."""
    import oracle
    from mfmg_b200 import hostsetup as hs

    P = hs.LaplaceProblem.create(3, 1, 16, "constant")
    its = []
    for cbs in ([], [(2, 2, 2)], [(2, 2, 2), (2, 2, 2)]):
        ops, res = hs.build_multilevel(P, (2, 2, 2), 1, cbs)
        for r in res[1:]:
            assert np.all(np.diff(r.rowptr) > 0) and np.array_equal(np.sort(r.col), np.arange(r.n_cols))
        a = ops[-1].to_scipy()
        assert abs(a - a.T).max() < 1e-12 * abs(a).max()
        H = oracle.Hierarchy([(o.n_rows, o.rowptr, o.col, o.val) for o in ops],
                             [(r.n_rows, r.n_cols, r.rowptr, r.col, r.val) for r in res], 1, True)
        x0 = oracle.std_uniform01(P.n, skip=P.constrained)
        _, it, _ = H.pcg(np.zeros(P.n), x0, 1e-8, 200)
        assert it > 0
        its.append(it)
    assert its[2] <= its[0] + 2
