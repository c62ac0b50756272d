"""Host setup: spectral AMGe restriction operator and Galerkin coarse operator.

Restates the reference's HOST setup path for block agglomerates -- it is not on the hot path; it
produces the CSR operators (R, A_c) that are uploaded once:

  * block agglomeration             include/mfmg/common/amge.templates.hpp:412-499
  * agglomerate sub-problems: same bilinear form on the agglomerate's cells, Dirichlet only on
    faces of the GLOBAL boundary    include/mfmg/common/amge.templates.hpp:645-698
  * local eigenproblem: diag_agg = diag(K_agg); shift by mean(diag); constrained diagonal parked
    far away; n_e algebraically smallest eigenpairs, unit 2-norm
                                     include/mfmg/dealii/amge_host.templates.hpp:378-394,446-467
  * R[row(a,k), g_j] += diag_agg_a[j] / diag_A[g_j] * v_{a,k}[j], rows agglomerate-major,
    eigenvector-minor               include/mfmg/common/amge.templates.hpp:300-321
  * A_c = R (A R^T)                 include/mfmg/common/hierarchy.hpp:225,230

Deviations (documented in SURVEY.md appendix C): agglomerates are visited lexicographically instead
of in deal.II's z-order (a permutation of coarse rows); the eigenproblem is solved on the
unconstrained block directly instead of through the (-0.5, 100] window of the `lapack` branch,
which returns too few pairs when kappa is large.
"""
from __future__ import annotations

import itertools

import numpy as np

from .problems import HostCSR, LaplaceProblem, assemble


def block_agglomerates(dim: int, cells, block):
    """List of (origin_cell, size_in_cells) per agglomerate, lexicographic (x fastest)."""
    ranges = []
    for d in range(dim):
        nb = -(-cells[d] // block[d])
        ranges.append([(b * block[d], min(block[d], cells[d] - b * block[d])) for b in range(nb)])
    aggs = []
    if dim == 2:
        for (oy, sy), (ox, sx) in itertools.product(ranges[1], ranges[0]):
            aggs.append(((ox, oy), (sx, sy)))
    else:
        for (oz, sz), (oy, sy), (ox, sx) in itertools.product(ranges[2], ranges[1], ranges[0]):
            aggs.append(((ox, oy, oz), (sx, sy, sz)))
    return aggs


def _local_global_nodes(problem: LaplaceProblem, origin, size) -> np.ndarray:
    p = problem.degree
    nodes = problem.nodes
    dim = problem.dim
    ax = [np.arange(size[d] * p + 1) + origin[d] * p for d in range(dim)]
    if dim == 2:
        g = ax[0][None, :] + nodes[0] * ax[1][:, None]
    else:
        g = ax[0][None, None, :] + nodes[0] * (ax[1][None, :, None] + nodes[1] * ax[2][:, None, None])
    return g.reshape(-1)


def _local_cells(problem: LaplaceProblem, origin, size) -> np.ndarray:
    dim = problem.dim
    cells = problem.cells
    ax = [np.arange(size[d]) + origin[d] for d in range(dim)]
    if dim == 2:
        c = ax[0][None, :] + cells[0] * ax[1][:, None]
    else:
        c = ax[0][None, None, :] + cells[0] * (ax[1][None, :, None] + cells[1] * ax[2][:, None, None])
    return c.reshape(-1)


def local_eigenvectors(problem: LaplaceProblem, coef_loc: np.ndarray, constr_loc: np.ndarray,
                       size, n_eigenvectors: int, mode: str = "free", dense_limit: int = 1500):
    """(eigvecs[n_e, nloc], diag_agg[nloc]) for one agglomerate class.

    mode:
      "free"          n_e smallest eigenpairs of the unconstrained block (robust default)
      "host_lapack"   literal amge_host.templates.hpp:378-394,446-467: shift by mean(diag), constrained
                      diagonal := 200, dense symmetric solve, eigenvalues in (-0.5, 100], ascending
      "device_lapack" literal amge_device.templates.cuh:217-323: dense sygvd of the local matrix AS
                      ASSEMBLED (no shift, constrained rows keep their assembled diagonal), n_e smallest.
                      Constrained unit vectors can win here; this is what the device golds of
                      tests/test_hierarchy_device.cu:365-381 were recorded with.
    """
    dim, p = problem.dim, problem.degree
    Aloc, diag_agg = assemble(dim, p, size, problem.G, coef_loc, constr_loc)
    nloc = Aloc.n_rows
    vecs = np.zeros((n_eigenvectors, nloc))
    K = Aloc.to_scipy()
    if mode == "device_lapack":
        w, v = np.linalg.eigh(K.toarray())
        ne = min(n_eigenvectors, nloc)
        vecs[:ne] = v[:, :ne].T
        return vecs, diag_agg
    if mode == "host_lapack":
        M = K.toarray()
        M[np.diag_indices(nloc)] += np.mean(diag_agg)
        cidx = np.flatnonzero(constr_loc != 0)
        M[cidx, cidx] = 200.0
        w, v = np.linalg.eigh(M)
        keep = np.flatnonzero((w > -0.5) & (w <= 100.0))
        if len(keep) < n_eigenvectors:
            raise RuntimeError("lapack branch: only %d eigenvalues in (-0.5, 100]" % len(keep))
        vecs[:] = v[:, keep[:n_eigenvectors]].T
        return vecs, diag_agg
    w, v = lowest_eigenpairs(K, constr_loc, n_eigenvectors, shift=float(np.mean(diag_agg)), dense_limit=dense_limit)
    vecs[:v.shape[0]] = v
    return vecs, diag_agg


def lowest_eigenpairs(K, constrained, n_eigenvectors: int, shift: float = 0.0, dense_limit: int = 1500):
    """The n algebraically smallest eigenpairs of the local matrix K restricted to its unconstrained DoFs, the
    post-processed form AMGe_host::compute_local_eigenvectors returns (include/mfmg/dealii/amge_host.templates.hpp:
    378-394,446-475): eigenvalues ascending, eigenvectors of unit 2-norm with zeros at the constrained DoFs.
    Returns (eigenvalues[k], eigenvectors[k, n_loc]) with k = min(n_eigenvectors, number of free DoFs).
    Small blocks: dense symmetric solve; large ones: shift-invert Lanczos around 0 after adding `shift` to the
    diagonal (the reference shifts by mean(diag), :384-388; the eigenvalues returned are un-shifted)."""
    import scipy.sparse as sp

    K = sp.csr_matrix(K)
    constrained = np.asarray(constrained)
    nloc = K.shape[0]
    free = np.flatnonzero(constrained == 0)
    ne = min(n_eigenvectors, len(free))
    vecs = np.zeros((ne, nloc))
    if ne == 0:
        return np.zeros(0), vecs
    Kff = K[free][:, free]
    if len(free) <= dense_limit:
        w, v = np.linalg.eigh(Kff.toarray())
        w, v = w[:ne], v[:, :ne]
    else:
        import scipy.sparse.linalg as spla

        Ks = (Kff + shift * sp.identity(len(free))).tocsc()
        rng = np.random.default_rng(0)
        w, v = spla.eigsh(Ks, k=ne, sigma=0.0, which="LM", v0=rng.standard_normal(len(free)),
                          tol=1e-13)
        order = np.argsort(w)
        w, v = w[order] - shift, v[:, order]
    # deterministic sign: make the entry of largest magnitude positive
    for k in range(ne):
        j = np.argmax(np.abs(v[:, k]))
        if v[j, k] < 0:
            v[:, k] = -v[:, k]
        vecs[k, free] = v[:, k] / np.linalg.norm(v[:, k])
    return w, vecs


def restriction_from_local(eigenvectors, diag_elements, dof_indices_maps, n_local_eigenvectors, global_diag,
                           n_cols: int) -> HostCSR:
    """AMGe::compute_restriction_sparse_matrix with its own argument list
    (include/mfmg/common/amge.templates.hpp:271-324): row `pos` (agglomerate-major, eigenvector-minor) gets
    `diag_elements[i][j] / global_diag[g] * eigenvectors[pos][j]` ADDED at column g = dof_indices_maps[i][j].
    Columns come out ascending; repeated columns within a row are summed (Trilinos `add` + `compress(add)`)."""
    import scipy.sparse as sp

    rows, cols, vals = [], [], []
    pos = 0
    for i, n_eig in enumerate(n_local_eigenvectors):
        g = np.asarray(dof_indices_maps[i], dtype=np.int64)
        w = np.asarray(diag_elements[i], dtype=np.float64) / np.asarray(global_diag, dtype=np.float64)[g]
        for _ in range(int(n_eig)):
            v = np.asarray(eigenvectors[pos], dtype=np.float64)
            if len(v) != len(g):
                raise ValueError(f"dof_indices_maps[{i}] has the wrong size: {len(g)} instead of {len(v)}")
            rows.append(np.full(len(g), pos, dtype=np.int64))
            cols.append(g)
            vals.append(w * v)
            pos += 1
    if pos == 0:
        return HostCSR(0, n_cols, np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int32), np.zeros(0))
    m = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(pos, n_cols)).tocsr()
    m.sum_duplicates()
    m.sort_indices()
    return HostCSR.from_scipy(m)


def build_restrictor(problem: LaplaceProblem, block, n_eigenvectors: int,
                     eigensolver: str = "free") -> HostCSR:
    """The restriction matrix R (n_c x n) of AMGe::setup_restrictor for block agglomerates."""
    dim = problem.dim
    aggs = block_agglomerates(dim, problem.cells, block)
    diag_A = problem.diag
    classes: dict = {}
    members: dict = {}
    for ia, (origin, size) in enumerate(aggs):
        cells_loc = _local_cells(problem, origin, size)
        g = _local_global_nodes(problem, origin, size)
        coef_loc = problem.coef[cells_loc]
        constr_loc = problem.constrained[g]
        key = (tuple(size), coef_loc.tobytes(), constr_loc.tobytes())
        if key not in classes:
            classes[key] = local_eigenvectors(problem, np.ascontiguousarray(coef_loc),
                                              np.ascontiguousarray(constr_loc), size,
                                              n_eigenvectors, mode=eigensolver)
            members[key] = []
        members[key].append((ia, g))

    n_agg = len(aggs)
    ne = n_eigenvectors
    nloc_of = np.zeros(n_agg, dtype=np.int64)
    for key, lst in members.items():
        for ia, g in lst:
            nloc_of[ia] = len(g)
    # row r = ia*ne + k has nloc_of[ia] entries
    row_len = np.repeat(nloc_of, ne)
    rowptr = np.zeros(n_agg * ne + 1, dtype=np.int64)
    np.cumsum(row_len, out=rowptr[1:])
    nnz = int(rowptr[-1])
    col = np.empty(nnz, dtype=np.int32)
    val = np.empty(nnz, dtype=np.float64)
    for key, lst in members.items():
        vecs, diag_agg = classes[key]
        for ia, g in lst:
            order = np.argsort(g, kind="stable")
            gs = g[order]
            w = diag_agg[order] / diag_A[gs]
            for k in range(ne):
                r = ia * ne + k
                col[rowptr[r]:rowptr[r + 1]] = gs
                val[rowptr[r]:rowptr[r + 1]] = w * vecs[k, order]
    return HostCSR(n_agg * ne, problem.n, rowptr, col, val)


def galerkin(A: HostCSR, R: HostCSR, max_chunk_nnz: int = 400_000_000) -> HostCSR:
    """A_c = R (A R^T): include/mfmg/common/hierarchy.hpp:225,230 (no dropping).

    Small operators: two sparse products.  Large ones (cfg3: 3.6e9 non-zeros, past scipy's int32 index range and
    ~100 GB of temporaries) are processed in blocks of coarse rows: a block of R only touches a contiguous range
    of fine rows, so it is multiplied with a zero-copy row slice of A and then with R^T."""
    import scipy.sparse as sp

    r = R.to_scipy()
    if A.nnz <= max_chunk_nnz:
        a = A.to_scipy()
        ap = a @ r.T.tocsr()
        ac = (r @ ap).tocsr()
        ac.sort_indices()
        return HostCSR.from_scipy(ac)
    rt = r.T.tocsr()
    ac = galerkin_rows(A, R, rt, 0, R.n_rows, max_chunk_nnz)
    return HostCSR.from_scipy(ac)


def galerkin_rows(A: HostCSR, R: HostCSR, rt, c0: int, c1: int, max_chunk_nnz: int = 400_000_000):
    """Rows [c0, c1) of R (A R^T) as a scipy CSR matrix (sorted columns); `rt` = R^T in CSR form.  Works in blocks
    of coarse rows on zero-copy row slices of A, so A may hold more than 2^31 non-zeros."""
    import scipy.sparse as sp

    n_c = R.n_rows
    span_nnz = int(A.rowptr[-1])
    n_chunks = int(min(max(c1 - c0, 1), max(1, -(-span_nnz // (max_chunk_nnz // 4)))))
    bounds = np.linspace(c0, c1, n_chunks + 1).astype(np.int64)
    blocks = []
    for b0, b1 in zip(bounds[:-1], bounds[1:]):
        if b1 <= b0:
            continue
        k0, k1 = int(R.rowptr[b0]), int(R.rowptr[b1])
        if k1 == k0:
            blocks.append(sp.csr_matrix((int(b1 - b0), n_c)))
            continue
        cols = R.col[k0:k1]
        p0, p1 = int(cols.min()), int(cols.max()) + 1
        a0, a1 = int(A.rowptr[p0]), int(A.rowptr[p1])
        if a1 - a0 >= 2 ** 31 - 1:
            raise MemoryError("galerkin: a block of coarse rows spans more than 2^31 non-zeros of A; "
                              "lower max_chunk_nnz")
        a_sub = sp.csr_matrix((A.val[a0:a1], A.col[a0:a1], (A.rowptr[p0:p1 + 1] - a0).astype(np.int32)),
                              shape=(p1 - p0, A.n_cols), copy=False)
        r_sub = sp.csr_matrix((R.val[k0:k1], (cols - p0).astype(np.int32),
                               (R.rowptr[b0:b1 + 1] - k0).astype(np.int32)), shape=(int(b1 - b0), p1 - p0))
        blocks.append((r_sub @ (a_sub @ rt)).tocsr())  # the reference's association: R (A R^T)
    ac = sp.vstack(blocks, format="csr") if blocks else sp.csr_matrix((0, n_c))
    ac.sort_indices()
    return ac


def transpose(R: HostCSR) -> HostCSR:
    """Explicit transpose with ascending column indices (stable counting sort)."""
    t = R.to_scipy().T.tocsr()
    t.sort_indices()
    return HostCSR.from_scipy(t)


def aggregate_restrictor(agg_grid, block, n_eigenvectors: int = 1) -> HostCSR:
    """Restrictor of a FURTHER level ("max levels" > 2 -- which the reference itself cannot
    build: include/mfmg/common/hierarchy.hpp:209-210 re-uses the fine evaluator).  The coarse DoFs of a structured
    block agglomeration are numbered (agglomerate, eigenvector) with lexicographic agglomerates `agg_grid`; this groups
    `block` agglomerates per direction and sums DoFs of equal eigenvector index (unsmoothed aggregation: entries 1).
    Rows = (super-agglomerate, eigenvector), columns = the coarse DoFs of the level above.  Unpinned by the reference:
    the oracle's generic multi-level cycle is the contract for the device path, convergence is reported separately."""
    agg_grid = tuple(int(a) for a in agg_grid)
    block = tuple(int(b) for b in block)
    dim = len(agg_grid)
    sup = tuple(-(-a // b) for a, b in zip(agg_grid, block))
    idx = np.indices(agg_grid[::-1])[::-1]                 # idx[d][..., j, i] = coordinate d (x fastest in memory)
    sup_id = np.zeros(agg_grid[::-1], dtype=np.int64)
    stride = 1
    for d in range(dim):
        sup_id += (idx[d] // block[d]) * stride
        stride *= sup[d]
    sup_id = sup_id.reshape(-1)                            # per agglomerate, lexicographic
    n_agg, n_sup, ne = sup_id.size, int(np.prod(sup)), n_eigenvectors
    rows = (sup_id[:, None] * ne + np.arange(ne)[None, :]).reshape(-1)
    cols = np.arange(n_agg * ne, dtype=np.int64)
    import scipy.sparse as sp

    m = sp.csr_matrix((np.ones(n_agg * ne), (rows, cols)), shape=(n_sup * ne, n_agg * ne))
    m.sort_indices()
    return HostCSR.from_scipy(m)


def build_multilevel(problem: LaplaceProblem, block, n_eigenvectors: int, coarse_blocks, eigensolver: str = "free"):
    """Operators [A_0, A_1, ...] and restrictors [R_1, R_2, ...] of a hierarchy with len(coarse_blocks) + 2 levels:
    spectral AMGe from the mesh for the first transition (build_restrictor), aggregate_restrictor for the others,
    Galerkin operators A_{l+1} = R (A_l R^T) throughout (hierarchy.hpp:225,230)."""
    R = build_restrictor(problem, block, n_eigenvectors, eigensolver=eigensolver)
    ops, res = [problem.A], [R]
    ops.append(galerkin(problem.A, R))
    grid = tuple(-(-c // b) for c, b in zip(problem.cells, block))
    for cb in coarse_blocks:
        R2 = aggregate_restrictor(grid, cb, n_eigenvectors)
        assert R2.n_cols == ops[-1].n_rows
        res.append(R2)
        ops.append(galerkin(ops[-1], R2))
        grid = tuple(-(-g // b) for g, b in zip(grid, cb))
    return ops, res
