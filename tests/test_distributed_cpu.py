"""CPU (gloo, world_size 2 and 3) tests of the multi-GPU host logic: slab partition, [owned | ghost] renumbering, halo
plan wiring, coarse slices.  The partitioned V-cycle (oracle kernels + gloo exchange) must reproduce the serial
oracle V-cycle on every rank's rows."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from helpers import oracle_hierarchy, two_level_problem


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, out_q):
    import torch.distributed as dist

    import oracle
    from dist_emulation import dist_vcycle_cpu
    from mfmg_b200 import hostsetup as hs

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dim, degree, cells, block, ne, mat = case
        P, R, Ac = two_level_problem(dim, degree, cells, block, ne, mat)
        (part,), row_off, coarse_off = hs.make_parts(P, R, Ac, (block,) * dim, ne, world, ranks=[rank])
        # the plan must agree with what the other ranks computed: exchange ghost lists for real and re-wire
        lists = [None] * world
        dist.all_gather_object(lists, part.ghost_global)
        nb, send = hs.partition.wire_send_lists(lists, row_off, rank)
        assert sorted(nb) == [q for q, s in zip(part.neighbors, part.send_indices) if len(s)]
        rng = np.random.default_rng(11)
        b = rng.standard_normal(P.n)
        lu, piv, info = oracle.lu_factor_csr(Ac.n_rows, Ac.rowptr, Ac.col, Ac.val)
        x_loc = dist_vcycle_cpu(part, lu, piv, b[part.row_begin:part.row_end])
        x_ref = oracle_hierarchy(P, R, Ac, 1, True).vmult(b)[part.row_begin:part.row_end]
        err = np.linalg.norm(x_loc - x_ref) / np.linalg.norm(x_ref)
        # the same cycle with the domain-decomposed coarse solve (one all-reduce, nothing gathered)
        plan = hs.coarse_dd_plan(part.Ac, part.coarse_offsets, rank)
        if plan is not None:
            x_dd = dist_vcycle_cpu(part, lu, piv, b[part.row_begin:part.row_end], dd_plan=plan)
            err = max(err, np.linalg.norm(x_dd - x_ref) / np.linalg.norm(x_ref))
        out_q.put((rank, float(err), part.n_owned, part.n_ghost, part.boundary_lo, part.boundary_hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,case", [(2, (3, 1, 8, 2, 1, "constant")), (2, (3, 2, 4, 2, 2, "linear")),
                                        (3, (3, 1, 12, 2, 2, "discontinuous")), (2, (2, 1, 16, 4, 2, "constant"))])
def test_partitioned_vcycle_matches_serial(world, case):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    results = sorted(q.get() for _ in range(world))
    for rank, err, n_owned, n_ghost, blo, bhi in results:
        assert err < 1e-13, (rank, err)
        assert n_ghost > 0 and 0 <= blo <= bhi <= n_owned
        assert (bhi - blo) > 0  # there is an interior to overlap the exchange with
    assert sum(r[2] for r in results) == int(np.prod([case[2] * case[1] + 1] * case[0]))


def test_partition_properties_single_process():
    from mfmg_b200 import hostsetup as hs

    P, R, Ac = two_level_problem(3, 1, 8, 2, 2, "constant")
    parts, row_off, coarse_off = hs.make_parts(P, R, Ac, (2, 2, 2), 2, 4)
    assert row_off[0] == 0 and row_off[-1] == P.n and coarse_off[-1] == R.n_rows
    rng = np.random.default_rng(0)
    x = rng.standard_normal(P.n)
    y = P.A.to_scipy() @ x
    for p in parts:
        # local SpMV with exact ghost values reproduces the global rows bit for bit
        xl = np.concatenate([x[p.row_begin:p.row_end], x[p.ghost_global]])
        yl = p.A.to_scipy() @ xl
        assert np.array_equal(yl, y[p.row_begin:p.row_end])
        # rows outside [boundary_lo, boundary_hi) are exactly the rows that touch ghosts
        cols_interior = p.A.col[p.A.rowptr[p.boundary_lo]:p.A.rowptr[p.boundary_hi]]
        assert np.all(cols_interior < p.n_owned)
        # send lists: what I send to q is what q expects from me, in q's ghost order
        for q, sidx in zip(p.neighbors, p.send_indices):
            other = parts[q]
            expect = other.ghost_global[(other.ghost_global >= p.row_begin) & (other.ghost_global < p.row_end)]
            assert np.array_equal(sidx + p.row_begin, expect)
        # slabs send contiguous planes (no packing kernel needed)
        for sidx in p.send_indices:
            if len(sidx):
                assert np.array_equal(sidx, np.arange(sidx[0], sidx[0] + len(sidx)))


@pytest.mark.parametrize("world,cells,block,ne,degree,mat", [(2, (6, 6, 8), (2, 2, 2), 2, 1, "constant"),
                                                             (3, (4, 6, 12), (2, 3, 2), 1, 1, "linear"),
                                                             (2, (4, 4, 8), (2, 2, 4), 2, 2, "discontinuous"),
                                                             (1, (4, 4, 4), (2, 2, 2), 2, 1, "constant")])
def test_slab_local_setup_equals_sliced_global(world, cells, block, ne, degree, mat):
    """hostsetup.build_slab_part (each rank builds only its slab + one agglomerate layer) reproduces the parts
    obtained by slicing the globally built operators."""
    from mfmg_b200 import hostsetup as hs

    h = (0.05, 0.04, 0.03)
    P = hs.LaplaceProblem.create_box(3, degree, cells, h, mat)
    R = hs.build_restrictor(P, block, ne)
    Ac = hs.galerkin(P.A, R)
    ref_parts, row_off, coarse_off = hs.make_parts(P, R, Ac, block, ne, world)

    # emulate the setup-time gathers: two passes (the first collects what every rank would contribute)
    contributions = {}

    def run(rank, record):
        calls = [0]

        def gather(obj):
            k = calls[0]
            calls[0] += 1
            if record:
                contributions.setdefault(k, {})[rank] = obj
                return [obj] * world           # placeholder, results of this pass are discarded
            return [contributions[k][r] for r in range(world)]

        return hs.build_slab_part(degree, cells, h, mat, block, ne, world, rank, gather)

    for r in range(world):
        run(r, True)
    for r in range(world):
        part, ref = run(r, False), ref_parts[r]
        assert (part.row_begin, part.row_end, part.n_ghost) == (ref.row_begin, ref.row_end, ref.n_ghost)
        assert np.array_equal(part.ghost_global, ref.ghost_global)
        for name in ("A", "R", "P"):
            a, b = getattr(part, name), getattr(ref, name)
            assert (a.n_rows, a.n_cols) == (b.n_rows, b.n_cols), name
            assert np.array_equal(a.rowptr, b.rowptr) and np.array_equal(a.col, b.col), name
            assert np.allclose(a.val, b.val, rtol=1e-12, atol=1e-15), name
        da, db = part.Ac.to_scipy().toarray(), ref.Ac.to_scipy().toarray()
        assert np.allclose(da, db, rtol=1e-11, atol=1e-13 * np.abs(db).max())
        assert part.neighbors == ref.neighbors and part.recv_counts == ref.recv_counts
        for s1, s2 in zip(part.send_indices, ref.send_indices):
            assert np.array_equal(s1, s2)
        assert (part.boundary_lo, part.boundary_hi) == (ref.boundary_lo, ref.boundary_hi)


@pytest.mark.parametrize("world,cells,block,mat", [(2, (6, 5, 8), (2, 2, 2), "linear"), (3, (4, 6, 12), (2, 3, 2), "discontinuous"),
                                                   (1, (4, 4, 4), (2, 2, 2), "constant")])
def test_matrix_free_slab_data(world, cells, block, mat):
    """hostsetup.slab attaches the matrix-free level-0 data of each rank (local box, coefficient rows, constraint flags
    in [owned | ghost below | ghost above] order).  Checked with the oracle's cell loop on the local box: its rows
    for the owned nodes, fed with the rank's [owned | ghost] copy of a global vector, equal the global operator's."""
    import oracle
    from helpers import slab_parts, slab_vector
    from mfmg_b200 import hostsetup as hs

    h = (0.05, 0.04, 0.03)
    P = hs.LaplaceProblem.create_box(3, 1, cells, h, mat)
    Mo = oracle.MatrixFreeLaplace(3, 1, P.cells, P.h, P.coef_per_q(), P.constrained)
    rng = np.random.default_rng(1)
    x = rng.standard_normal(P.n)
    y_ref = Mo.apply(x)
    plane = (cells[0] + 1) * (cells[1] + 1)
    for part in slab_parts(world, cells, h, block, 1, 1, mat):
        mf = part.mf
        assert mf is not None
        own0, own1 = mf["own_planes"]
        nz = mf["cells"][2] + 1
        assert (own1 - own0) * plane == part.n_owned and nz * plane == part.n_owned + part.n_ghost
        # vector layout -> natural plane order of the local box
        order = list(range(own0, own1)) + list(range(0, own0)) + list(range(own1, nz))
        nat = np.empty(nz * plane, dtype=np.int64)   # nat[natural index] = vector index
        for pos, g in enumerate(order):
            nat[g * plane:(g + 1) * plane] = np.arange(pos * plane, (pos + 1) * plane)
        Ml = oracle.MatrixFreeLaplace(3, 1, mf["cells"], mf["h"], mf["coef"], mf["constrained"][nat])
        xl = slab_vector(part, x)
        yl = Ml.apply(xl[nat])
        got = yl[own0 * plane:own1 * plane]
        assert np.max(np.abs(got - y_ref[part.row_begin:part.row_end])) <= 1e-12 * np.abs(y_ref).max()
        dl = Ml.diag()[own0 * plane:own1 * plane]
        assert np.max(np.abs(dl - Mo.diag()[part.row_begin:part.row_end])) <= 1e-12 * np.abs(Mo.diag()).max()


@pytest.mark.parametrize("world,cells,block,ne", [(2, (4, 4, 8), (2, 2, 2), 1), (3, (6, 4, 18), (2, 2, 3), 2), (4, (4, 4, 8), (2, 2, 2), 1)])
def test_coarse_dd_plan_block_elimination(world, cells, block, ne):
    """hostsetup.coarse_dd_plan: the interior / separator blocks it hands to csrc/coarse_dd.cu reproduce A_c^-1 b by
    block elimination (numpy restatement of the device algorithm), and the prolongation rows of every rank only read
    coarse entries the solver leaves valid there."""
    from helpers import slab_parts
    from mfmg_b200 import hostsetup as hs

    h = (0.1, 0.1, 0.1)
    parts = slab_parts(world, cells, h, block, ne, 1, "linear")
    Ac = parts[0].Ac.to_scipy().toarray()
    n_c = Ac.shape[0]
    rng = np.random.default_rng(0)
    b = rng.standard_normal(n_c)
    x_ref = np.linalg.solve(Ac, b)
    plans = [hs.coarse_dd_plan(p.Ac, p.coarse_offsets, p.rank) for p in parts]
    assert all(pl is not None for pl in plans)
    n_S, sep = plans[0]["n_S"], plans[0]["sep_index"]
    S = plans[0]["A_SS"].to_scipy().toarray().copy()
    E, Minv = [], []
    for pl in plans:
        AII, AIS, ASI = (pl[k].to_scipy().toarray() for k in ("A_II", "A_IS", "A_SI"))
        Mi = np.linalg.inv(AII) if AII.shape[0] else AII
        a0, na = pl["adj_begin"], AIS.shape[1]
        S[a0:a0 + na, a0:a0 + na] -= ASI @ (Mi @ AIS)
        E.append(Mi @ AIS)
        Minv.append(Mi)
    t, ys = np.zeros(n_S), []
    for pl, Mi in zip(plans, Minv):
        y = Mi @ b[pl["own_begin"]:pl["own_begin"] + pl["n_I"]]
        ys.append(y)
        a0, na = pl["adj_begin"], pl["A_IS"].n_cols
        t[a0:a0 + na] -= pl["A_SI"].to_scipy() @ y
        o0, on = pl["own_sep_begin"], pl["own_sep_n"]
        t[o0:o0 + on] += b[sep[o0:o0 + on]]
    xs = np.linalg.solve(S, t)
    x = np.zeros(n_c)
    x[sep] = xs
    for pl, y, Er in zip(plans, ys, E):
        a0, na = pl["adj_begin"], pl["A_IS"].n_cols
        x[pl["own_begin"]:pl["own_begin"] + pl["n_I"]] = y - Er @ xs[a0:a0 + na]
    assert np.linalg.norm(x - x_ref) <= 1e-12 * np.linalg.norm(x_ref)
    # the two-stream form of csrc/coarse_dd.cu: W = A_SI A_II^-1 formed at setup, so the separator right-hand side
    # t = b_S - sum_r W_r b_I,r needs the restricted residual only (it runs NEXT TO y = A_II^-1 b_I on the device)
    t2 = np.zeros(n_S)
    for pl, Mi in zip(plans, Minv):
        W = pl["A_SI"].to_scipy().toarray() @ Mi
        a0, na = pl["adj_begin"], pl["A_IS"].n_cols
        t2[a0:a0 + na] -= W @ b[pl["own_begin"]:pl["own_begin"] + pl["n_I"]]
        o0, on = pl["own_sep_begin"], pl["own_sep_n"]
        t2[o0:o0 + on] += b[sep[o0:o0 + on]]
    assert np.linalg.norm(t2 - t) <= 1e-13 * np.linalg.norm(t)
    xs2 = np.linalg.solve(S, t2)
    x2 = np.zeros(n_c)
    x2[sep] = xs2
    for pl, y, Er in zip(plans, ys, E):
        a0, na = pl["adj_begin"], pl["A_IS"].n_cols
        x2[pl["own_begin"]:pl["own_begin"] + pl["n_I"]] = y - Er @ xs2[a0:a0 + na]
    assert np.linalg.norm(x2 - x_ref) <= 1e-12 * np.linalg.norm(x_ref)
    for p, pl in zip(parts, plans):
        assert np.all(np.isin(p.P.col, pl["valid_cols"]))


def test_coarse_dd_plan_rejects_operators_that_are_not_block_tridiagonal():
    """The plan is only offered when interiors of different ranks are decoupled once the separators are removed;
    otherwise from_partition falls back to the dense inverse.  All ranks must reach the same decision."""
    import scipy.sparse as sp

    from mfmg_b200 import hostsetup as hs

    n, world = 40, 4
    co = np.array([0, 10, 20, 30, 40])
    tri = sp.diags([np.ones(n - 1), 4 * np.ones(n), np.ones(n - 1)], [-1, 0, 1]).tocsr()
    ok = [hs.coarse_dd_plan(hs.HostCSR.from_scipy(tri), co, r) for r in range(world)]
    assert all(p is not None for p in ok)
    assert [p["n_S"] for p in ok] == [3] * world and ok[0]["n_sep_below"] == 0 and ok[2]["n_sep_below"] == 1
    far = tri.tolil()
    far[2, 35] = far[35, 2] = 0.5                     # rank 0 coupled to rank 3: not block tridiagonal
    bad = [hs.coarse_dd_plan(hs.HostCSR.from_scipy(far.tocsr()), co, r) for r in range(world)]
    assert all(p is None for p in bad)
    assert hs.coarse_dd_plan(hs.HostCSR.from_scipy(tri), np.array([0, n]), 0) is None   # one rank: nothing to do


def test_rows_restricted():
    from mfmg_b200 import hostsetup as hs
    from mfmg_b200.hostsetup.partition import rows_restricted

    rng = np.random.default_rng(0)
    import scipy.sparse as sp

    M = sp.random(12, 30, density=0.3, random_state=1, format="csr")
    M.sort_indices()
    H = hs.HostCSR.from_scipy(M)
    sub = rows_restricted(H, 3, 9, 10, 22, 40, pad_rows_before=2, n_rows=10)
    ref = np.zeros((10, 40))
    ref[2:8, :12] = M.toarray()[3:9, 10:22]
    assert (sub.n_rows, sub.n_cols) == (10, 40) and np.array_equal(sub.to_scipy().toarray(), ref)


@pytest.mark.parametrize("world", [2, 3, 4])
def test_distributed_mv_kat_halo_plan(world):
    """The reference's distributed SpMV KAT (tests/test_sparse_matrix_device.cu:116-223) on the host-side plan: 10 rows
    per rank, five columns per row drawn from std::default_random_engine(local row) over ALL global columns (so every
    rank is every other rank's neighbour), value i + j with set() semantics, x = local index.  The reference
    all-gathers the source vector (sparse_matrix_device.templates.cuh:104-138); here the rows are renumbered
    [owned | ghost] and only the listed entries travel.  Checked: ghost order == send-list order for every pair of
    ranks, and local SpMV on [owned | received] == the global product, bit for bit."""
    import oracle
    from mfmg_b200.hostsetup.partition import _localise, wire_send_lists
    from mfmg_b200.hostsetup.problems import HostCSR

    n_local = 10
    size = world * n_local
    rowptr, col, val = [0], [], []
    for r in range(world):
        for i in range(n_local):
            cols = oracle.std_uniform_int(5, 0, size - 1, seed=i)
            entries = {}
            for j, c in enumerate(cols):
                entries[int(c)] = float(i + j)          # set(): the last value written to an entry stays
            for c in sorted(entries):
                col.append(c)
                val.append(entries[c])
            rowptr.append(len(col))
    A = HostCSR(size, size, np.array(rowptr, dtype=np.int64), np.array(col, dtype=np.int32), np.array(val))
    x = (np.arange(size) % n_local).astype(np.float64)      # vector.local_element(i) = i on every rank
    y_ref = oracle.spmv(size, A.rowptr, A.col, A.val, x)
    assert np.array_equal(y_ref, A.to_scipy() @ x)
    offsets = np.arange(world + 1) * n_local
    ghosts, locals_ = [], []
    for r in range(world):
        rb, re_ = int(offsets[r]), int(offsets[r + 1])
        c = A.col[A.rowptr[rb]:A.rowptr[re_]].astype(np.int64)
        g = np.unique(c[(c < rb) | (c >= re_)])
        ghosts.append(g)
        locals_.append(_localise(A, slice(rb, re_), rb, re_, g))
    n_pairs = 0
    for r in range(world):
        rb, re_ = int(offsets[r]), int(offsets[r + 1])
        v = np.full(n_local + len(ghosts[r]), np.nan)
        v[:n_local] = x[rb:re_]
        owner = np.searchsorted(offsets, ghosts[r], side="right") - 1
        off = 0
        for q in np.unique(owner):
            cnt = int(np.sum(owner == q))
            nb, send_idx = wire_send_lists(ghosts, offsets, int(q))       # what rank q sends, per destination
            sent = x[offsets[q]:offsets[q + 1]][send_idx[nb.index(r)]]     # ... to rank r, in q's send order
            assert len(sent) == cnt
            v[n_local + off:n_local + off + cnt] = sent                    # lands in r's ghost slots in ghost order
            off += cnt
            n_pairs += 1
        assert off == len(ghosts[r]) and not np.isnan(v).any()
        assert np.array_equal(v[n_local:], x[ghosts[r]])
        L = locals_[r]
        y_loc = oracle.spmv(L.n_rows, L.rowptr, L.col, L.val, v)
        assert np.array_equal(y_loc, y_ref[rb:re_])
    assert n_pairs >= world    # the random columns make (nearly) every rank a neighbour of every other
