"""CPU emulation of the row-partitioned V-cycle: the oracle's kernels on the LOCAL operators of each rank, with the
halo exchange / gathers carried by torch.distributed (gloo).  It exercises exactly the host-side data the GPU path
consumes (hostsetup.partition.LocalPart): local CSR numbering, ghost order, send lists, boundary row ranges,
coarse slices.  Test infrastructure only."""
import numpy as np
import torch
import torch.distributed as dist

import oracle


def halo_exchange(part, v):
    """Fill v[n_owned:] from the neighbours (v: numpy array of n_owned + n_ghost)."""
    reqs, recv_bufs = [], []
    off = 0
    for q, cnt, sidx in zip(part.neighbors, part.recv_counts, part.send_indices):
        if len(sidx):
            t = torch.from_numpy(np.ascontiguousarray(v[sidx]))
            reqs.append(dist.isend(t, q))
        if cnt:
            buf = torch.empty(cnt, dtype=torch.float64)
            reqs.append(dist.irecv(buf, q))
            recv_bufs.append((off, cnt, buf))
        off += cnt
    for r in reqs:
        r.wait()
    for o, c, buf in recv_bufs:
        v[part.n_owned + o:part.n_owned + o + c] = buf.numpy()


def gather_coarse(part, bc_full):
    world = part.world
    co = part.coarse_offsets
    mine = torch.from_numpy(np.ascontiguousarray(bc_full[co[part.rank]:co[part.rank + 1]]))
    sizes = [int(co[r + 1] - co[r]) for r in range(world)]
    outs = [torch.empty(s, dtype=torch.float64) for s in sizes]
    if len(set(sizes)) == 1:
        dist.all_gather(outs, mine)
    else:
        objs = [None] * world
        dist.all_gather_object(objs, mine.numpy())
        outs = [torch.from_numpy(o) for o in objs]
    for r in range(world):
        bc_full[co[r]:co[r + 1]] = outs[r].numpy()


def coarse_dd_solve_cpu(plan, bc_full, g_below=None):
    """csrc/coarse_dd.cu restated with numpy + gloo: bc_full is valid on this rank's coarse rows only; the result is
    valid on this rank's rows and on all separator rows (everything else is NaN to catch illegal reads)."""
    AII, AIS, ASI = (plan[k].to_scipy().toarray() for k in ("A_II", "A_IS", "A_SI"))
    n_I, n_S, sep = plan["n_I"], plan["n_S"], plan["sep_index"]
    a0, na = plan["adj_begin"], AIS.shape[1]
    Mi = np.linalg.inv(AII) if n_I else AII
    E = Mi @ AIS
    # setup: Schur complement summed over the ranks
    S = plan["A_SS"].to_scipy().toarray() if dist.get_rank() == 0 else np.zeros((n_S, n_S))
    S[a0:a0 + na, a0:a0 + na] -= ASI @ E
    St = torch.from_numpy(S)
    dist.all_reduce(St)
    # apply
    y = Mi @ bc_full[plan["own_begin"]:plan["own_begin"] + n_I]
    t = np.zeros(n_S)
    t[a0:a0 + na] -= ASI @ y
    o0, on = plan["own_sep_begin"], plan["own_sep_n"]
    t[o0:o0 + on] += bc_full[sep[o0:o0 + on]]
    if g_below is not None and len(g_below):   # this rank's share of the restriction onto the separator below
        t[a0:a0 + len(g_below)] += g_below
    tt = torch.from_numpy(t)
    dist.all_reduce(tt)                      # the only communication of the solve
    xs = np.linalg.solve(St.numpy(), tt.numpy())
    x = np.full(plan["n_c"], np.nan)
    x[sep] = xs
    x[plan["own_begin"]:plan["own_begin"] + n_I] = y - E @ xs[a0:a0 + na]
    return x


def dist_vcycle_cpu(part, lu, piv, b_loc, nu=1, dd_plan=None):
    """Preconditioner-mode V(nu,nu) on this rank's rows; returns x_loc (owned entries).  dd_plan: use the
    domain-decomposed coarse solve (no gather of the coarse right-hand side or solution) instead of the dense LU."""
    n, ng = part.n_owned, part.n_ghost
    A, R, P = part.A, part.R, part.P
    dinv = oracle.inv_diag(n, A.rowptr, A.col, A.val)
    x = np.zeros(n + ng)
    for s in range(nu):
        halo_exchange(part, x)
        r = oracle.spmv(n, A.rowptr, A.col, A.val, x) - b_loc
        x[:n] = x[:n] - dinv * r
    res = np.zeros(n + ng)
    halo_exchange(part, x)
    res[:n] = oracle.spmv(n, A.rowptr, A.col, A.val, x) - b_loc
    co = part.coarse_offsets
    bc = np.zeros(part.Ac.n_rows)
    if dd_plan is not None:
        # no exchange of the residual's halo: the entries of R on the ghost plane are dropped here and come in
        # through the neighbour's R_below share, summed by the all-reduce of the separator right-hand sides
        res[n:] = np.nan
        Rs = R.to_scipy().tocsc()[:, :n].tocsr()
        bc[co[part.rank]:co[part.rank + 1]] = Rs @ res[:n]
        g_below = None
        if part.rank > 0:
            f0 = dd_plan["sep_below_first_row"]
            Rb = part.R_below.to_scipy()
            assert Rb[:f0].nnz == 0
            g_below = Rb[f0:, :n] @ res[:n]
            assert len(g_below) == dd_plan["n_sep_below"]
        xc = coarse_dd_solve_cpu(dd_plan, bc, g_below)
        assert not np.any(np.isnan(xc[np.unique(P.col)])), "P reads coarse entries the DD solve does not provide"
        xc = np.nan_to_num(xc)
    else:
        halo_exchange(part, res)
        bc[co[part.rank]:co[part.rank + 1]] = oracle.spmv(R.n_rows, R.rowptr, R.col, R.val, res)
        gather_coarse(part, bc)
        xc = oracle.lu_solve(lu, piv, bc)
    x[:n] = x[:n] - oracle.spmv(n, P.rowptr, P.col, P.val, xc)
    for s in range(nu):
        halo_exchange(part, x)
        r = oracle.spmv(n, A.rowptr, A.col, A.val, x) - b_loc
        x[:n] = x[:n] - dinv * r
    return x[:n].copy()
