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


def dist_vcycle_cpu(part, lu, piv, b_loc, nu=1):
    """Preconditioner-mode V(nu,nu) on this rank's rows; returns x_loc (owned entries)."""
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
    halo_exchange(part, res)
    co = part.coarse_offsets
    bc = np.zeros(part.Ac.n_rows)
    bc[co[part.rank]:co[part.rank + 1]] = oracle.spmv(R.n_rows, R.rowptr, R.col, R.val, res)
    gather_coarse(part, bc)
    xc = oracle.lu_solve(lu, piv, bc)
    x[:n] = x[:n] - oracle.spmv(n, P.rowptr, P.col, P.val, xc)
    for s in range(nu):
        halo_exchange(part, x)
        r = oracle.spmv(n, A.rowptr, A.col, A.val, x) - b_loc
        x[:n] = x[:n] - dinv * r
    return x[:n].copy()
