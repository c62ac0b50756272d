"""Host setup: row partition of the level operators for the multi-GPU V-cycle.

The reference partitions rows through deal.II's distributed triangulation (`locally_owned_dofs`, agglomerates never
cross ranks, include/mfmg/common/amge.templates.hpp:453-478) and keeps GLOBAL column indices with a full-length
source vector that is all-gathered on every SpMV (include/mfmg/cuda/sparse_matrix_device.templates.cuh:104-138).
Here every level is partitioned in contiguous row blocks (z-slabs of the lexicographic numbering, aligned with the
agglomerate layers), columns are renumbered `[owned | ghost]`, and a halo plan lists exactly the entries that cross
rank boundaries (the information cuda_solver.cu:306-426 derives for AMGX's one-ring maps).

Pure host code (numpy); used once at setup.  Nothing here runs in the V-cycle.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .problems import HostCSR


@dataclass
class LocalPart:
    rank: int
    world: int
    row_begin: int          # first owned fine row (global numbering)
    row_end: int
    n_owned: int
    n_ghost: int
    ghost_global: np.ndarray  # global fine index of every ghost slot, ascending (hence grouped by owner)
    A: HostCSR                # n_owned x (n_owned + n_ghost), local columns
    R: HostCSR                # nc_owned x (n_owned + n_ghost), local columns
    P: HostCSR                # n_owned x n_c (GLOBAL coarse columns: x_c is replicated on every rank)
    Ac: HostCSR               # the full coarse operator (replicated dense solve)
    coarse_offsets: np.ndarray  # world + 1 offsets of the owned coarse rows
    neighbors: list = field(default_factory=list)      # ranks exchanged with, ascending
    recv_counts: list = field(default_factory=list)    # ghost entries received from each neighbour (ghost order)
    send_indices: list = field(default_factory=list)   # per neighbour: LOCAL owned indices to send, ascending
    # rows [0, n_owned) whose columns are all owned form the interior; with a slab partition the rows that
    # touch ghosts are a prefix and a suffix of the owned range:
    boundary_lo: int = 0     # rows [0, boundary_lo) reference ghosts below
    boundary_hi: int = 0     # rows [boundary_hi, n_owned) reference ghosts above
    # rows of R of the rank BELOW restricted to this rank's owned fine entries (local columns, one row per coarse row
    # of that rank): what this rank contributes to the restriction onto the neighbour's top agglomerate layer.  With
    # it the residual's halo need not be exchanged (device.Hierarchy.from_partition, coarse_dd).
    R_below: HostCSR = None


def slab_row_ranges(nodes, degree: int, cells_z: int, block_z: int, world: int):
    """Split the cell layers along z into `world` contiguous slabs aligned with the agglomerate layers.
    Returns (cell_layer_offsets[world+1], row_offsets[world+1])."""
    n_blocks = -(-cells_z // block_z)
    if n_blocks < world:
        raise ValueError(f"cannot give each of {world} ranks an agglomerate layer ({n_blocks} layers)")
    base, rem = divmod(n_blocks, world)
    layers = [0]
    for r in range(world):
        nb = base + (1 if r < rem else 0)
        layers.append(min(cells_z, layers[-1] + nb * block_z))
    layers[-1] = cells_z
    plane = int(np.prod(nodes[:-1]))
    rows = [layers[r] * degree * plane for r in range(world)]
    rows.append(int(np.prod(nodes)))
    return np.array(layers, dtype=np.int64), np.array(rows, dtype=np.int64)


def _localise(M: HostCSR, rows: slice, row_begin: int, row_end: int, ghost_global: np.ndarray) -> HostCSR:
    """Rows `rows` of M with columns renumbered [owned | ghost]."""
    rp = M.rowptr[rows.start:rows.stop + 1]
    k0, k1 = int(rp[0]), int(rp[-1])
    col = M.col[k0:k1].astype(np.int64)
    owned = (col >= row_begin) & (col < row_end)
    loc = np.empty_like(col)
    loc[owned] = col[owned] - row_begin
    gpos = np.searchsorted(ghost_global, col[~owned])
    assert np.all(ghost_global[gpos] == col[~owned])
    loc[~owned] = (row_end - row_begin) + gpos
    return HostCSR(rows.stop - rows.start, (row_end - row_begin) + len(ghost_global),
                   np.ascontiguousarray(rp - k0), loc.astype(np.int32), np.ascontiguousarray(M.val[k0:k1]))


def rows_restricted(M: HostCSR, r0: int, r1: int, lo: int, hi: int, n_cols: int, pad_rows_before: int = 0,
                    n_rows: int | None = None) -> HostCSR:
    """Rows [r0, r1) of M keeping the entries with lo <= col < hi (renumbered col - lo), as a matrix with `n_cols`
    columns and `n_rows` rows of which the first `pad_rows_before` are empty."""
    k0, k1 = int(M.rowptr[r0]), int(M.rowptr[r1])
    col = M.col[k0:k1].astype(np.int64)
    keep = (col >= lo) & (col < hi)
    row_of = np.repeat(np.arange(r1 - r0), np.diff(M.rowptr[r0:r1 + 1]))
    n_rows = (r1 - r0) + pad_rows_before if n_rows is None else n_rows
    counts = np.bincount(row_of[keep] + pad_rows_before, minlength=n_rows)
    rowptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    return HostCSR(n_rows, n_cols, rowptr, (col[keep] - lo).astype(np.int32), np.ascontiguousarray(M.val[k0:k1][keep]))


def partition_two_level(A: HostCSR, R: HostCSR, Ac: HostCSR, row_offsets, coarse_offsets, rank: int) -> LocalPart:
    """Local operators of `rank` from the global ones (the global-redundant setup mode: every rank runs the host
    setup and slices; a slab-local setup is the next step)."""
    world = len(row_offsets) - 1
    rb, re_ = int(row_offsets[rank]), int(row_offsets[rank + 1])
    cb, ce = int(coarse_offsets[rank]), int(coarse_offsets[rank + 1])
    n_owned = re_ - rb
    # ghosts: every column of the owned rows of A and of the owned rows of R that is not owned
    ka0, ka1 = int(A.rowptr[rb]), int(A.rowptr[re_])
    kr0, kr1 = int(R.rowptr[cb]), int(R.rowptr[ce])
    cols = np.concatenate([A.col[ka0:ka1], R.col[kr0:kr1]]).astype(np.int64)
    ghost_global = np.unique(cols[(cols < rb) | (cols >= re_)])
    A_loc = _localise(A, slice(rb, re_), rb, re_, ghost_global)
    R_loc = _localise(R, slice(cb, ce), rb, re_, ghost_global)
    # P = R^T rows of the owned fine nodes, global coarse columns
    Rt = R.to_scipy().T.tocsr()
    Rt.sort_indices()
    P_loc = HostCSR.from_scipy(Rt[rb:re_])
    part = LocalPart(rank, world, rb, re_, n_owned, len(ghost_global), ghost_global, A_loc, R_loc, P_loc, Ac,
                     np.asarray(coarse_offsets, dtype=np.int64))
    owner = np.searchsorted(np.asarray(row_offsets), ghost_global, side="right") - 1
    for q in np.unique(owner):
        part.neighbors.append(int(q))
        part.recv_counts.append(int(np.sum(owner == q)))
    # boundary row ranges (rows referencing ghost columns)
    has_ghost = np.zeros(n_owned, dtype=bool)
    gmask = A_loc.col >= n_owned
    if gmask.any():
        row_of = np.repeat(np.arange(n_owned), np.diff(A_loc.rowptr))
        has_ghost[np.unique(row_of[gmask])] = True
    idx = np.flatnonzero(has_ghost)
    lo_rows = idx[idx < n_owned // 2]
    hi_rows = idx[idx >= n_owned // 2]
    part.boundary_lo = int(lo_rows.max() + 1) if len(lo_rows) else 0
    part.boundary_hi = int(hi_rows.min()) if len(hi_rows) else n_owned
    if rank > 0:
        part.R_below = rows_restricted(R, int(coarse_offsets[rank - 1]), cb, rb, re_, n_owned + len(ghost_global))
    return part


def wire_send_lists(parts_or_ghost_lists, row_offsets, rank: int):
    """Send lists of `rank`: for every other rank q, the owned entries q has as ghosts (ascending global order,
    which is the order of q's ghost slots).  `parts_or_ghost_lists[q]` is q's ghost_global array (gathered at setup
    with torch.distributed.all_gather_object, or computed redundantly)."""
    rb, re_ = int(row_offsets[rank]), int(row_offsets[rank + 1])
    neighbors, send_indices = [], []
    for q, gl in enumerate(parts_or_ghost_lists):
        if q == rank:
            continue
        gl = np.asarray(gl, dtype=np.int64)
        mine = gl[(gl >= rb) & (gl < re_)]
        if len(mine):
            neighbors.append(q)
            send_indices.append((mine - rb).astype(np.int32))
    return neighbors, send_indices


def finalize_plan(part: LocalPart, all_ghost_lists, row_offsets) -> LocalPart:
    nb_send, send_idx = wire_send_lists(all_ghost_lists, row_offsets, part.rank)
    # one neighbour list for both directions (a rank may only send to or only receive from a neighbour)
    nbs = sorted(set(part.neighbors) | set(nb_send))
    recv = {q: c for q, c in zip(part.neighbors, part.recv_counts)}
    send = {q: s for q, s in zip(nb_send, send_idx)}
    part.neighbors = nbs
    part.recv_counts = [recv.get(q, 0) for q in nbs]
    part.send_indices = [send.get(q, np.zeros(0, dtype=np.int32)) for q in nbs]
    return part


def make_parts(problem, R: HostCSR, Ac: HostCSR, block, n_eigenvectors: int, world: int, ranks=None):
    """LocalParts (wired halo plans) of the slab partition of a two-level hierarchy.  `ranks`: which ranks to
    build (default all); the ghost lists of all ranks are derived from the global operators, so no setup-time
    communication is needed in this global-redundant mode."""
    dim = problem.dim
    layers, row_off = slab_row_ranges(problem.nodes, problem.degree, problem.cells[dim - 1], block[dim - 1], world)
    aggs_per_layer = 1
    for d in range(dim - 1):
        aggs_per_layer *= -(-problem.cells[d] // block[d])
    coarse_off = np.array([(-(-int(l) // block[dim - 1])) * aggs_per_layer * n_eigenvectors for l in layers],
                          dtype=np.int64)
    coarse_off[-1] = R.n_rows
    parts = [partition_two_level(problem.A, R, Ac, row_off, coarse_off, r) for r in range(world)]
    ghost_lists = [p.ghost_global for p in parts]
    wanted = range(world) if ranks is None else ranks
    return [finalize_plan(parts[r], ghost_lists, row_off) for r in wanted], row_off, coarse_off


def coarse_dd_plan(Ac: HostCSR, coarse_offsets, rank: int):
    """Index sets and blocks of the domain-decomposed coarse solve (csrc/coarse_dd.cu) for `rank`, or None when the
    coarse operator is not block tridiagonal in the ranks' row blocks (then the dense inverse is used).

    Separator of rank r < N-1: its coarse rows from the first one coupled to rank r+1 to the end of its block (for
    slab partitions: the last agglomerate layer).  The decision only uses the replicated A_c and the offsets, so every
    rank reaches the same one without communication."""
    import scipy.sparse as sp

    co = np.asarray(coarse_offsets, dtype=np.int64)
    world = len(co) - 1
    if world < 2:
        return None
    A = Ac.to_scipy().tocsr()
    n_c = A.shape[0]
    row_of = np.repeat(np.arange(n_c), np.diff(A.indptr))
    sep_begin = co[1:].copy()                      # S_r = [sep_begin[r], co[r+1]); empty on the last rank
    for r in range(world - 1):
        mask = (row_of >= co[r]) & (row_of < co[r + 1]) & (A.indices >= co[r + 1])
        if mask.any():
            sep_begin[r] = row_of[mask].min()
        if (A.indices[(row_of >= co[r]) & (row_of < co[r + 1])] >= (co[r + 2] if r + 2 <= world else n_c)).any():
            return None                             # couples beyond the next rank
    s_off = np.concatenate([[0], np.cumsum(co[1:] - sep_begin)]).astype(np.int64)   # S numbering offsets per rank
    n_S = int(s_off[-1])
    if n_S == 0:
        return None
    sep_index = np.concatenate([np.arange(sep_begin[r], co[r + 1]) for r in range(world)]).astype(np.int32)
    is_sep = np.zeros(n_c, dtype=bool)
    is_sep[sep_index] = True
    # every interior may only couple to itself and to the separators right below / above it
    for r in range(world):
        lo, hi = co[r], sep_begin[r]
        if hi <= lo:
            continue
        cols = A.indices[A.indptr[lo]:A.indptr[hi]]
        inside = (cols >= lo) & (cols < hi)
        below = (cols >= (sep_begin[r - 1] if r > 0 else lo)) & (cols < lo) if r > 0 else np.zeros_like(inside)
        above = (cols >= hi) & (cols < co[r + 1])
        if not np.all(inside | below | above):
            return None
    lo, hi = int(co[rank]), int(sep_begin[rank])
    adj_begin = int(s_off[rank - 1]) if rank > 0 else 0
    adj_end = int(s_off[rank + 1])
    adj_cols = sep_index[adj_begin:adj_end].astype(np.int64)
    I = np.arange(lo, hi)
    blocks = {
        "A_II": HostCSR.from_scipy(A[lo:hi][:, lo:hi]),
        "A_IS": HostCSR.from_scipy(A[lo:hi][:, adj_cols] if len(adj_cols) else sp.csr_matrix((hi - lo, 0))),
        "A_SI": HostCSR.from_scipy(A[adj_cols][:, lo:hi] if len(adj_cols) else sp.csr_matrix((0, hi - lo))),
        "A_SS": HostCSR.from_scipy(A[sep_index.astype(np.int64)][:, sep_index.astype(np.int64)]),
    }
    for m in blocks.values():
        m.to_scipy().sort_indices()
    return {"n_c": n_c, "own_begin": lo, "n_I": len(I), "n_S": n_S, "adj_begin": adj_begin,
            "n_sep_below": int(s_off[rank] - s_off[rank - 1]) if rank > 0 else 0,
            "sep_below_first_row": int(sep_begin[rank - 1] - co[rank - 1]) if rank > 0 else 0,
            "own_sep_begin": int(s_off[rank]), "own_sep_n": int(s_off[rank + 1] - s_off[rank]),
            "sep_index": sep_index, "valid_cols": np.concatenate([np.arange(lo, co[rank + 1]), sep_index]), **blocks}
