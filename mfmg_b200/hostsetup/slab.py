"""Host setup, slab-local: each rank builds ONLY its z-slab of a two-level hierarchy (plus one agglomerate layer
on each side), so the multi-GPU configurations never materialise the global operators on one host process.

Rank r owns the cell layers [cz0, cz1) (aligned with the agglomerate layers) and the node planes [cz0 p, cz1 p)
(the last rank also owns the top plane).  It assembles the sub-box of cell layers [cz0 - bz, cz1 + bz) -- cut planes
are NOT Dirichlet; rows on the cut planes are incomplete and never used -- builds the restrictor rows of the
agglomerates in that sub-box, and extracts

  A_loc  rows of the owned nodes, columns [owned | ghost]
  R_loc  rows of the owned agglomerates
  P_loc  rows of R^T for the owned nodes (global coarse columns; nodes on the bottom shared plane pick up the
         agglomerate layer below, which is why that layer is built too)
  A_c    rows of R A R^T for the owned agglomerates (exact because every node within reach of an owned
         agglomerate lies strictly inside the sub-box when the agglomerates are >= 2 cells thick); the row blocks
         are gathered so that every rank holds the whole coarse operator (replicated dense solve).

`gather(obj) -> [obj of rank 0, ..., obj of rank world-1]` is the only setup-time communication
(torch.distributed.all_gather_object in bench.py; a list comprehension in the single-process tests).
The result is identical to slicing the global operators (tests/test_distributed_cpu.py).
"""
from __future__ import annotations

import numpy as np
from .amge import build_restrictor, galerkin_rows
from .partition import LocalPart, finalize_plan, rows_restricted, slab_row_ranges
from .problems import HostCSR, LaplaceProblem


_BLOCK = 1 << 26  # non-zeros per block of the streaming helpers below


def _outside(col: np.ndarray, k0: int, k1: int, lo: int, hi: int) -> np.ndarray:
    """Sorted unique column indices in col[k0:k1] that are not in [lo, hi)."""
    found = [np.zeros(0, dtype=col.dtype)]
    for b0 in range(k0, k1, _BLOCK):
        c = col[b0:min(k1, b0 + _BLOCK)]
        found.append(np.unique(c[(c < lo) | (c >= hi)]))
    return np.unique(np.concatenate(found))


def _localise_ext(M: HostCSR, r0: int, r1: int, lo: int, hi: int, ghost: np.ndarray) -> HostCSR:
    """Rows [r0, r1) of M; column c becomes c - lo when lo <= c < hi (owned), else (hi - lo) + position in `ghost`."""
    rp = M.rowptr[r0:r1 + 1]
    k0, k1 = int(rp[0]), int(rp[-1])
    loc = np.empty(k1 - k0, dtype=np.int32)
    for b0 in range(k0, k1, _BLOCK):
        b1 = min(k1, b0 + _BLOCK)
        c = M.col[b0:b1]
        owned = (c >= lo) & (c < hi)
        out = (c - lo).astype(np.int32)
        if not owned.all():
            g = c[~owned]
            pos = np.searchsorted(ghost, g)
            assert np.all(ghost[np.minimum(pos, len(ghost) - 1)] == g)
            out[~owned] = (hi - lo) + pos
        loc[b0 - k0:b1 - k0] = out
    return HostCSR(r1 - r0, (hi - lo) + len(ghost), np.ascontiguousarray(rp - k0), loc, M.val[k0:k1])


def _rows_with_ghosts(A_loc: HostCSR, n_owned: int) -> np.ndarray:
    """Sorted rows of A_loc that reference a ghost column (>= n_owned)."""
    rows = [np.zeros(0, dtype=np.int64)]
    nnz = A_loc.nnz
    for b0 in range(0, nnz, _BLOCK):
        k = np.flatnonzero(A_loc.col[b0:min(nnz, b0 + _BLOCK)] >= n_owned) + b0
        if len(k):
            rows.append(np.unique(np.searchsorted(A_loc.rowptr, k, side="right") - 1))
    return np.unique(np.concatenate(rows))


def build_slab_part(degree: int, cells, h, material: str, block, n_eigenvectors: int, world: int, rank: int, gather,
                    eigensolver: str = "free") -> LocalPart:
    dim = 3
    cells = tuple(int(c) for c in cells)
    block = tuple(int(b) for b in block)
    p = degree
    if block[2] < 2 and world > 1:
        raise ValueError("slab-local setup needs agglomerates at least 2 cells thick in z")
    nodes = tuple(c * p + 1 for c in cells)
    plane = nodes[0] * nodes[1]
    layers, row_off = slab_row_ranges(nodes, p, cells[2], block[2], world)
    aggs_per_layer = (-(-cells[0] // block[0])) * (-(-cells[1] // block[1]))
    coarse_off = np.array([(-(-int(l) // block[2])) * aggs_per_layer * n_eigenvectors for l in layers], dtype=np.int64)
    cz0, cz1 = int(layers[rank]), int(layers[rank + 1])
    e0, e1 = max(0, cz0 - block[2]), min(cells[2], cz1 + block[2])
    ext = LaplaceProblem.create_box(dim, p, (cells[0], cells[1], e1 - e0), h, material,
                                    origin=(0.0, 0.0, e0 * h[2]),
                                    faces=[(True, True), (True, True), (e0 == 0, e1 == cells[2])])
    R_ext = build_restrictor(ext, block, n_eigenvectors, eigensolver=eigensolver)
    off = e0 * p * plane                                  # ext node index + off = global node index
    coff = (e0 // block[2]) * aggs_per_layer * n_eigenvectors   # ext coarse index + coff = global coarse index
    rb, re_ = int(row_off[rank]), int(row_off[rank + 1])
    cb, ce = int(coarse_off[rank]), int(coarse_off[rank + 1])
    n_global = int(np.prod(nodes))
    nc_global = int(coarse_off[-1])

    # Owned rows of the sub-box operators with columns renumbered [owned | ghost].  Everything works on the sub-box's
    # own int32 columns, in bounded blocks of non-zeros: at 2 ranks of cfg3 the sub-box operator has 2e9 non-zeros
    # and whole-array int64 temporaries would cost tens of GB per rank.
    lo, hi = rb - off, re_ - off            # owned node range in sub-box numbering
    ka0, ka1 = int(ext.A.rowptr[lo]), int(ext.A.rowptr[hi])
    kr0, kr1 = int(R_ext.rowptr[cb - coff]), int(R_ext.rowptr[ce - coff])
    ghost_ext = np.union1d(_outside(ext.A.col, ka0, ka1, lo, hi), _outside(R_ext.col, kr0, kr1, lo, hi))
    ghost_global = ghost_ext.astype(np.int64) + off
    A_loc = _localise_ext(ext.A, lo, hi, lo, hi, ghost_ext)
    R_loc = _localise_ext(R_ext, cb - coff, ce - coff, lo, hi, ghost_ext)

    # P rows of the owned nodes, global coarse columns
    Rs = R_ext.to_scipy()
    Rt = Rs.T.tocsr()
    Rt.sort_indices()
    Pl = Rt[rb - off:re_ - off].tocsr()
    P_loc = HostCSR(Pl.shape[0], nc_global, np.ascontiguousarray(Pl.indptr, dtype=np.int64),
                    np.ascontiguousarray(Pl.indices + coff, dtype=np.int32), np.ascontiguousarray(Pl.data))

    # owned rows of A_c = R (A R^T), in blocks of coarse rows (the sub-box operator can exceed 2^31 non-zeros)
    ac_rows = galerkin_rows(ext.A, R_ext, Rt, cb - coff, ce - coff)
    blocks = gather((ac_rows.indptr.astype(np.int64), (ac_rows.indices.astype(np.int64) + coff), ac_rows.data))
    rp = [np.zeros(1, dtype=np.int64)]
    cols_all, vals_all = [], []
    for indptr, indices, data in blocks:
        rp.append(indptr[1:] + rp[-1][-1])
        cols_all.append(indices)
        vals_all.append(data)
    Ac = HostCSR(nc_global, nc_global, np.concatenate(rp), np.concatenate(cols_all).astype(np.int32),
                 np.concatenate(vals_all))

    part = LocalPart(rank, world, rb, re_, re_ - rb, len(ghost_global), ghost_global, A_loc, R_loc, P_loc, Ac, coarse_off)
    row_offsets = np.asarray(row_off)
    owner = np.searchsorted(row_offsets, ghost_global, side="right") - 1
    for q in np.unique(owner):
        part.neighbors.append(int(q))
        part.recv_counts.append(int(np.sum(owner == q)))
    n_owned = re_ - rb
    idx = _rows_with_ghosts(A_loc, n_owned)
    lo_rows, hi_rows = idx[idx < n_owned // 2], idx[idx >= n_owned // 2]
    part.boundary_lo = int(lo_rows.max() + 1) if len(lo_rows) else 0
    part.boundary_hi = int(hi_rows.min()) if len(hi_rows) else n_owned
    ghost_lists = gather(ghost_global)
    finalize_plan(part, ghost_lists, row_offsets)
    # extras for the driver
    part.constrained = ext.constrained[rb - off:re_ - off]
    part.n_global = n_global
    if rank > 0:
        # rows of the rank below present in the sub-box (its top agglomerate layer) restricted to the owned nodes
        c_prev = int(coarse_off[rank - 1])
        part.R_below = rows_restricted(R_ext, 0, cb - coff, lo, hi, n_owned + len(ghost_global),
                                       pad_rows_before=coff - c_prev, n_rows=cb - c_prev)
    part.mf = _matrix_free_slab(ext, p, cells, h, e0, cz0, cz1, rank, world, plane, ghost_global, rb, re_)
    return part


def _matrix_free_slab(ext, p, cells, h, e0, cz0, cz1, rank, world, plane, ghost_global, rb, re_):
    """Data of the matrix-free level-0 operator on this rank's slab (3D Q1): the local box is the owned cell layers
    plus the layer below (its contributions complete the bottom owned plane); node planes of the box are ordered
    [owned | ghost below | ghost above] like every vector of the partitioned level.  None when the ghost set of the
    assembled rows is not exactly those planes (then the halo plan could not serve both)."""
    if p != 1:
        return None
    l0 = cz0 - 1 if rank > 0 else cz0            # first cell layer of the local box (global numbering)
    l1 = cz1
    g_lo, g_hi = l0, l1                          # node planes l0 .. l1 (inclusive) of the local box
    own_lo = cz0
    own_hi = cz1 + 1 if rank == world - 1 else cz1   # owned planes [own_lo, own_hi)
    expect = np.concatenate([np.arange(g * plane, (g + 1) * plane, dtype=np.int64)
                             for g in list(range(g_lo, own_lo)) + list(range(own_hi, g_hi + 1))] + [np.zeros(0, np.int64)])
    if not np.array_equal(expect, ghost_global):
        return None
    cxy = cells[0] * cells[1]
    coef = ext.coef[(l0 - e0) * cxy:(l1 - e0) * cxy]
    nq = (p + 1) ** 3
    if coef.shape[1] != nq:
        coef = np.repeat(coef, nq, axis=1)
    c_ext = ext.constrained.reshape(-1, plane)       # [ext plane][node in plane]
    order = list(range(own_lo, own_hi)) + list(range(g_lo, own_lo)) + list(range(own_hi, g_hi + 1))
    constrained = np.concatenate([c_ext[g - e0] for g in order])
    return {"degree": p, "cells": (cells[0], cells[1], l1 - l0), "h": tuple(h), "coef": np.ascontiguousarray(coef),
            "constrained": np.ascontiguousarray(constrained), "own_planes": (own_lo - g_lo, own_hi - g_lo)}
