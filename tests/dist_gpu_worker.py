"""Multi-GPU parity worker (launched by tests/test_gpu_multi.py under torch.distributed.run, one rank per GPU):
the row-partitioned V-cycle and PCG through the C ABI (NCCL halo exchange overlapped with interior rows, NCCL
all-reduce for the dots) against the SERIAL CPU oracle on the same global problem."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import oracle  # noqa: E402
from helpers import oracle_hierarchy, two_level_problem  # noqa: E402
from mfmg_b200 import device as d  # noqa: E402
from mfmg_b200 import hostsetup as hs  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    handle = d.CudaHandle(local)
    handle.init_comm_from_torch()
    transport = handle.lib.mfmgb_comm_transport(handle.ctx).decode()
    if rank == 0:
        print(f"transport: {transport}", flush=True)
    want_peer = os.environ.get("MFMGB_PEER", "1") != "0"
    assert ("peer-memory" in transport) == want_peer, transport
    # small all-reduce through peer memory (or NCCL): exact integer sums, back to back without synchronisation
    tri = world * (world + 1) // 2
    for n in (1, 3, 100, 5000, 9000):      # 9000 > the peer slot capacity: NCCL serves it
        pattern = (np.arange(n) % 13).astype(np.float64)
        vecs = [d.DeviceVector.from_host(handle, (rank + 1) * pattern * (k + 1)) for k in range(12)]
        for v in vecs:
            d.check(handle.ctx, handle.lib.mfmgb_allreduce_sum(handle.ctx, v.ptr, n))
        handle.synchronize()
        for k, v in enumerate(vecs):
            assert np.array_equal(v.to_host(), tri * pattern * (k + 1)), (rank, "allreduce", n, k)
    cases = [(3, 1, 16, 2, 2, "constant", 1), (3, 1, 24, 4, 1, "discontinuous", 2), (3, 2, 8, 2, 2, "linear", 1),
             (2, 1, 64, 2, 2, "constant", 1)]
    if world > 4:   # every rank needs at least one agglomerate layer
        cases = [(3, 1, 16, 2, 2, "constant", 1), (3, 1, 32, 4, 1, "discontinuous", 2), (3, 2, 16, 2, 1, "linear", 1),
                 (2, 1, 64, 2, 2, "constant", 1)]
    # second pass: force the row-split (multi-GPU) dense coarse solve, which is normally used from n_c = 8192 on
    # coarse-solver variants: "dd" = domain-decomposed direct solve (the default for slab partitions), "split" = dense
    # inverse split by rows over the ranks, "replicated" = dense inverse on every rank
    # "dd_halo": the same with the residual's halo still exchanged (by default it is not: the neighbour's share of the
    # restriction joins the coarse solver's all-reduce)
    cases = [c + ("dd",) for c in cases] + [c + ("split",) for c in cases[:2]] + [cases[0] + ("replicated",)] + \
        [cases[1] + ("dd_halo",)]
    for dim, degree, cells, block, ne, mat, nu, coarse in cases:
        os.environ["MFMGB_DENSE_SPLIT_MIN"] = "1" if coarse == "split" else "8192"
        os.environ["MFMGB_COARSE_DD"] = "1" if coarse.startswith("dd") else "0"
        os.environ["MFMGB_RESTRICT_NO_HALO"] = "0" if coarse == "dd_halo" else "1"
        P, R, Ac = two_level_problem(dim, degree, cells, block, ne, mat)
        (part,), row_off, coarse_off = hs.make_parts(P, R, Ac, (block,) * dim, ne, world, ranks=[rank])
        H = d.Hierarchy.from_partition(handle, part, {"is preconditioner": True, "smoother": {"n_smoothing_steps": nu}})
        assert (H.coarse_dd is not None) == coarse.startswith("dd"), (rank, coarse)
        Ho = oracle_hierarchy(P, R, Ac, nu, True)
        rng = np.random.default_rng(5)
        b_h = rng.standard_normal(P.n)
        sl = slice(part.row_begin, part.row_end)
        b = H.build_vector()
        x = H.build_vector()
        bl = np.zeros(H.vector_size)
        bl[:part.n_owned] = b_h[sl]
        b.upload(bl)
        H.vmult(x, b)
        x_ref = Ho.vmult(b_h)[sl]
        err = np.linalg.norm(x.to_host()[:part.n_owned] - x_ref) / np.linalg.norm(x_ref)
        assert err < 1e-12, (rank, "vcycle", err)
        # CUDA-graph replay of the partitioned cycle (NCCL + two streams captured) == eager launches, bit for bit
        if os.environ.get("MFMGB_DIST_GRAPH", "1") != "0":
            x_eager = x.to_host()[:part.n_owned].copy()
            H.use_graph(True)
            for _ in range(40):     # back to back, no host synchronisation: ranks may run ahead of each other
                H.vmult(x, b)
            assert np.array_equal(x.to_host()[:part.n_owned], x_eager), (rank, "graph replay")
            H.use_graph(False)
        # halo exchange alone: ghosts equal the owners' values
        v = H.build_vector()
        vl = np.zeros(H.vector_size)
        vl[:part.n_owned] = b_h[sl]
        v.upload(vl)
        H.halo.exchange(v)
        handle.synchronize()
        assert np.array_equal(v.to_host()[part.n_owned:], b_h[part.ghost_global]), (rank, "halo")
        # many exchanges of different vectors in flight (mailbox parity / sequence flags), checked afterwards
        vs = []
        for k in range(9):
            w = H.build_vector()
            wl = np.zeros(H.vector_size)
            wl[:part.n_owned] = b_h[sl] * (k + 2)
            w.upload(wl)
            vs.append(w)
        for w in vs:
            H.halo.exchange(w)
        handle.synchronize()
        for k, w in enumerate(vs):
            assert np.array_equal(w.to_host()[part.n_owned:], b_h[part.ghost_global] * (k + 2)), (rank, "halo burst", k)
        # PCG: equal iteration counts, residual history within 1e-10 (north star)
        x0 = oracle.std_uniform01(P.n, skip=P.constrained)
        x_o, it_ref, hist_ref = Ho.pcg(np.zeros(P.n), x0, 1e-8, 500)
        xl = np.zeros(H.vector_size)
        xl[:part.n_owned] = x0[sl]
        x.upload(xl)
        b.fill(0.0)
        # dot products cover owned entries only: pass views of length n_owned through the raw ABI
        import ctypes

        hist = np.zeros(501)
        it = ctypes.c_int(0)
        rc = handle.lib.mfmgb_pcg(handle.ctx, H.ptr, H.operators[0].ptr, b.ptr, x.ptr, 1e-8, 500, ctypes.byref(it),
                                  hist.ctypes.data)
        d.check(handle.ctx, rc)
        assert it.value == it_ref, (rank, it.value, it_ref)
        hist = hist[:it.value + 1]
        assert np.max(np.abs(hist - hist_ref) / hist_ref) < 1e-10, (rank, "pcg history")
        err = np.linalg.norm(x.to_host()[:part.n_owned] - x_o[sl]) / max(np.linalg.norm(x_o[sl]), 1e-300)
        assert err < 1e-8 or np.linalg.norm(x_o[sl]) < 1e-6, (rank, "pcg x", err)
        if rank == 0:
            print(f"case {dim}D Q{degree} {cells}^{dim} {mat} coarse={coarse}: vcycle OK, PCG {it.value} its (oracle {it_ref})", flush=True)
    # matrix-free level 0 on z-slabs (cfg4): slab-local setup with real gathers, assembled R / P / A_c
    cells, h, block = (12, 10, 8 * world), (0.05, 0.04, 0.03), (4, 5, 4)

    def gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    os.environ["MFMGB_COARSE_DD"] = "1"
    os.environ["MFMGB_RESTRICT_NO_HALO"] = "1"
    for mat in ("discontinuous", "linear"):
        part = hs.build_slab_part(1, cells, h, mat, block, 1, world, rank, gather)
        assert part.mf is not None
        H = d.Hierarchy.from_partition(handle, part, {"is preconditioner": True}, matrix_free=True)
        Pg = hs.LaplaceProblem.create_box(3, 1, cells, h, mat)
        Rg = hs.build_restrictor(Pg, block, 1)
        Acg = hs.galerkin(Pg.A, Rg)
        Ho = oracle_hierarchy(Pg, Rg, Acg, 1, True)
        rng = np.random.default_rng(11)
        b_h = rng.standard_normal(Pg.n)
        b_h[Pg.constrained != 0] = 0.0
        sl = slice(part.row_begin, part.row_end)
        b, x = H.build_vector(), H.build_vector()
        bl = np.zeros(H.vector_size)
        bl[:part.n_owned] = b_h[sl]
        b.upload(bl)
        H.vmult(x, b)
        x_ref = Ho.vmult(b_h)[sl]
        err = np.linalg.norm(x.to_host()[:part.n_owned] - x_ref) / np.linalg.norm(x_ref)
        assert err < 1e-9, (rank, "matrix-free vcycle", err)   # matrix-free == assembled to 1e-9 (test_hierarchy.cc:644-695)
        if rank == 0:
            print(f"matrix-free slab hierarchy {mat}: vcycle OK ({H.operators[0].kernel})", flush=True)
    dist.barrier()
    print(f"RANK {rank} OK", flush=True)
    handle.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
