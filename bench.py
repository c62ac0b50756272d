#!/usr/bin/env python
"""bench.py -- V-cycle-apply throughput on B200 (BASELINE.json metric: V-cycles/s and SpMV HBM GB/s).

  python bench.py --gpus N --steps K --warmup W                    # our arm (CUDA path through the C ABI)
  python bench.py --impl reference --gpus N --steps K --warmup W   # the reference arm: CPU oracle port, all host cores

A "step" is one multigrid V-cycle apply (mfmg::Hierarchy::vmult, preconditioner mode, V(1,1) Jacobi, direct coarse
solve).  Default workload = BASELINE configs[1]: 3D Q1 Laplace on 128^3 cells (2 146 689 DoFs, 57 066 625 nnz),
spectral-AMGe two-level hierarchy with 8^3-cell agglomerates x 1 eigenvector (n_c = 4096); with --gpus N it is weak
scaled: N such cubes stacked along z, one z-slab per GPU.  Operators come from the host setup path and are uploaded
once; the timed region contains only V-cycles.  Prints ONE JSON line, which also carries

  parity      the timed hierarchy's V-cycle and PCG against the SERIAL CPU oracle on the same global problem (every N);
              the process exits non-zero when a north-star tolerance is exceeded
  north_star  BASELINE configs[3] (3D Q1, 512^3 cells, 135 M DoFs): one GPU at N = 1, row-partitioned strong scaling
              at N > 1 (skipped with a reason when host memory or time do not allow it)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_emit = print
METRIC = "vcycle_apply_throughput"
UNIT = "V-cycles/s"
TOL_VCYCLE, TOL_HIST, TOL_PCG = 1e-12, 1e-10, 1e-8     # north star: per-application 1e-12, histories 1e-10


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=128, help="cells per direction (cfg1: 128)")
    ap.add_argument("--block", type=int, default=8, help="agglomerate edge in cells")
    ap.add_argument("--neig", type=int, default=1, help="eigenvectors per agglomerate")
    ap.add_argument("--degree", type=int, default=1)
    ap.add_argument("--material", default="constant")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--lanes", type=int, default=0, help="override lanes per row of A (0 = automatic)")
    ap.add_argument("--repeats", type=int, default=10, help="extra K-step regions timed for min / median / spread")
    ap.add_argument("--parity", default="oracle", choices=["oracle", "props", "none"],
                    help="oracle: V-cycle + PCG against the serial CPU oracle on the same global problem; props: "
                         "size-independent properties only (configurations the oracle cannot finish in the budget)")
    ap.add_argument("--pcg", action="store_true", help="(kept for compatibility: PCG is part of --parity oracle)")
    ap.add_argument("--matrix-free", action="store_true",
                    help="level 0 is the matrix-free operator (BASELINE configs[4]); R, P and A_c stay assembled")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = N cubes of --cells^3 stacked along z (the default the driver runs); "
                         "strong = one --cells^3 cube row-partitioned over the N GPUs (BASELINE configs[3])")
    ap.add_argument("--levels", type=int, default=2,
                    help="N = 1: hierarchy depth; levels beyond the second aggregate 2x2x2 agglomerates each "
                         "(hostsetup.build_multilevel; the reference cannot build more than two levels itself)")
    ap.add_argument("--north-star", default="auto", choices=["auto", "on", "off"],
                    help="also measure BASELINE configs[3] (512^3 cells): auto = only for the default workload")
    ap.add_argument("--north-star-cells", type=int, default=512)
    ap.add_argument("--north-star-timeout", type=float, default=480.0)
    ap.add_argument("--other-configs", default="auto", choices=["auto", "on", "off"],
                    help="N = 1: also measure BASELINE configs[2] (Q2) and configs[4] (matrix-free fine level) in child "
                         "processes and report them under 'other_configs': auto = only for the default workload")
    ap.add_argument("--other-configs-timeout", type=float, default=200.0)
    return ap.parse_args(argv)


# ---------------------------------------------------------------------------------------------------------------
# the workload, described identically by both arms
# ---------------------------------------------------------------------------------------------------------------
def global_cells(args, world):
    c = args.cells
    return (c, c, c * world if args.scaling == "weak" else c)


def canonical_config(args, world, n, nnz, n_c):
    """`config` of the JSON line: a function of the command line and the problem sizes only, so that the reference arm
    and ours print the same dict for the same problem."""
    cx, cy, cz = global_cells(args, world)
    units = world if args.scaling == "weak" else 1
    return {
        "workload": (f"3D Q{args.degree} {args.material} Laplace, {cx}x{cy}x{cz} cells (h=1/{args.cells}), n={n}, "
                     f"nnz={nnz}; two-level spectral AMGe, {args.block}^3-cell agglomerates x {args.neig} eigvec "
                     f"(n_c={n_c}); V(1,1) Jacobi omega=1, direct LU coarse solve, preconditioner mode; level 0 "
                     + ("matrix-free (laplace_matrix_free)" if args.matrix_free else "assembled CSR")
                     + (f"; {args.levels} levels (2x2x2 aggregation below level 1)" if args.levels > 2 else "")),
        "scaling": args.scaling,
        "units_per_step": units,
        "value_definition": ("single-GPU-sized V-cycle units per second = units_per_step x (V-cycles of the global "
                             "problem per second); the global problem is units_per_step cubes stacked along z"
                             if args.scaling == "weak" else "V-cycles per second of the one global problem"),
        "l2_policy": ("inputs larger than L2, no flush: one cycle streams the level-0 operator twice, "
                      "%.2f GB per GPU-sized unit vs 126 MB L2" % (12.0 * nnz / max(world, 1) / 1e9)
                      if not args.matrix_free else
                      "inputs larger than L2, no flush: vectors of %.2f GB per GPU-sized unit vs 126 MB L2"
                      % (8.0 * n / max(world, 1) / 1e9)),
    }


def build_global(args, world):
    """Global operators on this host process (N = 1 workload, the reference arm and the parity oracle at every N)."""
    from mfmg_b200 import hostsetup as hs

    cells = global_cells(args, world)
    nodes = [c * args.degree + 1 for c in cells]
    est = 3.0 * 12.0 * (2 * args.degree + 1) ** 3 * float(np.prod(nodes))
    try:
        import psutil

        avail = psutil.virtual_memory().available
        if est > 0.9 * avail:
            raise MemoryError(f"host setup of {np.prod(nodes)} DoFs needs ~{est / 1e9:.0f} GB, {avail / 1e9:.0f} GB available")
    except ImportError:
        pass
    t0 = time.time()
    h = (1.0 / args.cells,) * 3
    if world == 1 or args.scaling == "strong":
        P = hs.LaplaceProblem.create(3, args.degree, args.cells, args.material)
    else:
        P = hs.LaplaceProblem.create_box(3, args.degree, cells, h, args.material)
    R = hs.build_restrictor(P, (args.block,) * 3, args.neig)
    Ac = hs.galerkin(P.A, R)
    if args.levels > 2:   # deeper hierarchy: (operators, restrictors) lists ride along on the problem object
        grid = tuple(-(-c // args.block) for c in cells)
        ops, res = [P.A, Ac], [R]
        for _ in range(args.levels - 2):
            R2 = hs.aggregate_restrictor(grid, (2, 2, 2), args.neig)
            res.append(R2)
            ops.append(hs.galerkin(ops[-1], R2))
            grid = tuple(-(-g // 2) for g in grid)
        P.levels = (ops, res)
    return P, R, Ac, time.time() - t0


def algorithmic_bytes(P, R, Ac, mf_cells=None, mf_nq=8):
    """bytes per kernel / per V-cycle (8 B values, 4 B columns, 4 B row offsets).
    mf_cells: level 0 is matrix-free with that many cells and mf_nq coefficient entries per cell."""
    n, nnz, nc, nnzr = P.n, P.A.nnz, Ac.n_rows, R.nnz
    ro = 8 if nnz >= 2 ** 31 - 64 else 4   # row offsets are 64-bit past 2^31 non-zeros
    spmv = 12 * nnz + ro * (n + 1) + 8 * n + 8 * n
    jacobi = 12 * nnz + ro * (n + 1) + 8 * n + 8 * n + 8 * n + 8 * n
    if mf_cells is not None:
        spmv = 16 * n + 8 * mf_cells * mf_nq + n       # x in, y out, coefficient table, constraint flags
        jacobi = spmv + 8 * n + 8 * n                   # + b, D^-1 (x_old is the gathered x)
    resid = spmv + 8 * n
    restrict = 12 * nnzr + 4 * (nc + 1) + 8 * n + 8 * nc
    prolong = 12 * nnzr + 4 * (n + 1) + 8 * nc + 16 * n
    dense = 8 * nc * nc + 16 * nc
    zero_guess = 24 * n
    vcycle = zero_guess + resid + restrict + dense + prolong + jacobi
    return dict(spmv=spmv, jacobi=jacobi, resid=resid, restrict=restrict, prolong=prolong, dense=dense,
                zero_guess=zero_guess, vcycle=vcycle)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t_begin=None, t_end=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ts, line in self.rows:
            if t_begin is not None and not (t_begin - 0.05 <= ts <= t_end + 0.15):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}


NCU_TILE_CSV = ("profiles/r02_ncu_full_csr_tile_raw.csv", "profiles/r01_ncu_full_csr_tile_raw.csv")


def ncu_traffic_of_dominant_kernel(args, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of the fused Jacobi sweep from the committed `ncu --set full`
    capture of this very command line.  Only valid for the default single-GPU cfg1 workload; None otherwise."""
    if world != 1 or args.matrix_free or (args.cells, args.block, args.neig, args.degree, args.material) != \
            (128, 8, 1, 1, "constant"):
        return None, None
    import csv

    for rel in NCU_TILE_CSV:
        try:
            rows = list(csv.reader(open(os.path.join(ROOT, rel))))
            hdr, units = rows[0], rows[1]
            for r in rows[2:]:
                if "csr_tile_kernel<4, 2," in r[hdr.index("Kernel Name")]:
                    total = 0.0
                    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                        i = hdr.index(key)
                        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
                        total += float(r[i]) * scale
                    return total, rel + " (ncu --set full, same command line)"
        except (OSError, ValueError, KeyError, IndexError):
            continue
    return None, None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def oracle_hierarchy(P, R, Ac, matrix_free=False):
    import oracle

    fine = (P.n, P.A.rowptr, P.A.col, P.A.val)
    if matrix_free:
        fine = oracle.MatrixFreeLaplace(3, P.degree, P.cells, P.h, P.coef_per_q(), P.constrained)
    if getattr(P, "levels", None):
        ops, res = P.levels
        return oracle.Hierarchy([fine] + [(o.n_rows, o.rowptr, o.col, o.val) for o in ops[1:]],
                                [(r.n_rows, r.n_cols, r.rowptr, r.col, r.val) for r in res], 1, True)
    return oracle.Hierarchy([fine, (Ac.n_rows, Ac.rowptr, Ac.col, Ac.val)],
                            [(R.n_rows, R.n_cols, R.rowptr, R.col, R.val)], 1, True)


def cpu_vcycle_rate(Ho, n, budget_s=12.0, max_cycles=40):
    """The oracle (CPU port of the reference host path) on a bounded sample of the same workload."""
    import oracle

    cores = oracle.num_threads()
    rng = np.random.default_rng(0)
    b = rng.standard_normal(n)
    Ho.vmult(b)  # warm-up
    t0 = time.perf_counter()
    k = 0
    while k < max_cycles and (time.perf_counter() - t0 < budget_s or k < 2):
        Ho.vmult(b)
        k += 1
    dt = time.perf_counter() - t0
    return k / dt, cores, k


def run_reference(args):
    """The reference arm: the CPU restatement of the reference host path (oracle/, OpenMP over all host threads) on
    the SAME global problem as our arm at --gpus N / --scaling (rank 0 alone works)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle

    world = max(1, args.gpus)
    avail = host_threads()   # launchers like torchrun export OMP_NUM_THREADS=1
    oracle.set_num_threads(avail)
    from mfmg_b200 import hostsetup as hs

    hs.set_num_threads(avail)
    P, R, Ac, setup_s = build_global(args, world)
    t0 = time.time()
    Ho = oracle_hierarchy(P, R, Ac, args.matrix_free)
    factor_s = time.time() - t0
    cores = oracle.num_threads()
    rng = np.random.default_rng(0)
    b = rng.standard_normal(P.n)
    for _ in range(max(args.warmup, 1)):
        Ho.vmult(b)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        Ho.vmult(b)
    dt = time.perf_counter() - t0
    cfg = canonical_config(args, world, P.n, P.A.nnz, Ac.n_rows)
    value = cfg["units_per_step"] * args.steps / dt
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} full V-cycles of the same global problem; CPU restatement of the "
                                   f"reference host path (oracle/mfmg_oracle.c, OpenMP row-parallel, coarse LU on band "
                                   f"storage), not the reference binary (deal.II/Trilinos/MPI unavailable)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "setup_s": {"host_setup": setup_s, "oracle_finalize": factor_s},
    }
    _emit(json.dumps(out))


# ---------------------------------------------------------------------------------------------------------------
# parity of the timed hierarchy against the serial oracle
# ---------------------------------------------------------------------------------------------------------------
def parity_oracle(args, d, handle, H, dist, rank, world, n_local, row_begin, global_ops, n_global):
    """V-cycle output and PCG (iteration count, residual history) of the hierarchy that was just timed against the
    serial CPU oracle on the same global problem.  global_ops = (P, R, Ac) on rank 0 (None elsewhere).
    Every rank calls this; rank 0 returns the dict."""
    import ctypes

    import oracle

    t_start = time.time()
    rng = np.random.default_rng(20261018)
    b_glob = rng.standard_normal(n_global)          # the same stream on every rank
    sl = slice(row_begin, row_begin + n_local)
    bl = np.zeros(H.vector_size)
    bl[:n_local] = b_glob[sl]
    b = d.DeviceVector.from_host(handle, bl)
    x = d.DeviceVector(handle, H.vector_size)
    H.vmult(x, b)
    handle.synchronize()
    x_loc = x.to_host()[:n_local].copy()
    Ho = x0_glob = None
    if rank == 0:
        P, R, Ac = global_ops
        Ho = oracle_hierarchy(P, R, Ac, args.matrix_free)
        x0_glob = oracle.std_uniform01(P.n, skip=P.constrained)   # tests/hierarchy_driver.cc:153-164
    if dist is not None:
        parts = [None] * world if rank == 0 else None
        dist.gather_object((row_begin, x_loc), parts, dst=0)
        x0_list = [None]
        sizes = [None] * world
        dist.all_gather_object(sizes, (row_begin, n_local))
        scatter = [x0_glob[rb:rb + nl] for rb, nl in sizes] if rank == 0 else None
        dist.scatter_object_list(x0_list, scatter, src=0)
        x0_loc = x0_list[0]
    else:
        parts = [(row_begin, x_loc)]
        x0_loc = x0_glob
    # PCG on the GPU(s): b = 0, x0 = the driver's random guess, absolute tolerance 1e-8
    # (the oracle's matrix-free cell loop is serial: beyond a few million DoFs its PCG does not fit the bench budget --
    # the V-cycle is still compared with the oracle, the PCG is checked through its true residual instead)
    oracle_pcg = not (args.matrix_free and n_global > 3_000_000)
    max_it = 200
    xl = np.zeros(H.vector_size)
    xl[:n_local] = x0_loc
    x.upload(xl)
    b.fill(0.0)
    hist = np.zeros(max_it + 1)
    it = ctypes.c_int(0)
    A0 = H.operators[0]
    a_ptr = A0.ptr if isinstance(A0, d.SparseMatrixDevice) else None
    H.use_graph(not args.no_graph)
    t0 = time.perf_counter()
    rc = handle.lib.mfmgb_pcg(handle.ctx, H.ptr, a_ptr, b.ptr, x.ptr, TOL_PCG, max_it, ctypes.byref(it),
                              hist.ctypes.data)
    handle.synchronize()
    gpu_s = time.perf_counter() - t0
    d.check(handle.ctx, rc)
    if rank != 0:
        return None, None
    x_gpu = np.empty(n_global)
    for rb, xs in parts:
        x_gpu[rb:rb + len(xs)] = xs
    x_ref = Ho.vmult(b_glob)
    v_err = float(np.linalg.norm(x_gpu - x_ref) / np.linalg.norm(x_ref))
    hist = hist[:it.value + 1]
    if oracle_pcg:
        t0 = time.perf_counter()
        _, it_ref, hist_ref = Ho.pcg(np.zeros(n_global), x0_glob, TOL_PCG, max_it)
        cpu_s = time.perf_counter() - t0
        m = min(len(hist), len(hist_ref))
        h_err = float(np.max(np.abs(hist[:m] - hist_ref[:m]) / hist_ref[:m]))
        ok = v_err <= TOL_VCYCLE and it.value == it_ref and h_err <= TOL_HIST
    else:
        it_ref, h_err, cpu_s = None, None, None
        ok = v_err <= TOL_VCYCLE and hist[-1] <= TOL_PCG
    return Ho, {"against": "serial CPU oracle (oracle/mfmg_oracle.c) on the same global problem, n=%d" % n_global,
            "vcycle_rel_err": v_err, "vcycle_tol": TOL_VCYCLE,
            "pcg_tol_abs": TOL_PCG, "pcg_iters_gpu": int(it.value),
            "pcg_iters_oracle": int(it_ref) if it_ref is not None else "skipped (serial matrix-free oracle, n > 3e6)",
            "hist_max_rel": h_err, "hist_tol": TOL_HIST, "residual_0": float(hist[0]), "residual_last": float(hist[-1]),
            "gpu_solve_s": gpu_s, "oracle_solve_s": cpu_s, "oracle_threads": oracle.num_threads(), "ok": bool(ok),
            "wall_s": time.time() - t_start}


def parity_props(args, d, handle, H, dist, rank, world, n_local):
    """Size-independent properties for configurations whose oracle does not finish in the bench budget: the V-cycle is
    linear and symmetric (it is a CG preconditioner), CUDA-graph replay equals eager launches bit for bit, and PCG
    reaches the tolerance with a true residual below it."""
    import ctypes

    import torch

    def dot(u, v):
        s = float(np.dot(u, v))
        if dist is not None:
            t = torch.tensor([s], dtype=torch.float64, device="cuda")
            dist.all_reduce(t)
            s = float(t[0])
        return s

    rng = np.random.default_rng(77 + rank)
    u_h, v_h = rng.standard_normal(n_local), rng.standard_normal(n_local)

    def apply(w_h, graph):
        H.use_graph(graph)
        wl = np.zeros(H.vector_size)
        wl[:n_local] = w_h
        b = d.DeviceVector.from_host(handle, wl)
        x = d.DeviceVector(handle, H.vector_size)
        H.vmult(x, b)
        if graph:
            H.vmult(x, b)
        handle.synchronize()
        return x.to_host()[:n_local].copy()

    Mu, Mv = apply(u_h, False), apply(v_h, False)
    Muv = apply(2.0 * u_h - 3.0 * v_h, False)
    lin = np.sqrt(dot(Muv - (2.0 * Mu - 3.0 * Mv), Muv - (2.0 * Mu - 3.0 * Mv)) / dot(Muv, Muv))
    sym = abs(dot(u_h, Mv) - dot(Mu, v_h)) / np.sqrt(dot(u_h, u_h) * dot(Mv, Mv))
    replay_equal = bool(np.array_equal(apply(u_h, True), Mu))
    if dist is not None:
        t = torch.tensor([1.0 if replay_equal else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        replay_equal = bool(t[0] > 0.5)
    # PCG from a random guess with b = 0
    max_it = 200
    xl = np.zeros(H.vector_size)
    xl[:n_local] = rng.random(n_local)
    x = d.DeviceVector.from_host(handle, xl)
    b = d.DeviceVector(handle, H.vector_size)
    b.fill(0.0)
    hist = np.zeros(max_it + 1)
    it = ctypes.c_int(0)
    A0 = H.operators[0]
    H.use_graph(not args.no_graph)
    rc = handle.lib.mfmgb_pcg(handle.ctx, H.ptr, A0.ptr if isinstance(A0, d.SparseMatrixDevice) else None, b.ptr,
                              x.ptr, TOL_PCG, max_it, ctypes.byref(it), hist.ctypes.data)
    handle.synchronize()
    d.check(handle.ctx, rc)
    ok = lin < 1e-12 and sym < 1e-12 and replay_equal and hist[it.value] <= TOL_PCG
    return {"against": "size-independent properties (the oracle does not finish in the bench budget at this size)",
            "linearity_rel": float(lin), "symmetry_rel": float(sym), "graph_replay_bitwise": replay_equal,
            "pcg_iters_gpu": int(it.value), "residual_0": float(hist[0]), "residual_last": float(hist[it.value]),
            "ok": bool(ok)}


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
def measure(args, d, handle, stream, H, dist, rank, world, local_rank, n_local, nbytes, units):
    """Timed region + repeats + stage times + SpMV + e2e for one finalized hierarchy; returns a dict of raw numbers."""
    import torch

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(values):
        if dist is None:
            return [float(v) for v in values]
        t = torch.tensor(values, device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    H.use_graph(not args.no_graph)
    rng = np.random.default_rng(rank)
    b_h = np.zeros(H.vector_size)
    b_h[:n_local] = rng.standard_normal(n_local)
    b = d.DeviceVector.from_host(handle, b_h)
    x = d.DeviceVector(handle, H.vector_size)
    K, W = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(local_rank)
    res = {}
    with torch.cuda.stream(stream):
        for _ in range(W):
            H.vmult(x, b)
        barrier()
        sampler.start()
        time.sleep(0.25)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = handle.launch_count
        t_begin = time.time()
        barrier()
        ev0.record(stream)
        for _ in range(K):
            H.vmult(x, b)
        ev1.record(stream)
        barrier()
        launches = handle.launch_count - launches0
        ms_total = ev0.elapsed_time(ev1)
        # the same K-step region repeated: min / median / max of the per-step time (max over ranks each)
        reps = []
        for _ in range(max(0, args.repeats)):
            barrier()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record(stream)
            for _ in range(K):
                H.vmult(x, b)
            r1.record(stream)
            barrier()
            reps.append(r0.elapsed_time(r1) / K)
        t_end = time.time()
        clocks = sampler.stop(t_begin, t_end)

        # per-stage durations of the same cycle (CUDA events between the level-0 stages, un-captured launches); the
        # ranks are aligned by a barrier before every profiled cycle, so a stage holds its own wait for the
        # neighbours, not the skew accumulated over earlier cycles
        n_prof = min(K, 20)
        stages = {k: 0.0 for k in d.Hierarchy.STAGES}
        H.profile(x, b)
        for _ in range(n_prof):
            barrier()
            for k, v in H.profile(x, b).items():
                stages[k] += v / n_prof

        # the same cycle resolved per kernel group, timed INSIDE a CUDA-graph replay (event-record nodes); rank 0's view
        n_tl = 5
        timeline = None
        for _ in range(n_tl):
            barrier()
            tl = H.timeline(x, b, graph=not args.no_graph)
            if timeline is None:
                timeline = [[name, 0.0] for name, _ in tl]
            for k, (_, v) in enumerate(tl):
                timeline[k][1] += v / n_tl

        # plain SpMV y = A x (the BASELINE's "SpMV HBM GB/s")
        Ad = H.operators[0]
        y = d.DeviceVector(handle, n_local)
        if world > 1:
            H.halo.exchange(b)
        fine_apply = (lambda: Ad.apply(b, y)) if args.matrix_free else (lambda: Ad.vmult(y, b))
        for _ in range(3):
            fine_apply()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(K):
            fine_apply()
        e1.record(stream)
        torch.cuda.synchronize()
        spmv_ms = e0.elapsed_time(e1) / K

        # end to end through the host-vector entry point: pinned host b -> H2D -> V-cycle -> D2H -> pinned host x
        b_pin = torch.from_numpy(b_h[:n_local].copy()).pin_memory()
        x_pin = torch.empty(n_local, dtype=torch.float64).pin_memory()
        for _ in range(3):
            H.vmult_host_ptr(x_pin.data_ptr(), b_pin.data_ptr())
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tw0 = time.perf_counter()
        g0.record(stream)
        for _ in range(K):
            H.vmult_host_ptr(x_pin.data_ptr(), b_pin.data_ptr())
        g1.record(stream)
        barrier()
        tw1 = time.perf_counter()
        e2e_seq_ms = max(g0.elapsed_time(g1), 1e3 * (tw1 - tw0)) / K
        checksum = float(x_pin.double().abs().sum())
        # the same K steps through the batch entry point: every step still copies its own right-hand side H2D from
        # pinned memory and its own result D2H, but the three stages of consecutive steps overlap (full-duplex PCIe)
        n_pin = min(4, K)
        b_pins = [b_pin] + [torch.from_numpy(b_h[:n_local].copy()).pin_memory() for _ in range(n_pin - 1)]
        x_pins = [x_pin] + [torch.empty(n_local, dtype=torch.float64).pin_memory() for _ in range(n_pin - 1)]
        bp = [b_pins[j % n_pin].data_ptr() for j in range(K)]
        xp = [x_pins[j % n_pin].data_ptr() for j in range(K)]
        H.vmult_host_batch_ptr(xp[:n_pin], bp[:n_pin])
        barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tw0 = time.perf_counter()
        h0.record(stream)
        H.vmult_host_batch_ptr(xp, bp)
        h1.record(stream)
        barrier()
        tw1 = time.perf_counter()
        e2e_ms = max(h0.elapsed_time(h1), 1e3 * (tw1 - tw0)) / K
        checksum_batch = float(x_pins[(K - 1) % n_pin].double().abs().sum())
        assert checksum_batch == checksum, "batch and sequential host-vector entry points disagree"

    vals = max_over_ranks([ms_total, e2e_ms, e2e_seq_ms] + reps + [stages[k] for k in d.Hierarchy.STAGES])
    ms_total, e2e_ms, e2e_seq_ms = vals[0], vals[1], vals[2]
    reps = vals[3:3 + len(reps)]
    stages = dict(zip(d.Hierarchy.STAGES, vals[3 + len(reps):]))
    res.update(ms_per_step=ms_total / K, e2e_ms=e2e_ms, e2e_seq_ms=e2e_seq_ms, reps=reps, stages=stages, spmv_ms=spmv_ms, clocks=clocks,
               timeline=timeline,
               launches=int(launches), checksum=checksum, K=K, W=W)
    return res


def build_ours(args, d, handle, dist, rank, world):
    """Host setup + upload of this rank's share.  Returns (H, info)."""
    from mfmg_b200 import hostsetup as hs

    info = {}
    if world == 1:
        P, R, Ac, setup_s = build_global(args, 1)
        t0 = time.time()
        if args.matrix_free:
            M = d.MatrixFreeLaplaceDevice(handle, 3, args.degree, P.cells, P.h, P.coef_per_q(), P.constrained)
            H = d.Hierarchy(handle, [M, d.SparseMatrixDevice.from_host(handle, Ac)],
                            [d.SparseMatrixDevice.from_host(handle, R)], {"is preconditioner": True})
            mf_nq = 0 if "stencil" in M.kernel else (1 if "per-cell" in M.kernel else (args.degree + 1) ** 3)
            nbytes = algorithmic_bytes(P, R, Ac, int(np.prod(P.cells)), mf_nq)
        elif getattr(P, "levels", None):
            ops, res = P.levels
            H = d.Hierarchy(handle, [d.SparseMatrixDevice.from_host(handle, o) for o in ops],
                            [d.SparseMatrixDevice.from_host(handle, r) for r in res], {"is preconditioner": True})
            nbytes = algorithmic_bytes(P, R, Ac)
            # deeper levels: the dense solve moves to the last level, levels 1 .. L-2 cost a zero-guess sweep, a
            # residual, a fused Jacobi sweep and their transfers each
            extra = -nbytes["dense"] + 8 * ops[-1].n_rows ** 2 + 16 * ops[-1].n_rows
            for lvl in range(1, len(ops) - 1):
                nl, nz, rr = ops[lvl].n_rows, ops[lvl].nnz, res[lvl]
                extra += 24 * nl + 2 * (12 * nz + 4 * (nl + 1) + 16 * nl) + 8 * nl + 16 * nl
                extra += 2 * 12 * rr.nnz + 4 * (rr.n_rows + 1) + 4 * (nl + 1) + 8 * nl + 8 * rr.n_rows + 8 * rr.n_rows + 16 * nl
            nbytes["vcycle"] += extra
            nbytes["dense"] = 8 * ops[-1].n_rows ** 2 + 16 * ops[-1].n_rows
        else:
            H = d.Hierarchy.from_host(handle, P.A, R, Ac, {"is preconditioner": True})
            nbytes = algorithmic_bytes(P, R, Ac)
        handle.synchronize()
        info.update(setup_s=setup_s, upload_s=time.time() - t0, n_local=P.n, row_begin=0, n_global=P.n,
                    nnz_global=P.A.nnz, n_c=Ac.n_rows, global_ops=(P, R, Ac), nbytes=nbytes,
                    parallelism="1 GPU", coarse="dense LU (explicit inverse, one GEMV per apply)")
        return H, info

    if getattr(handle, "nranks", None) is None:   # (the north-star leg re-uses the communicator of the main leg)
        handle.init_comm_from_torch()
    hs.set_num_threads(max(1, host_threads() // world))  # torchrun exports OMP_NUM_THREADS=1

    def gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    t0 = time.time()
    c = args.cells
    cells = global_cells(args, world)
    part = hs.build_slab_part(args.degree, cells, (1.0 / c,) * 3, args.material, (args.block,) * 3, args.neig, world,
                              rank, gather)
    setup_s = time.time() - t0
    t0 = time.time()
    H = d.Hierarchy.from_partition(handle, part, {"is preconditioner": True}, matrix_free=args.matrix_free)
    handle.synchronize()
    if H.coarse_dd is not None:
        coarse_desc = (f"domain-decomposed direct coarse solve (interior {H.coarse_dd.n_interior} rows per rank, "
                       f"{H.coarse_dd.n_separator} separator rows, one all-reduce)")
    else:
        coarse_desc = ("dense coarse solve " + ("split by rows over the ranks (one all-gather)"
                                                if part.Ac.n_rows >= 8192 else "replicated"))
    upload_s = time.time() - t0

    class _Shape:  # sizes of this rank's share for the byte counts
        pass

    Pl, Rl, Acl = _Shape(), _Shape(), _Shape()
    Pl.n, Pl.A = part.n_owned, part.A
    Rl.nnz = part.R.nnz
    Acl.n_rows = part.Ac.n_rows
    if args.matrix_free:
        M = H.operators[0]
        nbytes = algorithmic_bytes(Pl, Rl, Acl, int(np.prod(part.mf["cells"])),
                                   0 if "stencil" in M.kernel else (1 if "per-cell" in M.kernel else 8))
    else:
        nbytes = algorithmic_bytes(Pl, Rl, Acl)
    if H.coarse_dd is not None:   # per-GPU bytes of the domain-decomposed coarse solve instead of 8 n_c^2
        nbytes["vcycle"] += H.coarse_dd.bytes_per_solve - nbytes["dense"]
        nbytes["dense"] = H.coarse_dd.bytes_per_solve
    import torch

    t = torch.tensor([part.A.nnz], dtype=torch.int64, device="cuda")
    dist.all_reduce(t)
    info.update(setup_s=setup_s, upload_s=upload_s, n_local=part.n_owned, row_begin=part.row_begin,
                n_global=part.n_global, nnz_global=int(t[0]), n_c=part.Ac.n_rows, global_ops=None, nbytes=nbytes,
                coarse=coarse_desc,
                parallelism=(f"{world} GPUs, rows partitioned in z-slabs ({part.n_owned} rows, {part.A.nnz} nnz, "
                             f"{part.n_ghost} ghosts on rank 0), halo exchange of 1 node plane each way overlapped with "
                             f"interior rows ({H.transport}), {coarse_desc}"))
    return H, info


def north_star_subprocess(args):
    """N = 1: BASELINE configs[3] on one GPU in a child process (its own host setup of 135 M DoFs), bounded in time."""
    try:
        import psutil

        avail = psutil.virtual_memory().available
    except ImportError:
        avail = 0
    nodes = args.north_star_cells + 1
    need = 3.0 * 12.0 * 27.0 * nodes ** 3 * 1.15
    if avail and need > avail:
        return {"skipped": f"host setup of {nodes}^3 DoFs needs ~{need / 1e9:.0f} GB, {avail / 1e9:.0f} GB available"}
    cmd = [sys.executable, os.path.abspath(__file__), "--cells", str(args.north_star_cells), "--block", "16",
           "--steps", str(max(3, min(args.steps, 10))), "--warmup", "3", "--repeats", "3", "--no-cpu-baseline",
           "--parity", "props", "--north-star", "off"]
    t0 = time.time()
    try:
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                             timeout=args.north_star_timeout)
    except subprocess.TimeoutExpired:
        return {"skipped": f"did not finish within {args.north_star_timeout:.0f} s"}
    if res.returncode != 0:
        return {"skipped": f"child exited {res.returncode}: {res.stderr[-300:]}"}
    try:
        line = json.loads(res.stdout.strip().splitlines()[-1])
    except (ValueError, IndexError):
        return {"skipped": "child printed no JSON line"}
    return summarise_north_star(line, time.time() - t0)


OTHER_CONFIGS = (
    ("cfg2", "BASELINE configs[2]: 3D Q2 discontinuous diffusion, 100^3 cells",
     ["--cells", "100", "--degree", "2", "--block", "10", "--material", "discontinuous"]),
    ("cfg4", "BASELINE configs[4]: matrix-free Q1 fine level, 256^3 cells, assembled R / A_c",
     ["--cells", "256", "--block", "16", "--matrix-free"]),
)


def other_configs_subprocess(args):
    """N = 1: the two remaining single-GPU configurations of BASELINE.json, each in a bounded child process (host setup,
    upload, 20 timed cycles, size-independent parity properties; their oracle parity lives in tests/ and profiles/)."""
    out = {}
    for key, desc, flags in OTHER_CONFIGS:
        cmd = [sys.executable, os.path.abspath(__file__)] + flags + [
            "--steps", "20", "--warmup", "3", "--repeats", "3", "--no-cpu-baseline", "--parity", "props",
            "--north-star", "off", "--other-configs", "off"]
        t0 = time.time()
        try:
            res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                                 timeout=args.other_configs_timeout)
        except subprocess.TimeoutExpired:
            out[key] = {"config": desc, "skipped": f"did not finish within {args.other_configs_timeout:.0f} s"}
            continue
        if res.returncode != 0:
            out[key] = {"config": desc, "skipped": f"child exited {res.returncode}: {res.stderr[-300:]}"}
            continue
        try:
            line = json.loads(res.stdout.strip().splitlines()[-1])
        except (ValueError, IndexError):
            out[key] = {"config": desc, "skipped": "child printed no JSON line"}
            continue
        out[key] = {"config": desc + " -- " + line["config"]["workload"], "vcycles_per_s": line["value"],
                    "ms_per_step": line["ms_per_step"], "timing": line.get("timing"),
                    "vcycle_roofline": line["vcycle_roofline"], "roofline": line["roofline"],
                    "timeline_in_graph_ms": line.get("timeline_in_graph_ms"), "parity": line.get("parity"),
                    "e2e": {k: line["e2e"][k] for k in ("value", "unit", "ms_per_step") if k in line["e2e"]},
                    "clocks": line["clocks"], "wall_s": time.time() - t0}
    return out


def summarise_north_star(line, wall_s):
    return {"config": "BASELINE configs[3]: " + line["config"]["workload"], "n_gpus": line["n_gpus"],
            "scaling": "strong", "vcycles_per_s": line["value"], "ms_per_step": line["ms_per_step"],
            "steps": line["steps"], "timing": line.get("timing"),
            "vcycle_roofline": line["vcycle_roofline"], "roofline": line["roofline"], "spmv": line["spmv"],
            "e2e": line["e2e"], "clocks": line["clocks"], "parity": line.get("parity"),
            "setup_s": line["details"]["setup_s"], "wall_s": wall_s,
            "target": ">= 70 % of HBM roofline per GPU at 1 GPU; >= 80 % parallel efficiency at 8 GPUs = "
                      "vcycles_per_s(8) / (8 x vcycles_per_s(1))"}


def assemble_line(args, d, H, info, m, world, units):
    """The JSON line of one measured hierarchy."""
    peak, peak_src = measured_peak()
    nbytes = info["nbytes"]
    ms_per_step = m["ms_per_step"]
    value = units * 1e3 / ms_per_step
    e2e_value = units * 1e3 / m["e2e_ms"]
    stages = m["stages"]
    dom_ms = stages["post_smooth"]
    achieved = nbytes["jacobi"] / (dom_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic_of_dominant_kernel(args, world)
    reps = sorted(m["reps"])
    timing = None
    if reps:
        med = float(np.median(reps))
        timing = {"repeats": len(reps), "steps_per_repeat": m["K"], "ms_per_step_min": reps[0],
                  "ms_per_step_median": med, "ms_per_step_max": reps[-1],
                  "spread_pct": 100.0 * (reps[-1] - reps[0]) / med,
                  "value_from_median": units * 1e3 / med}
    n_local = info["n_local"]
    kern0 = H.operators[0].kernel
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": m["K"], "warmup": m["W"],
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": canonical_config(args, world, info["n_global"], info["nnz_global"], info["n_c"]),
        "clocks": m["clocks"],
        "timing": timing,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * n_local * world,
                "d2h_bytes_per_step": 8 * n_local * world, "ms_per_step": m["e2e_ms"], "checksum_abs_x": m["checksum"],
                "api": "mfmgb_vcycle_host_batch: the K steps as K independent host right-hand sides, each copied H2D from "
                       "pinned memory and its result copied D2H; H2D of step j+1, the cycle of step j and D2H of step j-1 "
                       "overlap",
                "sequential": {"value": units * 1e3 / m["e2e_seq_ms"], "ms_per_step": m["e2e_seq_ms"],
                               "api": "mfmgb_vcycle_host: one synchronous call per step (H2D, cycle, D2H back to back)"}},
        "gpu_launches": m["launches"],
        "roofline": {"bound": "hbm",
                     "kernel": (f"mf_q1_kernel<Jacobi epilogue> ({kern0})" if args.matrix_free else
                                f"csr_{kern0}_kernel<Jacobi epilogue>") + " (post-smoothing sweep on A)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": nbytes["jacobi"], "launch_ms": dom_ms},
        "vcycle_roofline": {"algorithmic_bytes_per_cycle": nbytes["vcycle"],
                            "achieved_gbs": nbytes["vcycle"] / (ms_per_step * 1e-3) / 1e9,
                            "frac_of_measured_peak": nbytes["vcycle"] / (ms_per_step * 1e-3) / 1e9 / peak,
                            "frac_of_nominal_8TBs": nbytes["vcycle"] / (ms_per_step * 1e-3) / 1e9 / 8000.0},
        "spmv": {"gbs": nbytes["spmv"] / (m["spmv_ms"] * 1e-3) / 1e9, "ms": m["spmv_ms"],
                 "frac_of_measured_peak": nbytes["spmv"] / (m["spmv_ms"] * 1e-3) / 1e9 / peak,
                 "frac_of_nominal_8TBs": nbytes["spmv"] / (m["spmv_ms"] * 1e-3) / 1e9 / 8000.0},
        "stage_ms": stages,
        "timeline_in_graph_ms": [[name, round(v, 5)] for name, v in (m.get("timeline") or [])],
        "stage_gbs": {"residual": nbytes["resid"] / (stages["residual"] * 1e-3) / 1e9,
                      "restrict": nbytes["restrict"] / (stages["restrict"] * 1e-3) / 1e9,
                      "coarse": nbytes["dense"] / (stages["coarse"] * 1e-3) / 1e9,
                      "prolong_correct": nbytes["prolong"] / (stages["prolong_correct"] * 1e-3) / 1e9,
                      "post_smooth": achieved},
        "details": {"parallelism": info["parallelism"], "coarse_solver": info["coarse"], "cuda_graph": not args.no_graph,
                    "lanes_per_row_A": None if args.matrix_free else H.operators[0].lanes_per_row, "kernel_A": kern0,
                    "lanes_per_row_R": H.restrictors[0].lanes_per_row,
                    "launches_per_cycle": m["launches"] // max(m["K"], 1),
                    "setup_s": {"host_setup": info["setup_s"], "upload_and_factor": info["upload_s"]},
                    "global_vcycles_per_s": 1e3 / ms_per_step},
    }
    return out


def run_ours(args):
    import torch

    from mfmg_b200 import device as d

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    stream = torch.cuda.Stream()
    handle = d.CudaHandle(local_rank, stream=stream.cuda_stream)
    units = world if args.scaling == "weak" else 1
    H, info = build_ours(args, d, handle, dist, rank, world)
    if args.lanes and not args.matrix_free:
        H.operators[0].set_lanes_per_row(args.lanes)
    m = measure(args, d, handle, stream, H, dist, rank, world, local_rank, info["n_local"], info["nbytes"], units)
    out = assemble_line(args, d, H, info, m, world, units) if rank == 0 else None

    # ---- parity of the hierarchy that was just timed ----
    parity = None
    Ho = gops = None
    if args.parity == "oracle":
        import oracle

        oracle.set_num_threads(host_threads())
        gops = info["global_ops"]
        if world > 1 and rank == 0:
            from mfmg_b200 import hostsetup as hs

            hs.set_num_threads(host_threads())
            P, R, Ac, _ = build_global(args, world)
            gops = (P, R, Ac)
        Ho, parity = parity_oracle(args, d, handle, H, dist, rank, world, info["n_local"], info["row_begin"], gops,
                                   info["n_global"])
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            rate, cores, k = cpu_vcycle_rate(Ho, info["n_global"])
            out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": f"{k} full V-cycles of the same workload on the host cores (oracle port "
                                             f"of the reference host path; the reference binary cannot be built here)"}
    elif args.parity == "props":
        parity = parity_props(args, d, handle, H, dist, rank, world, info["n_local"])
    if rank == 0:
        out["parity"] = parity
        if "cpu_baseline" not in out and world == 1:
            out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
                                   "sample": "skipped (--no-cpu-baseline or --parity != oracle)"}

    # ---- BASELINE configs[3], driver-visible ----
    default_workload = (args.cells, args.block, args.neig, args.degree, args.material, args.scaling) == \
        (128, 8, 1, 1, "constant", "weak") and not args.matrix_free
    want_ns = args.north_star == "on" or (args.north_star == "auto" and default_workload)
    if want_ns:
        del H   # the cfg1 operators leave the device and the host before the 135 M-DoF leg
        Ho = None
        info["global_ops"] = gops = None
        import gc

        gc.collect()
        try:
            if world == 1:
                ns = north_star_subprocess(args)
            else:
                ns = north_star_in_process(args, d, handle, stream, dist, rank, world, local_rank)
        except Exception as exc:  # never lose the main line to a secondary leg
            ns = {"skipped": f"{type(exc).__name__}: {exc}"}
        if rank == 0:
            out["north_star"] = ns
    want_other = world == 1 and (args.other_configs == "on" or (args.other_configs == "auto" and default_workload))
    if want_other:
        if not want_ns:
            del H
            Ho = None
            info["global_ops"] = gops = None
            import gc

            gc.collect()
        try:
            out["other_configs"] = other_configs_subprocess(args)
        except Exception as exc:  # never lose the main line to a secondary leg
            out["other_configs"] = {"skipped": f"{type(exc).__name__}: {exc}"}
    if rank == 0:
        _emit(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and parity is not None and not parity["ok"]:
        sys.stderr.write("bench.py: PARITY FAILED: %s\n" % json.dumps(parity))
        sys.exit(3)


def north_star_in_process(args, d, handle, stream, dist, rank, world, local_rank):
    """N > 1: BASELINE configs[3] row-partitioned over the N GPUs (strong scaling), same processes."""
    import copy

    import torch

    a = copy.copy(args)
    a.cells, a.block, a.neig, a.degree, a.material = args.north_star_cells, 16, 1, 1, "constant"
    a.scaling, a.matrix_free, a.lanes = "strong", False, 0
    a.steps, a.warmup, a.repeats = max(3, min(args.steps, 10)), 3, 3
    try:
        import psutil

        avail = psutil.virtual_memory().available
    except ImportError:
        avail = 0
    nodes = a.cells + 1
    need = 3.0 * 12.0 * 27.0 * nodes ** 3 * (1.0 + 2.0 * 16 * world / a.cells) * 1.15   # slabs + their overlap layers
    flag = torch.tensor([1 if (not avail or need <= avail) else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag[0]) == 0:
        return {"skipped": f"host setup of {nodes}^3 DoFs over {world} ranks needs ~{need / 1e9:.0f} GB, "
                           f"{avail / 1e9:.0f} GB available"}
    t0 = time.time()
    H, info = build_ours(a, d, handle, dist, rank, world)
    m = measure(a, d, handle, stream, H, dist, rank, world, local_rank, info["n_local"], info["nbytes"], 1)
    par = parity_props(a, d, handle, H, dist, rank, world, info["n_local"])
    if rank != 0:
        return None
    line = assemble_line(a, d, H, info, m, world, 1)
    line["parity"] = par
    return summarise_north_star(line, time.time() - t0)


def main():
    args = parse_args()
    # stdout carries exactly ONE JSON line: anything libraries print there (e.g. "NCCL version ...") goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    real_stdout = os.fdopen(saved_stdout, "w")
    global _emit
    _emit = lambda text: (real_stdout.write(text + "\n"), real_stdout.flush())  # noqa: E731
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
