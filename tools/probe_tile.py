"""GPU probe: tile-streamed (TMA) vs vector CSR kernels on the cfg1 / Q2 operators; sweeps lanes, ring depth and
CTAs per SM.  JSON to stdout.  Usage: python tools/probe_tile.py [cells] [degree] [quick]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mfmg_b200 import device as d  # noqa: E402
from mfmg_b200 import hostsetup as hs  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 128
degree = int(sys.argv[2]) if len(sys.argv) > 2 else 1
quick = len(sys.argv) > 3

P = hs.LaplaceProblem.create(3, degree, cells)
stream = torch.cuda.Stream()
h = d.CudaHandle(0, stream=stream.cuda_stream)
Ad = d.SparseMatrixDevice.from_host(h, P.A)
n = P.n
rng = np.random.default_rng(0)
x = d.DeviceVector.from_host(h, rng.standard_normal(n))
b = d.DeviceVector.from_host(h, rng.standard_normal(n))
y = d.DeviceVector(h, n)
sm = d.CudaSmoother(d.CudaMatrixOperator(Ad), {})
lib, ctx = h.lib, h.ctx


def timeit(fn, reps=20):
    with torch.cuda.stream(stream):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


b_spmv = 12 * P.A.nnz + 4 * (n + 1) + 16 * n
b_res = b_spmv + 8 * n
b_jac = b_spmv + 16 * n
ops = {
    "spmv": (lambda: Ad.vmult(y, x), b_spmv),
    "resid": (lambda: d.check(ctx, lib.mfmgb_residual_neg(ctx, Ad.ptr, x.ptr, b.ptr, y.ptr)), b_res),
    "jacobi": (lambda: d.check(ctx, lib.mfmgb_jacobi_apply_oop(ctx, sm.ptr, Ad.ptr, b.ptr, x.ptr, y.ptr)), b_jac),
}
out = {"n": n, "nnz": P.A.nnz, "degree": degree, "cells": cells}


def run(tag):
    r = {}
    for name, (fn, nbytes) in ops.items():
        ms = timeit(fn)
        r[name] = {"ms": round(ms, 5), "gbs": round(nbytes / ms / 1e6, 1)}
    out[tag] = r
    print(tag, r, file=sys.stderr, flush=True)


ref = None
for lanes in ((4, 8) if degree == 1 else (8, 16)):
    Ad.set_kernel(0)
    Ad.set_lanes_per_row(lanes)
    run(f"vec_l{lanes}")
    if ref is None:
        Ad.vmult(y, x)
        ref = y.to_host()
lane_set = (2, 4, 8) if degree == 1 else (4, 8, 16)
stage_set = (2, 3) if not quick else (2,)
cta_set = (2, 3, 4) if not quick else (4,)
for lanes in lane_set:
    for stages in stage_set:
        for ctas in cta_set:
            os.environ["MFMGB_TILE_STAGES"] = str(stages)
            os.environ["MFMGB_TILE_CTAS"] = str(ctas)
            Ad.set_lanes_per_row(lanes)  # re-plans the ring
            try:
                Ad.set_kernel(1)
            except d.MfmgError:
                continue
            run(f"tile_l{lanes}_s{stages}_c{ctas}")
Ad.set_lanes_per_row(lane_set[1])
Ad.vmult(y, x)
out["tile_equals_vec_bitwise"] = bool(np.array_equal(np.asarray(y.to_host()), ref)) if lane_set[1] in (4, 8) else None
a = torch.empty(1 << 28, dtype=torch.float64, device="cuda")
bb = torch.empty_like(a)
with torch.cuda.stream(stream):
    ms = timeit(lambda: bb.copy_(a), 10)
out["copy_gbs"] = 2 * a.numel() * 8 / ms / 1e6
print(json.dumps(out, indent=1))
