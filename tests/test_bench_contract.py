"""bench.py contract checks that need no GPU: the reference arm (the CPU oracle port timed on the host cores) prints
exactly one JSON line on stdout with the keys the driver reads, also under a launcher that exports OMP_NUM_THREADS=1
and on ranks > 0 (which must stay silent); the byte model of SURVEY.md section 8(d)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, *args):
    env = dict(os.environ, OMP_NUM_THREADS="1")
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--cells", "16",
                           "--block", "4", "--steps", "2", "--warmup", "1", *args], env=env, stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, text=True, timeout=600)


def test_reference_arm_prints_one_json_line():
    res = _run()
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    out = json.loads(lines[0])
    assert out["impl"] == "reference" and out["metric"] == "vcycle_apply_throughput" and out["unit"] == "V-cycles/s"
    assert out["higher_is_better"] is True and out["dtype"] == "f64" and out["steps"] == 2
    assert out["value"] > 0 and abs(out["ms_per_step"] * out["value"] - 1e3) < 1e-6 * 1e3
    assert out["cpu_baseline"]["kind"] == "port" and out["cpu_baseline"]["value"] == out["value"]
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    assert out["cpu_baseline"]["cores"] == avail          # all host threads despite OMP_NUM_THREADS=1
    assert out["e2e"] == {"value": out["value"], "unit": "V-cycles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in out["config"]


def test_reference_arm_is_silent_on_other_ranks():
    res = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2")
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_algorithmic_byte_model():
    sys.path.insert(0, ROOT)
    import bench

    class S:
        pass

    P, R, Ac = S(), S(), S()
    P.n, P.A = 1000, S()
    P.A.nnz, R.nnz, Ac.n_rows = 27000, 5000, 10
    b = bench.algorithmic_bytes(P, R, Ac)
    n, nnz, nr, nc = 1000, 27000, 5000, 10
    assert b["spmv"] == 12 * nnz + 4 * (n + 1) + 16 * n
    assert b["vcycle"] == 24 * n + (b["spmv"] + 8 * n) + (12 * nr + 4 * (nc + 1) + 8 * n + 8 * nc) + \
        (8 * nc * nc + 16 * nc) + (12 * nr + 4 * (n + 1) + 8 * nc + 16 * n) + (b["spmv"] + 16 * n)
    # B_V = 24 nnz_A + 24 nnz_R + 116 n + 20 n_c + B_dense (+ the two "+1" row-offset entries)
    assert b["vcycle"] == 24 * nnz + 24 * nr + 116 * n + 20 * nc + 8 * nc * nc + 16 * nc + 4 * 3 + 4 * 0 + 0 or \
        abs(b["vcycle"] - (24 * nnz + 24 * nr + 116 * n + 20 * nc + 8 * nc * nc + 16 * nc)) <= 16
    mf = bench.algorithmic_bytes(P, R, Ac, mf_cells=900, mf_nq=1)
    assert mf["spmv"] == 16 * n + 8 * 900 + n


def test_secondary_legs_degrade_to_skipped_entries():
    """The `other_configs` leg (BASELINE configs[2] and [4] in child processes) never raises: without a CUDA device the
    children exit non-zero and every entry becomes {"config": ..., "skipped": reason} -- the main line is not lost."""
    sys.path.insert(0, ROOT)
    import bench

    args = bench.parse_args([])
    args.other_configs_timeout = 300.0
    out = bench.other_configs_subprocess(args)
    assert set(out) == {"cfg2", "cfg4"}
    for entry in out.values():
        assert "config" in entry
        assert ("skipped" in entry) != ("vcycles_per_s" in entry)
    import torch

    if not torch.cuda.is_available():
        assert all("skipped" in e for e in out.values())
