"""Multi-GPU parity (-m gpu, needs >= 2 GPUs; skipped otherwise): runs tests/dist_gpu_worker.py with one rank per GPU."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.gpu
@pytest.mark.parametrize("kernel,transport", [("auto", "peer"), ("tile", "peer"), ("tile", "peer_unfused"),
                                              ("tile", "peer_serial_coarse"), ("auto", "nccl")])
@pytest.mark.parametrize("world", [2, 4])
def test_partitioned_hierarchy_on_gpus(world, kernel, transport):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world),
           os.path.join(ROOT, "tests", "dist_gpu_worker.py")]
    # "tile": interior/boundary row ranges through csr_tile.cu; "peer": ghost entries and small reductions stored
    # into peer memory over NVLink by our own kernels, "nccl": the NCCL send/recv + all-reduce path
    # "peer_unfused": the interior rows / wait kernel / boundary rows form instead of the one-launch form in which
    # the tile kernel waits for the neighbours' flags itself and gathers ghost columns from the mailbox
    # "peer_serial_coarse": the domain-decomposed coarse solve as a chain (GEMV, all-reduce, Schur GEMV) instead of the
    # default two-stream form  y = A_II^-1 b_I  ||  [W b_I -> all-reduce -> Schur GEMV]
    env = dict(os.environ, MFMGB_CSR_KERNEL=kernel, MFMGB_PEER="0" if transport == "nccl" else "1",
               MFMGB_HALO_FUSED="0" if transport == "peer_unfused" else "1",
               MFMGB_COARSE_OVERLAP="0" if transport == "peer_serial_coarse" else "1")
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900, env=env)
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):   # keep the worker's log as evidence (copied to profiles/ by the builder)
        with open(os.path.join(out_dir, f"multi_gpu_parity_w{world}_{kernel}_{transport}.log"), "w") as f:
            f.write(res.stdout)
    assert res.returncode == 0, res.stdout[-4000:]
    for r in range(world):
        assert f"RANK {r} OK" in res.stdout, res.stdout[-4000:]
