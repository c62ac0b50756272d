"""Multi-GPU probe (torchrun, one rank per GPU): the cfg1 weak-scaling V-cycle under different launch parameters of the
fused compute + exchange kernel (mfmgb_tunable_set), one hierarchy, one process per rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/probe_weak.py
prints one JSON line per setting on rank 0: ms per cycle (CUDA events on the launching stream around K graph replays,
barrier + synchronize on both sides, max over the ranks) and a check that the result did not change."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from mfmg_b200 import device as d  # noqa: E402

rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local_rank)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
args = bench.parse_args(["--gpus", str(world)] + sys.argv[1:])
stream = torch.cuda.Stream()
handle = d.CudaHandle(local_rank, stream=stream.cuda_stream)
H, info = bench.build_ours(args, d, handle, dist, rank, world)
n_local = info["n_local"]
n_vec = H.vector_size
rng = np.random.default_rng(100 + rank)
b_h = np.zeros(n_vec)
b_h[:n_local] = rng.standard_normal(n_local)
b = d.DeviceVector.from_host(handle, b_h)
H.use_graph(True)
K = 40


def run(setting):
    for k, v in setting.items():
        assert handle.lib.mfmgb_tunable_set(k.encode(), int(v)) == 0, k
    x = d.DeviceVector(handle, n_vec)          # a fresh (b, x) pair: a fresh graph capture with these parameters
    with torch.cuda.stream(stream):
        for _ in range(5):
            H.vmult(x, b)
        times = []
        for _ in range(5):
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(K):
                H.vmult(x, b)
            e1.record(stream)
            torch.cuda.synchronize()
            dist.barrier()
            times.append(e0.elapsed_time(e1) / K)
    t = torch.tensor(times, dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    cs = float(np.abs(x.to_host()[:n_local]).sum())
    return sorted(t.tolist()), cs


settings = [{"halo_push_ctas": 16, "halo_push_penalty": 0}]
for ctas in (8, 16, 32, 64):
    for pen in (2, 4, 6):
        settings.append({"halo_push_ctas": ctas, "halo_push_penalty": pen})
settings.append({"halo_push_ctas": 16, "halo_push_penalty": 0})
ref = None
for s in settings:
    times, cs = run(s)
    ref = cs if ref is None else ref
    if rank == 0:
        print(json.dumps({**s, "ms_min": times[0], "ms_median": times[len(times) // 2], "same_result": cs == ref}), flush=True)
handle.synchronize()
dist.barrier()
dist.destroy_process_group()
