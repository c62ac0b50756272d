"""Diagnostic for the CUDA-graph capture of the partitioned V-cycle (NCCL + two streams).  Run under torchrun with
MFMGB_DIST_GRAPH=1 and a short `timeout`; prints progress markers and dumps the Python stacks if it stalls."""
import faulthandler
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from mfmg_b200 import device as d  # noqa: E402
from mfmg_b200 import hostsetup as hs  # noqa: E402

faulthandler.dump_traceback_later(25, exit=True)
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
handle = d.CudaHandle(int(os.environ["LOCAL_RANK"]))
handle.init_comm_from_torch()


def say(msg):
    print(f"[rank {rank}] {msg}", file=sys.stderr, flush=True)


def gather(obj):
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out


c = int(os.environ.get("DIAG_CELLS", "32"))
part = hs.build_slab_part(1, (c, c, c * world), (1.0 / c,) * 3, "constant", (8, 8, 8), 1, world, rank, gather)
H = d.Hierarchy.from_partition(handle, part, {"is preconditioner": True})
b, x = H.build_vector(), H.build_vector()
bl = np.zeros(H.vector_size)
bl[:part.n_owned] = np.random.default_rng(rank).standard_normal(part.n_owned)
b.upload(bl)
for _ in range(3):
    H.vmult(x, b)
handle.synchronize()
x_eager = x.to_host()[:part.n_owned].copy()
say("eager cycles done")
H.use_graph(True)
H.vmult(x, b)
say("graph captured + first replay enqueued")
handle.synchronize()
say("first replay finished")
for _ in range(5):
    H.vmult(x, b)
handle.synchronize()
say("5 more replays finished; bitwise equal to eager: %s" % np.array_equal(x.to_host()[:part.n_owned], x_eager))
dist.barrier()
handle.close()
say("done")
