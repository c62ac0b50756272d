"""Host<->device copy rates of this box with pinned buffers of the e2e leg's size (17 MB vectors), one direction at a
time and both at once -- the bound of bench.py's e2e number (every step moves one vector in and one out)."""
import json
import sys

import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2146689
reps = 40
dev = torch.device("cuda:0")
h_in = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(4)]
h_out = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(4)]
d_in = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(4)]
d_out = [torch.ones(n, dtype=torch.float64, device=dev) for _ in range(4)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_event(e0)
    s2.wait_event(e0)
    for _ in range(reps):
        fn()
    a, b = torch.cuda.Event(), torch.cuda.Event()
    a.record(s1)
    b.record(s2)
    torch.cuda.current_stream().wait_event(a)
    torch.cuda.current_stream().wait_event(b)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


k = [0]


def h2d():
    with torch.cuda.stream(s1):
        d_in[k[0] % 4].copy_(h_in[k[0] % 4], non_blocking=True)
    k[0] += 1


def d2h():
    with torch.cuda.stream(s2):
        h_out[k[0] % 4].copy_(d_out[k[0] % 4], non_blocking=True)
    k[0] += 1


def both():
    with torch.cuda.stream(s1):
        d_in[k[0] % 4].copy_(h_in[k[0] % 4], non_blocking=True)
    with torch.cuda.stream(s2):
        h_out[k[0] % 4].copy_(d_out[k[0] % 4], non_blocking=True)
    k[0] += 1


mb = n * 8 / 1e6
out = {"bytes": n * 8}
for name, fn in (("h2d", h2d), ("d2h", d2h), ("both", both)):
    ms = timed(fn)
    out[name] = {"ms": ms, "GBs_per_direction": mb / ms}
print(json.dumps(out))
