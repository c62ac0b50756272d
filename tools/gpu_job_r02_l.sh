# round 2, call L (2 GPUs): validate the in-process north-star leg (strong scaling) at a reduced size
set -x
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29691 bench.py --gpus 2 --steps 10 --warmup 3 --north-star on --north-star-cells 256 > gpurun_out/r02_bench_n2_ns256.json 2> gpurun_out/r02_bench_n2_ns256.err); echo "rc=$?"; tail -3 gpurun_out/r02_bench_n2_ns256.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n2_ns256.json')); print(d['value'], d['parity']['ok']); print(json.dumps(d['north_star'])[:1500])"
