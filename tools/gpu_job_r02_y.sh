# round 2, call Y (8 GPUs): the weak-scaling line at HEAD (two-stream coarse solve, stacked restriction, short-row P kernel)
set -x
(time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29682 bench.py --gpus 8 --steps 20 --warmup 5 --north-star off > gpurun_out/r02_bench_cfg1_weak_n8_head.json 2> gpurun_out/r02_bench_cfg1_weak_n8_head.err); echo "bench8 rc=$?"; tail -3 gpurun_out/r02_bench_cfg1_weak_n8_head.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_cfg1_weak_n8_head.json')); print(d['value'], d['ms_per_step'], d['timing']); print(d['parity']); print(d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['sequential']); print(d.get('timeline_in_graph_ms'))"
