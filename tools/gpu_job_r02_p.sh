# round 2, call P (2 GPUs): overlapped coarse solve -- parity worker on every transport / kernel form, then the cfg1 weak
# bench A/B (overlapped vs chained coarse solve)
set -x
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r02_pytest_multi_n2_overlap.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest_multi_n2_overlap.log
for ov in 1 0; do
  (MFMGB_COARSE_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2969$ov bench.py --gpus 2 --steps 20 --warmup 5 --north-star off > gpurun_out/r02_bench_cfg1_weak_n2_overlap$ov.json 2> gpurun_out/r02_bench_cfg1_weak_n2_overlap$ov.err); echo "rc=$?"; tail -2 gpurun_out/r02_bench_cfg1_weak_n2_overlap$ov.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_cfg1_weak_n2_overlap$ov.json')); print('overlap=$ov', d['value'], d['ms_per_step'], d['parity']); print(d.get('timeline_in_graph_ms'))"
done
