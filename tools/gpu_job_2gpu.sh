set -x
timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "2 and auto" 2>&1 | tail -6
for nh in 1 0; do
MFMGB_RESTRICT_NO_HALO=$nh timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2959$nh bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_w2_nh$nh.json 2> gpurun_out/bench_w2_nh$nh.err
echo "rc=$?"; head -c 200 gpurun_out/bench_w2_nh$nh.json; echo
done
