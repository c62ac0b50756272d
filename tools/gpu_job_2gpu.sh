set -x
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "2" 2>&1 | tail -15
for g in 1 0; do
  MFMGB_DIST_GRAPH=$g timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$g bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_w2_g$g.json 2> gpurun_out/bench_w2_g$g.err
  echo "rc=$?"; head -c 260 gpurun_out/bench_w2_g$g.json; echo
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --cells 256 --block 16 --matrix-free --steps 20 --warmup 3 > gpurun_out/bench_mf256_n2.json 2> gpurun_out/bench_mf256_n2.err
echo "rc=$?"; head -c 260 gpurun_out/bench_mf256_n2.json
