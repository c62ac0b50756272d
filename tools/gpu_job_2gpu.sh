set -x
for ov in 1 0; do
  MFMGB_HALO_OVERLAP=$ov timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2957$ov bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_w2_ov$ov.json 2> gpurun_out/bench_w2_ov$ov.err
  echo "rc=$?"; head -c 200 gpurun_out/bench_w2_ov$ov.json; echo
done
