set -x
# cfg4: matrix-free Q1 fine level, 256^3 cells per GPU, weak scaling at 8 GPUs
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --cells 256 --block 16 --matrix-free --steps 20 --warmup 3 > gpurun_out/bench_mf256_n8.json 2> gpurun_out/bench_mf256_n8.err
echo "rc=$?"; head -c 220 gpurun_out/bench_mf256_n8.json; echo; tail -2 gpurun_out/bench_mf256_n8.err
