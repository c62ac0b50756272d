set -x
# driver-style weak-scaling bench at 8 and 4 GPUs (cfg1 per GPU), then the cfg3 strong-scaling point with the DD coarse solve
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/bench_w8.json 2> gpurun_out/bench_w8.err
echo "rc=$?"; head -c 220 gpurun_out/bench_w8.json; echo; tail -2 gpurun_out/bench_w8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 4 --steps 30 --warmup 5 > gpurun_out/bench_w4.json 2> gpurun_out/bench_w4.err
echo "rc=$?"; head -c 220 gpurun_out/bench_w4.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29603 bench.py --gpus 8 --cells 512 --block 16 --scaling strong --steps 20 --warmup 3 > gpurun_out/bench_cfg3_n8_dd.json 2> gpurun_out/bench_cfg3_n8_dd.err
echo "rc=$?"; head -c 220 gpurun_out/bench_cfg3_n8_dd.json; echo; tail -2 gpurun_out/bench_cfg3_n8_dd.err
