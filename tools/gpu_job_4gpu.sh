set -x
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "4 and auto" 2>&1 | tail -6
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 4 --steps 30 --warmup 5 > gpurun_out/bench_w4.json 2> gpurun_out/bench_w4.err
echo "rc=$?"; head -c 220 gpurun_out/bench_w4.json; echo
