set -x
timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "2" 2>&1 | tail -15
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_w2_new.json 2> gpurun_out/bench_w2_new.err
echo "rc=$?"; head -c 200 gpurun_out/bench_w2_new.json; echo
