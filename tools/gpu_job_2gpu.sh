set -x
timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "2" 2>&1 | tail -6
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --cells 256 --block 16 --matrix-free --steps 20 --warmup 3 > gpurun_out/bench_mf256_n2.json 2> gpurun_out/bench_mf256_n2.err
echo "rc=$?"; head -c 200 gpurun_out/bench_mf256_n2.json; echo
