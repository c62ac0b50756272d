# round 2, call B (2 GPUs): multi-GPU parity (peer-memory and NCCL transports) + bench N=2 with oracle parity
set -x
nvidia-smi topo -m > gpurun_out/r02_topo_n2.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02_pytest_multi_n2.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02_pytest_multi_n2.log
for peer in 1 0; do
MFMGB_PEER=$peer timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2961$peer bench.py --gpus 2 --steps 30 --warmup 5 --north-star off > gpurun_out/r02_bench_n2_peer$peer.json 2> gpurun_out/r02_bench_n2_peer$peer.err
echo "bench peer=$peer rc=$?"; tail -3 gpurun_out/r02_bench_n2_peer$peer.err; head -c 300 gpurun_out/r02_bench_n2_peer$peer.json; echo
done
