# round 2, call D (2 GPUs): multi-GPU parity at HEAD + bench N=2 (in-graph timeline) with exchange overlapped / not
set -x
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02_pytest_multi_n2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest_multi_n2.log
for ov in 1 0; do
MFMGB_HALO_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2962$ov bench.py --gpus 2 --steps 30 --warmup 5 --north-star off > gpurun_out/r02_bench_n2_ov$ov.json 2> gpurun_out/r02_bench_n2_ov$ov.err
echo "bench overlap=$ov rc=$?"; tail -2 gpurun_out/r02_bench_n2_ov$ov.err; head -c 200 gpurun_out/r02_bench_n2_ov$ov.json; echo
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --cells 256 --block 16 --matrix-free --steps 20 --warmup 5 --north-star off > gpurun_out/r02_bench_mf256_n2.json 2> gpurun_out/r02_bench_mf256_n2.err
echo "bench mf rc=$?"; tail -2 gpurun_out/r02_bench_mf256_n2.err; head -c 200 gpurun_out/r02_bench_mf256_n2.json; echo
