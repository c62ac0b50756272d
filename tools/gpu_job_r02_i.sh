# round 2, call I (2 GPUs): parity with the fused compute+exchange A-kernel, bench N=2 in its three forms
set -x
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02_pytest_multi_n2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest_multi_n2.log
for cfg in "1 1" "1 0" "0 0"; do
set -- $cfg
MFMGB_HALO_FUSED=$1 MFMGB_HALO_PUSH_IN_KERNEL=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2965$1 bench.py --gpus 2 --steps 30 --warmup 5 --north-star off > gpurun_out/r02_bench_n2_fused$1$2.json 2> gpurun_out/r02_bench_n2_fused$1$2.err
echo "bench fused=$1 push_in_kernel=$2 rc=$?"; tail -2 gpurun_out/r02_bench_n2_fused$1$2.err; head -c 200 gpurun_out/r02_bench_n2_fused$1$2.json; echo
done
