# round 2, call O (2 GPUs): parity + bench N=2 with the stacked restriction
set -x
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02_pytest_multi_n2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest_multi_n2.log
for st in 1 0; do
MFMGB_RESTRICT_STACK=$st timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2967$st bench.py --gpus 2 --steps 30 --warmup 5 --north-star off > gpurun_out/r02_bench_n2_stack$st.json 2> gpurun_out/r02_bench_n2_stack$st.err
echo "bench stack=$st rc=$?"; tail -2 gpurun_out/r02_bench_n2_stack$st.err; head -c 200 gpurun_out/r02_bench_n2_stack$st.json; echo
done
