# round 2, call J (2 GPUs): GPU suite on rank-0 GPU (streamed GEMV etc.), multi-GPU parity, bench N=2 fused / unfused
set -x
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_j.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_pytest_gpu_j.log
for cfg in "1" "0"; do
MFMGB_HALO_FUSED=$cfg timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2966$cfg bench.py --gpus 2 --steps 30 --warmup 5 --north-star off > gpurun_out/r02_bench_n2_j$cfg.json 2> gpurun_out/r02_bench_n2_j$cfg.err
echo "bench fused=$cfg rc=$?"; tail -2 gpurun_out/r02_bench_n2_j$cfg.err; head -c 200 gpurun_out/r02_bench_n2_j$cfg.json; echo
done
timeout 300 python bench.py --steps 30 --warmup 5 --north-star off > gpurun_out/r02_bench_n1_j.json 2> gpurun_out/r02_bench_n1_j.err; echo "n1 rc=$?"; head -c 200 gpurun_out/r02_bench_n1_j.json; echo
MFMGB_TILE_MIN_ROWS=1024 MFMGB_TILE_MIN_ROW_NNZ=1 timeout 300 python bench.py --steps 30 --warmup 5 --north-star off --no-cpu-baseline > gpurun_out/r02_bench_n1_j_tileRP.json 2> gpurun_out/r02_bench_n1_j_tileRP.err; echo "n1 tileRP rc=$?"; head -c 200 gpurun_out/r02_bench_n1_j_tileRP.json; echo
MFMGB_GEMV_STREAM=0 timeout 300 python bench.py --steps 30 --warmup 5 --north-star off --no-cpu-baseline > gpurun_out/r02_bench_n1_j_oldgemv.json 2> gpurun_out/r02_bench_n1_j_oldgemv.err; echo "n1 oldgemv rc=$?"; head -c 200 gpurun_out/r02_bench_n1_j_oldgemv.json; echo
