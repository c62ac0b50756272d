# round 2, call J1 (1 GPU): GPU suite, N=1 bench variants (streamed GEMV on/off, tile kernel for R and P)
set -x
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_j.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_pytest_gpu_j.log
timeout 300 python bench.py --steps 30 --warmup 5 --north-star off > gpurun_out/r02_bench_n1_j.json 2> gpurun_out/r02_bench_n1_j.err; echo "n1 rc=$?"; head -c 200 gpurun_out/r02_bench_n1_j.json; echo
MFMGB_TILE_MIN_ROWS=1024 MFMGB_TILE_MIN_ROW_NNZ=1 timeout 300 python bench.py --steps 30 --warmup 5 --north-star off --no-cpu-baseline > gpurun_out/r02_bench_n1_j_tileRP.json 2> gpurun_out/r02_bench_n1_j_tileRP.err; echo "n1 tileRP rc=$?"; head -c 200 gpurun_out/r02_bench_n1_j_tileRP.json; echo
MFMGB_TILE_MIN_ROWS=1024 timeout 300 python bench.py --steps 30 --warmup 5 --north-star off --no-cpu-baseline > gpurun_out/r02_bench_n1_j_tileR.json 2> gpurun_out/r02_bench_n1_j_tileR.err; echo "n1 tileR rc=$?"; head -c 200 gpurun_out/r02_bench_n1_j_tileR.json; echo
MFMGB_GEMV_STREAM=0 timeout 300 python bench.py --steps 30 --warmup 5 --north-star off --no-cpu-baseline > gpurun_out/r02_bench_n1_j_oldgemv.json 2> gpurun_out/r02_bench_n1_j_oldgemv.err; echo "n1 oldgemv rc=$?"; head -c 200 gpurun_out/r02_bench_n1_j_oldgemv.json; echo
