# round 2, call C (1 GPU): full GPU suite (stencil z-sweep, adopted arrays), matrix-free probes, cfg4 bench at N=1
set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_c.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_pytest_gpu_c.log
for cfg in "constant 0" "constant 1" "constant 3" "constant 4" "discontinuous 0"; do
  set -- $cfg
  MFMGB_MF_SEGMENTS=$2 timeout 300 python tools/probe_mf.py 256 1 $1 2>&1 | tail -1 | sed "s/^/mat=$1 seg=$2 /"
done | tee gpurun_out/r02_probe_mf.txt
(time timeout 900 python bench.py --cells 256 --block 16 --matrix-free --steps 20 --warmup 5 > gpurun_out/r02_bench_mf256_n1.json 2> gpurun_out/r02_bench_mf256_n1.err); echo "rc=$?"; tail -3 gpurun_out/r02_bench_mf256_n1.err; head -c 400 gpurun_out/r02_bench_mf256_n1.json; echo
timeout 600 python bench.py --steps 20 --warmup 5 --north-star off > gpurun_out/r02_bench_n1_c.json 2> gpurun_out/r02_bench_n1_c.err; echo "rc=$?"; head -c 200 gpurun_out/r02_bench_n1_c.json; echo
