# round 2, call H (1 GPU): full GPU suite, stencil-kernel sweeps (running pointers, arithmetic flags), cfg4 bench
set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_h.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02_pytest_gpu_h.log
for cfg in "3 0 1" "3 3 1" "3 6 1" "3 9 1" "3 12 1" "2 6 1" "2 9 1" "3 6 0"; do
  set -- $cfg
  MFMGB_MF_MINB=$1 MFMGB_MF_SEGMENTS=$2 MFMGB_MF_ARITH_FLAGS=$3 timeout 300 python tools/probe_mf.py 256 1 constant 2>&1 | tail -1 | sed "s/^/minb=$1 seg=$2 arith=$3 /"
done | tee gpurun_out/r02_probe_mf_h.txt
(time timeout 600 python bench.py --cells 256 --block 16 --matrix-free --steps 20 --warmup 5 > gpurun_out/r02_bench_mf256_n1.json 2> gpurun_out/r02_bench_mf256_n1.err); echo "rc=$?"; tail -3 gpurun_out/r02_bench_mf256_n1.err; head -c 300 gpurun_out/r02_bench_mf256_n1.json; echo
