# round 2, call E (1 GPU): full GPU suite (Chebyshev, stencil ring), stencil-kernel sweep, cfg4 + default bench
set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_e.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_pytest_gpu_e.log
for cfg in "3 0" "3 2" "3 3" "3 4" "3 6" "2 0" "2 3" "2 4"; do
  set -- $cfg
  MFMGB_MF_MINB=$1 MFMGB_MF_SEGMENTS=$2 timeout 300 python tools/probe_mf.py 256 1 constant 2>&1 | tail -1 | sed "s/^/minb=$1 seg=$2 /"
done | tee gpurun_out/r02_probe_mf_e.txt
(time timeout 900 python bench.py --cells 256 --block 16 --matrix-free --steps 20 --warmup 5 > gpurun_out/r02_bench_mf256_n1.json 2> gpurun_out/r02_bench_mf256_n1.err); echo "rc=$?"; tail -3 gpurun_out/r02_bench_mf256_n1.err; head -c 400 gpurun_out/r02_bench_mf256_n1.json; echo
timeout 600 python bench.py --steps 20 --warmup 5 --north-star off > gpurun_out/r02_bench_n1_e.json 2> gpurun_out/r02_bench_n1_e.err; echo "rc=$?"; tail -3 gpurun_out/r02_bench_n1_e.err; head -c 200 gpurun_out/r02_bench_n1_e.json; echo
