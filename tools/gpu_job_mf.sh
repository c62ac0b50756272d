set -x
timeout 900 python -m pytest tests/test_gpu_matrix_free.py -m gpu -x -q 2>&1 | tail -3
for cfg in "constant 0" "constant 12" "linear 0"; do
  set -- $cfg
  MFMGB_MF_TZ=$2 timeout 300 python tools/probe_mf.py 256 1 $1 2>&1 | tail -1 | sed "s/^/mat=$1 tz=$2 /"
done
timeout 600 python bench.py --cells 256 --block 16 --matrix-free --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mf256.json 2> gpurun_out/bench_mf256.err
tail -2 gpurun_out/bench_mf256.err
