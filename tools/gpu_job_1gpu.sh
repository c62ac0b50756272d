set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --pcg > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"; tail -2 gpurun_out/bench_default.err
