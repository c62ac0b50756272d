set -x
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 100 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "rc=$?"; head -c 150 gpurun_out/bench_final.json
