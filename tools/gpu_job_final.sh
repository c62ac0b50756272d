set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"; tail -2 gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "rc=$?"; head -c 300 gpurun_out/bench_reference.json
(time timeout 900 python bench.py --cells 100 --degree 2 --block 10 --material discontinuous --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg2.json) 2> gpurun_out/bench_cfg2.err; tail -4 gpurun_out/bench_cfg2.err
