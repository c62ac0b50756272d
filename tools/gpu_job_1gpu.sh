set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python tools/probe_tile.py 48 2 quick > gpurun_out/probe_tile_q2.json 2> gpurun_out/probe_tile_q2.err; tail -4 gpurun_out/probe_tile_q2.err
(time timeout 900 python bench.py --cells 100 --degree 2 --block 10 --material discontinuous --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg2.json) 2> gpurun_out/bench_cfg2.err
tail -4 gpurun_out/bench_cfg2.err
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err; tail -2 gpurun_out/bench_cfg1.err
