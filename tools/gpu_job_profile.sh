# ncu launch list of the V-cycle kernels (setup kernels filtered out) + full captures of the small-stage kernels
set -x
K='regex:csr_tile_kernel|csr_vec_kernel|gemv_kernel|zero_guess_kernel'
timeout 300 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 120 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:csr_vec_kernel|gemv_kernel|zero_guess_kernel' -s 8 -c 4 -o gpurun_out/prof_small python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu3.log 2>&1
tail -3 gpurun_out/ncu3.log
# cfg2: 3D Q2 discontinuous diffusion, 100^3 cells (8.1 M DoFs)
(time timeout 900 python bench.py --cells 100 --degree 2 --block 10 --material discontinuous --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg2.json) 2> gpurun_out/bench_cfg2.err
tail -4 gpurun_out/bench_cfg2.err
