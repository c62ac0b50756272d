# round 2, call N (1 GPU): final default bench line, ncu launch list of the cycle, ncu --set full of the A-kernel and of
# the matrix-free stencil sweep at HEAD
set -x
(time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_final_n1.json 2> gpurun_out/r02_bench_final_n1.err); echo "bench rc=$?"; head -c 300 gpurun_out/r02_bench_final_n1.json; echo
K='regex:csr_tile_kernel|csr_vec_kernel|gemv_stream_kernel|zero_guess_kernel'
timeout 300 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --parity none --north-star off --repeats 0 > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 120 --csv --log-file gpurun_out/r02_launches_vcycle.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --parity none --north-star off --repeats 0 > gpurun_out/ncu1.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --parity none --north-star off --repeats 0 > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:csr_tile_kernel|gemv_stream_kernel' -s 8 -c 4 -o gpurun_out/r02_prof_cycle python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --parity none --north-star off --repeats 0 > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
timeout 300 python tools/probe_mf.py 256 1 constant > gpurun_out/r02_plain_mf.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mf_q1_stencil_kernel -s 3 -c 1 -o gpurun_out/r02_prof_stencil_head python tools/probe_mf.py 256 1 constant > gpurun_out/r02_ncu_stencil.log 2>&1
tail -2 gpurun_out/r02_ncu_stencil.log; cat gpurun_out/r02_plain_mf.log | tail -1
