# round 2, call Q (1 GPU): host<->device copy rates (bound of the e2e leg); cfg2 (Q2) bench line and ncu --set full of its A-kernel
set -x
timeout 120 python tools/probe_pcie.py > gpurun_out/r02_probe_pcie.txt 2>&1; cat gpurun_out/r02_probe_pcie.txt
C2="--cells 100 --degree 2 --block 10 --material discontinuous --steps 20 --warmup 3 --no-cpu-baseline --north-star off"
(time timeout 600 python bench.py $C2 > gpurun_out/r02_bench_cfg2_n1.json 2> gpurun_out/r02_bench_cfg2_n1.err); echo "rc=$?"; tail -4 gpurun_out/r02_bench_cfg2_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_cfg2_n1.json')); print(d['value'], d['ms_per_step'], d['roofline'], d['parity']); print(d.get('timeline_in_graph_ms'))"
C2N="--cells 100 --degree 2 --block 10 --material discontinuous --steps 2 --warmup 3 --no-graph --no-cpu-baseline --parity none --north-star off --repeats 0"
timeout 600 python bench.py $C2N > gpurun_out/plain_cfg2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:csr_tile_kernel' -s 6 -c 2 -o gpurun_out/r02_prof_cfg2 python bench.py $C2N > gpurun_out/ncu_cfg2.log 2>&1; tail -2 gpurun_out/ncu_cfg2.log
