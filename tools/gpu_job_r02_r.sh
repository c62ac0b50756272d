# round 2, call R (1 GPU): node-pair stencil sweep -- matrix-free tests, apply probe (both forms), cfg4 bench line;
# host<->device copy rates; cfg2 bench + ncu of its A-kernel
set -x
timeout 900 python -m pytest tests/test_gpu_matrix_free.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_pytest_gpu_r.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_gpu_r.log
timeout 300 python tools/probe_mf.py 256 1 constant > gpurun_out/r02_probe_mf_pairs.txt 2>&1; tail -1 gpurun_out/r02_probe_mf_pairs.txt
MFMGB_MF_STENCIL_FORM=1 timeout 300 python tools/probe_mf.py 256 1 constant 2>&1 | tail -1 | tee -a gpurun_out/r02_probe_mf_pairs.txt
for sg in 4 6 10 14; do echo "segments=$sg" >> gpurun_out/r02_probe_mf_pairs.txt; MFMGB_MF_SEGMENTS=$sg timeout 300 python tools/probe_mf.py 256 1 constant 2>&1 | tail -1 | tee -a gpurun_out/r02_probe_mf_pairs.txt; done
(time timeout 900 python bench.py --cells 256 --block 16 --matrix-free --steps 20 --warmup 5 --no-cpu-baseline --north-star off > gpurun_out/r02_bench_cfg4_mf256_n1_pairs.json 2> gpurun_out/r02_bench_cfg4_mf256_n1_pairs.err); echo "rc=$?"; tail -4 gpurun_out/r02_bench_cfg4_mf256_n1_pairs.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_cfg4_mf256_n1_pairs.json')); print(d['value'], d['ms_per_step'], d['roofline'], d['parity']); print(d.get('timeline_in_graph_ms'))"
timeout 120 python tools/probe_pcie.py > gpurun_out/r02_probe_pcie.txt 2>&1; cat gpurun_out/r02_probe_pcie.txt
C2="--cells 100 --degree 2 --block 10 --material discontinuous --steps 20 --warmup 3 --no-cpu-baseline --north-star off"
(time timeout 600 python bench.py $C2 > gpurun_out/r02_bench_cfg2_n1.json 2> gpurun_out/r02_bench_cfg2_n1.err); echo "rc=$?"; tail -4 gpurun_out/r02_bench_cfg2_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_cfg2_n1.json')); print(d['value'], d['ms_per_step'], d['roofline'], d['parity']); print(d.get('timeline_in_graph_ms'))"
C2N="--cells 100 --degree 2 --block 10 --material discontinuous --steps 2 --warmup 3 --no-graph --no-cpu-baseline --parity none --north-star off --repeats 0"
timeout 600 python bench.py $C2N > gpurun_out/plain_cfg2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:csr_tile_kernel' -s 6 -c 2 -o gpurun_out/r02_prof_cfg2 python bench.py $C2N > gpurun_out/ncu_cfg2.log 2>&1; tail -2 gpurun_out/ncu_cfg2.log
