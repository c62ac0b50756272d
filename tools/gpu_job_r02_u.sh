# round 2, call U (1 GPU): strip sweep with deeper rings -- tests, probe, ncu --set full, cfg4 bench line with oracle parity
set -x
timeout 900 python -m pytest tests/test_gpu_matrix_free.py -m gpu -q -x > gpurun_out/r02_pytest_gpu_u.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_gpu_u.log
: > gpurun_out/r02_probe_mf_strips2.txt
for sg in 0 8 16; do echo "form=3 deep rings segments=$sg (0 = default)" >> gpurun_out/r02_probe_mf_strips2.txt; MFMGB_MF_SEGMENTS=$sg timeout 300 python tools/probe_mf.py 256 1 constant 2>&1 | tail -1 | tee -a gpurun_out/r02_probe_mf_strips2.txt; done
timeout 300 python tools/probe_mf.py 256 1 constant > gpurun_out/r02_plain_mf3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mf_q1_stencil3_kernel -s 3 -c 1 -o gpurun_out/r02_prof_stencil_strips python tools/probe_mf.py 256 1 constant > gpurun_out/r02_ncu_stencil3.log 2>&1
tail -2 gpurun_out/r02_ncu_stencil3.log
(time timeout 900 python bench.py --cells 256 --block 16 --matrix-free --steps 20 --warmup 5 --no-cpu-baseline --north-star off > gpurun_out/r02_bench_cfg4_mf256_n1_strips.json 2> gpurun_out/r02_bench_cfg4_mf256_n1_strips.err); echo "rc=$?"; tail -4 gpurun_out/r02_bench_cfg4_mf256_n1_strips.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_cfg4_mf256_n1_strips.json')); print(d['value'], d['ms_per_step'], d['roofline'], d['parity']); print(d.get('timeline_in_graph_ms'))"
