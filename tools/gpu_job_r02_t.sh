# round 2, call T (1 GPU): strip form of the stencil sweep -- matrix-free tests + parity suite (new gather unroll for long
# rows included), apply probe for the three forms, cfg4 and cfg2 bench lines
set -x
timeout 900 python -m pytest tests/test_gpu_matrix_free.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_pytest_gpu_t.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_gpu_t.log
: > gpurun_out/r02_probe_mf_strips.txt
for form in 3 2 1; do echo "form=$form" >> gpurun_out/r02_probe_mf_strips.txt; MFMGB_MF_STENCIL_FORM=$form timeout 300 python tools/probe_mf.py 256 1 constant 2>&1 | tail -1 | tee -a gpurun_out/r02_probe_mf_strips.txt; done
for sg in 3 5 8 12; do echo "form=3 segments=$sg" >> gpurun_out/r02_probe_mf_strips.txt; MFMGB_MF_SEGMENTS=$sg timeout 300 python tools/probe_mf.py 256 1 constant 2>&1 | tail -1 | tee -a gpurun_out/r02_probe_mf_strips.txt; done
(time timeout 900 python bench.py --cells 256 --block 16 --matrix-free --steps 20 --warmup 5 --no-cpu-baseline --north-star off --parity props > gpurun_out/r02_bench_cfg4_mf256_n1_strips.json 2> gpurun_out/r02_bench_cfg4_mf256_n1_strips.err); echo "rc=$?"; tail -4 gpurun_out/r02_bench_cfg4_mf256_n1_strips.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_cfg4_mf256_n1_strips.json')); print(d['value'], d['ms_per_step'], d['roofline'], d['parity']); print(d.get('timeline_in_graph_ms'))"
C2="--cells 100 --degree 2 --block 10 --material discontinuous --steps 20 --warmup 3 --no-cpu-baseline --north-star off --parity props"
(time timeout 600 python bench.py $C2 > gpurun_out/r02_bench_cfg2_n1_unroll16.json 2> gpurun_out/r02_bench_cfg2_n1_unroll16.err); echo "rc=$?"; tail -4 gpurun_out/r02_bench_cfg2_n1_unroll16.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_cfg2_n1_unroll16.json')); print(d['value'], d['ms_per_step'], d['roofline'], d['parity']); print(d.get('timeline_in_graph_ms'))"
