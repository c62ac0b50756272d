# round 2, call W (1 GPU): short-row kernel for P (4 rows per thread in flight) -- parity suite, cfg1 and cfg4 lines
set -x
timeout 900 python -m pytest tests/test_gpu_matrix_free.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_pytest_gpu_w.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_gpu_w.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --north-star off --other-configs off > gpurun_out/r02_cfg1_w.json 2> gpurun_out/r02_cfg1_w.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_cfg1_w.json')); print('cfg1', d['value'], d['ms_per_step'], d['timing']['ms_per_step_median'], d['parity']['ok'], d['e2e']['value'], d['e2e']['ms_per_step']); print(d.get('timeline_in_graph_ms'))"
timeout 600 python bench.py --cells 256 --block 16 --matrix-free --steps 20 --warmup 5 --no-cpu-baseline --north-star off --parity props --repeats 3 > gpurun_out/r02_cfg4_w.json 2> gpurun_out/r02_cfg4_w.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_cfg4_w.json')); print('cfg4', d['value'], d['ms_per_step'], d['parity']['ok']); print(d.get('timeline_in_graph_ms'))"
