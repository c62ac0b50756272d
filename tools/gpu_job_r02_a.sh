# round 2, call A (1 GPU): GPU test suite, default bench (with parity + north-star leg), reference arm
set -x
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r02_machine.txt; free -g >> gpurun_out/r02_machine.txt; nproc >> gpurun_out/r02_machine.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_n1.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest_gpu_n1.log
(time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err); echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n1.err; head -c 600 gpurun_out/r02_bench_n1.json; echo
(time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref_n1.json 2> gpurun_out/r02_bench_ref_n1.err); echo "ref rc=$?"; head -c 300 gpurun_out/r02_bench_ref_n1.json; echo
