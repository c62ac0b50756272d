# round 2, call X (2 GPUs): parity worker on every transport / kernel form at HEAD, then the launch-parameter probe of
# the fused compute + exchange kernel
set -x
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r02_pytest_multi_n2_head.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_multi_n2_head.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29671 tools/probe_weak.py > gpurun_out/r02_probe_weak_n2.txt 2> gpurun_out/r02_probe_weak_n2.err; echo "rc=$?"; tail -3 gpurun_out/r02_probe_weak_n2.err; cat gpurun_out/r02_probe_weak_n2.txt
