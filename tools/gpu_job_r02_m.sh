# round 2, call M (1 GPU): full GPU suite at HEAD, then compute-sanitizer memcheck of smoke() (one tool per call)
set -x
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_head.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_pytest_gpu_head.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_memcheck_smoke.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/r02_sanitizer_memcheck_smoke.log
