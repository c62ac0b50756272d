# round 2, call M (1 GPU): compute-sanitizer memcheck of smoke() (one tool per call)
set -x
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_memcheck_smoke.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/r02_sanitizer_memcheck_smoke.log
