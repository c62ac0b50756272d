# round 2, last call (1 GPU): smoke() with the NVTX-instrumented library
timeout 40 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
