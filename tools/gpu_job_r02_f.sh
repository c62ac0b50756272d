# round 2, call F (1 GPU): Chebyshev gold-config probe, ncu --set full of the stencil z-sweep kernel
set -x
timeout 300 python tools/debug_cheb.py > gpurun_out/r02_debug_cheb.txt 2>&1; tail -50 gpurun_out/r02_debug_cheb.txt
export MFMGB_MF_SEGMENTS=3
timeout 300 python tools/probe_mf.py 256 1 constant > gpurun_out/r02_plain_mf.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mf_q1_stencil_kernel -s 3 -c 2 -o gpurun_out/r02_prof_stencil python tools/probe_mf.py 256 1 constant > gpurun_out/r02_ncu_stencil.log 2>&1
tail -3 gpurun_out/r02_ncu_stencil.log
