# round 2, call S (1 GPU): ncu --set full of the node-pair stencil sweep (apply form)
set -x
timeout 300 python tools/probe_mf.py 256 1 constant > gpurun_out/r02_plain_mf2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mf_q1_stencil2_kernel -s 3 -c 1 -o gpurun_out/r02_prof_stencil_pairs python tools/probe_mf.py 256 1 constant > gpurun_out/r02_ncu_stencil2.log 2>&1
tail -2 gpurun_out/r02_ncu_stencil2.log; tail -1 gpurun_out/r02_plain_mf2.log
