# round 2, call V (1 GPU): z-segment count of the strip sweep inside the cfg4 cycle (fused residual / Jacobi forms)
set -x
for sg in 0 6 12; do
  MFMGB_MF_SEGMENTS=$sg timeout 600 python bench.py --cells 256 --block 16 --matrix-free --steps 20 --warmup 5 --no-cpu-baseline --north-star off --parity none --repeats 3 > gpurun_out/r02_cfg4_seg$sg.json 2> gpurun_out/r02_cfg4_seg$sg.err; echo "rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r02_cfg4_seg$sg.json')); print('segments=$sg', d['value'], d['ms_per_step']); print(d.get('timeline_in_graph_ms'))"
done
