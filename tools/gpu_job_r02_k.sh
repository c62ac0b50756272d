# round 2, call K (8 GPUs): parity worker at 8 ranks (fused compute+exchange path), bench N=8 (weak cfg1), then the full
# default line incl. the 512^3 strong-scaling north-star leg
set -x
nvidia-smi topo -m > gpurun_out/r02_topo_n8.txt 2>&1; free -g >> gpurun_out/r02_topo_n8.txt; nproc >> gpurun_out/r02_topo_n8.txt
MFMGB_CSR_KERNEL=tile timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29681 tests/dist_gpu_worker.py > gpurun_out/r02_multi_gpu_parity_w8_tile_peer.log 2>&1; echo "worker8 rc=$?"; grep -c "RANK .* OK" gpurun_out/r02_multi_gpu_parity_w8_tile_peer.log; tail -3 gpurun_out/r02_multi_gpu_parity_w8_tile_peer.log
(time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29682 bench.py --gpus 8 --steps 20 --warmup 5 --north-star off > gpurun_out/r02_bench_n8_main.json 2> gpurun_out/r02_bench_n8_main.err); echo "bench8 main rc=$?"; tail -2 gpurun_out/r02_bench_n8_main.err; head -c 300 gpurun_out/r02_bench_n8_main.json; echo
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29683 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err); echo "bench8 rc=$?"; tail -2 gpurun_out/r02_bench_n8.err; head -c 300 gpurun_out/r02_bench_n8.json; echo
