# round 2, call Z (1 GPU): the default line without the 512^3 leg (main leg + cpu baseline + other_configs children)
set -x
(time timeout 200 python bench.py --steps 20 --warmup 5 --north-star off > gpurun_out/r02_bench_default_n1_head_no_ns.json 2> gpurun_out/r02_bench_default_n1_head_no_ns.err); echo "rc=$?"; tail -3 gpurun_out/r02_bench_default_n1_head_no_ns.err; head -c 400 gpurun_out/r02_bench_default_n1_head_no_ns.json
