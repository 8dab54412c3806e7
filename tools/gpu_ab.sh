#!/bin/bash
# GPU box helper: final validation of the committed state (full GPU test tier, smoke, context runs, default bench)
cd /root/repo; T=r02
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3 | cut -c1-200
echo "=== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
for CTX in 1024 4096 16384; do
  echo "=== decode at context $CTX"; timeout 400 python bench.py --context $CTX --steps 32 --no-cpu --no-ref-ext > gpurun_out/${T}_bench_ctx$CTX.json 2> gpurun_out/${T}_bench_ctx$CTX.err; cut -c1-230 gpurun_out/${T}_bench_ctx$CTX.json; tail -2 gpurun_out/${T}_bench_ctx$CTX.err | cut -c1-200
done
echo "=== default bench"; timeout 600 python bench.py > gpurun_out/${T}_bench_n1_final.json 2> gpurun_out/${T}_bench_n1_final.err; cut -c1-230 gpurun_out/${T}_bench_n1_final.json; tail -2 gpurun_out/${T}_bench_n1_final.err
echo "=== rows microbench 4096x4096"; timeout 300 python tools/microbench.py --ref --shapes qkvo --m 8,16 2>&1 | cut -c1-330 | tail -2
