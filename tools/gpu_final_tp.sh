#!/bin/bash
# 2-GPU check of the column-sharded path; every command has its own short timeout
mkdir -p gpurun_out
echo "=== tp_check"; timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/tp_check.py test-small 2>&1 | grep -E "step|TP_CHECK|rror|failed" | head -12
echo "=== bench N=2"; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 32 --warmup 3 > gpurun_out/bench_tp2.log 2>&1; echo "rc=$?"; grep '"metric"' gpurun_out/bench_tp2.log | cut -c1-700; grep -iE "error|Traceback" gpurun_out/bench_tp2.log | head -5
