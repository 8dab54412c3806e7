#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
echo "=== tp_check"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/tp_check.py test-small 2>&1 | grep -v "^W\|OMP_NUM" | tail -12
echo "=== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 32 --warmup 3 > gpurun_out/bench_tp2.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench_tp2.log | cut -c1-1200
