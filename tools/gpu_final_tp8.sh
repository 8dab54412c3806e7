#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4; do
echo "=== bench N=$n"; timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2962$n bench.py --gpus $n --steps 32 --warmup 3 > gpurun_out/bench_tp$n.log 2>&1; echo "rc=$?"; grep '"metric"' gpurun_out/bench_tp$n.log | cut -c1-330; grep -iE "error|Traceback" gpurun_out/bench_tp$n.log | head -3
done
