#!/bin/bash
# GPU box helper (2 GPUs): tensor-parallel parity check, repeated
cd /root/repo
for V in 1 2 3 4 5 6; do
env A=$V timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tools/tp_check.py test-small 2>&1 | grep "step 3\|TP_CHECK\|rror" | tail -2 | tr '\n' ' ' | cut -c1-200; echo
done
timeout 300 python -m pytest tests/test_gpu_tp.py -m gpu -q -x 2>&1 | tail -2 | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 64 --warmup 8 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo rc=$?; cut -c1-300 gpurun_out/r02_bench_n2.json
