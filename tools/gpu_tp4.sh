#!/bin/bash
# GPU box helper (4 GPUs): tensor-parallel parity check + N=4 bench line
cd /root/repo
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29655 tools/tp_check.py test-small 2>&1 | grep "step\|TP_CHECK\|rror" | tail -5 | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 64 --warmup 8 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err; echo rc=$?; cut -c1-300 gpurun_out/r02_bench_n4.json; tail -2 gpurun_out/r02_bench_n4.err | cut -c1-200
