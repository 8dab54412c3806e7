#!/bin/bash
# GPU box helper (2 GPUs): tensor-parallel parity test + N=2 bench line
cd /root/repo
timeout 300 python -m pytest tests/test_gpu_tp.py -m gpu -q -x 2>&1 | tail -2 | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 64 --warmup 8 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo rc=$?; cut -c1-400 gpurun_out/r02_bench_n2.json
