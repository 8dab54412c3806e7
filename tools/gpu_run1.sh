#!/bin/bash
# first GPU session: smoke, parity tests, golden vectors from the reference ext, micro-benchmark
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "=== smoke"; timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short --maxfail=25 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
echo "=== golden"; timeout 600 python oracle/gen_golden.py gpurun_out/golden > gpurun_out/golden.log 2>&1; echo "golden rc=$?"; tail -5 gpurun_out/golden.log
echo "=== microbench"; timeout 900 python tools/microbench.py --ref --shapes qkvo,gateup,down,head,qkvo54,down43,b3,gptq --m 1,8 --json gpurun_out/microbench1.json > gpurun_out/microbench1.log 2>&1; echo "mb rc=$?"; tail -30 gpurun_out/microbench1.log
