#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=line --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu3.log | cut -c1-300
echo "=== microbench phases (2 CTA/SM)"; timeout 600 python tools/microbench.py --phases --shapes qkvo,gateup,down43,head --m 1 > gpurun_out/mb3_c2.log 2>&1; echo "rc=$?"; cat gpurun_out/mb3_c2.log | cut -c1-400
echo "=== microbench (1 CTA/SM)"; timeout 600 python tools/microbench.py --ctas-per-sm 1 --phases --shapes qkvo,gateup --m 1,8 > gpurun_out/mb3_c1.log 2>&1; echo "rc=$?"; cat gpurun_out/mb3_c1.log | cut -c1-400
echo "=== bench 7b"; timeout 1200 python bench.py --steps 64 --warmup 4 --no-cpu > gpurun_out/bench_7b_3.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench_7b_3.log | cut -c1-1200
