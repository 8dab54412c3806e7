#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=line --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu5.log | cut -c1-300
echo "=== microbench 2 CTA/SM"; timeout 600 python tools/microbench.py --ref --shapes qkvo,gateup,down,down43,head,gptq --m 1,8 2>&1 | grep -v launch | cut -c1-420
echo "=== timeline qkvo"; timeout 300 python tools/microbench.py --phases --shapes qkvo --m 1 2>&1 | grep '"cta": 0' | head -8 | cut -c1-330
echo "=== microbench 1 CTA/SM"; timeout 600 python tools/microbench.py --ctas-per-sm 1 --shapes qkvo,gateup,down --m 1 2>&1 | cut -c1-420
echo "=== bench 7b"; timeout 1200 python bench.py --steps 64 --warmup 4 --no-cpu > gpurun_out/bench_7b_5.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench_7b_5.log | cut -c1-400
