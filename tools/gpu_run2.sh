#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short --maxfail=25 -p no:cacheprovider -x > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu2.log
echo "=== microbench phases (1 CTA/SM)"; timeout 600 python tools/microbench.py --phases --shapes qkvo,gateup,head --m 1,8 > gpurun_out/mb2_c1.log 2>&1; echo "rc=$?"; cat gpurun_out/mb2_c1.log | cut -c1-400
echo "=== microbench (2 CTA/SM)"; timeout 600 python tools/microbench.py --ctas-per-sm 2 --phases --shapes qkvo,gateup --m 1 > gpurun_out/mb2_c2.log 2>&1; echo "rc=$?"; cat gpurun_out/mb2_c2.log | cut -c1-400
echo "=== bench tiny"; timeout 900 python bench.py --model tinyllama-1.1b-4.0bpw --steps 32 --warmup 4 > gpurun_out/bench_tiny.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/bench_tiny.log | cut -c1-1500
echo "=== bench 7b"; timeout 1200 python bench.py --steps 64 --warmup 4 > gpurun_out/bench_7b.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/bench_7b.log | cut -c1-2500
