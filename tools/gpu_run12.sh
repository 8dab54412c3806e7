#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest tc"; timeout 900 python -m pytest tests/test_gpu_linear.py tests/test_gpu_ops.py tests/test_gpu_golden.py -m gpu -q --tb=short --maxfail=4 -p no:cacheprovider -x 2>&1 | tail -4 | cut -c1-250
echo "=== microbench tc 2 CTA/SM"; timeout 300 python tools/microbench.py --phases --shapes qkvo,gateup,down,head --m 1,8 2>&1 | grep -v '"launch": [013-9]' | grep -v '"launch": 1[01]' | grep -v '"cta": 100' | cut -c1-640
echo "=== bench 7b"; timeout 900 python bench.py --steps 64 --warmup 4 --no-cpu > gpurun_out/bench_7b_12.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_7b_12.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ['value','ms_per_step','launches_per_step']}, d['e2e'], {k:d['roofline'][k] for k in ['achieved','frac','avg_launch_us']})"
