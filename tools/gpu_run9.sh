#!/bin/bash
echo "=== pytest tc (linear)"; timeout 900 python -m pytest tests/test_gpu_linear.py tests/test_gpu_ops.py -m gpu -q --tb=short --maxfail=4 -p no:cacheprovider -x 2>&1 | tail -3 | cut -c1-250
echo "=== microbench tc 1 CTA/SM"; timeout 300 python tools/microbench.py --phases --shapes qkvo,gateup,down,head --m 1 2>&1 | grep -v '"launch": [02-9]' | grep -v '"launch": 1[01]' | grep -v '"cta": 100' | cut -c1-640
echo "=== microbench tc 2 CTA/SM"; timeout 300 python tools/microbench.py --ctas-per-sm 2 --phases --shapes qkvo,gateup,down,head --m 1,8 2>&1 | grep -v '"launch": [02-9]' | grep -v '"launch": 1[01]' | grep -v '"cta": 100' | cut -c1-640
