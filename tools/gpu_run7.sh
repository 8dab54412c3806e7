#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest tc"; timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=6 -p no:cacheprovider -x > gpurun_out/pytest_tc7.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_tc7.log | cut -c1-250
echo "=== microbench tc"; timeout 300 python tools/microbench.py --phases --shapes qkvo,gateup,down,head --m 1,8 2>&1 | grep -v '"launch": [02-9]' | grep -v '"launch": 1[01]' | cut -c1-420
