#!/bin/bash
mkdir -p gpurun_out
echo "=== smoke (tc layout)"; timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -5
echo "=== pytest linear tc"; timeout 600 python -m pytest tests/test_gpu_linear.py -m gpu -q --tb=short --maxfail=6 -p no:cacheprovider -x > gpurun_out/pytest_tc.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_tc.log | cut -c1-250
echo "=== pytest mma layout"; EXL2B_LAYOUT=mma timeout 600 python -m pytest tests -m gpu -q --tb=line --maxfail=6 -p no:cacheprovider 2>&1 | tail -3 | cut -c1-250
echo "=== microbench tc"; timeout 300 python tools/microbench.py --shapes qkvo,gateup,down,head --m 1,8 2>&1 | cut -c1-420
