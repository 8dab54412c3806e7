#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 1500 python -m pytest tests/test_gpu_linear.py tests/test_gpu_ops.py tests/test_gpu_golden.py -m gpu -q --tb=short --maxfail=5 -p no:cacheprovider -x 2>&1 | tail -8 | cut -c1-300
for c in 2 1; do
echo "=== microbench TC_CTAS=$c"; EXL2B_TC_CTAS=$c timeout 300 python tools/microbench.py --shapes qkvo,gateup,down,head --m 1 2>&1 | cut -c1-330
echo "=== bench 7b TC_CTAS=$c"; EXL2B_TC_CTAS=$c timeout 900 python bench.py --steps 64 --warmup 4 --no-cpu > gpurun_out/bench_7b_23_$c.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_7b_23_$c.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ['value','ms_per_step','launches_per_step']}, d['e2e'], {k:d['roofline'][k] for k in ['achieved','frac','avg_launch_us']})" || tail -20 gpurun_out/bench_7b_23_$c.log
done
echo "=== timeline"; timeout 600 python tools/model_timeline.py 2 2>&1 | tail -11
