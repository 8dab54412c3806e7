#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 600 python -m pytest tests/test_gpu_linear.py tests/test_gpu_ops.py tests/test_gpu_decoder.py -m gpu -q --tb=short --maxfail=5 -p no:cacheprovider -x 2>&1 | tail -3 | cut -c1-300
echo "=== bench 7b"; timeout 400 python bench.py --steps 64 --warmup 4 --no-cpu > gpurun_out/bench_7b_27.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_7b_27.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ['value','ms_per_step','launches_per_step']}, d['e2e'], {k:d['roofline'][k] for k in ['achieved','frac','avg_launch_us']})" || tail -20 gpurun_out/bench_7b_27.log
echo "=== timeline"; timeout 200 python tools/model_timeline.py 2 2>&1 | tail -7
