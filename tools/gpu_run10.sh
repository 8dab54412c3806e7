#!/bin/bash
mkdir -p gpurun_out
echo "=== bench 7b (TC, 2 CTA/SM)"; timeout 900 python bench.py --steps 64 --warmup 4 --no-cpu > gpurun_out/bench_7b_10a.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_7b_10a.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ['value','ms_per_step','launches_per_step']}, d['e2e'], {k:d['roofline'][k] for k in ['achieved','frac','avg_launch_us']})"
echo "=== bench 7b (TC, 1 CTA/SM)"; EXL2B_TC_CTAS=1 timeout 900 python bench.py --steps 64 --warmup 4 --no-cpu > gpurun_out/bench_7b_10b.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_7b_10b.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ['value','ms_per_step','launches_per_step']}, d['e2e'], {k:d['roofline'][k] for k in ['achieved','frac','avg_launch_us']})"
echo "=== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 300 --csv --log-file gpurun_out/launches_r1a.csv python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/ncu_b.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/launches_r1a.csv | cut -c1-300
