#!/bin/bash
mkdir -p gpurun_out
EXL2B_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_decode.csv python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/ncu_b.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_b.log | cut -c1-300; wc -l gpurun_out/launches_decode.csv
