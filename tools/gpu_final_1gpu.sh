#!/bin/bash
# round-end verification on one B200: full GPU test tier, smoke, the default bench (+ reference arm), ncu launch list,
# ncu --set full of the dominant kernel, in-model timeline.  Every command has its own timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/r01_gpu.csv 2>&1
echo "=== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -6 | cut -c1-300
echo "=== smoke"; timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "=== bench default"; timeout 600 python bench.py > gpurun_out/r01_bench_n1.json 2> gpurun_out/r01_bench_n1.err; echo "rc=$?"; cut -c1-900 gpurun_out/r01_bench_n1.json
echo "=== bench reference arm"; timeout 400 python bench.py --impl reference --steps 12 --warmup 3 > gpurun_out/r01_bench_ref.json 2>&1; echo "rc=$?"; cut -c1-600 gpurun_out/r01_bench_ref.json
echo "=== ncu launch list"; EXL2B_PROFILE=1 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_decode.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r01_ncu_launch.log 2>&1; echo "rc=$?"; wc -l gpurun_out/r01_launches_decode.csv
echo "=== ncu full"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 12 -c 1 -f -o gpurun_out/r01_gemm_tc python tools/microbench.py --shapes gateup --m 1 --total-mb 96 2>&1 | tail -2
echo "=== timeline"; timeout 300 python tools/model_timeline.py 2 > gpurun_out/r01_timeline.txt 2>&1; tail -11 gpurun_out/r01_timeline.txt
echo "=== microbench (+reference kernels)"; timeout 400 python tools/microbench.py --ref --shapes qkvo,gateup,down,head --m 1,8 > gpurun_out/r01_microbench.jsonl 2>&1; tail -8 gpurun_out/r01_microbench.jsonl | cut -c1-420
