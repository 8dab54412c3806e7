#!/bin/bash
# Round-2 evidence run on one B200 (each command has its own timeout; outputs under gpurun_out/, copied to profiles/ afterwards):
#   full GPU test tier, smoke, default bench (+ reference arm), ncu launch list of a decode step, ncu --set full of the batch-1
#   GEMV per shape class, in-model timeline, microbench vs the reference kernels at 1 row and at many rows, prefill mode,
#   TinyLlama / GPTQ / 70B presets.
mkdir -p gpurun_out
T=r02
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${T}_gpu.csv 2>&1
echo "=== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -6 | cut -c1-300
echo "=== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "=== bench default"; timeout 900 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "rc=$?"; cut -c1-1500 gpurun_out/${T}_bench_n1.json; tail -3 gpurun_out/${T}_bench_n1.err
echo "=== bench reference arm"; timeout 400 python bench.py --impl reference --steps 12 --warmup 3 > gpurun_out/${T}_bench_ref.json 2>&1; echo "rc=$?"; cut -c1-600 gpurun_out/${T}_bench_ref.json
echo "=== microbench 1 row (+reference kernels)"; timeout 500 python tools/microbench.py --ref --shapes qkvo,gateup,down,head,qkvo54,gateup54,down43,gptq --m 1 > gpurun_out/${T}_microbench.jsonl 2>&1; cut -c1-500 gpurun_out/${T}_microbench.jsonl | tail -9
echo "=== microbench many rows (+reference reconstruct+cuBLAS)"; timeout 500 python tools/microbench.py --ref --shapes qkvo,gateup,down --m 8,16,64,256,2048 > gpurun_out/${T}_microbench_rows.jsonl 2>&1; cut -c1-500 gpurun_out/${T}_microbench_rows.jsonl | tail -16
echo "=== ncu launch list"; EXL2B_PROFILE=1 timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_decode.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-ref-ext > gpurun_out/${T}_ncu_launch.log 2>&1; echo "rc=$?"; wc -l gpurun_out/${T}_launches_decode.csv
for SH in qkvo54 gateup54 down43 head; do
  echo "=== ncu full $SH"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemv_i8 -s 12 -c 1 -f -o gpurun_out/${T}_gemv_i8_$SH python tools/microbench.py --shapes $SH --m 1 --total-mb 96 2>&1 | tail -1 | cut -c1-200
done
echo "=== timeline"; timeout 300 python tools/model_timeline.py 2 > gpurun_out/${T}_timeline.txt 2>&1; tail -14 gpurun_out/${T}_timeline.txt
for CTX in 1024 4096 16384; do
  echo "=== decode at context $CTX"; timeout 400 python bench.py --context $CTX --steps 32 --no-cpu --no-ref-ext > gpurun_out/${T}_bench_ctx$CTX.json 2> gpurun_out/${T}_bench_ctx$CTX.err; cut -c1-400 gpurun_out/${T}_bench_ctx$CTX.json; tail -2 gpurun_out/${T}_bench_ctx$CTX.err
done
echo "=== prefill"; timeout 600 python bench.py --mode prefill --steps 4 > gpurun_out/${T}_bench_prefill.json 2> gpurun_out/${T}_bench_prefill.err; cut -c1-900 gpurun_out/${T}_bench_prefill.json; tail -3 gpurun_out/${T}_bench_prefill.err
echo "=== tinyllama"; timeout 600 python bench.py --model tinyllama-1.1b-4.0bpw --steps 64 --no-cpu --no-ref-ext > gpurun_out/${T}_bench_tinyllama.json 2> gpurun_out/${T}_bench_tinyllama.err; cut -c1-600 gpurun_out/${T}_bench_tinyllama.json; tail -2 gpurun_out/${T}_bench_tinyllama.err
echo "=== gptq 7b"; timeout 600 python bench.py --model llama2-7b-gptq-g128-act --steps 64 --no-cpu --no-ref-ext > gpurun_out/${T}_bench_gptq.json 2> gpurun_out/${T}_bench_gptq.err; cut -c1-600 gpurun_out/${T}_bench_gptq.json; tail -2 gpurun_out/${T}_bench_gptq.err
echo "=== 70b"; timeout 900 python bench.py --model llama2-70b-2.5bpw --steps 16 --warmup 3 --no-cpu --no-ref-ext > gpurun_out/${T}_bench_70b.json 2> gpurun_out/${T}_bench_70b.err; cut -c1-600 gpurun_out/${T}_bench_70b.json; tail -2 gpurun_out/${T}_bench_70b.err
