#!/bin/bash
# GPU box helper: one-off runs (edit per experiment)
cd /root/repo; T=r02
for CTX in 4096 16384; do
  echo "=== decode at context $CTX"; timeout 400 python bench.py --context $CTX --steps 32 --no-cpu --no-ref-ext > gpurun_out/${T}_bench_ctx$CTX.json 2> gpurun_out/${T}_bench_ctx$CTX.err; cut -c1-330 gpurun_out/${T}_bench_ctx$CTX.json; tail -2 gpurun_out/${T}_bench_ctx$CTX.err | cut -c1-200
done
echo "=== prefill"; timeout 600 python bench.py --mode prefill --steps 4 > gpurun_out/${T}_bench_prefill.json 2> gpurun_out/${T}_bench_prefill.err; cut -c1-1200 gpurun_out/${T}_bench_prefill.json; tail -3 gpurun_out/${T}_bench_prefill.err | cut -c1-200
echo "=== attention tests"; timeout 300 python -m pytest tests/test_gpu_attn_long.py tests/test_gpu_decoder.py -m gpu -q 2>&1 | tail -2
