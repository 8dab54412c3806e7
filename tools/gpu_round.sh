#!/bin/bash
# GPU box helper, one call = one iteration of the batch-1 GEMV work: parity, warp sweep, phase stamps, ncu, timeline, bench.
TAG=${1:-r02}
mkdir -p gpurun_out
echo "== pytest"; timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -12 | cut -c1-250
echo "== sweep"; tools/gpu_sweep.sh 2>&1 | tail -34
echo "== microbench"; tools/gpu_mb.sh gpurun_out/${TAG}_mb.json 2>&1 | grep -v "passed\|warnings\|Docs\|^$\|exp(" | tail -30
for W in 16; do
  echo "== ncu warps $W"; EXL2B_I8_WARPS=$W timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemv_i8 -s 40 -c 1 -f -o gpurun_out/${TAG}_i8_qkvo54_w$W python tools/microbench.py --shapes qkvo54 --m 1 2>&1 | tail -1 | cut -c1-150
done
echo "== ncu head"; timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemv_i8 -s 8 -c 1 -f -o gpurun_out/${TAG}_i8_head python tools/microbench.py --shapes head --m 1 --total-mb 300 2>&1 | tail -1 | cut -c1-150
echo "== timeline"; timeout 200 python tools/model_timeline.py 4 2>&1 | tail -24
echo "== bench"; timeout 500 python bench.py --steps 64 --warmup 8 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 2500 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
