#!/bin/bash
# GPU box helper: warp-count / shared-memory-budget sweep of the batch-1 GEMV (microbench, graph replay).  usage: tools/gpu_sweep.sh <tag>
TAG=${1:-sweep}
for W in 16 12; do for SM in 114688 98304; do
  echo "== warps $W smem $SM"
  EXL2B_I8_WARPS=$W EXL2B_I8_SMEM=$SM timeout 200 python tools/microbench.py --shapes qkvo54,gateup54,down43,head --m 1 2>&1 | python -c '
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print("  ", d["shape"], "us %.2f" % d["new_graph_us"], "TB/s %.2f" % (d["new_graph_gbs"] / 1e3))
'
done; done
