#!/bin/bash
# GPU box helper: parity subset + microbench (with phase stamps) of the batch-1 GEMV.  usage: tools/gpu_mb.sh <out.json> [shapes]
OUT=${1:-gpurun_out/mb.json}
SHAPES=${2:-qkvo,qkvo54,gateup54,down,down43,head}
timeout 200 python -m pytest tests/test_gpu_linear.py tests/test_gpu_vs_reference.py tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -4
timeout 250 python tools/microbench.py --ref --shapes $SHAPES --m 1 --phases --json $OUT 2>&1 | python -c '
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l)
    except Exception:
        print(l.strip()[:200]); continue
    if "cta" in d:
        if d["launch"] in (2, 5) and d["cta"] == 100:
            print(d["shape"], "cta", d["cta"], "launch", d["launch"], d["rel_to_grid_start[start,requested,wait_done,staged,warp0_done,all_warps_done]"], "grid_ns", d["grid_ns"])
    else:
        print(d["shape"], "new_graph_us %.2f" % d["new_graph_us"], "TB/s %.2f" % (d["new_graph_gbs"] / 1e3), "ref_graph_us %.2f" % d.get("ref_graph_us", 0))
'
