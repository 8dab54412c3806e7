#!/bin/bash
# GPU box helper: full parity suite, microbench (phase stamps), decode timeline, short bench.  usage: tools/gpu_round.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
echo "== pytest"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== microbench"; tools/gpu_mb.sh gpurun_out/${TAG}_mb.json 2>&1 | grep -v "passed\|warnings\|Docs\|^$" | tail -30
echo "== timeline"; timeout 200 python tools/model_timeline.py 4 2>&1 | tail -24
echo "== bench"; timeout 500 python bench.py --steps 64 --warmup 8 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 3000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
