#!/bin/bash
cd /root/repo
for PF in 0 1; do
  echo "== L2PF=$PF"
  EXL2B_I8_L2PF=$PF timeout 300 python bench.py --steps 64 --warmup 8 --no-ref-ext 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().split('\n')[-1]); print('tok/s', d['value'], 'roofline', d['roofline']['frac'], 'avg_launch_us', d['roofline']['avg_launch_us'])"
  EXL2B_I8_L2PF=$PF timeout 200 python tools/model_timeline.py 4 2>&1 | grep -E "^ (15|16|17|18|19) "
done
