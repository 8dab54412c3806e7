#!/bin/bash
# GPU box helper: one-off A/B runs (edit per experiment)
cd /root/repo
run_mb() { timeout 200 python tools/microbench.py --shapes $1 --m 1 2>&1 | python -c '
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print("  ", d["shape"], "us %.2f" % d["new_graph_us"], "TB/s %.2f" % (d["new_graph_gbs"] / 1e3))
'; }
run_bench() { timeout 300 python bench.py --steps 64 --warmup 8 --no-ref-ext --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().split('\n')[-1]); print('  tok/s %.1f' % d['value'], 'e2e %.1f' % d['e2e']['value'], 'roofline %.3f' % d['roofline']['frac'], 'avg_launch_us %.2f' % d['roofline']['avg_launch_us'], 'parity %.2e' % d['parity']['timed_vs_unchained_logits_rel_l2'])"; }
echo "== tests (i8 paths)"; timeout 400 python -m pytest tests/test_gpu_linear.py tests/test_gpu_row_blocks.py tests/test_gpu_decoder.py -m gpu -q -x 2>&1 | tail -2 | cut -c1-200
echo "== base"; run_bench; run_mb head,gateup54,down43,qkvo54
echo "== L1HINT=1"; export EXL2B_I8_L1HINT=1; run_bench; run_mb head,gateup54,down43,qkvo54; unset EXL2B_I8_L1HINT
echo "== DOUBLE_MB=90 (head only)"; export EXL2B_I8_DOUBLE_MB=90; run_bench; run_mb head; unset EXL2B_I8_DOUBLE_MB
echo "== DOUBLE_MB=40 (gate|up and head)"; export EXL2B_I8_DOUBLE_MB=40; run_bench; unset EXL2B_I8_DOUBLE_MB
echo "== DOUBLE_MB=20 (qkv, gate|up, down, head)"; export EXL2B_I8_DOUBLE_MB=20; run_bench; unset EXL2B_I8_DOUBLE_MB
echo "== DOUBLE_MB=1 (all) microbench"; export EXL2B_I8_DOUBLE_MB=1; run_mb head,gateup54,down43,qkvo54; unset EXL2B_I8_DOUBLE_MB
