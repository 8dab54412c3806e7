#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest mlp"; timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=line -k "mlp or attn" -p no:cacheprovider 2>&1 | tail -3
echo "=== timeline 2 CTA/SM PDL"; timeout 300 python tools/microbench.py --phases --shapes qkvo --m 1 2>&1 | cut -c1-330
echo "=== timeline 1 CTA/SM PDL"; timeout 300 python tools/microbench.py --ctas-per-sm 1 --phases --shapes qkvo --m 1 2>&1 | cut -c1-330
echo "=== timeline 2 CTA/SM no PDL"; EXL2B_NO_PDL=1 timeout 300 python tools/microbench.py --phases --shapes qkvo --m 1 2>&1 | cut -c1-330
