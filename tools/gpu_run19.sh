#!/bin/bash
# phase profile build (clock64 accumulators compiled in), then restore the normal library
EXL2B_NVCC_EXTRA="-DEXL2B_TC_PROFILE" python -c "from exllamav2_b200 import build; build.build(force=True)" > /dev/null
echo "=== phases"; timeout 300 python tools/microbench.py --phases --shapes gateup,head --m 1 --total-mb 128 2>&1 | grep '"launch": [4-6]' | cut -c1-700
python -c "from exllamav2_b200 import build; build.build(force=True)" > /dev/null
