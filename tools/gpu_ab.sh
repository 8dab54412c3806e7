#!/bin/bash
# GPU box helper: one-off A/B runs (edit per experiment)
cd /root/repo
EXL2B_NO_PDL=1 timeout 200 python tools/pdl_check.py /tmp/nopdl.pt test-small 2>&1 | tail -3
timeout 200 python tools/pdl_check.py /tmp/pdl.pt test-small /tmp/nopdl.pt 2>&1 | tail -4
timeout 200 python tools/pdl_check.py /tmp/pdl2.pt test-small /tmp/nopdl.pt 2>&1 | tail -4
