#!/bin/bash
# round 2, run 15: warp-specialised attention backward (attn_bwd_ws.cu), first light
L=lrp-explains-transformers_b200/lxt_b200/lib
export LD_LIBRARY_PATH=$L
echo "== selftest (ws)"; timeout 120 $L/selftest_attn 2>&1 | tail -22
echo "== perf ws"; timeout 120 $L/selftest_attn --perf 2>&1 | grep "^perf" | head -2
echo "== perf pipe"; LRP_ATTN_BWD=pipe timeout 120 $L/selftest_attn --perf 2>&1 | grep "^perf" | head -2
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout=300 2>&1 | tail -5
