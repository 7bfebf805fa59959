#!/bin/bash
# round 2: timeline + perf of the warp-specialised attention backward
L=lrp-explains-transformers_b200/lxt_b200/lib
export LD_LIBRARY_PATH=$L
LRP_ATTN_DEBUG=1 timeout 120 $L/selftest_attn --perf 2>&1 | grep -v "^ok\|^perf" | head -18
timeout 120 $L/selftest_attn 2>&1 | grep -v "^ok" | tail -3
timeout 120 $L/selftest_attn --perf 2>&1 | grep "^perf" | head -2
