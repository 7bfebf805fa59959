#!/bin/bash
# round 2, run 12: CTA order (heavy-first 1-D grid) and polynomial exp2 in the attention backward
cd lrp-explains-transformers_b200/lxt_b200/lib
export LD_LIBRARY_PATH=$PWD
echo "== correctness default"; timeout 300 ./selftest_attn 2>&1 | tail -4
echo "== correctness POLY=2"; LRP_ATTN_POLY=2 timeout 300 ./selftest_attn 2>&1 | tail -4
for g in -1 0 1 2 4; do echo "== SCHED_GROUP=$g"; LRP_ATTN_SCHED_GROUP=$g timeout 300 ./selftest_attn --perf 2>&1 | grep "^perf"; done
for pl in 2 3 4; do echo "== POLY=$pl (group 2)"; LRP_ATTN_POLY=$pl timeout 300 ./selftest_attn --perf 2>&1 | grep "^perf"; done
