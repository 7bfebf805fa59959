#!/bin/bash
# round 2, run 14: in-kernel dQ finish variants (LRP_ATTN_FIN 0..3)
L=lrp-explains-transformers_b200/lxt_b200/lib
export LD_LIBRARY_PATH=$L
for f in 0 1 2 3; do
  echo "== FIN=$f"; LRP_ATTN_FIN=$f timeout 300 $L/selftest_attn 2>&1 | grep -v "^ok" | tail -3
  LRP_ATTN_FIN=$f timeout 300 $L/selftest_attn --perf 2>&1 | grep "^perf" | head -2
done
echo "== FIN=2 POLY=3"; LRP_ATTN_FIN=2 LRP_ATTN_POLY=3 timeout 300 $L/selftest_attn --perf 2>&1 | grep "^perf" | head -2
