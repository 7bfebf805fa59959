#!/bin/bash
# round 2, run: polynomial exp2 in the attention forward
L=lrp-explains-transformers_b200/lxt_b200/lib
export LD_LIBRARY_PATH=$L
for pl in 0 4 3 2; do
  echo "== FWD_POLY=$pl"; LRP_ATTN_FWD_POLY=$pl timeout 120 $L/selftest_attn 2>&1 | grep "D=128 causal=1 window=0 packed=1" | head -2
  LRP_ATTN_FWD_POLY=$pl timeout 120 $L/selftest_attn 2>&1 | tail -1
  for i in 1 2; do LRP_ATTN_FWD_POLY=$pl timeout 120 $L/selftest_attn --perf 2>&1 | grep "^perf" | head -2 | cut -c1-90; done
done
