#!/bin/bash
# round 2, run 13: dQ finished inside the pipelined attention backward (tile completion counters)
L=lrp-explains-transformers_b200/lxt_b200/lib
echo "== selftest"; LD_LIBRARY_PATH=$L timeout 300 $L/selftest_attn 2>&1 | grep -v "^ok" | tail -5
for i in 1 2; do LD_LIBRARY_PATH=$L timeout 300 $L/selftest_attn --perf 2>&1 | grep "^perf"; done
echo "== POLY=3"; LRP_ATTN_POLY=3 LD_LIBRARY_PATH=$L timeout 300 $L/selftest_attn --perf 2>&1 | grep "^perf"
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_rules_gpu.py tests/test_engine_gpu.py tests/test_monkey_patch_gpu.py tests/test_precision_gpu.py -m gpu -q -x --timeout=600 2>&1 | tail -8
