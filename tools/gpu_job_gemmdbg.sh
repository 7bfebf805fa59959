#!/bin/bash
# streaming (evict-first) epilogue stores: same-box A/B, stand-alone GEMMs and the step
L=lrp-explains-transformers_b200/lxt_b200/lib
export LD_LIBRARY_PATH=$L
timeout 300 $L/selftest_gemm 2>&1 | tail -1
for s in 0 1 0 1; do
  echo "== STREAM_STORES=$s"
  LRP_GEMM_STREAM_STORES=$s timeout 300 $L/selftest_gemm --perf 2>&1 | grep "^perf" | head -7 | cut -c6-100
  LRP_GEMM_STREAM_STORES=$s timeout 300 python bench.py --no-cpu-baseline --dropin 0 --no-kernels --steps 3 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read());print('step',d['value'],d['clocks']['sm_mhz'],d['roofline']['achieved'],d['roofline']['share_of_step'])"
done
