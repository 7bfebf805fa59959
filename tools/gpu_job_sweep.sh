#!/bin/bash
# rasterisation group of the pair GEMM, re-swept with evict-first epilogue stores (same box)
for g in 32 16 64 32 24 48; do
  LRP_GROUP_M=$g timeout 300 python bench.py --no-cpu-baseline --dropin 0 --no-kernels --steps 3 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read());print('group',$g,'step',round(d['value'],3),d['clocks']['sm_mhz'],round(d['roofline']['achieved'],1),round(d['roofline']['share_of_step'],4))"
done
