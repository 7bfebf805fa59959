#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
for v in "mlp,rms,lin,attn" "none" "lin" "attn" "mlp" "rms"; do
  timeout 600 python tools/debug_dropin.py "$v" 2>&1 | grep "variant\[" 
done 2>&1 | tee gpurun_out/debug_dropin.log
