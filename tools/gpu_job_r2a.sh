#!/bin/bash
# GPU job: native attention selftest (+perf) then the GPU test-suite.  Run from the repo root on the GPU box.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
( cd lrp-explains-transformers_b200/lxt_b200/lib && timeout 300 ./selftest_attn --perf ) > gpurun_out/st_attn_r2a.log 2>&1
echo "selftest rc=$?" >> gpurun_out/st_attn_r2a.log
tail -32 gpurun_out/st_attn_r2a.log
timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
