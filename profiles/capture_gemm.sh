#!/bin/bash
# one `ncu --set full` capture of eight consecutive GEMM launches of a timed bench step (run through gpurun)
mkdir -p gpurun_out
ncu --nvtx --nvtx-include lrp_timed/ --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 10 -c 8 -f -o gpurun_out/prof_gemm \
  python bench.py --global-batch 8 --micro-batch 8 --steps 1 --warmup 3 --no-cpu-baseline --dropin 0 --no-kernels > gpurun_out/prof_gemm.log 2>&1
ls -la gpurun_out/prof_gemm.ncu-rep
