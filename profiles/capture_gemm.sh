#!/bin/bash
# one `ncu --set full` capture of three consecutive GEMM launches of a bench step (run through gpurun)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_pair_kernel -s 40 -c 3 -o gpurun_out/prof_gemm -f \
  python bench.py --per-gpu-batch 8 --micro-batch 8 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_gemm.log 2>&1
ls -la gpurun_out/prof_gemm.ncu-rep
