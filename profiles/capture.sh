#!/bin/bash
# ncu captures behind the numbers in profiles/ (run on the GPU box via gpurun; never a multi-rank command).
#   launch list of one bench step (cold-cache, serialised: compare SHARES)  + full captures of the top kernels
set -x
mkdir -p gpurun_out
BENCH="python bench.py --per-gpu-batch 8 --micro-batch 8 --steps 1 --warmup 3 --no-cpu-baseline"
# one micro-batch of 8 prompts = 581 launches per step; skip the 3 warm-up steps
ncu --metrics gpu__time_duration.sum --clock-control none -s 1743 -c 581 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_pair_kernel -s 40 -c 3 -o gpurun_out/prof_gemm $BENCH > gpurun_out/prof_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 4 -c 1 -o gpurun_out/prof_attn_fwd $BENCH > gpurun_out/prof_attn_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_pipe_kernel -s 4 -c 1 -o gpurun_out/prof_attn_bwd $BENCH > gpurun_out/prof_attn_bwd.log 2>&1
ls -la gpurun_out
