#!/bin/bash
# ncu captures behind the numbers in profiles/ (run on the GPU box via gpurun; never a multi-rank command).
#   launch list of one bench step (cold-cache, serialised: compare SHARES)  + full captures of the top kernels.
# bench.py brackets its timed device steps with the NVTX range "lrp_timed": only those launches are profiled.
set -x
mkdir -p gpurun_out
BENCH="python bench.py --global-batch 8 --micro-batch 8 --steps 1 --warmup 3 --no-cpu-baseline --dropin 0 --no-kernels"
NV='--nvtx --nvtx-include lrp_timed/'
ncu $NV --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/launches_bench.log 2>&1
ncu $NV --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 10 -c 8 -f -o gpurun_out/prof_gemm $BENCH > gpurun_out/prof_gemm.log 2>&1
ncu $NV --set full --clock-control none --import-source on -k regex:attn_fwd_ws_kernel -s 4 -c 1 -f -o gpurun_out/prof_attn_fwd $BENCH > gpurun_out/prof_attn_fwd.log 2>&1
ncu $NV --set full --clock-control none --import-source on -k regex:attn_bwd_pipe_kernel -s 4 -c 1 -f -o gpurun_out/prof_attn_bwd $BENCH > gpurun_out/prof_attn_bwd.log 2>&1
ls -la gpurun_out | tail -8
