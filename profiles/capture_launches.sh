#!/bin/bash
# per-launch device time of one micro-batch (8 prompts) of the bench step: 581 launches (cold-cache, serialised)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 1743 -c 581 --csv --log-file gpurun_out/launches.csv \
  python bench.py --per-gpu-batch 8 --micro-batch 8 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
tail -2 gpurun_out/launches_bench.log | cut -c1-300
wc -l gpurun_out/launches.csv
