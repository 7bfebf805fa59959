#!/bin/bash
# round 2 final, part 3: whole GPU suite with per-test durations (fixtures for the full-width oracle), ViT-L/16 result lines
timeout 2400 python -m pytest tests -m gpu -q --timeout=900 --durations=40 2>&1 | tail -70 > gpurun_out/pytest_gpu_final.txt; tail -60 gpurun_out/pytest_gpu_final.txt
timeout 600 python tools/bench_vit.py > gpurun_out/bench_vit_l16.jsonl 2> gpurun_out/bench_vit.err; cat gpurun_out/bench_vit_l16.jsonl | cut -c1-400; tail -3 gpurun_out/bench_vit.err
