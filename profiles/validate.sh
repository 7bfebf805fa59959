#!/bin/bash
# end-of-change validation on the GPU box: the -m gpu suite, then one default bench line (kept in gpurun_out/)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_final.json
cut -c1-400 gpurun_out/bench_final.json
