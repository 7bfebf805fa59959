#!/bin/bash
# round 2 final measurements, part 1: quick regression tests, bench lines, A/B of the CTA order inside the step, sqrt / nf4 variants
mkdir -p gpurun_out/final
timeout 600 python -m pytest tests/test_rules_gpu.py tests/test_gamma_gpu.py -m gpu -q --timeout=300 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/final/bench_n1.json 2> gpurun_out/final/bench_n1.err; tail -c 600 gpurun_out/final/bench_n1.json
for g in -1 2 -1 2; do LRP_ATTN_SCHED_GROUP=$g timeout 300 python bench.py --no-cpu-baseline --dropin 0 --no-kernels --steps 3 > gpurun_out/final/ab_sched_$g.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/final/ab_sched_$g.json'));print('sched',$g,d['value'],d['clocks'],d['roofline']['achieved'])"; done
timeout 400 python bench.py --no-cpu-baseline --dropin 0 --no-kernels --store sqrt > gpurun_out/final/bench_sqrt.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/final/bench_sqrt.json'));print('sqrt',d['value'],d['hbm_peak_gb'])"
timeout 400 python bench.py --no-cpu-baseline --dropin 0 --no-kernels --quant nf4 > gpurun_out/final/bench_nf4.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/final/bench_nf4.json'));print('nf4',d['value'],d['hbm_peak_gb'])"
timeout 400 python bench.py --no-cpu-baseline --dropin 0 --model tinyllama-1.1b --seq 512 > gpurun_out/final/bench_tinyllama.json 2>/dev/null; tail -c 300 gpurun_out/final/bench_tinyllama.json
timeout 600 python bench.py --no-cpu-baseline --dropin 0 --model gemma3-4b --seq 8192 --per-gpu-batch 4 --micro-batch 1 > gpurun_out/final/bench_gemma3.json 2>/dev/null; tail -c 300 gpurun_out/final/bench_gemma3.json
