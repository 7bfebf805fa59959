#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_dropin_gpu.py tests/test_precision_gpu.py tests/test_kernels_gpu.py tests/test_vit_gpu.py tests/test_qwen_gpu.py tests/test_monkey_patch_gpu.py tests/test_gemma3_gpu.py tests/test_engine_gpu.py tests/test_baseline_configs_gpu.py -m gpu -q --timeout=900 -s 2>&1 | grep -v "Warning\|warn(" > gpurun_out/pytest_gpu_d.log
grep -E "PARITY|passed|failed|rel-L2|padded|checkpoint|^E  |FAILED" gpurun_out/pytest_gpu_d.log | cut -c1-300 | tail -70
( cd lrp-explains-transformers_b200/lxt_b200/lib && timeout 600 ./selftest_gemm --sweep > ../../../gpurun_out/gemm_sweep_r2.log 2>&1 )
cat gpurun_out/gemm_sweep_r2.log | tail -50
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err
tail -c 3000 gpurun_out/bench_r2d.json; tail -5 gpurun_out/bench_r2d.err
