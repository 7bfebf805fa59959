#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_rules_gpu.py tests/test_nf4_gpu.py tests/test_engine_gpu.py tests/test_kernels_gpu.py tests/test_full_size_properties_gpu.py -m gpu -q --timeout=900 -s 2>&1 | grep -v "Warning\|warn(" > gpurun_out/pytest_gpu_f.log
grep -E "passed|failed|NF4|^E  |FAILED|rope|trace" gpurun_out/pytest_gpu_f.log | cut -c1-300 | tail -30
bash profiles/capture.sh > gpurun_out/capture.log 2>&1
tail -12 gpurun_out/capture.log
