#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
bash tools/gpu_job_debug.sh
( cd lrp-explains-transformers_b200/lxt_b200/lib && timeout 300 ./selftest_attn > ../../../gpurun_out/st_attn_r2c.log 2>&1; timeout 300 ./selftest_gemm > ../../../gpurun_out/st_gemm_r2c.log 2>&1 )
tail -3 gpurun_out/st_attn_r2c.log gpurun_out/st_gemm_r2c.log
timeout 1500 python -m pytest tests/test_dropin_gpu.py tests/test_engine_gpu.py tests/test_precision_gpu.py tests/test_kernels_gpu.py tests/test_full_size_properties_gpu.py tests/test_monkey_patch_gpu.py tests/test_gemma3_gpu.py tests/test_vit_gpu.py -m gpu -q --timeout=900 -s 2>&1 | grep -v "Warning\|warn(" > gpurun_out/pytest_gpu_c.log
grep -E "PARITY|passed|failed|rel-L2|Error|error|padded|checkpoint|rope" gpurun_out/pytest_gpu_c.log | tail -50
