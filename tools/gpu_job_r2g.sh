#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_full_size_properties_gpu.py tests/test_precision_gpu.py -m gpu -q --timeout=900 -s 2>&1 | grep -v "Warning\|warn(" > gpurun_out/pytest_gpu_g.log
grep -E "passed|failed|^E  |FAILED" gpurun_out/pytest_gpu_g.log | cut -c1-300 | tail -20
bash profiles/capture_gemm.sh
for v in "1 1" "0 1" "1 0"; do set -- $v; echo "== LRP_FUSE_ACT=$1 LRP_FUSE_DELTA=$2"; LRP_FUSE_ACT=$1 LRP_FUSE_DELTA=$2 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernels --dropin 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['share_of_step'], d['clocks']['sm_mhz'], d['gpu_launches'])"; done
