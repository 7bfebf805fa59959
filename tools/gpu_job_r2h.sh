#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -m gpu -q --timeout=600 2>&1 | tail -3
for v in "1 1" "0 1"; do set -- $v; echo "== LRP_FUSE_ACT=$1 LRP_FUSE_DELTA=$2"; LRP_FUSE_ACT=$1 LRP_FUSE_DELTA=$2 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernels --dropin 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['share_of_step'], d['clocks']['sm_mhz'], d['gpu_launches'])"; done
bash profiles/capture_gemm.sh
