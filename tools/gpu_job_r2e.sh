#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_dropin_gpu.py tests/test_rules_gpu.py tests/test_qwen_gpu.py tests/test_monkey_patch_gpu.py tests/test_gemma3_gpu.py tests/test_gpt2_gpu.py tests/test_engine_gpu.py tests/test_baseline_configs_gpu.py::test_llama3_8b_width_seq2048_engine_and_dropin -m gpu -q --timeout=900 -s 2>&1 | grep -v "Warning\|warn(" > gpurun_out/pytest_gpu_e.log
grep -E "PARITY|passed|failed|rel-L2|padded|checkpoint|^E  |FAILED|rope" gpurun_out/pytest_gpu_e.log | cut -c1-300 | tail -60
timeout 1200 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2e.json'))
print({k: d[k] for k in ("value","ms_per_step","gpu_launches","clocks")})
print("roofline", d["roofline"]["achieved"], d["roofline"]["share_of_step"])
print("dropin", json.dumps(d.get("dropin_api"))[:500])
print("cpu", json.dumps(d.get("cpu_baseline"))[:900])
print("attn", json.dumps(d["kernels"].get("flash_attnlrp"))[:600])
PY
tail -3 gpurun_out/bench_r2e.err
