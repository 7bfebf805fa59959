#!/bin/bash
# GPU job: the whole GPU test-suite (parity proper), then a short bench.  Run from the repo root on the GPU box.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
rm -f gpurun_out/parity_r02.jsonl
timeout 2400 python -m pytest tests -m gpu -q --timeout=900 --durations=15 -s 2>&1 | grep -v "Warning\|warn(" > gpurun_out/pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
grep -E "PARITY|passed|failed|rc=|rel-L2|Error|error" gpurun_out/pytest_gpu.log | tail -60
