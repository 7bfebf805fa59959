#!/bin/bash
# last run of the round: whole GPU suite on the final tree
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 2>&1 | tail -50 > gpurun_out/pytest_gpu_final.txt; tail -3 gpurun_out/pytest_gpu_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
