#!/bin/bash
timeout 600 python tools/debug_modules.py 2>&1 | grep -v Warning | tail -30
timeout 600 python -m pytest tests/test_modules_gpu.py -m gpu -q --timeout=300 --tb=line 2>&1 | tail -8
