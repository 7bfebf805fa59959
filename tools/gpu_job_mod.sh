#!/bin/bash
timeout 600 python -m pytest tests/test_modules_gpu.py tests/test_rules_gpu.py -m gpu -q --timeout=300 --tb=short 2>&1 | tail -6
