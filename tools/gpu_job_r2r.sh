#!/bin/bash
timeout 300 python tools/debug_taylor.py 2>&1 | tail -12
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 2>&1 | tail -12
