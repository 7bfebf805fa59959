#!/bin/bash
timeout 300 python tools/debug_repeat.py 2>&1 | tail -4
LRP_GEMM_STREAM_STORES=0 timeout 300 python tools/debug_repeat.py 2>&1 | tail -4
LRP_ATTN_SCHED_GROUP=-1 timeout 300 python tools/debug_repeat.py 2>&1 | tail -4
