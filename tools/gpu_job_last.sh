#!/bin/bash
# sanity of the clean-built library: native self-tests + engine / kernel tests
timeout 600 python -m pytest tests/test_native_selftests_gpu.py tests/test_engine_gpu.py tests/test_kernels_gpu.py tests/test_monkey_patch_gpu.py -m gpu -q --timeout=300 2>&1 | tail -3
