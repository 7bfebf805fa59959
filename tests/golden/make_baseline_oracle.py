"""Fixtures for tests/test_baseline_configs_gpu.py: the CPU oracle (oracle/attnlrp_oracle.py, pinned to the real reference on the small
goldens by tests/test_oracle_vs_golden.py) evaluated ONCE at the BASELINE.json configurations at full width, so that the GPU parity
tests do not spend ~10 minutes of host time per run on it.  Weights and token ids are regenerated from their seeds by the test
(tests/helpers.py:baseline_case); only the oracle's outputs are stored: fp32 relevance, the reference-style bf16 relevance, the
arg-max token and every (S/32)-th row of the input-embedding gradient.

    python tests/golden/make_baseline_oracle.py [case ...]        # needs no GPU and no /root/reference; ~3-10 min per case
    LRP_FULL_ORACLE=1 pytest tests/test_baseline_configs_gpu.py    # ignores the fixtures and recomputes the oracle live
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from helpers import BASELINE_CASES, baseline_oracle  # noqa: E402

if __name__ == "__main__":
    out_dir = os.environ.get("LRP_GOLDEN_OUT", HERE)
    for name in (sys.argv[1:] or BASELINE_CASES):
        t0 = time.time()
        out = baseline_oracle(name)
        np.savez_compressed(os.path.join(out_dir, f"baseline_{name}.npz"), **out)
        print(f"baseline_{name}.npz  S={int(out['S'][0])}  |ref|={float(np.linalg.norm(out['ref'])):.6e}  {time.time() - t0:.0f} s", flush=True)
