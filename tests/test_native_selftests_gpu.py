"""GPU: the native (no Python, no torch) self-tests of liblrp_b200.so through the C ABI: tcgen05 GEMM incl. the fused
eps-LRP Linear kernel, and flash AttnLRP forward/backward in all three backward variants (the GEMM self-test covers the one-CTA and the CTA-pair kernels)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lrp-explains-transformers_b200", "lxt_b200", "lib")


def _run(exe, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([os.path.join(LIB, exe)], cwd=LIB, env=e, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "SELFTEST PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-1000:]


def test_native_gemm_and_fused_eps_linear():
    _run("selftest_gemm")


# "pipe" = the default software-pipelined kernel with bulk-tensor dQ reductions; v1 = first-generation red.global kernel;
# v2 = two-pass atomic-free kernels (always used for head_dim 256)
@pytest.mark.parametrize("variant", ["pipe", "v1", "v2"])
def test_native_flash_attnlrp(variant):
    _run("selftest_attn", {"LRP_ATTN_BWD": variant})
