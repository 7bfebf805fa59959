"""CPU: the C-ABI library loads and exports every symbol include/lrp_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "lrp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lrp_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path_entry_points():
    syms = _declared_symbols()
    for must in ("lrp_linear_fwd", "lrp_linear_dgrad_fused", "lrp_linear_eps_bwd", "lrp_attn_fwd", "lrp_attn_bwd",
                 "lrp_rmsnorm_fwd", "lrp_rmsnorm_bwd", "lrp_gated_act_fwd", "lrp_gated_act_bwd", "lrp_rope_inplace",
                 "lrp_gxi_reduce", "lrp_layernorm_fwd", "lrp_layernorm_bwd", "lrp_last_error", "lrp_version"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from lxt_b200 import _capi
    assert os.path.exists(_capi.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/lrp_b200.h but not exported"


def test_python_binding_covers_the_header():
    from lxt_b200 import _capi
    assert sorted(_capi.SIGNATURES) == _declared_symbols()
    lib = _capi.lib()
    assert lib.lrp_version() >= 1000
    assert isinstance(lib.lrp_launch_count(), int)


def test_no_device_is_reported_not_faked():
    import torch
    from lxt_b200 import _capi
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    rc = _capi.lib().lrp_check_device()
    assert rc == -3
    assert b"CUDA" in _capi.lib().lrp_last_error() or b"device" in _capi.lib().lrp_last_error()
    with pytest.raises(_capi.LrpError):
        _capi.require_device()


def test_every_entry_point_cites_the_reference_code_it_replaces():
    """include/lrp_b200.h: the comment in front of each declaration names a reference (or call-site) file:line"""
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "lrp_b200.h")).read()
    decls = list(re.finditer(r"^(?:int|int64_t|const char\*)\s+(lrp_[a-z0-9_]+)\s*\(", hdr, re.M))
    from lxt_b200 import _capi
    assert len(decls) == len(_capi.SIGNATURES) >= 44
    prev, missing = 0, []
    for m in decls:
        if not re.search(r"[A-Za-z0-9_/\.]+\.py:\d+", hdr[prev:m.start()]):
            missing.append(m.group(1))
        prev = m.end()
    assert not missing, missing
