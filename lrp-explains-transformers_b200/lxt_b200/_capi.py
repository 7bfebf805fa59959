"""ctypes binding of liblrp_b200.so (the C ABI declared in include/lrp_b200.h).

The library is the product: there is NO Python/CPU fallback.  `lib()` raises `RuntimeError` when the shared
object is missing or a symbol is absent, and every wrapper raises `LrpError` on a non-zero return code.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "liblrp_b200.so")


class LrpError(RuntimeError):
    pass


class Epilogue(C.Structure):
    """mirror of `lrp_epilogue_t`"""

    _fields_ = [
        ("out", C.c_void_p),
        ("out_is_f32", C.c_int32),
        ("shadow_bf16", C.c_void_p),
        ("resid_f32", C.c_void_p),
        ("rowscale", C.c_void_p),
        ("colscale", C.c_void_p),
        ("bias", C.c_void_p),
        ("alpha", C.c_float),
        ("ldc", C.c_int64),
        ("gated_gu", C.c_void_p),
        ("gated_out", C.c_void_p),
        ("gated_act", C.c_int32),
        ("gated_cp", C.c_int32),
        ("act_out", C.c_void_p),
        ("delta_o", C.c_void_p),
        ("delta_out", C.c_void_p),
        ("delta_head_dim", C.c_int32),
        ("delta_seq", C.c_int32),
        ("gated_layout", C.c_int32),
    ]


_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
_EP = C.POINTER(Epilogue)

# name -> (restype, argtypes); must list every symbol of include/lrp_b200.h (tests check this)
SIGNATURES = {
    "lrp_version": (_i, []),
    "lrp_last_error": (C.c_char_p, []),
    "lrp_check_device": (_i, []),
    "lrp_launch_count": (_i64, []),
    "lrp_gemm_bf16": (_i, [_vp, _i64, _vp, _i64, _i, _i, _i, _i, _EP, _i, _vp]),
    "lrp_gemm_bf16_batched": (_i, [_vp, _i64, _i64, _i, _vp, _i64, _i64, _i, _i, _i, _i, _i, _EP, _i64, _vp]),
    "lrp_linear_fwd": (_i, [_vp, _i64, _vp, _i64, _i, _i, _i, _EP, _vp]),
    "lrp_linear_dgrad_fused": (_i, [_vp, _i64, _vp, _i64, _i, _i, _i, _EP, _vp]),
    "lrp_linear_eps_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _f, _vp]),
    "lrp_linear_eps_flags_count": (_i64, [_i]),
    "lrp_rmsnorm_fwd": (_i, [_vp, _i, _vp, _f, _f, _vp, _vp, _i, _i, _vp]),
    "lrp_rmsnorm_bwd": (_i, [_vp, _vp, _f, _vp, _vp, _i, _i, _i, _i, _vp]),
    "lrp_rmsnorm_fwd_residual": (_i, [_vp, _vp, _f, _f, _vp, _vp, _i, _i, _vp]),
    "lrp_headnorm_inplace": (_i, [_vp, _i64, _i, _i, _i, _vp, _vp, _f, _f, _vp, _i, _i, _vp]),
    "lrp_layernorm_fwd": (_i, [_vp, _vp, _vp, _f, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "lrp_layernorm_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "lrp_rope_inplace": (_i, [_vp, _i64, _i, _i, _vp, _vp, _i, _i, _i, _vp]),
    "lrp_gated_act_fwd": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "lrp_gated_act_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "lrp_act_identity_fwd": (_i, [_vp, _vp, _i64, _i, _i, _vp]),
    "lrp_act_identity_bwd": (_i, [_vp, _vp, _vp, _i64, _i, _i, _vp]),
    "lrp_attn_fwd": (_i, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _i, _i, _i, _i, _i, _f, _i, _i, _vp]),
    "lrp_attn_bwd": (
        _i,
        [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp,
         _i, _i, _i, _i, _i, _f, _i, _i, _f, _f, _f, _vp],
    ),
    "lrp_embed_gather": (_i, [_vp, _vp, _f, _vp, _i, _i, _vp]),
    "lrp_argmax_rows": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "lrp_gather_rows_f32": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "lrp_seed_gradient": (_i, [_vp, _vp, _vp, _f, _vp, _i, _vp, _vp, _i, _i, _vp]),
    "lrp_gxi_reduce": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "lrp_gxi_reduce_bf16": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "lrp_gxi_reduce_mixed": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "lrp_cast_f32_to_bf16": (_i, [_vp, _vp, _i64, _vp]),
    "lrp_eps_div": (_i, [_vp, _vp, _vp, _i64, _f, _f, _i, _vp]),
    "lrp_mul": (_i, [_vp, _vp, _vp, _i64, _f, _i, _vp]),
    "lrp_gamma_split": (_i, [_vp, _vp, _i64, _i, _i, _vp]),
    "lrp_gamma_s": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i, _f, _i, _vp]),
    "lrp_gamma_combine": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _vp]),
    "lrp_scale": (_i, [_vp, _vp, _i64, _f, _i, _vp]),
    "lrp_identity_rule_bwd": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _vp]),
    "lrp_softmax_dt_bwd": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _vp]),
    "lrp_add2_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _f, _i, _vp]),
    "lrp_softmax_fwd": (_i, [_vp, _vp, _i64, _i, _f, _i, _vp]),
    # validation-precision mode (fp32 activations)
    "lrp_rmsnorm_fwd_t": (_i, [_vp, _i, _vp, _f, _f, _vp, _i, _vp, _i, _i, _vp]),
    "lrp_rmsnorm_bwd_t": (_i, [_vp, _i, _vp, _f, _vp, _vp, _i, _i, _i, _i, _vp]),
    "lrp_rmsnorm_fwd_residual_t": (_i, [_vp, _i, _vp, _f, _f, _vp, _vp, _i, _i, _vp]),
    "lrp_headnorm_inplace_t": (_i, [_vp, _i, _i64, _i, _i, _i, _vp, _vp, _f, _f, _vp, _i, _i, _vp]),
    "lrp_rope_inplace_t": (_i, [_vp, _i, _i64, _i, _i, _vp, _vp, _i, _i, _i, _vp]),
    "lrp_gated_act_fwd_t": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "lrp_gated_act_bwd_t": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "lrp_split_bf16x2": (_i, [_vp, _vp, _vp, _i64, _vp]),
    "lrp_attn_fwd_f32": (_i, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _i, _i, _vp]),
    "lrp_attn_bwd_f32": (
        _i,
        [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp,
         _i, _i, _i, _i, _i, _f, _i, _i, _f, _f, _f, _vp],
    ),
    "lrp_attn_fwd_varlen": (_i, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _i, _i, _vp]),
    "lrp_attn_bwd_varlen": (
        _i,
        [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _i,
         _i, _i, _i, _i, _i, _f, _i, _i, _f, _f, _f, _vp],
    ),
    "lrp_attn_bwd_workspace_bytes": (_i, [_i, _i, _i, _i, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "lrp_quant_nf4": (_i, [_vp, _vp, _vp, _i64, _i, _vp]),
    "lrp_dequant_nf4": (_i, [_vp, _vp, _vp, _i64, _i, _vp]),
}

_lib = None
_lock = threading.Lock()


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"liblrp_b200.so not found at {LIB_PATH}: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (or `make -C lrp-explains-transformers_b200/csrc`). There is no CPU fallback."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(handle, name)
            except AttributeError as e:  # pragma: no cover - build/ABI mismatch
                raise RuntimeError(f"liblrp_b200.so does not export {name}") from e
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().lrp_last_error()
        raise LrpError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


def require_device() -> None:
    check(lib().lrp_check_device(), "lrp_check_device")
