"""Thin torch-tensor wrappers over the C ABI.  PyTorch is used for device memory and streams only.

Every function launches hand-written sm_100a kernels from liblrp_b200.so on the CURRENT torch CUDA stream and
raises if the tensors are not CUDA tensors of the documented dtype/layout — there is no eager fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _capi
from ._capi import Epilogue, check

ACT_SILU, ACT_GELU_TANH, ACT_GELU_ERF = 0, 1, 2

# Optional per-launch timing of the dominant kernel (bench.py's roofline): when set to a list, every GEMM launch
# appends (flops, start_event, end_event) recorded on the launching stream.
GEMM_PROFILE = None


def launch_count() -> int:
    """kernels enqueued by liblrp_b200.so in this process so far"""
    return int(_capi.lib().lrp_launch_count())


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise _capi.LrpError(f"{name}: expected a CUDA tensor (the B200 path has no CPU fallback)")
    if t.dtype != dtype:
        raise _capi.LrpError(f"{name}: expected dtype {dtype}, got {t.dtype}")


def _rowmajor2d(t: torch.Tensor, name: str) -> int:
    if t.dim() != 2 or t.stride(1) != 1:
        raise _capi.LrpError(f"{name}: expected a 2-D tensor with contiguous rows")
    return t.stride(0)


def make_epilogue(out: torch.Tensor, *, resid: Optional[torch.Tensor] = None, rowscale=None, colscale=None, bias=None,
                  shadow: Optional[torch.Tensor] = None, alpha: float = 1.0, act_out: Optional[torch.Tensor] = None,
                  act: int = ACT_SILU, delta=None) -> Epilogue:
    ldc = _rowmajor2d(out, "out")
    e = Epilogue()
    e.out = out.data_ptr()
    e.out_is_f32 = 1 if out.dtype == torch.float32 else 0
    if out.dtype not in (torch.float32, torch.bfloat16):
        raise _capi.LrpError("gemm output must be bf16 or fp32")
    for name, t in (("resid", resid), ("rowscale", rowscale), ("colscale", colscale), ("bias", bias)):
        if t is not None:
            _need(t, torch.float32, name)
    if resid is not None and _rowmajor2d(resid, "resid") != ldc:
        raise _capi.LrpError("resid must share the output's leading dimension")
    if shadow is not None:
        _need(shadow, torch.bfloat16, "shadow")
        if _rowmajor2d(shadow, "shadow") != ldc:
            raise _capi.LrpError("shadow must share the output's leading dimension")
    e.shadow_bf16 = _p(shadow)
    e.resid_f32 = _p(resid)
    e.rowscale = _p(rowscale)
    e.colscale = _p(colscale)
    e.bias = _p(bias)
    e.alpha = alpha
    e.ldc = ldc
    if act_out is not None:   # fused gated-MLP forward: out = interleaved (gate | up) blocks, act_out = act(gate) * up
        _need(act_out, torch.bfloat16, "act_out")
        if out.dtype != torch.bfloat16 or not act_out.is_contiguous() or tuple(act_out.shape) != (out.shape[0], out.shape[1] // 2):
            raise _capi.LrpError("act_out must be a contiguous bf16 [M, N/2] tensor next to a bf16 output")
        e.act_out = act_out.data_ptr()
        e.gated_act = act
    if delta is not None:   # (o bf16 [M, N], delta_out fp32 [B,H,S], head_dim, seq): fused attention-backward prologue
        o_t, d_out, D, S = delta
        _need(o_t, torch.bfloat16, "delta o")
        _need(d_out, torch.float32, "delta out")
        if out.dtype != torch.bfloat16 or o_t.shape != out.shape or _rowmajor2d(o_t, "delta o") != ldc or not d_out.is_contiguous() \
                or d_out.numel() != out.shape[0] * (out.shape[1] // D):
            raise _capi.LrpError("fused delta: o must share the bf16 output's shape / layout, delta_out holds B*H*S floats")
        e.delta_o, e.delta_out, e.delta_head_dim, e.delta_seq = o_t.data_ptr(), d_out.data_ptr(), int(D), int(S)
    return e


def split_bf16x2(x: torch.Tensor):
    """fp32 x -> (hi, lo) bf16 with x ~= hi + lo to ~16 mantissa bits (validation-precision GEMM operands)"""
    _need(x, torch.float32, "x")
    x = x.contiguous()
    if x.numel() % 8:
        raise _capi.LrpError("split_bf16x2: element count must be a multiple of 8")
    hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    check(_capi.lib().lrp_split_bf16x2(x.data_ptr(), hi.data_ptr(), lo.data_ptr(), x.numel(), _stream()), "lrp_split_bf16x2")
    return hi, lo


def weight_split(w: torch.Tensor):
    """(hi, lo) bf16 split of an fp32 weight, memoised ON the tensor object (a module's Parameter): it lives and dies with that
    object and is invalidated by in-place updates.  (Never keyed by address: the caching allocator hands the storage of a
    freed model to the next one.)"""
    ent = getattr(w, "_lrp_split", None)
    if ent is not None and ent[0] == w._version and ent[1].shape == w.shape:
        return ent[1], ent[2]
    hi, lo = split_bf16x2(w.detach())
    try:
        w._lrp_split = (w._version, hi, lo)
    except Exception:   # pragma: no cover  (tensor subclasses without a __dict__)
        pass
    return hi, lo


def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, *, b_layout: int, tile_n: int = 0, b_split=None, **epi) -> torch.Tensor:
    """out = epilogue(a @ b.T) for b_layout 0 (b is [N,K]) or epilogue(a @ b) for b_layout 1 (b is [K,N]).
    An fp32 `a` selects the validation-precision form: a = hi + lo (two bf16 terms), out = epilogue(hi b) then
    out += alpha * rowscale * colscale * (lo b), both on the same bf16 tcgen05 kernel with fp32 accumulation; `out` must be fp32."""
    if a.dtype == torch.float32:
        if out.dtype != torch.float32:
            raise _capi.LrpError("gemm: an fp32 A operand (validation precision) needs an fp32 output")
        if epi.get("shadow") is not None:
            raise _capi.LrpError("gemm: no bf16 shadow in validation precision")
        hi, lo = split_bf16x2(a)
        epi2 = {k: v for k, v in epi.items() if k in ("rowscale", "colscale", "alpha")}
        if b.dtype == torch.float32:   # fp32 weights (an fp32 HF model): W = Wh + Wl as well; the lo*lo term (2^-18) is dropped
            bh, bl = b_split if b_split is not None else split_bf16x2(b)
            gemm(hi, bh, out, b_layout=b_layout, tile_n=tile_n, **epi)
            gemm(lo, bh, out, b_layout=b_layout, tile_n=tile_n, resid=out, **epi2)
            return gemm(hi, bl, out, b_layout=b_layout, tile_n=tile_n, resid=out, **epi2)
        gemm(hi, b, out, b_layout=b_layout, tile_n=tile_n, **epi)
        return gemm(lo, b, out, b_layout=b_layout, tile_n=tile_n, resid=out, **epi2)
    _need(a, torch.bfloat16, "a")
    _need(b, torch.bfloat16, "b")
    lda, ldb = _rowmajor2d(a, "a"), _rowmajor2d(b, "b")
    M, K = a.shape
    N = b.shape[0] if b_layout == 0 else b.shape[1]
    if (b.shape[1] if b_layout == 0 else b.shape[0]) != K:
        raise _capi.LrpError(f"gemm: inner dimensions differ ({tuple(a.shape)} vs {tuple(b.shape)}, layout {b_layout})")
    if tuple(out.shape) != (M, N):
        raise _capi.LrpError(f"gemm: output shape {tuple(out.shape)} != {(M, N)}")
    e = make_epilogue(out, **epi)
    prof = GEMM_PROFILE
    if prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    check(_capi.lib().lrp_gemm_bf16(a.data_ptr(), lda, b.data_ptr(), ldb, b_layout, M, N, K, C.byref(e), tile_n, _stream()),
          "lrp_gemm_bf16")
    if prof is not None:
        ev1.record()
        prof.append((2.0 * M * N * K, ev0, ev1))
    return out


def gemm_batched(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, *, a_layout: int = 0, b_layout: int = 0) -> torch.Tensor:
    """One launch for G independent products.  a [G,M,K] (a_layout 0) or [G,K,M] (a_layout 1: the stored tensor is A^T);
    b [G,N,K] (b_layout 0) or [G,K,N] (b_layout 1); out [G,M,N] bf16 or fp32, all contiguous.
    fp32 operands select validation precision: both sides are split into two bf16 terms and the three significant products
    (hi*hi, lo*hi, hi*lo) accumulate in fp32 on the same kernel."""
    if a.dim() != 3 or b.dim() != 3 or out.dim() != 3 or not (a.is_contiguous() and b.is_contiguous() and out.is_contiguous()):
        raise _capi.LrpError("gemm_batched: contiguous 3-D tensors expected")
    G = a.shape[0]
    M, K = (a.shape[1], a.shape[2]) if a_layout == 0 else (a.shape[2], a.shape[1])
    N = b.shape[1] if b_layout == 0 else b.shape[2]
    if b.shape[0] != G or (b.shape[2] if b_layout == 0 else b.shape[1]) != K or tuple(out.shape) != (G, M, N):
        raise _capi.LrpError(f"gemm_batched: shapes {tuple(a.shape)}, {tuple(b.shape)} -> {tuple(out.shape)} do not match")
    if a.dtype == torch.float32 or b.dtype == torch.float32:
        if out.dtype != torch.float32:
            raise _capi.LrpError("gemm_batched: fp32 operands need an fp32 output")
        ah, al = split_bf16x2(a.float())
        bh, bl = split_bf16x2(b.float())
        _gemm_batched_bf16(ah, bh, out, a_layout, b_layout, False)
        _gemm_batched_bf16(al, bh, out, a_layout, b_layout, True)
        return _gemm_batched_bf16(ah, bl, out, a_layout, b_layout, True)
    return _gemm_batched_bf16(a, b, out, a_layout, b_layout, False)


def _gemm_batched_bf16(a, b, out, a_layout, b_layout, accumulate):
    _need(a, torch.bfloat16, "a")
    _need(b, torch.bfloat16, "b")
    G = a.shape[0]
    M, K = (a.shape[1], a.shape[2]) if a_layout == 0 else (a.shape[2], a.shape[1])
    N = b.shape[1] if b_layout == 0 else b.shape[2]
    e = make_epilogue(out[0], resid=out[0] if accumulate else None)
    check(_capi.lib().lrp_gemm_bf16_batched(a.data_ptr(), a.stride(1), a.stride(0), a_layout, b.data_ptr(), b.stride(1), b.stride(0),
                                            b_layout, G, M, N, K, C.byref(e), out.stride(0), _stream()), "lrp_gemm_bf16_batched")
    return out


def linear_dgrad_gated_bwd(gy: torch.Tensor, w: torch.Tensor, gu: torch.Tensor, ggu: torch.Tensor, act: int = ACT_SILU,
                           cp: bool = False, layout: int = 0) -> torch.Tensor:
    """Down-projection LRP dgrad with the gated-MLP point-wise rules fused into the epilogue:
    g_a = gy @ w  (never written);  ggu = [g_gate | g_up] as `gated_act_bwd(g_a, gu)` would produce.
    gy [T,d] bf16, w [d,I] bf16 (nn.Linear layout of down_proj), gu/ggu [T,2I] bf16 contiguous."""
    _need(gy, torch.bfloat16, "gy")
    _need(w, torch.bfloat16, "w")
    _need(gu, torch.bfloat16, "gu")
    _need(ggu, torch.bfloat16, "ggu")
    T, K = gy.shape
    I = w.shape[1]
    if w.shape[0] != K or tuple(gu.shape) != (T, 2 * I) or tuple(ggu.shape) != (T, 2 * I) or not (gu.is_contiguous() and ggu.is_contiguous()):
        raise _capi.LrpError("linear_dgrad_gated_bwd: shape mismatch")
    e = Epilogue()
    e.alpha = 1.0
    e.ldc = I
    e.gated_gu, e.gated_out, e.gated_act, e.gated_cp, e.gated_layout = gu.data_ptr(), ggu.data_ptr(), act, int(cp), int(layout)
    prof = GEMM_PROFILE
    if prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    check(_capi.lib().lrp_gemm_bf16(gy.data_ptr(), _rowmajor2d(gy, "gy"), w.data_ptr(), _rowmajor2d(w, "w"), 1, T, I, K,
                                    C.byref(e), 0, _stream()), "lrp_gemm_bf16(gated)")
    if prof is not None:
        ev1.record()
        prof.append((2.0 * T * I * K, ev0, ev1))
    return ggu


def linear_fwd(x: torch.Tensor, w: torch.Tensor, out: torch.Tensor, **epi) -> torch.Tensor:
    """y = x W^T (+ fused epilogue).  x [T,K] bf16, w [N,K] bf16  (fp32 x / w: validation precision, see `gemm`)."""
    return gemm(x, w, out, b_layout=0, **epi)


def linear_dgrad(gy: torch.Tensor, w: torch.Tensor, out: torch.Tensor, **epi) -> torch.Tensor:
    """g_x = g_y W (+ fused epilogue).  gy [T,N] bf16, w [N,K] bf16 read in place (no transpose copy)."""
    return gemm(gy, w, out, b_layout=1, **epi)


def rmsnorm_fwd(x: torch.Tensor, w: torch.Tensor, eps: float, *, w_offset: float = 0.0, want_rstd: bool = True,
                out_dtype=torch.bfloat16, out: Optional[torch.Tensor] = None, rstd: Optional[torch.Tensor] = None):
    """x [T,d] (bf16 or fp32) -> (y [T,d] bf16 (default) or fp32, rstd fp32 [T])"""
    if x.dtype not in (torch.float32, torch.bfloat16):
        raise _capi.LrpError("rmsnorm_fwd: x must be bf16 or fp32")
    _need(x, x.dtype, "x")
    _need(w, torch.bfloat16, "w")
    T, d = x.shape
    if not x.is_contiguous():
        raise _capi.LrpError("rmsnorm_fwd: x must be contiguous")
    y = out if out is not None else torch.empty((T, d), dtype=out_dtype, device=x.device)
    if y.dtype not in (torch.float32, torch.bfloat16) or not y.is_contiguous():
        raise _capi.LrpError("rmsnorm_fwd: y must be a contiguous bf16 or fp32 tensor")
    if rstd is None and want_rstd:
        rstd = torch.empty((T,), dtype=torch.float32, device=x.device)
    check(_capi.lib().lrp_rmsnorm_fwd_t(x.data_ptr(), int(x.dtype == torch.float32), w.data_ptr(), w_offset, eps,
                                        y.data_ptr(), int(y.dtype == torch.float32), _p(rstd), T, d, _stream()), "lrp_rmsnorm_fwd")
    return y, rstd


def rmsnorm_bwd(gy: torch.Tensor, w: torch.Tensor, rstd: torch.Tensor, *, w_offset: float = 0.0,
                out: Optional[torch.Tensor] = None, out_dtype=torch.bfloat16, accumulate: bool = False) -> torch.Tensor:
    if gy.dtype not in (torch.float32, torch.bfloat16):
        raise _capi.LrpError("rmsnorm_bwd: gy must be bf16 or fp32")
    _need(gy, gy.dtype, "gy")
    _need(w, torch.bfloat16, "w")
    _need(rstd, torch.float32, "rstd")
    T, d = gy.shape
    if not gy.is_contiguous():
        raise _capi.LrpError("rmsnorm_bwd: gy must be contiguous")
    if out is None:
        out = torch.empty((T, d), dtype=out_dtype, device=gy.device)
    check(_capi.lib().lrp_rmsnorm_bwd_t(gy.data_ptr(), int(gy.dtype == torch.float32), w.data_ptr(), w_offset, rstd.data_ptr(),
                                        out.data_ptr(), int(out.dtype == torch.float32), int(accumulate), T, d, _stream()),
          "lrp_rmsnorm_bwd")
    return out


def layernorm_fwd(x: torch.Tensor, w: Optional[torch.Tensor], b: Optional[torch.Tensor], eps: float):
    if x.dtype not in (torch.float32, torch.bfloat16) or not x.is_cuda or not x.is_contiguous():
        raise _capi.LrpError("layernorm_fwd: x must be a contiguous CUDA bf16/fp32 tensor")
    T, d = x.shape
    y = torch.empty_like(x)
    mean = torch.empty((T,), dtype=torch.float32, device=x.device)
    rstd = torch.empty((T,), dtype=torch.float32, device=x.device)
    check(_capi.lib().lrp_layernorm_fwd(x.data_ptr(), _p(w), _p(b), eps, y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), T, d,
                                        int(x.dtype == torch.float32), _stream()), "lrp_layernorm_fwd")
    return y, mean, rstd


def layernorm_bwd(gy: torch.Tensor, w: Optional[torch.Tensor], rstd: torch.Tensor) -> torch.Tensor:
    if gy.dtype not in (torch.float32, torch.bfloat16) or not gy.is_cuda or not gy.is_contiguous():
        raise _capi.LrpError("layernorm_bwd: gy must be a contiguous CUDA bf16/fp32 tensor")
    T, d = gy.shape
    gx = torch.empty_like(gy)
    check(_capi.lib().lrp_layernorm_bwd(gy.data_ptr(), _p(w), rstd.data_ptr(), gx.data_ptr(), T, d,
                                        int(gy.dtype == torch.float32), _stream()), "lrp_layernorm_bwd")
    return gx


def rope_inplace(qk: torch.Tensor, n_heads: int, D: int, cos: torch.Tensor, sin: torch.Tensor, S: int, *, inverse: bool = False):
    """rotate the first n_heads*D columns of qk [T, ld] in place; cos/sin fp32 [S, D/2]."""
    if qk.dtype not in (torch.float32, torch.bfloat16):
        raise _capi.LrpError("rope_inplace: qk must be bf16 or fp32")
    _need(qk, qk.dtype, "qk")
    _need(cos, torch.float32, "cos")
    _need(sin, torch.float32, "sin")
    ld = _rowmajor2d(qk, "qk")
    check(_capi.lib().lrp_rope_inplace_t(qk.data_ptr(), int(qk.dtype == torch.float32), ld, n_heads, D, cos.data_ptr(),
                                         sin.data_ptr(), qk.shape[0], S, int(inverse), _stream()), "lrp_rope_inplace")
    return qk


def gated_act_fwd(gu: torch.Tensor, act: int = ACT_SILU, out: Optional[torch.Tensor] = None, layout: int = 0) -> torch.Tensor:
    if gu.dtype not in (torch.float32, torch.bfloat16):
        raise _capi.LrpError("gated_act_fwd: gu must be bf16 or fp32")
    _need(gu, gu.dtype, "gu")
    T, I2 = gu.shape
    if not gu.is_contiguous():
        raise _capi.LrpError("gated_act_fwd: gu must be contiguous")
    a = out if out is not None else torch.empty((T, I2 // 2), dtype=gu.dtype, device=gu.device)
    if a.dtype != gu.dtype or not a.is_contiguous():
        raise _capi.LrpError("gated_act_fwd: output must be contiguous and share gu's dtype")
    check(_capi.lib().lrp_gated_act_fwd_t(gu.data_ptr(), a.data_ptr(), int(gu.dtype == torch.float32), int(layout), T, I2 // 2, act,
                                          _stream()), "lrp_gated_act_fwd")
    return a


def gated_act_bwd(ga: torch.Tensor, gu: torch.Tensor, act: int = ACT_SILU, out: Optional[torch.Tensor] = None, cp: bool = False,
                  layout: int = 0) -> torch.Tensor:
    if gu.dtype not in (torch.float32, torch.bfloat16):
        raise _capi.LrpError("gated_act_bwd: gu must be bf16 or fp32")
    _need(ga, gu.dtype, "ga")
    _need(gu, gu.dtype, "gu")
    T, I2 = gu.shape
    if not (ga.is_contiguous() and gu.is_contiguous()):
        raise _capi.LrpError("gated_act_bwd: inputs must be contiguous")
    if out is None:
        out = torch.empty_like(gu)
    if out.dtype != gu.dtype or not out.is_contiguous():
        raise _capi.LrpError("gated_act_bwd: output must be contiguous and share gu's dtype")
    check(_capi.lib().lrp_gated_act_bwd_t(ga.data_ptr(), gu.data_ptr(), out.data_ptr(), int(gu.dtype == torch.float32), int(layout), T,
                                          I2 // 2, act, int(cp), _stream()), "lrp_gated_act_bwd")
    return out


def act_identity_fwd(x: torch.Tensor, act: int) -> torch.Tensor:
    if x.dtype not in (torch.float32, torch.bfloat16) or not x.is_cuda or not x.is_contiguous():
        raise _capi.LrpError("act_identity_fwd: x must be a contiguous CUDA bf16/fp32 tensor")
    y = torch.empty_like(x)
    check(_capi.lib().lrp_act_identity_fwd(x.data_ptr(), y.data_ptr(), x.numel(), act, int(x.dtype == torch.float32), _stream()),
          "lrp_act_identity_fwd")
    return y


def act_identity_bwd(gy: torch.Tensor, x: torch.Tensor, act: int) -> torch.Tensor:
    if gy.dtype != x.dtype or not gy.is_cuda or not gy.is_contiguous() or not x.is_contiguous():
        raise _capi.LrpError("act_identity_bwd: gy/x must be contiguous CUDA tensors of the same dtype")
    gx = torch.empty_like(x)
    check(_capi.lib().lrp_act_identity_bwd(gy.data_ptr(), x.data_ptr(), gx.data_ptr(), x.numel(), act,
                                           int(x.dtype == torch.float32), _stream()), "lrp_act_identity_bwd")
    return gx


def _bshd(t: torch.Tensor, name: str, dtype=torch.bfloat16):
    """accept [B,S,H,D] tensors whose last dim is contiguous, head stride D, batch stride S*token stride"""
    _need(t, dtype, name)
    B, S, H, D = t.shape
    if t.stride(3) != 1 or (H > 1 and t.stride(2) != D) or (B > 1 and t.stride(0) != S * t.stride(1)):
        raise _capi.LrpError(f"{name}: expected [B,S,H,D] with strides (S*ld, ld, D, 1), got {t.stride()}")
    return t.stride(1)


def _kv_range_ptr(kv_range, B: int):
    if kv_range is None:
        return None
    if not (kv_range.is_cuda and kv_range.dtype == torch.int32 and kv_range.is_contiguous() and tuple(kv_range.shape) == (B, 2)):
        raise _capi.LrpError("kv_range must be a contiguous CUDA int32 tensor of shape [B, 2] (valid key range per sequence)")
    return kv_range.data_ptr()


def attn_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: float, *, causal: bool = True, window: int = 0, kv_range=None):
    """q [B,S,H,D], k/v [B,S,Hkv,D] (views into a packed buffer are fine) -> (o [B,S,H,D] bf16, lse fp32 [B,H,S]).
    kv_range: optional int32 [B,2] device tensor, keys outside [lo, hi) of each sequence are masked (padded batches)."""
    dt = q.dtype
    if dt not in (torch.float32, torch.bfloat16):
        raise _capi.LrpError("attn_fwd: q must be bf16 or fp32")
    ldq, ldk, ldv = _bshd(q, "q", dt), _bshd(k, "k", dt), _bshd(v, "v", dt)
    B, S, H, D = q.shape
    Hkv = k.shape[2]
    o = torch.empty((B, S, H, D), dtype=dt, device=q.device)
    lse = torch.empty((B, H, S), dtype=torch.float32, device=q.device)
    if dt == torch.float32:   # validation precision: fp32 CUDA-core kernel
        check(_capi.lib().lrp_attn_fwd_f32(q.data_ptr(), k.data_ptr(), v.data_ptr(), ldq, ldk, ldv, o.data_ptr(), lse.data_ptr(),
                                           _kv_range_ptr(kv_range, B), B, S, H, Hkv, D, scale, int(causal), window, _stream()),
              "lrp_attn_fwd_f32")
        return o, lse
    check(_capi.lib().lrp_attn_fwd_varlen(q.data_ptr(), k.data_ptr(), v.data_ptr(), ldq, ldk, ldv, o.data_ptr(), lse.data_ptr(),
                                          _kv_range_ptr(kv_range, B), B, S, H, Hkv, D, scale, int(causal), window, _stream()),
          "lrp_attn_fwd")
    return o, lse


_ATTN_WS: dict = {}


def attn_bwd_workspace_floats(B: int, S: int, H: int, D: int) -> int:
    """fp32 elements of the `dq_acc` workspace of attn_bwd (lrp_attn_bwd_workspace_bytes: [B,S,H,D] accumulator + counters)"""
    import ctypes
    nb = ctypes.c_int64(0)
    check(_capi.lib().lrp_attn_bwd_workspace_bytes(B, S, H, D, ctypes.byref(nb), None), "lrp_attn_bwd_workspace_bytes")
    return nb.value // 4


def attn_bwd(q, k, v, o, d_o, lse, scale: float, *, causal: bool = True, window: int = 0, q_div: float = 4.0,
             k_div: float = 4.0, v_div: float = 2.0, dq=None, dk=None, dv=None, dq_acc=None, delta=None, kv_range=None, flags: int = 0):
    """LRP backward of attention: returns (dq, dk, dv) already divided by (q_div, k_div, v_div)."""
    dt = q.dtype
    if dt not in (torch.float32, torch.bfloat16):
        raise _capi.LrpError("attn_bwd: q must be bf16 or fp32")
    ldq, ldk, ldv = _bshd(q, "q", dt), _bshd(k, "k", dt), _bshd(v, "v", dt)
    B, S, H, D = q.shape
    Hkv = k.shape[2]
    _need(o, dt, "o")
    _need(d_o, dt, "d_o")
    if not (o.is_contiguous() and d_o.is_contiguous()):
        raise _capi.LrpError("attn_bwd: o and d_o must be contiguous [B,S,H,D]")
    if dq is None:
        dq = torch.empty((B, S, H, D), dtype=dt, device=q.device)
    if dk is None:
        dk = torch.empty((B, S, Hkv, D), dtype=dt, device=q.device)
    if dv is None:
        dv = torch.empty((B, S, Hkv, D), dtype=dt, device=q.device)
    lddq, lddk, lddv = _bshd(dq, "dq", dt), _bshd(dk, "dk", dt), _bshd(dv, "dv", dt)
    if dt == torch.float32:   # validation precision: fp32 CUDA-core kernels (dQ pass + key-major dK/dV pass)
        if delta is None:
            delta = torch.empty((B, H, S), dtype=torch.float32, device=q.device)
        check(_capi.lib().lrp_attn_bwd_f32(q.data_ptr(), k.data_ptr(), v.data_ptr(), ldq, ldk, ldv, o.data_ptr(), d_o.data_ptr(),
                                           lse.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), lddq, lddk, lddv,
                                           delta.data_ptr(), _kv_range_ptr(kv_range, B), B, S, H, Hkv, D, scale, int(causal), window,
                                           q_div, k_div, v_div, _stream()), "lrp_attn_bwd_f32")
        return dq, dk, dv
    need = attn_bwd_workspace_floats(B, S, H, D)
    if dq_acc is None:
        # kept per (device, stream, shape): the kernel hands the workspace back zero, so only its first use pays the zero-fill
        key = (q.device.index, _stream(), B, S, H, D)
        dq_acc = _ATTN_WS.pop(key, None)
        if dq_acc is None:    # (head_dim 256 runs the atomic-free two-pass kernels: no accumulator, a token buffer satisfies the ABI)
            dq_acc = torch.zeros(need if D != 256 else 64, dtype=torch.float32, device=q.device)
        _ATTN_WS[key] = dq_acc                      # re-inserted last: dict order is the LRU order
        while len(_ATTN_WS) > 4:
            _ATTN_WS.pop(next(iter(_ATTN_WS)))
        flags = int(flags) | 2                      # LRP_ATTN_ACC_ZERO
    elif dq_acc.dtype != torch.float32 or (dq_acc.numel() < need and D != 256) or not dq_acc.is_contiguous():
        raise _capi.LrpError(f"attn_bwd: dq_acc must be a contiguous fp32 workspace of >= {need} elements "
                             "(attn_bwd_workspace_floats: accumulator + query-tile counters)")
    if delta is None:
        delta = torch.empty((B, H, S), dtype=torch.float32, device=q.device)
    try:
        check(_capi.lib().lrp_attn_bwd_varlen(q.data_ptr(), k.data_ptr(), v.data_ptr(), ldq, ldk, ldv, o.data_ptr(), d_o.data_ptr(),
                                              lse.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), lddq, lddk, lddv,
                                              dq_acc.data_ptr(), delta.data_ptr(), _kv_range_ptr(kv_range, B), int(flags), B, S, H, Hkv, D,
                                              scale, int(causal), window, q_div, k_div, v_div, _stream()), "lrp_attn_bwd")
    except Exception:
        _ATTN_WS.clear()      # a failed launch may leave a kept workspace dirty
        raise
    return dq, dk, dv


def embed_gather(ids: torch.Tensor, emb: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    _need(ids, torch.int64, "ids")
    _need(emb, torch.bfloat16, "emb")
    T = ids.numel()
    d = emb.shape[1]
    h = torch.empty((T, d), dtype=torch.float32, device=emb.device)
    check(_capi.lib().lrp_embed_gather(ids.data_ptr(), emb.data_ptr(), scale, h.data_ptr(), T, d, _stream()), "lrp_embed_gather")
    return h


def argmax_rows(logits: torch.Tensor):
    _need(logits, torch.float32, "logits")
    B, V = logits.shape
    idx = torch.empty((B,), dtype=torch.int32, device=logits.device)
    val = torch.empty((B,), dtype=torch.float32, device=logits.device)
    check(_capi.lib().lrp_argmax_rows(logits.data_ptr(), idx.data_ptr(), val.data_ptr(), B, V, _stream()), "lrp_argmax_rows")
    return idx, val


def gxi_reduce(x: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """relevance[t] = sum_d x[t,d]*g[t,d]; x,g both fp32 or both bf16, contiguous [T,d]"""
    mixed = x.dtype == torch.bfloat16 and g.dtype == torch.float32
    if (x.dtype != g.dtype and not mixed) or x.shape != g.shape or not (x.is_contiguous() and g.is_contiguous()):
        raise _capi.LrpError("gxi_reduce: x and g must be contiguous tensors of identical shape and dtype (or bf16 x with fp32 g)")
    T, d = x.shape
    rel = torch.empty((T,), dtype=torch.float32, device=x.device)
    if x.dtype == torch.bfloat16 and g.dtype == torch.float32:
        check(_capi.lib().lrp_gxi_reduce_mixed(x.data_ptr(), g.data_ptr(), rel.data_ptr(), T, d, _stream()), "lrp_gxi_reduce_mixed")
    elif x.dtype == torch.float32:
        check(_capi.lib().lrp_gxi_reduce(x.data_ptr(), g.data_ptr(), rel.data_ptr(), T, d, _stream()), "lrp_gxi_reduce")
    else:
        _need(x, torch.bfloat16, "x")
        check(_capi.lib().lrp_gxi_reduce_bf16(x.data_ptr(), g.data_ptr(), rel.data_ptr(), T, d, _stream()), "lrp_gxi_reduce_bf16")
    return rel


def cast_bf16(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need(x, torch.float32, "x")
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    check(_capi.lib().lrp_cast_f32_to_bf16(x.data_ptr(), out.data_ptr(), x.numel(), _stream()), "lrp_cast_f32_to_bf16")
    return out


# ---------------------------------------------------------------------------------------------------------
# generic element-wise rule kernels (any shape; tensors of one call share dtype bf16 or fp32)
# ---------------------------------------------------------------------------------------------------------
def _ew_prepare(*ts: torch.Tensor):
    t0 = ts[0]
    if not t0.is_cuda:
        raise _capi.LrpError("lxt_b200 rules run on CUDA tensors only (no CPU fallback)")
    if t0.dtype not in (torch.float32, torch.bfloat16):
        raise _capi.LrpError(f"lxt_b200 rules support bf16 and fp32 tensors, got {t0.dtype}")
    out = []
    for t in ts:
        if t.dtype != t0.dtype or t.shape != t0.shape or t.device != t0.device:
            t = t.to(device=t0.device, dtype=t0.dtype).expand(t0.shape)
        out.append(t.contiguous())
    return out, int(t0.dtype == torch.float32)


def eps_div(r: torch.Tensor, z: torch.Tensor, eps: float, alpha: float = 1.0) -> torch.Tensor:
    """r / (alpha*z + eps)"""
    (r, z), f32 = _ew_prepare(r, z)
    out = torch.empty_like(r)
    check(_capi.lib().lrp_eps_div(r.data_ptr(), z.data_ptr(), out.data_ptr(), r.numel(), alpha, eps, f32, _stream()), "lrp_eps_div")
    return out


def gamma_split(x2: torch.Tensor) -> torch.Tensor:
    """[max(x,0) | min(x,0)] side by side: x [T,K] -> [T,2K] (Gamma rule, see efficient/zennit_rules.py)"""
    (x2,), f32 = _ew_prepare(x2)
    T, K = x2.shape
    out = torch.empty((T, 2 * K), dtype=x2.dtype, device=x2.device)
    check(_capi.lib().lrp_gamma_split(x2.data_ptr(), out.data_ptr(), T, K, f32, _stream()), "lrp_gamma_split")
    return out


def gamma_s(g: torch.Tensor, y: torch.Tensor, zp: torch.Tensor, zn: torch.Tensor, eps: float) -> torch.Tensor:
    """[ [y>0] g*y/stab(zp) | [y<0] g*y/stab(zn) ]: [T,N] x4 -> [T,2N]"""
    (g, y, zp, zn), f32 = _ew_prepare(g, y, zp, zn)
    T, N = g.shape
    out = torch.empty((T, 2 * N), dtype=g.dtype, device=g.device)
    check(_capi.lib().lrp_gamma_s(g.data_ptr(), y.data_ptr(), zp.data_ptr(), zn.data_ptr(), out.data_ptr(), T, N, eps, f32, _stream()),
          "lrp_gamma_s")
    return out


def gamma_combine(x: torch.Tensor, g1: torch.Tensor, g2: torch.Tensor) -> torch.Tensor:
    """x * (x>0 ? g1 : g2) / stabilize(x, 1e-10)"""
    (x, g1, g2), f32 = _ew_prepare(x, g1, g2)
    out = torch.empty_like(x)
    check(_capi.lib().lrp_gamma_combine(x.data_ptr(), g1.data_ptr(), g2.data_ptr(), out.data_ptr(), x.numel(), f32, _stream()),
          "lrp_gamma_combine")
    return out


def mul(a: torch.Tensor, b: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    (a, b), f32 = _ew_prepare(a, b)
    out = torch.empty_like(a)
    check(_capi.lib().lrp_mul(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), scale, f32, _stream()), "lrp_mul")
    return out


def scale(x: torch.Tensor, factor: float) -> torch.Tensor:
    (x,), f32 = _ew_prepare(x)
    out = torch.empty_like(x)
    check(_capi.lib().lrp_scale(x.data_ptr(), out.data_ptr(), x.numel(), factor, f32, _stream()), "lrp_scale")
    return out


def identity_rule_bwd(gy: torch.Tensor, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    (gy, x, y), f32 = _ew_prepare(gy, x, y)
    gx = torch.empty_like(x)
    check(_capi.lib().lrp_identity_rule_bwd(gy.data_ptr(), x.data_ptr(), y.data_ptr(), gx.data_ptr(), x.numel(), f32, _stream()),
          "lrp_identity_rule_bwd")
    return gx


def softmax_dt_bwd(x: torch.Tensor, p: torch.Tensor, r: torch.Tensor) -> torch.Tensor:
    """Deep-Taylor softmax relevance over the LAST dimension"""
    (x, p, r), f32 = _ew_prepare(x, p, r)
    cols = x.shape[-1]
    out = torch.empty_like(x)
    check(_capi.lib().lrp_softmax_dt_bwd(x.data_ptr(), p.data_ptr(), r.data_ptr(), out.data_ptr(), x.numel() // cols, cols, f32,
                                         _stream()), "lrp_softmax_dt_bwd")
    return out


def softmax_fwd(x: torch.Tensor, temperature: float = 1.0) -> torch.Tensor:
    """softmax(x / temperature) over the LAST dimension (fp32 arithmetic; bf16 or fp32 storage)"""
    (x,), f32 = _ew_prepare(x)
    cols = x.shape[-1]
    out = torch.empty_like(x)
    check(_capi.lib().lrp_softmax_fwd(x.data_ptr(), out.data_ptr(), x.numel() // cols, cols, float(temperature), f32, _stream()),
          "lrp_softmax_fwd")
    return out


def add2_bwd(a: torch.Tensor, b: torch.Tensor, r: torch.Tensor, eps: float):
    (a, b, r), f32 = _ew_prepare(a, b, r)
    ra, rb = torch.empty_like(a), torch.empty_like(a)
    check(_capi.lib().lrp_add2_bwd(a.data_ptr(), b.data_ptr(), r.data_ptr(), ra.data_ptr(), rb.data_ptr(), a.numel(), eps, f32,
                                   _stream()), "lrp_add2_bwd")
    return ra, rb


def quant_nf4(w: torch.Tensor, blocksize: int = 64):
    """bf16 weight -> (packed uint8 [n/2], absmax fp32 [n/blocksize]) on the device (NF4, see csrc/quant.cu)"""
    _need(w, torch.bfloat16, "w")
    w = w.contiguous()
    n = w.numel()
    if n % blocksize or w.shape[-1] % blocksize:
        raise _capi.LrpError(f"quant_nf4: the contiguous dimension ({w.shape[-1]}) must be a multiple of the block size {blocksize}")
    packed = torch.empty(n // 2, dtype=torch.uint8, device=w.device)
    absmax = torch.empty(n // blocksize, dtype=torch.float32, device=w.device)
    check(_capi.lib().lrp_quant_nf4(w.data_ptr(), packed.data_ptr(), absmax.data_ptr(), n, blocksize, _stream()), "lrp_quant_nf4")
    return packed, absmax


def dequant_nf4(packed: torch.Tensor, absmax: torch.Tensor, out: torch.Tensor, blocksize: int = 64) -> torch.Tensor:
    """expand NF4 codes into the bf16 tensor `out` (contiguous, numel = 2 * packed.numel())"""
    _need(packed, torch.uint8, "packed")
    _need(absmax, torch.float32, "absmax")
    _need(out, torch.bfloat16, "out")
    if not out.is_contiguous() or out.numel() != 2 * packed.numel():
        raise _capi.LrpError("dequant_nf4: out must be contiguous with 2 * packed.numel() elements")
    check(_capi.lib().lrp_dequant_nf4(packed.data_ptr(), absmax.data_ptr(), out.data_ptr(), out.numel(), blocksize, _stream()),
          "lrp_dequant_nf4")
    return out


def pad_to8(t: torch.Tensor, dims) -> torch.Tensor:
    """zero-pad the given dims of `t` up to multiples of 8 (tensor-core tiles need 16-byte rows)"""
    pads = [0, 0] * t.dim()
    need = False
    for d in dims:
        d = d % t.dim()
        extra = (-t.shape[d]) % 8
        if extra:
            need = True
            pads[2 * (t.dim() - 1 - d) + 1] = extra
    return torch.nn.functional.pad(t, pads) if need else t


def linear_eps_bwd(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], r_out: torch.Tensor, eps: float) -> torch.Tensor:
    """Fused relevance-space epsilon rule of nn.Linear in one launch: R_in = x * ((R_out/(x W^T + b + eps)) W).
    x [T,K], w [N,K] are consumed as bf16 (fp32 inputs are rounded once); R_out / R_in keep R_out's dtype.
    Feature counts that are not multiples of 8 are zero-padded (padded outputs have R = 0, z = 0 -> s = 0)."""
    if not x.is_cuda:
        raise _capi.LrpError("linear_eps_bwd: CUDA tensors only")
    T, K0 = x.shape
    N0 = w.shape[0]
    xb = pad_to8(x.to(torch.bfloat16), [1]).contiguous()
    wb = pad_to8(w.to(torch.bfloat16), [0, 1]).contiguous()
    N, K = wb.shape
    r = pad_to8(r_out, [1]).contiguous()
    if r.dtype not in (torch.float32, torch.bfloat16):
        raise _capi.LrpError("linear_eps_bwd: relevance must be bf16 or fp32")
    bf = None if bias is None else pad_to8(bias.to(torch.float32), [0]).contiguous()
    r_in = torch.empty((T, K), dtype=r.dtype, device=x.device)
    s_ws = torch.empty((T, N), dtype=torch.bfloat16, device=x.device)
    flags = torch.zeros((int(_capi.lib().lrp_linear_eps_flags_count(T)),), dtype=torch.int32, device=x.device)
    check(_capi.lib().lrp_linear_eps_bwd(xb.data_ptr(), wb.data_ptr(), _p(bf), r.data_ptr(), int(r.dtype == torch.float32),
                                         r_in.data_ptr(), s_ws.data_ptr(), flags.data_ptr(), T, N, K, eps, _stream()),
          "lrp_linear_eps_bwd")
    return r_in[:, :K0] if K != K0 else r_in
