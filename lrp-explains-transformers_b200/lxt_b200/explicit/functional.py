"""Relevance-space rules — drop-in for `lxt.explicit.functional` (reference lxt/explicit/functional.py:44-158,
276-536).  Same names / argument order / defaults.  Users seed with `y.backward(relevance)` and read `x.grad`.

All arithmetic of the backward rules runs in liblrp_b200.so kernels:
  linear_epsilon   one fused tcgen05 launch  (z, R/(z+eps), contraction with W, * x)
  matmul           eps+uniform rule: eps_div kernel + two strided-batched tcgen05 launches (all [B*H] slices at once, a^T
                   consumed in place through an MN-major descriptor) + mul kernels; fp32 operands: two-term bf16 split
  softmax          forward kernel + Deep-Taylor rule kernel;  add2 / mul2 / rms_norm_identity element-wise kernels
CUDA tensors only (bf16 or fp32).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch.autograd import Function

from .. import ops
from .._capi import LrpError


# Debug switch of the reference (lxt/explicit/functional.py:10-37, lxt/explicit/check.py:6-15): when set, every rule
# redistributes the incoming relevance UNIFORMLY over its inputs so that conservation can be checked end to end.
CONSERVATION_CHECK_FLAG = [False]


def conservation_check_wrap(func):
    """decorator for rule backwards: pass-through normally, uniform redistribution under `conservation_check()`"""

    def wrapped(ctx, *out_relevance):
        inp_relevance = func(ctx, *out_relevance)
        if not CONSERVATION_CHECK_FLAG[0]:
            return inp_relevance
        total = sum(r.float().sum() for r in out_relevance if r is not None)
        count = sum(r.numel() for r in inp_relevance if r is not None)
        mean = total / count
        if torch.isnan(mean).any():
            raise ValueError(f"NaN at {func}")
        return tuple(torch.full(r.shape, float(mean), dtype=r.dtype, device=r.device) if r is not None else None
                     for r in inp_relevance)

    return wrapped


def _stabilize(input, epsilon=1e-6, inplace=False):
    """`z + eps` — no sign handling, exactly like the reference (functional.py:266-273)."""
    return input.add_(epsilon) if inplace else input + epsilon


def _fwd_linear(x2, w, b):
    """forward of linear on the tcgen05 GEMM; returns fp32 or bf16 following x (odd widths are zero-padded)"""
    N0 = w.shape[0]
    xb = ops.pad_to8(x2.to(torch.bfloat16), [1]).contiguous()
    wb = ops.pad_to8(w.to(torch.bfloat16), [0, 1]).contiguous()
    out = torch.empty((x2.shape[0], wb.shape[0]), dtype=x2.dtype if x2.dtype == torch.bfloat16 else torch.float32, device=x2.device)
    ops.linear_fwd(xb, wb, out, bias=None if b is None else ops.pad_to8(b.float(), [0]).contiguous())
    return out[:, :N0] if wb.shape[0] != N0 else out


class linear_epsilon_fn(Function):
    @staticmethod
    def forward(ctx, inputs, weight, bias=None, epsilon=1e-6):
        if not inputs.is_cuda:
            raise LrpError("linear_epsilon: CUDA tensors only (no CPU fallback)")
        x2 = inputs.reshape(-1, inputs.shape[-1])
        out = _fwd_linear(x2, weight, bias).to(inputs.dtype)
        ctx.save_for_backward(inputs, weight, bias if bias is not None else torch.empty(0, device=inputs.device))
        ctx.has_bias, ctx.epsilon = bias is not None, epsilon
        return out.view(*inputs.shape[:-1], weight.shape[0])

    @staticmethod
    @conservation_check_wrap
    def backward(ctx, *out_relevance):
        inputs, weight, bias = ctx.saved_tensors
        R = out_relevance[0].reshape(-1, weight.shape[0])
        if R.dtype not in (torch.float32, torch.bfloat16):
            R = R.float()
        r_in = ops.linear_eps_bwd(inputs.reshape(-1, inputs.shape[-1]), weight, bias if ctx.has_bias else None, R, ctx.epsilon)
        return r_in.to(inputs.dtype).view(inputs.shape), None, None, None


def _bmm_kernel(a3, b3, b_layout, a_layout=0):
    """ONE strided-batched tcgen05 launch over the leading batch dim (three for fp32 operands: two-term bf16 split).
    a3 [G,M,K] (a_layout 0) or [G,K,M] (a_layout 1, the stored tensor is a^T); b3 [G,N,K] (b_layout 0) or [G,K,N] (b_layout 1).
    Returns fp32 [G,M,N].  Extents that are not multiples of 8 are zero-padded (the products are unchanged)."""
    G = a3.shape[0]
    M0 = a3.shape[1] if a_layout == 0 else a3.shape[2]
    N0 = b3.shape[1] if b_layout == 0 else b3.shape[2]
    dt = torch.float32 if torch.float32 in (a3.dtype, b3.dtype) else torch.bfloat16
    ab = ops.pad_to8(a3.to(dt), [1, 2]).contiguous()
    bb = ops.pad_to8(b3.to(dt), [1, 2]).contiguous()
    M = ab.shape[1] if a_layout == 0 else ab.shape[2]
    N = bb.shape[1] if b_layout == 0 else bb.shape[2]
    out = torch.empty((G, M, N), dtype=torch.float32, device=a3.device)
    ops.gemm_batched(ab, bb, out, a_layout=a_layout, b_layout=b_layout)
    return out[:, :M0, :N0] if (M, N) != (M0, N0) else out


class matmul_fn(Function):
    """epsilon + uniform rule for torch.matmul (reference lxt/explicit/functional.py:367-408) at any batch shape, e.g. the
    [B,H,S,S] attention products: forward one batched launch, backward eps_div + two batched launches + two multiplies."""

    @staticmethod
    def forward(ctx, input_a, input_b, inplace=False, epsilon=1e-6):
        if not input_a.is_cuda:
            raise LrpError("matmul: CUDA tensors only (no CPU fallback)")
        M, K = input_a.shape[-2:]
        N = input_b.shape[-1]
        a3 = input_a.reshape(-1, M, K)
        b3 = input_b.expand(*input_a.shape[:-2], K, N).reshape(-1, K, N)
        out = _bmm_kernel(a3, b3, 1).to(input_a.dtype).view(*input_a.shape[:-2], M, N)
        ctx.save_for_backward(input_a, input_b, out)
        ctx.epsilon = epsilon
        return out

    @staticmethod
    @conservation_check_wrap
    def backward(ctx, *out_relevance):
        a, b, out = ctx.saved_tensors
        M, K = a.shape[-2:]
        N = b.shape[-1]
        s = ops.eps_div(out_relevance[0].to(out.dtype), out, ctx.epsilon, alpha=2.0)        # R / (2 O + eps)
        s3 = s.reshape(-1, M, N)
        a3 = a.reshape(-1, M, K)
        b3 = b.expand(*a.shape[:-2], K, N).reshape(-1, K, N)
        ra = ops.mul(_bmm_kernel(s3, b3, 0).to(a.dtype).view(a.shape), a)                   # (s b^T) * a : b3 is [G, N_out=K, contraction N]
        rb = _bmm_kernel(a3, s3, 1, a_layout=1).to(b.dtype)                                 # (a^T s): a3 consumed as the stored transpose
        rb = ops.mul(rb.view(*a.shape[:-2], K, N), b.expand(*a.shape[:-2], K, N))
        if rb.shape != b.shape:
            rb = rb.sum_to_size(b.shape)
        return ra, rb, None, None


class softmax_fn(Function):
    @staticmethod
    def forward(ctx, inputs, dim, dtype=None, temperature=1.0, inplace=False):
        if not inputs.is_cuda:
            raise LrpError("softmax: CUDA tensors only (no CPU fallback)")
        if dtype is not None:
            inputs = inputs.to(dtype)
        if inputs.dtype not in (torch.float32, torch.bfloat16):
            inputs = inputs.float()
        if temperature != 1.0:
            inputs = ops.scale(inputs, 1.0 / temperature)      # the rule is stated on x / temperature (functional.py:297-299)
        d = dim if dim >= 0 else inputs.dim() + dim
        if d != inputs.dim() - 1:
            outputs = ops.softmax_fwd(inputs.transpose(d, -1).contiguous()).transpose(d, -1)
        else:
            outputs = ops.softmax_fwd(inputs)
        ctx.save_for_backward(inputs, outputs)
        ctx.dim = dim
        return outputs

    @staticmethod
    @conservation_check_wrap
    def backward(ctx, *out_relevance):
        inputs, outputs = ctx.saved_tensors
        dim = ctx.dim if ctx.dim >= 0 else inputs.dim() + ctx.dim
        R = out_relevance[0].to(inputs.dtype)
        if dim != inputs.dim() - 1:
            x, p, r = (t.transpose(dim, -1).contiguous() for t in (inputs, outputs, R))
            return ops.softmax_dt_bwd(x, p, r).transpose(dim, -1), None, None, None, None
        return ops.softmax_dt_bwd(inputs, outputs, R), None, None, None, None


class flash_attention_fn(Function):
    """Relevance-space attention without the [B,H,S,S] tensors: the chain the reference's explicit models spell out as
        lf.matmul(q, k^T) -> lf.mul2(., scale) -> lf.add2(., mask) -> lf.softmax -> lf.matmul(p, v)
    (reference lxt/explicit/models/llama.py:378-391) as ONE flash forward and ONE flash backward.  For eps -> 0 the four rules
    telescope (every division by the scores / probabilities cancels against the factor the next rule multiplies with):
        g_O = R_O / (O + eps/2);   (dQ, dK, dV) = soft-max attention backward of g_O;
        R_Q = Q * dQ / 4,  R_K = K * dK / 4,  R_V = V * dV / 2
    i.e. the Gradient x Input form of the same uniform rules (lxt/efficient/patches.py:193-203), so the kernels of the efficient
    path serve the explicit API and nothing quadratic in S is stored.  q [B,H,S,D], k/v [B,Hkv,S,D] (grouped heads are consumed
    natively; the relevance of a shared K/V head is the sum over its query heads, as `repeat_kv` + autograd gives)."""

    @staticmethod
    def forward(ctx, query, key, value, scale=None, causal=True, window=0, epsilon=1e-6):
        if not query.is_cuda:
            raise LrpError("flash_attention: CUDA tensors only (no CPU fallback)")
        dt = query.dtype if query.dtype in (torch.bfloat16, torch.float32) else torch.float32
        q, k, v = (t.to(dt).transpose(1, 2).contiguous() for t in (query, key, value))      # [B,S,H,D]
        scale = float(q.shape[-1]) ** -0.5 if scale is None else float(scale)
        o, lse = ops.attn_fwd(q, k, v, scale, causal=bool(causal), window=int(window))
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.cfg = (scale, bool(causal), int(window), float(epsilon), query.dtype)
        return o.transpose(1, 2).to(query.dtype)

    @staticmethod
    @conservation_check_wrap
    def backward(ctx, *out_relevance):
        q, k, v, o, lse = ctx.saved_tensors
        scale, causal, window, epsilon, dt_in = ctx.cfg
        r = out_relevance[0].to(o.dtype).transpose(1, 2).contiguous()
        g = ops.eps_div(r, o, 0.5 * epsilon)                                                 # R_O / (O + eps/2)
        dq, dk, dv = ops.attn_bwd(q, k, v, o, g, lse, scale, causal=causal, window=window, q_div=4.0, k_div=4.0, v_div=2.0)
        rq, rk, rv = ops.mul(dq, q), ops.mul(dk, k), ops.mul(dv, v)
        return (rq.transpose(1, 2).to(dt_in), rk.transpose(1, 2).to(dt_in), rv.transpose(1, 2).to(dt_in), None, None, None, None)


class add2_tensors_fn(Function):
    @staticmethod
    def forward(ctx, input_a, input_b, inplace=False, epsilon=1e-6):
        outputs = input_a + input_b
        if input_a.requires_grad or input_b.requires_grad:
            ctx.save_for_backward(input_a, input_b)
            ctx.epsilon = epsilon
        return outputs

    @staticmethod
    @conservation_check_wrap
    def backward(ctx, *out_relevance):
        a, b = ctx.saved_tensors
        shape = torch.broadcast_shapes(a.shape, b.shape)
        ra, rb = ops.add2_bwd(a.expand(shape), b.expand(shape), out_relevance[0].to(a.dtype), ctx.epsilon)
        return ra.sum_to_size(a.shape), rb.sum_to_size(b.shape), None, None


class rms_norm_identity_fn(Function):
    @staticmethod
    def forward(ctx, hidden_states, weight, variance_epsilon):
        if hidden_states.is_cuda and hidden_states.dtype == torch.bfloat16:
            y, _ = ops.rmsnorm_fwd(hidden_states.reshape(-1, hidden_states.shape[-1]).contiguous(),
                                   weight.to(torch.bfloat16).contiguous(), variance_epsilon, want_rstd=False)
            return y.view(hidden_states.shape)
        hs = hidden_states.to(torch.float32)
        hs = hs * torch.rsqrt(hs.pow(2).mean(-1, keepdim=True) + variance_epsilon)
        return weight * hs.to(hidden_states.dtype)

    @staticmethod
    @conservation_check_wrap
    def backward(ctx, *out_relevance):
        return out_relevance + (None, None)


class mul2_fn(Function):
    @staticmethod
    def forward(ctx, input_a, input_b, inplace=False):
        ctx.requires_grads = [i for i, t in enumerate((input_a, input_b)) if isinstance(t, torch.Tensor) and t.requires_grad]
        return input_a * input_b

    @staticmethod
    @conservation_check_wrap
    def backward(ctx, *out_relevance):
        r = ops.scale(out_relevance[0], 1.0 / len(ctx.requires_grads))
        return tuple(r if i in ctx.requires_grads else None for i in range(2)) + (None,)


def _rowsum(x2: torch.Tensor) -> torch.Tensor:
    """sum over the last dim of a contiguous 2-D tensor on the reduction kernel (sum_d x * 1)"""
    return ops.gxi_reduce(x2, torch.ones_like(x2))


def _last(x: torch.Tensor, dim: int):
    d = dim if dim >= 0 else x.dim() + dim
    xt = x if d == x.dim() - 1 else x.transpose(d, -1)
    return d, xt.contiguous()


class mean_fn(Function):
    """epsilon rule for `x.mean(dim)` (reference lxt/explicit/functional.py:539-585): R_in = x * R_out / (sum(x) + eps)"""

    @staticmethod
    def forward(ctx, x, dim, keepdim, epsilon=1e-6):
        if not x.is_cuda or x.dtype not in (torch.float32, torch.bfloat16):
            raise LrpError("mean: CUDA bf16/fp32 tensors only (no CPU fallback)")
        d, xt = _last(x, dim)
        n = xt.shape[-1]
        xs = _rowsum(xt.reshape(-1, n)).view(xt.shape[:-1])            # fp32 sums, one per reduced row
        y = ops.scale(xs, 1.0 / n).to(x.dtype).unsqueeze(-1)
        y = y if d == x.dim() - 1 else y.transpose(d, -1)
        ctx.save_for_backward(x, xs)
        ctx.epsilon, ctx.dim, ctx.keepdim = epsilon, d, keepdim
        return y if keepdim else y.squeeze(d)

    @staticmethod
    @conservation_check_wrap
    def backward(ctx, *out_relevance):
        x, xs = ctx.saved_tensors
        d = ctx.dim
        R = out_relevance[0] if ctx.keepdim else out_relevance[0].unsqueeze(d)
        Rt = (R if d == x.dim() - 1 else R.transpose(d, -1)).contiguous().squeeze(-1)
        s = ops.eps_div(Rt.float(), xs, ctx.epsilon)                    # R / (sum + eps) per reduced row
        _, xt = _last(x, d)
        rel = ops.mul(xt, s.to(x.dtype).unsqueeze(-1).expand(xt.shape))
        rel = rel if d == x.dim() - 1 else rel.transpose(d, -1)
        return rel, None, None, None


class layer_norm_grad_fn(Function):
    """LayerNorm with the std detached, epsilon rule on the whole (reference functional.py:588-635): forward on the detached-std
    LayerNorm kernel, backward R / (y + eps) -> the same kernel's backward (g w rstd - mean(g w rstd)) -> * x."""

    @staticmethod
    def forward(ctx, x, weight, bias, variance_epsilon, epsilon=1e-6):
        if not x.is_cuda or x.dtype not in (torch.float32, torch.bfloat16):
            raise LrpError("layer_norm: CUDA bf16/fp32 tensors only (no CPU fallback)")
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        w = None if weight is None else weight.detach().to(x.dtype).contiguous()
        b = None if bias is None else bias.detach().to(x.dtype).contiguous()
        y, _, rstd = ops.layernorm_fwd(x2, w, b, variance_epsilon)
        ctx.save_for_backward(x, y, rstd)
        ctx.w, ctx.epsilon = w, epsilon
        return y.view(x.shape)

    @staticmethod
    @conservation_check_wrap
    def backward(ctx, *out_relevance):
        x, y, rstd = ctx.saved_tensors
        R = out_relevance[0].reshape(y.shape).to(y.dtype).contiguous()
        g = ops.layernorm_bwd(ops.eps_div(R, y, ctx.epsilon), ctx.w, rstd)
        return ops.mul(g.view(x.shape), x), None, None, None, None


class normalize_identity_fn(Function):
    """identity rule on F.normalize (reference functional.py:638-664): forward x / max(||x||_p, eps), relevance passes unchanged"""

    @staticmethod
    def forward(ctx, input, p, dim, eps):
        if not input.is_cuda or input.dtype not in (torch.float32, torch.bfloat16):
            raise LrpError("normalize: CUDA bf16/fp32 tensors only (no CPU fallback)")
        if p != 2.0:
            return F.normalize(input, p=p, dim=dim, eps=eps)
        d, xt = _last(input, dim)
        n = xt.shape[-1]
        x2 = xt.reshape(-1, n)
        inv = 1.0 / ops.gxi_reduce(x2, x2).sqrt_().clamp_min_(eps)        # [rows] fp32: a handful of scalars per row
        y = ops.mul(x2, inv.to(input.dtype).unsqueeze(-1).expand(x2.shape)).view(xt.shape)
        return y if d == input.dim() - 1 else y.transpose(d, -1)

    @staticmethod
    @conservation_check_wrap
    def backward(ctx, *out_relevance):
        return out_relevance + (None, None, None)


def mean(x, dim, keep_dim, epsilon=1e-6):
    """epsilon-LRP for the mean operation."""
    return mean_fn.apply(x, dim, keep_dim, epsilon)


def layer_norm(hidden_states, weight, bias, variance_epsilon):
    """identity rule on 1/std (detached) and on the weight, epsilon rule on (x - mean): standard nn.LayerNorm (AttnLRP Prop. 3.4)."""
    return layer_norm_grad_fn.apply(hidden_states, weight, bias, variance_epsilon)


def _layer_norm_slower(hidden_states, weight, bias, variance_epsilon):
    """the same LayerNorm composed from the primitive rules (reference functional.py:201-237), kept for the reference's own
    cross-check test (tests/test_functional.py:132-160)"""
    x_mean = mean(hidden_states, -1, keep_dim=True)
    var = ((hidden_states - x_mean) ** 2).mean(dim=-1, keepdim=True)
    std = (var + variance_epsilon).sqrt().detach()
    y = add2(hidden_states, mul2(x_mean, -1))
    y = mul2(y, 1 / std)
    y = mul2(y, weight)
    return add2(y, bias)


def normalize(input, p=2.0, dim=1, eps=1e-12, out=None):
    """identity rule on torch.nn.functional.normalize (AttnLRP Prop. 3.4)."""
    assert out is None, "out parameter is not supported"
    return normalize_identity_fn.apply(input, p, dim, eps)


def add2(input_a, input_b, inplace=False, epsilon=1e-8):
    """epsilon-LRP for a + b (AttnLRP Eq. 8)."""
    return add2_tensors_fn.apply(input_a, input_b, inplace, epsilon)


def softmax(input, dim, dtype=None, temperature=1.0, inplace=False):
    """Deep-Taylor-with-bias rule for softmax (AttnLRP Prop. 3.1)."""
    return softmax_fn.apply(input, dim, dtype, temperature, inplace)


def linear_epsilon(input, weight, bias=None, epsilon=1e-6):
    """epsilon-LRP for nn.functional.linear (AttnLRP Eq. 8) — the fused single-launch sm_100a kernel."""
    return linear_epsilon_fn.apply(input, weight, bias, epsilon)


def matmul(input_a, input_b, inplace=False, epsilon=1e-8):
    """epsilon + uniform rule for torch.matmul (AttnLRP Prop. 3.3)."""
    return matmul_fn.apply(input_a, input_b, inplace, epsilon)


def flash_attention(query, key, value, scale=None, causal=True, window=0, epsilon=1e-6):
    """soft-max attention under the explicit rules (eps + uniform matmuls, Deep-Taylor soft-max) with flash kernels: no [B,H,S,S]
    tensor in HBM.  Extension over the reference API (its explicit models materialise the scores)."""
    return flash_attention_fn.apply(query, key, value, scale, causal, window, epsilon)


def rms_norm_identity(hidden_states, weight, variance_epsilon):
    """identity rule for RMSNorm (AttnLRP Prop. 3.4)."""
    return rms_norm_identity_fn.apply(hidden_states, weight, variance_epsilon)


def mul2(input_a, input_b, inplace=False):
    """uniform rule for a * b (AttnLRP Prop. 3.2)."""
    return mul2_fn.apply(input_a, input_b, inplace)
