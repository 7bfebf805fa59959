"""Drop-in for `lxt.explicit.special` (reference lxt/explicit/special.py:34-140): CP-LRP multi-head attention in relevance space.

Rule: no relevance through the soft-max — q, k and the attention probabilities are constants; the product P·V is an epsilon rule in V
only (`rules.epsilon_lrp(torch.matmul, 1e-6, attention.detach(), v)`, special.py:124): R_V = V * (P^T (R_Y / (Y + 1e-6))).
B200 form: one flash forward; the backward is the flash backward with the gradient divisors (0, 0, 1) applied to R_Y / (Y + eps),
times V — the probabilities are never written to HBM unless the caller asks for them (`need_weights=True`).
"""
from __future__ import annotations

import math

import torch
from torch.autograd import Function

from .. import ops
from .._capi import LrpError


class _CPAttentionFn(Function):
    """y = softmax(q k^T * scale + key mask) v with q, k constant; backward: relevance of y -> relevance of v (epsilon rule)"""

    @staticmethod
    def forward(ctx, q, k, v, scale, kv_range, epsilon):
        o, lse = ops.attn_fwd(q, k, v, scale, causal=False, window=0, kv_range=kv_range)
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.cfg = (scale, kv_range, epsilon)
        return o

    @staticmethod
    def backward(ctx, r_out):
        q, k, v, o, lse = ctx.saved_tensors
        scale, kv_range, epsilon = ctx.cfg
        g = ops.eps_div(r_out.to(o.dtype).contiguous(), o, epsilon)                        # R / (Y + eps)
        _, _, dv = ops.attn_bwd(q, k, v, o, g, lse, scale, causal=False, window=0, q_div=0.0, k_div=0.0, v_div=1.0, kv_range=kv_range)
        return None, None, ops.mul(dv, v), None, None, None


def _project(x2: torch.Tensor, w: torch.Tensor, b) -> torch.Tensor:
    """x2 [T,E] @ w[N,E]^T + b on the tcgen05 GEMM; fp32 operands keep fp32 accuracy (two-term bf16 split), bf16 stays bf16"""
    if w.shape[0] % 8 or w.shape[1] % 8:
        raise LrpError(f"multi_head_attention_cp: projection widths must be multiples of 8 (got {tuple(w.shape)})")
    x2 = x2.contiguous()
    w = w.detach().to(x2.dtype).contiguous()
    out = torch.empty((x2.shape[0], w.shape[0]), dtype=x2.dtype, device=x2.device)
    ops.linear_fwd(x2, w, out, bias=None if b is None else b.detach().float().contiguous())
    return out


def _kv_range_from_key_padding(mask: torch.Tensor, B: int, S: int):
    """bool / float key_padding_mask [B,S] (True or -inf = ignore) -> int32 [B,2] valid key range; the flash kernels take one
    contiguous range per sequence (left or right padding), anything else is refused"""
    m = mask if mask.dtype == torch.bool else (mask != 0)
    if tuple(m.shape) != (B, S):
        raise LrpError(f"key_padding_mask: expected shape {(B, S)}, got {tuple(m.shape)}")
    valid = ~m
    n = valid.sum(1)
    pos = torch.arange(S, device=m.device)
    lo = torch.where(valid, pos, S).min(1).values.clamp(max=S)
    hi = lo + n
    contiguous = (valid == ((pos[None] >= lo[:, None]) & (pos[None] < hi[:, None]))).all()
    if not bool(contiguous) or bool((n == 0).any()):
        raise NotImplementedError("key_padding_mask: only one contiguous block of valid keys per sequence is supported")
    return torch.stack([lo, hi], 1).to(torch.int32).contiguous()


def multi_head_attention_cp(query, key, value, batch_first, num_heads, head_dim, q_proj_weight, bias_q, k_proj_weight, bias_k, v_proj, out_proj,
                            key_padding_mask=None, need_weights=True, attn_mask=None, average_attn_weights=True):
    """Same signature as the reference (special.py:34-36).  query / key / value [SeqLen, Batch, Embed] (or batch-first)."""
    if attn_mask is not None:
        raise NotImplementedError("multi_head_attention_cp: attn_mask is not supported by the flash kernels (key_padding_mask is)")
    if not query.is_cuda:
        raise LrpError("multi_head_attention_cp: CUDA tensors only (no CPU fallback)")
    if head_dim not in (64, 128, 256):
        raise LrpError(f"multi_head_attention_cp: head_dim must be 64, 128 or 256 (got {head_dim})")
    if batch_first is False:
        query, key, value = query.transpose(0, 1), key.transpose(0, 1), value.transpose(0, 1)
    B, Sq, E = query.shape
    Sk = value.shape[1]
    if Sq != Sk:
        raise NotImplementedError("multi_head_attention_cp: query and key lengths must match (self-attention shapes)")
    dt = query.dtype if query.dtype in (torch.bfloat16, torch.float32) else torch.float32
    with torch.no_grad():                                                              # q, k carry no relevance (special.py:100-102)
        q = _project(query.reshape(-1, E).to(dt), q_proj_weight, bias_q).view(B, Sq, num_heads, head_dim)
        k = _project(key.reshape(-1, E).to(dt), k_proj_weight, bias_k).view(B, Sk, num_heads, head_dim)
    v = v_proj(value).to(dt).view(B, Sk, num_heads, head_dim).contiguous()
    scale = 1.0 / math.sqrt(head_dim)
    kv_range = None if key_padding_mask is None else _kv_range_from_key_padding(key_padding_mask, B, Sk)
    y = _CPAttentionFn.apply(q, k, v, scale, kv_range, 1e-6)                            # [B,S,H,D]
    out = out_proj(y.reshape(B, Sq, E).to(query.dtype))
    if batch_first is False:
        out = out.transpose(0, 1)
    if not need_weights:
        return out, None
    with torch.no_grad():                                                              # probabilities only on request (special.py:134-137)
        Skp = (Sk + 7) // 8 * 8                                                        # the batched GEMM wants N in multiples of 8
        q3 = q.permute(0, 2, 1, 3).reshape(B * num_heads, Sq, head_dim).contiguous()
        k3 = torch.zeros((B * num_heads, Skp, head_dim), dtype=k.dtype, device=k.device)
        k3[:, :Sk] = k.permute(0, 2, 1, 3).reshape(B * num_heads, Sk, head_dim)
        logits = torch.empty((B * num_heads, Sq, Skp), dtype=torch.float32, device=q.device)
        ops.gemm_batched(q3, k3, logits, b_layout=0)
        logits = ops.scale(logits, scale).view(B, num_heads, Sq, Skp)
        pos = torch.arange(Skp, device=q.device)
        dead = (pos >= Sk)[None].expand(B, Skp)
        if kv_range is not None:
            dead = dead | (pos[None] < kv_range[:, :1]) | (pos[None] >= kv_range[:, 1:])
        logits = logits.masked_fill(dead[:, None, None, :], float("-inf"))
        attention = ops.softmax_fwd(logits.contiguous())[..., :Sk].to(query.dtype)
    return (out, attention.mean(dim=1)) if average_attn_weights else (out, attention)
