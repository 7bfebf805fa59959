"""Forward replacements and patch plumbing — drop-in for `lxt.efficient.patches`
(reference: lxt/efficient/patches.py:22-280).  Same names, same signatures, same return conventions; the
arithmetic runs in the sm_100a kernels of liblrp_b200.so through `torch.autograd.Function`s, so a patched
HuggingFace model keeps running unchanged Python (`loss.backward()`, `x * x.grad`).

Tensors must be CUDA bf16 (fp32 where noted): there is no CPU or eager fallback — unsupported inputs raise.
"""
from __future__ import annotations

import math
from warnings import warn

import torch
from torch.autograd import Function

from .. import ops
from .._capi import LrpError
from .rules import _act_code, divide_gradient, identity_rule_implicit, stop_gradient  # noqa: F401

# ---------------------------------------------------------------------------------------------------------
# patch plumbing (host logic; reference lxt/efficient/patches.py:22-104)
# ---------------------------------------------------------------------------------------------------------


def check_already_patched(target_fn, new_fn):
    """True (with a warning) when `target_fn` already lives in the module `new_fn` comes from — the reference's
    re-patch guard is a `__module__` string comparison (patches.py:40), kept bug-compatible."""
    if getattr(target_fn, "__module__", None) != getattr(new_fn, "__module__", object()):
        return False
    warn(f"{getattr(target_fn, '__name__', target_fn)} already patched.")
    return True


def patch_method(fn, module, method_name="forward", keep_original=False):
    """Replace `module.<method_name>` by `fn` (class-level, process-global).  Returns True when patched."""
    current = getattr(module, method_name)
    if check_already_patched(current, fn):
        return False
    if keep_original:
        setattr(module, f"original_{method_name}", current)
    setattr(module, method_name, fn)
    return True


def replace_module(patched_module, original_module):
    """Copy every public attribute of `patched_module` onto `original_module`."""
    if original_module == patched_module:
        return False
    for name in dir(patched_module):
        if name.startswith("__"):
            continue
        setattr(original_module, name, getattr(patched_module, name))
    return True


# ---------------------------------------------------------------------------------------------------------
# autograd bridges over the C-ABI kernels
# ---------------------------------------------------------------------------------------------------------


def _as_2d(x: torch.Tensor) -> torch.Tensor:
    x2 = x.reshape(-1, x.shape[-1])
    return x2 if x2.is_contiguous() else x2.contiguous()


def _need_bf16_cuda(x: torch.Tensor, what: str) -> None:
    """bf16 = production path; fp32 = validation precision (same kernels instantiated for fp32 activations, GEMMs as two-term
    bf16 splits on the tcgen05 kernel, fp32 CUDA-core attention).  Anything else raises: there is no fallback."""
    if not (x.is_cuda and x.dtype in (torch.bfloat16, torch.float32)):
        raise LrpError(f"{what}: the B200 path takes CUDA bfloat16 (or fp32, validation precision) tensors (got {x.device}, {x.dtype}); "
                       "no fallback exists")


class _RMSNormIdentityFn(Function):
    """y = (x * rsqrt(mean(x^2)+eps)) * (w + w_offset); backward with the variance detached (identity rule)."""

    @staticmethod
    def forward(ctx, x, weight, eps, w_offset):
        _need_bf16_cuda(x, "rms_norm_forward")
        x2 = _as_2d(x)
        if x.dtype == torch.float32:
            # validation precision: x_hat from the fp32 instantiation of the same kernel (zero weight, offset 1), then the fp32
            # scale vector (w + w_offset) applied by the element-wise kernel — no rounding of an fp32 model's norm weights
            wz = torch.zeros(x2.shape[1], dtype=torch.bfloat16, device=x.device)
            xh, rstd = ops.rmsnorm_fwd(x2, wz, eps, w_offset=1.0, out_dtype=torch.float32)
            wf = (weight.detach().float() + w_offset).contiguous()
            ctx.save_for_backward(wz, rstd, wf)
            ctx.w_offset = None
            return ops.mul(xh, wf).view(x.shape)
        w = weight.detach().to(torch.bfloat16).contiguous()
        y, rstd = ops.rmsnorm_fwd(x2, w, eps, w_offset=w_offset)
        ctx.save_for_backward(w, rstd)
        ctx.w_offset = w_offset
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, gy):
        if ctx.w_offset is None:
            wz, rstd, wf = ctx.saved_tensors
            g2 = ops.mul(_as_2d(gy), wf)
            return ops.rmsnorm_bwd(g2, wz, rstd, w_offset=1.0, out_dtype=torch.float32).view(gy.shape), None, None, None
        w, rstd = ctx.saved_tensors
        gx = ops.rmsnorm_bwd(_as_2d(gy), w, rstd, w_offset=ctx.w_offset)
        return gx.view(gy.shape), None, None, None


class _LayerNormIdentityFn(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        if not x.is_cuda or x.dtype not in (torch.bfloat16, torch.float32):
            raise LrpError("layer_norm_forward: CUDA bf16/fp32 tensors only")
        x2 = _as_2d(x)
        w = None if weight is None else weight.detach().to(x.dtype).contiguous()
        b = None if bias is None else bias.detach().to(x.dtype).contiguous()
        y, _, rstd = ops.layernorm_fwd(x2, w, b, eps)
        ctx.w = w
        ctx.save_for_backward(rstd)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, gy):
        (rstd,) = ctx.saved_tensors
        return ops.layernorm_bwd(_as_2d(gy), ctx.w, rstd).view(gy.shape), None, None, None


class _LinearFn(Function):
    """y = x W^T + b on the tcgen05 GEMM; backward g_x = g_y W (weights are frozen on this path)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x2 = _as_2d(x)
        w = weight.detach()
        w = w if w.is_contiguous() else w.contiguous()
        ctx.split = _wsplit(weight, x)
        y = torch.empty((x2.shape[0], w.shape[0]), dtype=x.dtype, device=x.device)
        ops.linear_fwd(x2, w, y, bias=None if bias is None else bias.detach().float().contiguous(), b_split=ctx.split)
        ctx.save_for_backward(w)
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, gy):
        (w,) = ctx.saved_tensors
        g2 = _as_2d(gy)
        gx = torch.empty((g2.shape[0], w.shape[1]), dtype=gy.dtype, device=gy.device)
        ops.linear_dgrad(g2, w, gx, b_split=ctx.split)
        return gx.view(*gy.shape[:-1], w.shape[1]), None, None


def _wsplit(weight: torch.Tensor, x: torch.Tensor):
    """validation precision with fp32 weights: the (hi, lo) bf16 split memoised on the Parameter object; None otherwise"""
    if x.dtype == torch.float32 and weight.dtype == torch.float32 and weight.is_contiguous():
        return ops.weight_split(weight)
    return None


def _linear_ok(x: torch.Tensor, weight: torch.Tensor) -> bool:
    """bf16 x bf16 (production) or fp32 x {fp32, bf16} (validation precision: split GEMM); feature counts multiples of 8"""
    if not (x.is_cuda and weight.is_cuda and weight.shape[0] % 8 == 0 and weight.shape[1] % 8 == 0 and x.numel() > 0):
        return False
    if x.dtype == torch.bfloat16:
        return weight.dtype == torch.bfloat16
    return x.dtype == torch.float32 and weight.dtype in (torch.float32, torch.bfloat16)


def _functorch_wrapped(x: torch.Tensor) -> bool:
    """the autograd bridges are plain autograd.Functions (no functorch rules): transformed tensors take the stock path"""
    f = getattr(torch._C, "_functorch", None)
    return f is not None and hasattr(f, "is_functorch_wrapped_tensor") and f.is_functorch_wrapped_tensor(x)


def _linear(mod, x):
    """run an nn.Linear through the B200 GEMM"""
    if not _linear_ok(x, mod.weight):
        raise LrpError(f"linear: needs CUDA bf16 (or fp32) with in/out features multiples of 8 (got {x.dtype}, {tuple(mod.weight.shape)})")
    return _LinearFn.apply(x, mod.weight, mod.bias)


class _GatedMLPFn(Function):
    """down( divide_gradient_2( identity_rule(act)(gate(x)) * up(x) ) ) as five launches forward, five backward."""

    @staticmethod
    def forward(ctx, x, wg, wu, wd, bg, bu, bd, act, cp=False):
        _need_bf16_cuda(x, "gated_mlp_forward")
        x2 = _as_2d(x)
        T, I = x2.shape[0], wg.shape[0]
        ctx.splits = sg, su, sd = [_wsplit(w, x) for w in (wg, wu, wd)]
        wg, wu, wd = (w.detach().contiguous() for w in (wg, wu, wd))
        f32 = lambda b: None if b is None else b.detach().float().contiguous()
        gu = torch.empty((T, 2 * I), dtype=x.dtype, device=x.device)
        ops.linear_fwd(x2, wg, gu[:, :I], bias=f32(bg), b_split=sg)
        ops.linear_fwd(x2, wu, gu[:, I:], bias=f32(bu), b_split=su)
        a = ops.gated_act_fwd(gu, act)
        y = torch.empty((T, wd.shape[0]), dtype=x.dtype, device=x.device)
        ops.linear_fwd(a, wd, y, bias=f32(bd), b_split=sd)
        ctx.save_for_backward(gu, wg, wu, wd)
        ctx.act, ctx.cp = act, cp
        return y.view(*x.shape[:-1], wd.shape[0])

    @staticmethod
    def backward(ctx, gy):
        gu, wg, wu, wd = ctx.saved_tensors
        T, I = gu.shape[0], wg.shape[0]
        g2 = _as_2d(gy)
        sg, su, sd = ctx.splits
        ga = torch.empty((T, I), dtype=gu.dtype, device=gy.device)
        ops.linear_dgrad(g2, wd, ga, b_split=sd)
        ggu = ops.gated_act_bwd(ga, gu, ctx.act, cp=ctx.cp)
        acc = torch.empty((T, wg.shape[1]), dtype=torch.float32, device=gy.device)
        ops.linear_dgrad(ggu[:, :I], wg, acc, b_split=sg)
        if gu.dtype == torch.float32:   # validation precision: the fp32 accumulator is the result
            ops.linear_dgrad(ggu[:, I:], wu, acc, resid=acc, b_split=su)
            gx = acc
        else:
            gx = torch.empty((T, wg.shape[1]), dtype=torch.bfloat16, device=gy.device)
            ops.linear_dgrad(ggu[:, I:], wu, acc, resid=acc, shadow=gx)
        return gx.view(*gy.shape[:-1], wg.shape[1]), None, None, None, None, None, None, None, None


class _FlashAttnLRPFn(Function):
    """soft-max attention with the AttnLRP backward (dQ/q_div, dK/k_div, dV/v_div); q,k,v in HF layout [B,H,S,D]."""

    @staticmethod
    def forward(ctx, q, k, v, scale, causal, window, q_div, k_div, v_div, kv_range=None):
        for t, n in ((q, "query"), (k, "key"), (v, "value")):
            _need_bf16_cuda(t, f"attention {n}")
        qs, ks, vs = (t.transpose(1, 2).contiguous() for t in (q, k, v))  # [B,S,H,D]; no copy if already so in memory
        o, lse = ops.attn_fwd(qs, ks, vs, scale, causal=causal, window=window, kv_range=kv_range)
        ctx.save_for_backward(qs, ks, vs, o, lse)
        ctx.kv_range = kv_range
        ctx.cfg = (scale, causal, window, q_div, k_div, v_div)
        return o  # [B,S,H,D] — what HF attention functions return after their transpose(1,2).contiguous()

    @staticmethod
    def backward(ctx, d_o):
        qs, ks, vs, o, lse = ctx.saved_tensors
        scale, causal, window, q_div, k_div, v_div = ctx.cfg
        dq, dk, dv = ops.attn_bwd(qs, ks, vs, o, d_o.contiguous(), lse, scale, causal=causal, window=window, q_div=q_div,
                                  k_div=k_div, v_div=v_div, kv_range=ctx.kv_range)
        return dq.transpose(1, 2), dk.transpose(1, 2), dv.transpose(1, 2), None, None, None, None, None, None, None


class _PackedSelfAttnFn(Function):
    """self-attention on a packed projection `qkv [B,S,3*H*D]` (q | k | v): the flash kernels read and write strided
    views of the packed buffers, so the in-projection GEMM output and its gradient are never split or copied."""

    @staticmethod
    def forward(ctx, qkv, H, D, scale, causal, q_div, k_div, v_div):
        _need_bf16_cuda(qkv, "packed self-attention")
        B, S, W = qkv.shape
        qkv = qkv.contiguous()
        q, k, v = (qkv[:, :, i * H * D:(i + 1) * H * D].view(B, S, H, D) for i in range(3))
        o, lse = ops.attn_fwd(q, k, v, scale, causal=causal, window=0)
        ctx.save_for_backward(qkv, o, lse)
        ctx.cfg = (H, D, scale, causal, q_div, k_div, v_div)
        return o.view(B, S, H * D)

    @staticmethod
    def backward(ctx, d_o):
        qkv, o, lse = ctx.saved_tensors
        H, D, scale, causal, q_div, k_div, v_div = ctx.cfg
        B, S, _ = qkv.shape
        q, k, v = (qkv[:, :, i * H * D:(i + 1) * H * D].view(B, S, H, D) for i in range(3))
        g = torch.empty_like(qkv)
        dq, dk, dv = (g[:, :, i * H * D:(i + 1) * H * D].view(B, S, H, D) for i in range(3))
        ops.attn_bwd(q, k, v, o, d_o.contiguous().view(B, S, H, D), lse, scale, causal=causal, window=0, q_div=q_div,
                     k_div=k_div, v_div=v_div, dq=dq, dk=dk, dv=dv)
        return g, None, None, None, None, None, None, None


_KV_RANGE_CACHE = {}


def _kv_range_from_mask(mask, B, S, causal, window):
    """HF hands every attention layer the same 4-D mask `[B or 1, 1, S, S]` (bool: True = attend; float: 0 = attend): causal
    (+ sliding window) AND the 2-D key-padding mask of the batch (transformers masking_utils.sdpa_mask / eager_mask).  The
    kernels take the causal / window part from flags and the padding part as one valid-key range per sequence, so the mask
    is reduced ON THE DEVICE (no host synchronisation) to `kv_range[b] = [first valid key, last valid key + 1)`: key j is
    valid iff some query attends to it.  The result is cached per mask tensor (the same object serves all layers).
    LRP_VERIFY_MASKS=1 re-expands the range and compares it with the mask (one host sync; raises on custom masks)."""
    if mask.dim() != 4 or mask.shape[-1] != S or mask.shape[-2] != S or mask.shape[1] != 1 or mask.shape[0] not in (1, B):
        raise LrpError(f"attention_mask of shape {tuple(mask.shape)} is not supported (expected [B,1,S,S] with S = {S})")
    # memoised per mask OBJECT (weak reference + version; never by address: freed masks' storage is handed out again)
    key = (id(mask), bool(causal), int(window))
    hit = _KV_RANGE_CACHE.get(key)
    if hit is not None and hit[0]() is mask and hit[1] == mask._version:
        return hit[2]
    allowed = mask[:, 0] if mask.dtype == torch.bool else (mask[:, 0] == 0)
    valid = allowed.any(dim=-2)                                   # [B or 1, S]
    idx = torch.arange(S, device=mask.device)
    lo = torch.where(valid, idx, S).amin(-1)
    hi = torch.where(valid, idx + 1, 0).amax(-1)
    kv_range = torch.stack([lo, hi], -1).to(torch.int32).expand(B, 2).contiguous()
    import os
    if os.environ.get("LRP_VERIFY_MASKS", "0") == "1":
        i = idx
        exp = (i[None, None, :] >= kv_range[:, 0, None, None]) & (i[None, None, :] < kv_range[:, 1, None, None])
        if causal:
            exp = exp & (i[None, None, :] <= i[None, :, None])
        if window:
            exp = exp & ((i[None, :, None] - i[None, None, :]) < window)
        got = allowed.expand(B, S, S)
        # rows that attend to nothing (padding queries) are don't-cares
        live = exp.any(-1, keepdim=True)
        if not bool(((got == exp) | ~live).all()):
            raise LrpError("attention_mask is not causal(+window) x contiguous key padding; custom masks are not supported")
    if len(_KV_RANGE_CACHE) > 8:
        _KV_RANGE_CACHE.clear()
    import weakref
    _KV_RANGE_CACHE[key] = (weakref.ref(mask), mask._version, kv_range)
    return kv_range


def _attention_scale(module, query, kwargs):
    """soft-max scale exactly as the wrapped HF function would use it: the `scaling` kwarg, else `module.scaling`, else
    head_dim^-0.5; GPT-2's non-default scaling switches are refused rather than silently dropped."""
    scaling = kwargs.get("scaling")
    if scaling is None:
        scaling = getattr(module, "scaling", None)
    if scaling is None:
        if getattr(module, "scale_attn_weights", True) is False or getattr(module, "scale_attn_by_inverse_layer_idx", False):
            raise LrpError("attention: GPT-2 scale_attn_weights=False / scale_attn_by_inverse_layer_idx are not supported")
        scaling = 1.0 / math.sqrt(query.shape[-1])
    return float(scaling)


def _lrp_attention(module, query, key, value, args, kwargs, q_div, k_div, v_div):
    mask = args[0] if len(args) > 0 else kwargs.get("attention_mask")
    if kwargs.get("softcap") is not None:
        raise LrpError("attention softcap is not supported by the B200 AttnLRP kernel")
    if kwargs.get("head_mask") is not None:
        raise LrpError("attention head_mask is not supported by the B200 AttnLRP kernel")
    window = kwargs.get("sliding_window") or 0
    is_causal = kwargs.get("is_causal")
    if is_causal is None:
        is_causal = getattr(module, "is_causal", True)
    B, _, S, _ = query.shape
    if key.shape[2] != S:
        raise LrpError("AttnLRP needs full-sequence self-attention (use_cache=False)")
    causal = bool(is_causal)
    if window and not causal:
        raise LrpError("a sliding window without causal attention is not supported by the B200 AttnLRP kernel")
    kv_range = None if mask is None else _kv_range_from_mask(mask, B, S, causal, int(window))
    out = _FlashAttnLRPFn.apply(query, key, value, _attention_scale(module, query, kwargs), causal, int(window), q_div, k_div, v_div,
                                kv_range)
    return out, None


# ---------------------------------------------------------------------------------------------------------
# AttnLRP patches (reference lxt/efficient/patches.py:111-220)
# ---------------------------------------------------------------------------------------------------------


def rms_norm_forward(self, hidden_states):
    """identity rule on RMSNorm: variance path detached (reference patches.py:111-123)."""
    return _RMSNormIdentityFn.apply(hidden_states, self.weight, float(getattr(self, "variance_epsilon", getattr(self, "eps", 1e-6))), 0.0)


def gemma_rms_norm_forward(self, x):
    """Gemma-style RMSNorm `(1 + w) * x_hat` with the identity rule (reference lxt/efficient/models/gemma3.py:11-12)."""
    return _RMSNormIdentityFn.apply(x, self.weight, float(self.eps), 1.0)


def layer_norm_forward(self, x):
    """identity rule on LayerNorm: std detached (reference patches.py:126-142)."""
    return _LayerNormIdentityFn.apply(x, self.weight, self.bias, float(self.eps))


def gated_mlp_forward(self, x):
    """identity rule on the activation, uniform rule on the product (reference patches.py:145-157)."""
    act = _act_code(self.act_fn)
    if act is None:
        raise LrpError(f"gated_mlp_forward: unsupported activation {type(self.act_fn).__name__}")
    return _GatedMLPFn.apply(x, self.gate_proj.weight, self.up_proj.weight, self.down_proj.weight, self.gate_proj.bias,
                             self.up_proj.bias, self.down_proj.bias, act)


def _find(mod, *names):
    for n in names:
        if hasattr(mod, n):
            return getattr(mod, n)
    raise AttributeError(f"{type(mod).__name__} has none of {names}")


def mlp_forward(self, x):
    """identity rule on the activation of a plain 2-layer MLP (reference patches.py:159-169)."""
    up, down = _find(self, "up_proj", "c_fc", "fc1"), _find(self, "down_proj", "c_proj", "fc2")
    act = _find(self, "act_fn", "act", "activation_fn")
    if isinstance(up, torch.nn.Linear):
        h = _linear(up, x)
    else:
        h = up(x)
    h = identity_rule_implicit(act, h)
    return _linear(down, h) if isinstance(down, torch.nn.Linear) else down(h)


def linear_forward(self, x):
    """nn.Linear on the tcgen05 GEMM (forward + LRP dgrad).  Not part of the reference's patch map — it leaves
    nn.Linear to cuBLAS — added so that a patched model runs every FLOP of the path on the B200 kernels.
    Inputs the kernel does not take (CPU, odd widths, tensors wrapped by a torch.func transform such as the jvp / vjp of
    `TaylorDecompositionRule`) go through the original forward unchanged."""
    if _linear_ok(x, self.weight) and not _functorch_wrapped(x):
        return _LinearFn.apply(x, self.weight, self.bias)
    return self.original_forward(x)


def wrap_attention_forward(forward_fn):
    """AttnLRP attention: uniform rule on both matmuls = dQ/4, dK/4, dV/2 (reference patches.py:193-203).
    `forward_fn` (the HF attention implementation being replaced) is kept only for introspection."""

    def attention_forward(module, query, key, value, *args, **kwargs):
        return _lrp_attention(module, query, key, value, args, kwargs, 4.0, 4.0, 2.0)

    attention_forward.wrapped = forward_fn
    return attention_forward


def _patch_attention_registry(module, wrapper):
    new_forward = wrapper(module.eager_attention_forward)
    if check_already_patched(module.eager_attention_forward, new_forward):
        return False
    module.eager_attention_forward = new_forward
    for key, value in list(module.ALL_ATTENTION_FUNCTIONS.items()):
        new_forward = wrapper(value)
        if check_already_patched(value, new_forward):
            return False  # registry shared with an already patched family (reference behaviour, patches.py:184-189)
        module.ALL_ATTENTION_FUNCTIONS[key] = new_forward
    return True


def patch_attention(module):
    """Patch `module.eager_attention_forward` and every entry of the shared `ALL_ATTENTION_FUNCTIONS` registry
    (reference patches.py:171-190)."""
    return _patch_attention_registry(module, wrap_attention_forward)


def non_linear_forward(self, x):
    """identity rule on an element-wise activation module (reference patches.py:206-211)."""
    code = _act_code(self)
    return identity_rule_implicit(self if code is not None else self.original_forward, x)


def dropout_forward(self, x):
    """Dropout is the identity on this path (reference patches.py:214-220: lets HF checkpointing run in train())."""
    return x


# ---------------------------------------------------------------------------------------------------------
# CP-LRP patches (reference lxt/efficient/patches.py:228-280)
# ---------------------------------------------------------------------------------------------------------


def cp_wrap_attention_forward(forward_fn):
    """CP-LRP: no relevance through the softmax (q, k detached), plain gradient through v."""

    def cp_attention_forward(module, query, key, value, *args, **kwargs):
        return _lrp_attention(module, query, key, value, args, kwargs, 0.0, 0.0, 1.0)

    cp_attention_forward.wrapped = forward_fn
    return cp_attention_forward


def patch_cp_attention(module):
    return _patch_attention_registry(module, cp_wrap_attention_forward)


def cp_multi_head_attention_forward(self, query, key, value, *args, **kwargs):
    """CP-LRP for torch.nn.MultiheadAttention: q, k detached (reference patches.py:261-269)."""
    return self.original_forward(stop_gradient(query), stop_gradient(key), value, *args, **kwargs)


def b200_cp_multi_head_attention_forward(self, query, key, value, *args, **kwargs):
    """CP-LRP `nn.MultiheadAttention` entirely on the B200 kernels: packed in-projection GEMM -> flash attention with
    q,k detached (dQ = dK = 0, plain dV) -> out-projection GEMM.  Same rule as `cp_multi_head_attention_forward`
    (reference patches.py:261-269); calls that the kernels do not cover (cross-attention, masks, fp32, seq-first
    layout, exotic head sizes) take that reference-equivalent route instead."""
    E, H = self.embed_dim, self.num_heads
    D = E // H
    plain = (query is key and key is value and query.dim() == 3 and self.batch_first and self._qkv_same_embed_dim
             and self.in_proj_weight is not None and query.is_cuda and query.dtype in (torch.bfloat16, torch.float32) and D in (64, 128)
             and query.dtype == self.in_proj_weight.dtype and kwargs.get("need_weights", True) is False
             and kwargs.get("attn_mask") is None and kwargs.get("key_padding_mask") is None and not kwargs.get("is_causal", False)
             and len(args) == 0 and self.bias_k is None and not self.add_zero_attn and (self.dropout == 0.0 or not self.training))
    if not plain:
        return cp_multi_head_attention_forward(self, query, key, value, *args, **kwargs)
    qkv = _LinearFn.apply(query, self.in_proj_weight, self.in_proj_bias)
    o = _PackedSelfAttnFn.apply(qkv, H, D, 1.0 / math.sqrt(D), False, 0.0, 0.0, 1.0)
    out = _LinearFn.apply(o, self.out_proj.weight, self.out_proj.bias)
    return out, None


def cp_gated_mlp_forward(self, x):
    """CP-LRP: no relevance through the gate, no uniform split (reference patches.py:272-280)."""
    act = _act_code(self.act_fn)
    if act is None:
        raise LrpError(f"cp_gated_mlp_forward: unsupported activation {type(self.act_fn).__name__}")
    return _GatedMLPFn.apply(x, self.gate_proj.weight, self.up_proj.weight, self.down_proj.weight, self.gate_proj.bias,
                             self.up_proj.bias, self.down_proj.bias, act, True)


class _MulConstFn(Function):
    """y = x * c with c constant (detached): g_x = g_y * c"""

    @staticmethod
    def forward(ctx, x, c):
        ctx.save_for_backward(c)
        return ops.mul(x, c)

    @staticmethod
    def backward(ctx, gy):
        (c,) = ctx.saved_tensors
        return ops.mul(gy, c), None
