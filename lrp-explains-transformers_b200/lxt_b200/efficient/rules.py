"""Gradient x Input rules — drop-in for `lxt.efficient.rules` (reference: lxt/efficient/rules.py:19-127), with the
backward arithmetic executed by the sm_100a kernels of liblrp_b200.so.

  identity_rule_implicit(fn, input)   fwd y = fn(x);  bwd g_x = g_y * y/(x + 1e-10)
  divide_gradient(input, factor=2)    fwd identity;   bwd g / factor
  stop_gradient(input)                detach
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import ops

_KNOWN_ACTS = {}


def _act_code(fn):
    """map well-known activations to the fused kernels' activation codes (None = generic callable)"""
    F = torch.nn.functional
    if fn is F.silu or isinstance(fn, torch.nn.SiLU) or getattr(fn, "__name__", "") == "silu":
        return ops.ACT_SILU
    if isinstance(fn, torch.nn.GELU):
        return ops.ACT_GELU_TANH if fn.approximate == "tanh" else ops.ACT_GELU_ERF
    cls = type(fn).__name__
    if cls in ("SiLUActivation",):
        return ops.ACT_SILU
    if cls in ("GELUTanh", "PytorchGELUTanh", "NewGELUActivation", "FastGELUActivation"):
        return ops.ACT_GELU_TANH
    if cls in ("GELUActivation",):
        return ops.ACT_GELU_ERF
    if fn is F.gelu:
        return ops.ACT_GELU_ERF
    return None


class identity_rule_implicit_fn(Function):
    @staticmethod
    def forward(ctx, fn, input, epsilon=1e-10):
        code = _act_code(fn)
        x = input.contiguous()
        if code is not None:
            out = ops.act_identity_fwd(x, code)
            ctx.code = code
            if input.requires_grad:
                ctx.save_for_backward(x)
        else:
            out = fn(input)
            ctx.code = None
            if input.requires_grad:
                ctx.save_for_backward(x, out)
        return out

    @staticmethod
    def backward(ctx, *out_relevance):
        g = out_relevance[0].contiguous()
        if ctx.code is not None:
            (x,) = ctx.saved_tensors
            return None, ops.act_identity_bwd(g, x, ctx.code), None
        x, y = ctx.saved_tensors
        return None, ops.identity_rule_bwd(g, x, y), None


class divide_gradient_fn(Function):
    @staticmethod
    def forward(ctx, input, factor=2):
        ctx.factor = factor
        return input

    @staticmethod
    def backward(ctx, *out_relevance):
        return ops.scale(out_relevance[0], 1.0 / ctx.factor), None


def identity_rule_implicit(fn, input):
    """identity rule (AttnLRP Eq. 9) on an element-wise non-linearity, Gradient x Input form."""
    return identity_rule_implicit_fn.apply(fn, input)


def divide_gradient(input, factor=2):
    """uniform rule (AttnLRP Eq. 7) after a matmul / element-wise product, Gradient x Input form."""
    return divide_gradient_fn.apply(input, factor)


def stop_gradient(input):
    """CP-LRP: no relevance flows through this tensor."""
    return input.detach()
