"""Module-wrapping rules — drop-in for `lxt.explicit.rules` (reference lxt/explicit/rules.py:8-417).

`EpsilonRule` / `UniformEpsilonRule` are generic (any differentiable module): the VJP is taken with
`torch.autograd.grad` exactly as the reference does (rules.py:212-222), while the rule arithmetic
(R/(out+eps), /n_inputs, * input) runs in liblrp_b200.so kernels.  `EpsilonRule` around an `nn.Linear`
short-circuits to the fused single-launch tcgen05 kernel (`linear_epsilon`).
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.autograd import Function

from .. import ops
from . import functional as lf
from .functional import conservation_check_wrap


class WrapModule(nn.Module):
    """base class: holds the wrapped module (reference rules.py:8-16)"""

    def __init__(self, module):
        super().__init__()
        self.module = module


class identity_fn(Function):
    @staticmethod
    def forward(ctx, fn, input):
        return fn(input)

    @staticmethod
    @conservation_check_wrap
    def backward(ctx, *out_relevance):
        return (None,) + out_relevance


class stop_relevance_fn(Function):
    @staticmethod
    def forward(ctx, fn, input):
        return fn(input)

    @staticmethod
    @conservation_check_wrap
    def backward(ctx, *out_relevance):
        return None, None


class IdentityRule(WrapModule):
    """relevance passes through unchanged (AttnLRP Eq. 9)"""

    def forward(self, input):
        return identity_fn.apply(self.module, input)


def identity(fn, input):
    return identity_fn.apply(fn, input)


class StopRelevanceRule(WrapModule):
    """no relevance to the input (CP-LRP)"""

    def forward(self, input):
        return stop_relevance_fn.apply(self.module, input)


class epsilon_lrp_fn(Function):
    n_div = False  # UniformEpsilon divides the normalised relevance by the number of saved inputs

    @staticmethod
    def forward(ctx, fn, epsilon, *inputs):
        requires_grads = [bool(inp.requires_grad) for inp in inputs]
        if not any(requires_grads):
            return fn(*inputs)  # nothing to explain / first pass of re-entrant checkpointing (reference rules.py:192-195)
        inputs = tuple(inp.detach().requires_grad_() if inp.requires_grad else inp for inp in inputs)
        with torch.enable_grad():
            outputs = fn(*inputs)
        ctx.epsilon, ctx.requires_grads = epsilon, requires_grads
        ctx.save_for_backward(*[inp for inp, r in zip(inputs, requires_grads) if r], outputs)
        return outputs.detach()

    @classmethod
    def _backward(cls, ctx, out_relevance, uniform):
        inputs, outputs = ctx.saved_tensors[:-1], ctx.saved_tensors[-1]
        s = ops.eps_div(out_relevance[0].to(outputs.dtype), outputs.detach(), ctx.epsilon)
        if uniform:
            s = ops.scale(s, 1.0 / len(inputs))
        grads = torch.autograd.grad(outputs, inputs, s)
        rel = iter(ops.mul(g, x.detach()) for g, x in zip(grads, inputs))
        return (None, None) + tuple(next(rel) if r else None for r in ctx.requires_grads)

    @staticmethod
    @conservation_check_wrap
    def backward(ctx, *out_relevance):
        return epsilon_lrp_fn._backward(ctx, out_relevance, False)


class uniform_epsilon_lrp_fn(epsilon_lrp_fn):
    @staticmethod
    @conservation_check_wrap
    def backward(ctx, *out_relevance):
        return epsilon_lrp_fn._backward(ctx, out_relevance, True)


def epsilon_lrp(fn, epsilon, *inputs):
    return epsilon_lrp_fn.apply(fn, epsilon, *inputs)


class EpsilonRule(WrapModule):
    """Gradient x Input / epsilon-LRP around a module (AttnLRP Eq. 4-5, 8)."""

    def __init__(self, module, epsilon=1e-8):
        super().__init__(module)
        self.epsilon = epsilon

    def forward(self, *inputs):
        m = self.module
        if (isinstance(m, nn.Linear) and len(inputs) == 1 and inputs[0].is_cuda and m.in_features % 8 == 0
                and m.out_features % 8 == 0):
            return lf.linear_epsilon(inputs[0], m.weight, m.bias, self.epsilon)
        return epsilon_lrp_fn.apply(m, self.epsilon, *inputs)


class UniformEpsilonRule(WrapModule):
    """epsilon rule followed by the uniform split over the inputs (AttnLRP §3.3.2, matmul)."""

    def __init__(self, module, epsilon=1e-6):
        super().__init__(module)
        self.epsilon = epsilon

    def forward(self, *inputs):
        return uniform_epsilon_lrp_fn.apply(self.module, self.epsilon, *inputs)


class taylor_decomposition_fn(Function):
    """Generalised Taylor decomposition without bias (AttnLRP Eq. 4-5; reference lxt/explicit/rules.py:317-372):
    output' = J(ref) . inputs (jvp at the reference point), R_i = inputs_i * vjp(ref)(R / (output' + eps)).
    The reference's `bias=True` branch reads an undefined variable (rules.py:354-360) and cannot run; it is refused here."""

    @staticmethod
    def forward(ctx, fn, ref, bias, distribute_bias, *inputs):
        if bias:
            raise NotImplementedError("TaylorDecompositionRule(bias=True) is unusable in the reference (undefined `output`, "
                                      "lxt/explicit/rules.py:354-360); only bias=False is provided")
        output = fn(*inputs)
        ctx.save_for_backward(*inputs)
        ctx.fn, ctx.ref = fn, ref
        return output

    @staticmethod
    @conservation_check_wrap
    def backward(ctx, *out_relevance):
        from torch.func import jvp, vjp
        inputs = ctx.saved_tensors
        ref = tuple(ctx.ref)
        _, jv = jvp(ctx.fn, ref, tuple(inputs))
        normed = ops.eps_div(out_relevance[0].to(jv.dtype), jv, 1e-6)
        _, vjpfunc = vjp(ctx.fn, *ref)
        grads = vjpfunc(normed)
        return (None, None, None, None) + tuple(ops.mul(g, x) for g, x in zip(grads, inputs))


class TaylorDecompositionRule(WrapModule):
    """Taylor decomposition at a reference point for any differentiable module (reference rules.py:285-314)."""

    def __init__(self, module, ref=0, bias=False, distribute_bias=None):
        super().__init__(module)
        self.ref, self.bias, self.distribute_bias = ref, bias, distribute_bias

    def forward(self, *inputs):
        return taylor_decomposition_fn.apply(self.module, self.ref, self.bias, self.distribute_bias, *inputs)


class uniform_rule_fn(Function):
    @staticmethod
    def forward(ctx, fn, *inputs):
        ctx.requires_grads = [bool(isinstance(i, torch.Tensor) and i.requires_grad) for i in inputs]
        return fn(*inputs)

    @staticmethod
    @conservation_check_wrap
    def backward(ctx, *out_relevance):
        n = max(1, sum(ctx.requires_grads))
        r = ops.scale(out_relevance[0], 1.0 / n)
        return (None,) + tuple(r if req else None for req in ctx.requires_grads)


class UniformRule(WrapModule):
    """relevance split uniformly over the inputs that require grad (AttnLRP Prop. 3.2; reference rules.py:391-417)"""

    def forward(self, *inputs):
        return uniform_rule_fn.apply(self.module, *inputs)
