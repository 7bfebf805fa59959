"""Drop-in for `lxt.explicit.modules` (reference lxt/explicit/modules.py:13-214): nn.Module wrappers that apply the relevance-space
rules of `lxt_b200.explicit.functional`, and the helpers that build them from stock torch modules.  Same class names, constructor
arguments and `INIT_MODULE_MAPPING` as the reference; the arithmetic is in liblrp_b200.so (see functional.py / special.py)."""
from __future__ import annotations

import inspect

import torch
import torch.nn as nn

from . import functional as lf
from . import special as ls


class SoftmaxDT(nn.Softmax):
    """Deep-Taylor soft-max rule (modules.py:13-22)"""

    def __init__(self, dim: int, dtype=None, temperature=1.0, inplace=False, **kwargs):
        super().__init__(dim)
        self.inplace, self.dtype, self.temperature = inplace, dtype, temperature

    def forward(self, inputs):
        return lf.softmax(inputs, self.dim, self.dtype, self.temperature, self.inplace)


class LinearEpsilon(nn.Linear):
    """epsilon rule on a Linear layer: one fused tcgen05 launch per backward (modules.py:25-32)"""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, device=None, dtype=None, epsilon=1e-6, **kwargs):
        super().__init__(in_features, out_features, bias, device, dtype)
        self.epsilon = epsilon

    def forward(self, inputs):
        return lf.linear_epsilon(inputs, self.weight, self.bias, self.epsilon)


class RMSNormIdentity(nn.Module):
    """identity rule on RMSNorm (modules.py:35-45)"""

    def __init__(self, hidden_size, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.variance_epsilon = eps

    def forward(self, hidden_states):
        return lf.rms_norm_identity(hidden_states, self.weight, self.variance_epsilon)


class LayerNormEpsilon(nn.LayerNorm):
    """epsilon rule on the mean, identity on the variance (modules.py:48-54)"""

    def __init__(self, normalized_shape, eps: float = 0.00001, elementwise_affine: bool = True, bias: bool = True, device=None, dtype=None):
        super().__init__(normalized_shape, eps, elementwise_affine, bias, device, dtype)

    def forward(self, x):
        return lf.layer_norm(x, self.weight, self.bias, self.eps)


class _PlainProjection(nn.Module):
    """weight / bias holder with a plain linear forward: exists so that rules can be attached to the in / out projections of
    MultiheadAttention_CP separately (modules.py:61-84)"""

    def __init__(self, weight, bias):
        super().__init__()
        self.weight, self.bias = weight, bias

    def forward(self, x):
        from ..efficient.patches import _LinearFn, _linear_ok      # the plain Linear of the drop-in path: tcgen05 GEMM forward + dgrad
        if _linear_ok(x, self.weight):
            return _LinearFn.apply(x, self.weight, self.bias)
        return torch.nn.functional.linear(x, self.weight, self.bias)


class LinearInProjection(_PlainProjection):
    pass


class LinearOutProjection(_PlainProjection):
    pass


class MultiheadAttention_CP(nn.Module):
    """CP-LRP attention: relevance flows through the value path only (modules.py:87-124; rule in special.py)"""

    def __init__(self):
        super().__init__()
        self.q_proj_weight = self.k_proj_weight = None
        self.v_proj = LinearInProjection(None, None)
        self.out_proj = LinearOutProjection(None, None)
        self.embed_dim = self.num_heads = self.head_dim = self.batch_first = None
        self.bias_q = self.bias_k = None

    def forward(self, query, key, value, key_padding_mask=None, need_weights=True, attn_mask=None, average_attn_weights=True, is_causal=False):
        assert is_causal is False or is_causal == False  # noqa: E712  (not supported by the reference either)
        return ls.multi_head_attention_cp(query, key, value, self.batch_first, self.num_heads, self.head_dim, self.q_proj_weight, self.bias_q,
                                          self.k_proj_weight, self.bias_k, self.v_proj, self.out_proj, key_padding_mask, need_weights,
                                          attn_mask, average_attn_weights)


def copy_parameters_and_buffers_(original, replacement):
    """share (not clone) the parameters and buffers of `original` with `replacement` (modules.py:127-136)"""
    for name, param in original.named_parameters():
        replacement.register_parameter(name, param)
    for name, buffer in original.named_buffers():
        replacement.register_buffer(name, buffer)


def _ctor_kwargs(original):
    return {a: getattr(original, a) for a in inspect.signature(original.__init__).parameters if hasattr(original, a)}


def initialize_generic(original, replacement):
    """build `replacement` with the constructor arguments found as attributes on `original` (modules.py:139-152)"""
    new = replacement(**_ctor_kwargs(original))
    copy_parameters_and_buffers_(original, new)
    return new


def initialize_bias(original, replacement):
    """as initialize_generic, with `bias` derived from whether the original has one (modules.py:155-170)"""
    kwargs = _ctor_kwargs(original)
    kwargs["bias"] = original.bias is not None
    new = replacement(**kwargs)
    copy_parameters_and_buffers_(original, new)
    return new


def initialize_MHA(original, replacement):
    """build a MultiheadAttention_CP from a torch nn.MultiheadAttention (modules.py:173-205): the packed in-projection is split into views"""
    new = replacement()
    E = original.embed_dim
    if not original._qkv_same_embed_dim:
        new.q_proj_weight, new.k_proj_weight, new.v_proj.weight = original.q_proj_weight, original.k_proj_weight, original.v_proj_weight
    else:
        w = original.in_proj_weight
        new.q_proj_weight, new.k_proj_weight, new.v_proj.weight = w[:E], w[E:2 * E], w[2 * E:3 * E]
    if original.in_proj_bias is not None:
        b = original.in_proj_bias
        new.bias_q, new.bias_k, new.v_proj.bias = b[:E], b[E:2 * E], b[2 * E:3 * E]
    if original.bias_k is not None:
        raise NotImplementedError("add_bias_kv=True is not supported yet.")
    new.out_proj.weight, new.out_proj.bias = original.out_proj.weight, original.out_proj.bias
    new.embed_dim, new.num_heads, new.head_dim, new.batch_first = original.embed_dim, original.num_heads, original.head_dim, original.batch_first
    return new


INIT_MODULE_MAPPING = {
    SoftmaxDT: initialize_generic,
    LinearEpsilon: initialize_bias,
    RMSNormIdentity: initialize_generic,
    LayerNormEpsilon: initialize_bias,
    MultiheadAttention_CP: initialize_MHA,
}
