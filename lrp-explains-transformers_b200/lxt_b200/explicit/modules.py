"""Drop-in for `lxt.explicit.modules` (reference lxt/explicit/modules.py:13-214): nn.Module wrappers that apply the relevance-space
rules of `lxt_b200.explicit.functional`, and the helpers that build them from stock torch modules.  Class names, constructor
arguments and `INIT_MODULE_MAPPING` are the reference's (that is the contract); the arithmetic is in liblrp_b200.so (functional.py,
special.py)."""
from __future__ import annotations

import inspect

import torch
import torch.nn as nn

from . import functional as lf
from . import special as ls


class SoftmaxDT(nn.Softmax):
    """Deep-Taylor soft-max rule (modules.py:13-22): `lf.softmax` with the stored dim / dtype / temperature"""

    def __init__(self, dim: int, dtype=None, temperature=1.0, inplace=False, **kwargs):
        nn.Softmax.__init__(self, dim)
        self.dtype, self.temperature, self.inplace = dtype, temperature, inplace

    def forward(self, inputs):
        return lf.softmax_fn.apply(inputs, self.dim, self.dtype, self.temperature, self.inplace)


class LinearEpsilon(nn.Linear):
    """epsilon rule on a Linear layer (modules.py:25-32): forward GEMM, backward ONE fused tcgen05 launch"""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, device=None, dtype=None, epsilon=1e-6, **kwargs):
        nn.Linear.__init__(self, in_features, out_features, bias=bias, device=device, dtype=dtype)
        self.epsilon = epsilon

    def forward(self, inputs):
        return lf.linear_epsilon_fn.apply(inputs, self.weight, self.bias, self.epsilon)


class RMSNormIdentity(nn.Module):
    """identity rule on RMSNorm (modules.py:35-45); the weight starts at one like the layer it stands in for"""

    def __init__(self, hidden_size, eps=1e-6):
        nn.Module.__init__(self)
        self.variance_epsilon = eps
        self.register_parameter("weight", nn.Parameter(torch.ones(hidden_size)))

    def forward(self, hidden_states):
        return lf.rms_norm_identity_fn.apply(hidden_states, self.weight, self.variance_epsilon)


class LayerNormEpsilon(nn.LayerNorm):
    """LayerNorm with the epsilon rule on the mean and the identity rule on the variance (modules.py:48-54)"""

    def __init__(self, normalized_shape, eps: float = 0.00001, elementwise_affine: bool = True, bias: bool = True, device=None, dtype=None):
        nn.LayerNorm.__init__(self, normalized_shape, eps=eps, elementwise_affine=elementwise_affine, bias=bias, device=device, dtype=dtype)

    def forward(self, x):
        return lf.layer_norm(x, self.weight, self.bias, self.eps)


class _Projection(nn.Module):
    """weight / bias holder with a plain linear forward (modules.py:61-84): the in / out projections of MultiheadAttention_CP are
    modules of their own so that rules can be attached to them separately"""

    def __init__(self, weight, bias):
        nn.Module.__init__(self)
        self.weight, self.bias = weight, bias

    def forward(self, x):
        from ..efficient.patches import _LinearFn, _linear_ok      # the plain Linear of the drop-in path: tcgen05 GEMM forward + dgrad
        if _linear_ok(x, self.weight):
            return _LinearFn.apply(x, self.weight, self.bias)
        return nn.functional.linear(x, self.weight, self.bias)


class LinearInProjection(_Projection):
    pass


class LinearOutProjection(_Projection):
    pass


class MultiheadAttention_CP(nn.Module):
    """CP-LRP attention: relevance flows through the value path only (modules.py:87-124; the rule itself: special.py)"""

    _GEOMETRY = ("embed_dim", "num_heads", "head_dim", "batch_first")

    def __init__(self):
        nn.Module.__init__(self)
        for name in ("q_proj_weight", "k_proj_weight", "bias_q", "bias_k") + self._GEOMETRY:
            setattr(self, name, None)
        self.v_proj, self.out_proj = LinearInProjection(None, None), LinearOutProjection(None, None)

    def forward(self, query, key, value, key_padding_mask=None, need_weights=True, attn_mask=None, average_attn_weights=True, is_causal=False):
        if is_causal:
            raise AssertionError("is_causal is not supported (neither by the reference, modules.py:119)")
        return ls.multi_head_attention_cp(query, key, value, self.batch_first, self.num_heads, self.head_dim, self.q_proj_weight, self.bias_q,
                                          self.k_proj_weight, self.bias_k, self.v_proj, self.out_proj, key_padding_mask=key_padding_mask,
                                          need_weights=need_weights, attn_mask=attn_mask, average_attn_weights=average_attn_weights)


def copy_parameters_and_buffers_(original, replacement):
    """`replacement` shares (not clones) every parameter and buffer of `original` (modules.py:127-136)"""
    for register, items in ((replacement.register_parameter, original.named_parameters()), (replacement.register_buffer, original.named_buffers())):
        for name, tensor in items:
            register(name, tensor)


def _build(original, replacement, **overrides):
    """instantiate `replacement` from the constructor arguments that `original` carries as attributes, then share its tensors"""
    kwargs = {name: getattr(original, name) for name in inspect.signature(type(original).__init__).parameters if hasattr(original, name)}
    kwargs.update(overrides)
    new = replacement(**kwargs)
    copy_parameters_and_buffers_(original, new)
    return new


def initialize_generic(original, replacement):
    """modules.py:139-152"""
    return _build(original, replacement)


def initialize_bias(original, replacement):
    """modules.py:155-170: `bias` is a tensor on the original and a flag in the constructor"""
    return _build(original, replacement, bias=original.bias is not None)


def initialize_MHA(original, replacement):
    """MultiheadAttention_CP from a torch nn.MultiheadAttention (modules.py:173-205): the packed in-projection becomes three views"""
    if original.bias_k is not None:
        raise NotImplementedError("add_bias_kv=True is not supported yet.")
    new = replacement()
    if original._qkv_same_embed_dim:
        wq, wk, wv = original.in_proj_weight.chunk(3, dim=0)
    else:
        wq, wk, wv = original.q_proj_weight, original.k_proj_weight, original.v_proj_weight
    new.q_proj_weight, new.k_proj_weight, new.v_proj.weight = wq, wk, wv
    if original.in_proj_bias is not None:
        new.bias_q, new.bias_k, new.v_proj.bias = original.in_proj_bias.chunk(3, dim=0)
    new.out_proj.weight, new.out_proj.bias = original.out_proj.weight, original.out_proj.bias
    for name in MultiheadAttention_CP._GEOMETRY:
        setattr(new, name, getattr(original, name))
    return new


INIT_MODULE_MAPPING = {
    SoftmaxDT: initialize_generic,
    LinearEpsilon: initialize_bias,
    RMSNormIdentity: initialize_generic,
    LayerNormEpsilon: initialize_bias,
    MultiheadAttention_CP: initialize_MHA,
}
