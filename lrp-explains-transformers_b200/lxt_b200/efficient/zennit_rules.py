"""The Gamma rule of the reference's vision-transformer recipe, natively in Gradient x Input space.

The reference gets the rule from the third-party `zennit` package (`zennit.rules.Gamma`, `zennit.composites.LayerMapComposite`,
examples/vit_torch.py:59-65) and rewrites zennit's hook so that it runs on modified gradients (`monkey_patch_zennit`,
lxt/efficient/zennit_patches.py:26-77): grad_output * output -> zennit rule -> / stabilize(input, 1e-10).  zennit is not part of
`/root/reference` and not installed here (setup.py:18 lists it unpinned), so this module restates the published rule — PARITY
UNPINNED: the oracle (oracle/gamma_oracle.py) is the same restatement evaluated with autograd on the CPU, checked against the
rule's closed forms only.

Gamma (zennit >= 0.5, generalised to signed inputs) for y = x W^T + b, with x+ = max(x,0), x- = min(x,0),
Wp = W + gamma max(W,0), Wm = W + gamma min(W,0) (biases alike, each counted once per branch):
    zp = x+ Wp^T + x- Wm^T + bp            zn = x+ Wm^T + x- Wp^T + bm
    sp = [y > 0] R / stabilize(zp)         sn = [y < 0] R / stabilize(zn)             R = grad_output * y
(every output takes exactly one branch, chosen by the sign of the unmodified pre-activation — the rule's fifth, unmodified pass)
    relevance = x+ * (sp Wp + sn Wm) + x- * (sp Wm + sn Wp)          grad_input = relevance / stabilize(x, 1e-10)
B200 form: the four modified forward passes are two GEMMs over [x+ | x-] (contraction 2K), the four backward passes two dgrads over
[sp | sn] (contraction 2N); clamping, normalisation and the final select are three small kernels (lrp_gamma_split / _s / _combine).
The stacked weights are built once per module at `register()` time.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.nn as nn
from torch.autograd import Function

from .. import ops
from .._capi import LrpError


class Gamma:
    """rule specification, same constructor as `zennit.rules.Gamma(gamma=0.25, stabilizer=1e-6)`"""

    def __init__(self, gamma: float = 0.25, stabilizer: float = 1e-6, zero_params=None):
        if zero_params is not None:
            raise NotImplementedError("Gamma(zero_params=...) is not supported")
        self.gamma = float(gamma)
        self.stabilizer = float(stabilizer)


class _Stacked:
    """[Wp | Wm], [Wm | Wp] (N x 2K: forward passes) and [Wp ; Wm], [Wm ; Wp] (2N x K: backward passes) of one weight matrix"""

    def __init__(self, weight: torch.Tensor, bias, gamma: float):
        w = weight.detach()
        wp = w + gamma * w.clamp(min=0)
        wm = w + gamma * w.clamp(max=0)
        self.f_pos = torch.cat([wp, wm], 1).contiguous()
        self.f_neg = torch.cat([wm, wp], 1).contiguous()
        self.b_pos = torch.cat([wp, wm], 0).contiguous()
        self.b_neg = torch.cat([wm, wp], 0).contiguous()
        self.w = w.contiguous()
        if bias is None:
            self.bias = self.bias_p = self.bias_m = None
        else:
            b = bias.detach().float()
            self.bias = b.contiguous()
            self.bias_p = (b + gamma * b.clamp(min=0)).contiguous()
            self.bias_m = (b + gamma * b.clamp(max=0)).contiguous()
        self.version = (weight._version, None if bias is None else bias._version)


class _GammaLinearFn(Function):
    @staticmethod
    def forward(ctx, x, st: _Stacked, stabilizer: float):
        x2 = x.reshape(-1, x.shape[-1])
        x2 = x2 if x2.is_contiguous() else x2.contiguous()
        y = torch.empty((x2.shape[0], st.w.shape[0]), dtype=x.dtype, device=x.device)
        ops.linear_fwd(x2, st.w, y, bias=st.bias)
        ctx.save_for_backward(x2, y)
        ctx.st, ctx.stabilizer, ctx.shape = st, stabilizer, x.shape
        return y.view(*x.shape[:-1], st.w.shape[0])

    @staticmethod
    def backward(ctx, gy):
        x2, y = ctx.saved_tensors
        st = ctx.st
        T, K = x2.shape
        N = st.w.shape[0]
        g2 = gy.reshape(T, N)
        xcat = ops.gamma_split(x2)                                          # [x+ | x-]
        zp = torch.empty((T, N), dtype=x2.dtype, device=x2.device)
        zn = torch.empty_like(zp)
        ops.linear_fwd(xcat, st.f_pos, zp, bias=st.bias_p)                  # x+ Wp^T + x- Wm^T + bp
        ops.linear_fwd(xcat, st.f_neg, zn, bias=st.bias_m)                  # x+ Wm^T + x- Wp^T + bm
        scat = ops.gamma_s(g2, y, zp, zn, ctx.stabilizer)                   # [sp | sn]
        g1 = torch.empty((T, K), dtype=x2.dtype, device=x2.device)
        gneg = torch.empty_like(g1)
        ops.linear_dgrad(scat, st.b_pos, g1)                                # sp Wp + sn Wm   (multiplies x+)
        ops.linear_dgrad(scat, st.b_neg, gneg)                              # sp Wm + sn Wp   (multiplies x-)
        return ops.gamma_combine(x2, g1, gneg).view(ctx.shape), None, None


def _check(x: torch.Tensor, weight: torch.Tensor) -> None:
    if not (x.is_cuda and weight.is_cuda):
        raise LrpError("Gamma rule: CUDA tensors only (no CPU fallback)")
    if x.dtype not in (torch.bfloat16, torch.float32) or weight.dtype != x.dtype:
        raise LrpError(f"Gamma rule: input and weight must both be bf16 or both fp32 (got {x.dtype}, {weight.dtype})")
    if weight.shape[0] % 8 or weight.shape[1] % 8:
        raise LrpError(f"Gamma rule: feature counts must be multiples of 8 (got {tuple(weight.shape)})")


def _stacked(module, weight2d, bias, gamma):
    st = getattr(module, "_lrp_gamma_stack", None)
    ver = (module.weight._version, None if bias is None else bias._version)
    if st is None or st.version != ver or st.gamma != gamma:
        st = _Stacked(weight2d, bias, gamma)
        st.version, st.gamma = ver, gamma
        module._lrp_gamma_stack = st
    return st


def _linear_forward(rule: Gamma):
    def forward(self, x):
        _check(x, self.weight)
        return _GammaLinearFn.apply(x, _stacked(self, self.weight, self.bias, rule.gamma), rule.stabilizer)
    return forward


def _conv_forward(rule: Gamma):
    """Conv2d whose stride equals its kernel (the ViT patch embedding): a Linear over the unfolded patches"""
    def forward(self, x):
        kh, kw = self.kernel_size
        if (tuple(self.stride) != (kh, kw) or tuple(self.padding) != (0, 0) or tuple(self.dilation) != (1, 1) or self.groups != 1
                or isinstance(self.padding, str)):
            raise NotImplementedError("Gamma rule on Conv2d: only non-overlapping patch embeddings (stride == kernel, no padding)")
        B, Cin, H, W = x.shape
        if H % kh or W % kw:
            raise NotImplementedError("Gamma rule on Conv2d: the image must tile into whole patches")
        w2 = self.weight.reshape(self.out_channels, Cin * kh * kw)
        _check(x, w2)
        nh, nw = H // kh, W // kw
        patches = x.reshape(B, Cin, nh, kh, nw, kw).permute(0, 2, 4, 1, 3, 5).reshape(B, nh * nw, Cin * kh * kw)
        y = _GammaLinearFn.apply(patches, _stacked(self, w2, self.bias, rule.gamma), rule.stabilizer)
        return y.view(B, nh, nw, self.out_channels).permute(0, 3, 1, 2)
    return forward


class LayerMapComposite:
    """`zennit.composites.LayerMapComposite([(nn.Conv2d, Gamma(g1)), (nn.Linear, Gamma(g2))])` for the module types the
    reference's recipe maps (examples/vit_torch.py:59-65): `register(model)` routes the matching modules through the rule,
    `remove()` restores them.  First matching entry wins, as in zennit."""

    def __init__(self, layer_map: List[Tuple[type, Gamma]], canonizers=None):
        if canonizers:
            raise NotImplementedError("canonizers are not supported")
        self.layer_map = list(layer_map)
        self._patched: List[nn.Module] = []

    def register(self, module: nn.Module) -> None:
        for child in module.modules():
            for types, rule in self.layer_map:
                if isinstance(child, types):
                    if not isinstance(rule, Gamma):
                        raise NotImplementedError(f"only the Gamma rule is provided, got {type(rule).__name__}")
                    if isinstance(child, nn.Conv2d):
                        fwd = _conv_forward(rule)
                    elif isinstance(child, nn.Linear):
                        fwd = _linear_forward(rule)
                    else:
                        raise NotImplementedError(f"Gamma rule on {type(child).__name__}")
                    child.forward = fwd.__get__(child, type(child))       # instance attribute: shadows the (possibly patched) class method
                    self._patched.append(child)
                    break

    def remove(self) -> None:
        for child in self._patched:
            child.__dict__.pop("forward", None)
            child.__dict__.pop("_lrp_gamma_stack", None)
        self._patched = []


def monkey_patch_zennit(verbose: bool = False) -> None:
    """Counterpart of `lxt.efficient.monkey_patch_zennit` (zennit_patches.py:65-77).  Nothing to patch: the rules of this module
    already take and return modified gradients."""
    if verbose:
        print("lxt_b200: Gamma / LayerMapComposite run in Gradient x Input space natively; no zennit hook to patch")
