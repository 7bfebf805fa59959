"""TEST INFRASTRUCTURE — CPU restatement of the Gamma rule as the reference runs it in Gradient x Input space.

PARITY UNPINNED.  The rule lives in the third-party package `zennit` (listed unpinned in the reference's setup.py:18; the rule
described here is the one of zennit 0.5.x, generalised to signed inputs), which is neither in /root/reference nor installed here, and
the reference holds no test or golden vector for it.  This file restates
  * zennit's published `Gamma(gamma, stabilizer)` rule: four (input modifier, parameter modifier) pairs plus one unmodified pass whose
    sign selects the branch, a gradient mapper and a reducer, evaluated by `BasicHook.backward` with `torch.autograd.grad`, and
  * the two lines the reference adds around it (lxt/efficient/zennit_patches.py:37-39 and 59-60): `grad_output * output` on the way
    in, `/ stabilize(input, 1e-10)` on the way out,
literally, pass by pass, with autograd on the CPU.  Call sites anchoring the semantics: examples/vit_torch.py:59-65.
Only tests/ may import this module.
"""
import torch
import torch.nn.functional as F


def stabilize(x: torch.Tensor, epsilon: float = 1e-6) -> torch.Tensor:
    """zennit.core.stabilize: x + ((x == 0) + sign(x)) * eps — eps is added with the sign of x, +eps at zero"""
    return x + ((x == 0.).to(x) + x.sign()) * epsilon


def _gamma_mod(p, gamma, lo=None, hi=None):
    """zennit GammaMod: param + gamma * param.clamp(min=lo, max=hi)"""
    return None if p is None else p + gamma * p.clamp(min=lo, max=hi)


def gamma_gxi(fn, x, weight, bias, grad_output, gamma: float = 0.25, stabilizer: float = 1e-6):
    """Modified gradient w.r.t. x of y = fn(x, weight, bias) under zennit's Gamma rule run through the reference's patched hook.
    fn: F.linear or a conv closure taking (input, weight, bias)."""
    out = fn(x, weight, bias)
    R = grad_output * out                                                   # zennit_patches.py:38
    in_mods = [lambda t: t.clamp(min=0), lambda t: t.clamp(max=0), lambda t: t.clamp(min=0), lambda t: t.clamp(max=0)]
    # (weight clamp bounds, keep bias?) of the four passes: the second pass of each branch zeroes the bias
    par_mods = [((0., None), True), ((None, 0.), False), ((None, 0.), True), ((0., None), False)]
    inputs, outputs = [], []
    for im, ((lo, hi), keep_bias) in zip(in_mods, par_mods):
        xi = im(x.detach()).requires_grad_()
        wi = _gamma_mod(weight, gamma, lo, hi)
        bi = _gamma_mod(bias, gamma, lo, hi) if (bias is not None and keep_bias) else None
        inputs.append(xi)
        with torch.enable_grad():                                           # (the hook runs inside a backward pass)
            outputs.append(fn(xi, wi, bi))
    zp, zn = outputs[0] + outputs[1], outputs[2] + outputs[3]
    plain = out.detach()                                                    # fifth pass: unmodified input and parameters
    grad_outputs = [(plain > 0.).to(R) * R / stabilize(zp, stabilizer)] * 2 + [(plain < 0.).to(R) * R / stabilize(zn, stabilizer)] * 2
    grads = torch.autograd.grad(outputs, inputs, grad_outputs=[g.detach() for g in grad_outputs])
    relevance = sum(i.detach() * g for i, g in zip(inputs, grads))          # the rule's reducer
    return relevance / stabilize(x.detach(), 1e-10)                         # zennit_patches.py:60


def gamma_linear_gxi(x, weight, bias, grad_output, gamma=0.25, stabilizer=1e-6):
    return gamma_gxi(F.linear, x, weight, bias, grad_output, gamma, stabilizer)


def gamma_conv2d_gxi(x, weight, bias, grad_output, stride, gamma=0.25, stabilizer=1e-6):
    return gamma_gxi(lambda i, w, b: F.conv2d(i, w, b, stride=stride), x, weight, bias, grad_output, gamma, stabilizer)


# ---- the reference's ViT recipe on the CPU (examples/vit_torch.py:15-16, 59-65, 84-91), for end-to-end checks ----
class _IdentityRuleFn(torch.autograd.Function):
    """lxt/efficient/rules.py:88-100: forward f(x), backward g * f(x) / (x + 1e-10)"""

    @staticmethod
    def forward(ctx, x, fn):
        y = fn(x)
        ctx.save_for_backward(x, y)
        return y

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        return g * (y / (x + 1e-10)), None


class _GammaFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, module, gamma, stabilizer):
        ctx.args = (module, gamma, stabilizer)
        ctx.save_for_backward(x)
        return module._gamma_fn(x, module.weight, module.bias)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        module, gamma, stabilizer = ctx.args
        return gamma_gxi(module._gamma_fn, x, module.weight.detach(), None if module.bias is None else module.bias.detach(), g, gamma,
                         stabilizer), None, None, None


def patch_vit_cpu(model, conv_gamma: float, lin_gamma: float, stabilizer: float = 1e-6):
    """torchvision VisionTransformer (CPU, any float dtype) under the reference's cp_LRP map (lxt/efficient/models/vit_torch.py:7-11:
    identity rule on GELU, LayerNorm with detached std, q / k detached in MultiheadAttention) plus zennit Gamma on every nn.Conv2d and
    nn.Linear MODULE call (the in / out projections inside nn.MultiheadAttention use the weights directly, no module call: no rule)."""
    import torch.nn as nn

    def ln_forward(self, x):                                               # lxt/efficient/patches.py:126-142
        mean = x.mean(-1, keepdim=True)
        std = ((x - mean) ** 2).mean(-1, keepdim=True).add(self.eps).sqrt().detach()
        return (x - mean) / std * self.weight + self.bias

    def mha_forward(self, query, key, value, *args, **kwargs):             # patches.py:261-269
        return nn.MultiheadAttention.forward(self, query.detach(), key.detach(), value, *args, **kwargs)

    for m in model.modules():
        if isinstance(m, nn.GELU):
            m.forward = (lambda self, x: _IdentityRuleFn.apply(x, F.gelu)).__get__(m)
        elif isinstance(m, nn.LayerNorm):
            m.forward = ln_forward.__get__(m)
        elif isinstance(m, nn.MultiheadAttention):
            m.forward = mha_forward.__get__(m)
        elif isinstance(m, nn.Conv2d):
            m._gamma_fn = (lambda stride: (lambda i, w, b: F.conv2d(i, w, b, stride=stride)))(m.stride)
            m.forward = (lambda self, x, g=conv_gamma: _GammaFn.apply(x, self, g, stabilizer)).__get__(m)
        elif isinstance(m, nn.Linear) and not isinstance(m, nn.modules.linear.NonDynamicallyQuantizableLinear):
            m._gamma_fn = F.linear
            m.forward = (lambda self, x, g=lin_gamma: _GammaFn.apply(x, self, g, stabilizer)).__get__(m)
    return model
