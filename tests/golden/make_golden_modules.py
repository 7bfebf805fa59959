"""Golden vectors for the module layer of the explicit API, from the REAL reference (`lxt.explicit.modules`, `lxt.explicit.special`).

    PYTHONPATH=/root/reference python tests/golden/make_golden_modules.py      # build container only; writes explicit_modules.npz

Cases (fp32, CPU, seeded): SoftmaxDT, LinearEpsilon, RMSNormIdentity, LayerNormEpsilon on [2,17,128] activations; MultiheadAttention_CP
initialised from a torch nn.MultiheadAttention(128, 2 heads) by `initialize_MHA`, self-attention on [2,17,128] — plain, with a
key-padding mask (right padding) and with need_weights — each with the relevance it hands to its input.
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))

import lxt.explicit.modules as lm  # noqa: E402  (the reference)


def main():
    g = torch.Generator().manual_seed(77)
    rn = lambda *s: torch.randn(*s, generator=g)
    out = {}
    x, R = rn(2, 17, 128), rn(2, 17, 128)

    def run(mod, inp, seed, **kw):
        xi = inp.clone().requires_grad_()
        y = mod(xi, **kw)
        y0 = y[0] if isinstance(y, tuple) else y
        y0.backward(seed)
        return y, xi.grad.clone()

    y, r = run(lm.SoftmaxDT(dim=-1, temperature=2.0), x, R)
    out.update(x=x, R=R, softmax_y=y.detach(), softmax_R=r)

    # epsilon rules divide by the layer output: positive, bf16-representable operands keep z = x W^T + b away from zero (with z ~ 0 the
    # quotient R / (z + 1e-6) amplifies any rounding of z without bound, in the reference as much as anywhere else)
    lin = nn.Linear(128, 64)
    xl = (torch.rand(2, 17, 128, generator=g) + 0.5).bfloat16().float()
    with torch.no_grad():
        lin.weight.copy_((torch.rand(64, 128, generator=g) * 0.1 + 0.02).bfloat16().float())
        lin.bias.copy_((torch.rand(64, generator=g) * 0.1).bfloat16().float())
    le = lm.initialize_bias(lin, lm.LinearEpsilon)
    Rl = rn(2, 17, 64)
    y, r = run(le, xl, Rl)
    out.update(lin_x=xl, lin_w=lin.weight.detach(), lin_b=lin.bias.detach(), lin_Rout=Rl, lin_y=y.detach(), lin_R=r)

    rms = lm.RMSNormIdentity(128, eps=1e-5)
    with torch.no_grad():
        rms.weight.copy_(1.0 + 0.1 * rn(128))
    y, r = run(rms, x, R)
    out.update(rms_w=rms.weight.detach(), rms_y=y.detach(), rms_R=r)

    ln = nn.LayerNorm(128)
    with torch.no_grad():
        ln.weight.copy_(1.0 + 0.1 * rn(128))
        ln.bias.copy_(0.1 * rn(128))
    lne = lm.initialize_bias(ln, lm.LayerNormEpsilon)
    y, r = run(lne, x, R)
    out.update(ln_w=ln.weight.detach(), ln_b=ln.bias.detach(), ln_y=y.detach(), ln_R=r)

    mha = nn.MultiheadAttention(128, 2, batch_first=True)
    with torch.no_grad():
        mha.in_proj_bias[256:] = 2.0          # value path well away from zero: the epsilon rule on P V divides by the attention output
    cp = lm.initialize_MHA(mha, lm.MultiheadAttention_CP)
    out.update(mha_in_w=mha.in_proj_weight.detach(), mha_in_b=mha.in_proj_bias.detach(), mha_out_w=mha.out_proj.weight.detach(),
               mha_out_b=mha.out_proj.bias.detach())

    def run_mha(**kw):
        xi = x.clone().requires_grad_()
        res = cp(xi, xi, xi, **kw)
        res[0].backward(R)
        return res, xi.grad.clone()

    (y, w), r = run_mha(need_weights=True)
    out.update(mha_y=y.detach(), mha_w=w.detach(), mha_R=r)
    (y, w), r = run_mha(need_weights=False)
    assert w is None
    out.update(mha_y_nw=y.detach(), mha_R_nw=r)
    kpm = torch.zeros(2, 17, dtype=torch.bool)
    kpm[0, 12:] = True
    kpm[1, 15:] = True
    (y, w), r = run_mha(need_weights=False, key_padding_mask=kpm)
    out.update(mha_kpm=kpm, mha_y_kpm=y.detach(), mha_R_kpm=r)
    np.savez_compressed(os.path.join(HERE, "explicit_modules.npz"), **{k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()})
    print("explicit_modules.npz", {k: tuple(v.shape) for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
