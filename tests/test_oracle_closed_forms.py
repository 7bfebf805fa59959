"""CPU: the reference's own unit tests hold no golden vectors — they compare `lxt.explicit` against closed-form
einsum restatements of the AttnLRP propositions on unseeded random inputs (reference tests/test_functional.py:8-26
softmax Prop 3.1, :29-54 matmul Prop 3.3, :57-76 linear Eq. 8, :79-106 sum, :163-178 rms-norm identity;
tests/test_rules.py:9-24 EpsilonRule == linear_epsilon).  The same closed forms, with the same shapes and tolerances,
are applied here to the oracle's rule functions (on seeded inputs), plus the conservation properties the rules are built on."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import attnlrp_oracle as O  # noqa: E402


def _g(seed=0):
    return torch.Generator().manual_seed(seed)


def test_softmax_prop_3_1():
    x = torch.randn(16, 10, 32, generator=_g(1))
    r = torch.randn(16, 10, 32, generator=_g(2))
    p = torch.softmax(x, -1)
    gt = x * (r - p * r.sum(-1, keepdim=True))
    assert torch.allclose(O.softmax_relevance(x, r, -1), gt, rtol=0, atol=1e-5)
    # -inf (masked) inputs carry no relevance
    xm = x.clone()
    xm[..., 20:] = float("-inf")
    out = O.softmax_relevance(xm, r, -1)
    assert torch.isfinite(out).all() and float(out[..., 20:].abs().max()) == 0.0


def test_matmul_prop_3_3():
    eps = 1e-9
    a = torch.randn(2, 10, 32, generator=_g(3))
    b = torch.randn(2, 32, 5, generator=_g(4))
    r = torch.randn(2, 10, 5, generator=_g(5))
    y = torch.matmul(a, b)
    ra_gt = torch.einsum("bji, bip, bjp -> bji", a, b, r / (2 * y + eps))
    rb_gt = torch.einsum("bji, bip, bjp -> bip", a, b, r / (2 * y + eps))
    ra, rb = O.matmul_relevance(a, b, r, eps)
    assert torch.allclose(ra, ra_gt, rtol=0, atol=1e-4) and torch.allclose(rb, rb_gt, rtol=0, atol=1e-4)
    # conservation: the two halves together carry R_out (Prop 3.3: each operand gets one half); fp64 so that
    # near-zero denominators do not hide it
    ra, rb = O.matmul_relevance(a.double(), b.double(), r.double(), 1e-12)
    assert abs(float(ra.sum() + rb.sum() - r.double().sum())) < 1e-6 * float(r.abs().sum())
    assert abs(float(ra.sum() - rb.sum())) < 1e-6 * float(r.abs().sum())


def test_linear_equation_8_and_epsilon_rule_module():
    eps = 1e-9
    x = torch.randn(16, 10, generator=_g(6))
    w = torch.randn(5, 10, generator=_g(7))
    bias = torch.randn(5, generator=_g(8))
    r = torch.randn(16, 5, generator=_g(9))
    y = torch.nn.functional.linear(x, w, bias)
    gt = torch.einsum("ji, bi, bj -> bi", w, x, r / (y + eps))
    assert torch.allclose(O.linear_epsilon_relevance(x, w, bias, r, eps), gt, rtol=0, atol=1e-3)
    # reference tests/test_rules.py: EpsilonRule(F.linear) == lf.linear_epsilon  (shapes [1,5] x [5,5])
    x1, w1, b1, r1 = (torch.randn(1, 5, generator=_g(10)), torch.randn(5, 5, generator=_g(11)),
                      torch.randn(5, generator=_g(12)), torch.randn(1, 5, generator=_g(13)))
    assert torch.allclose(O.epsilon_rule_relevance(x1, w1, b1, r1, 1e-6), O.linear_epsilon_relevance(x1, w1, b1, r1, 1e-6),
                          rtol=0, atol=1e-3)
    # conservation without bias: sum_i R_in[b,i] = sum_j R_out[b,j]
    rin = O.linear_epsilon_relevance(x.double(), w.double(), None, r.double(), 1e-12)
    assert torch.allclose(rin.sum(-1), r.double().sum(-1), rtol=0, atol=1e-6)


def test_sum_epsilon_rule():
    eps = 1e-9
    a = torch.randn(16, 10, 32, generator=_g(14))
    b = torch.randn(16, 10, 32, generator=_g(15))
    r = torch.randn(16, 10, 32, generator=_g(16))
    ra, rb = O.add2_relevance(a, b, r, eps)
    assert torch.allclose(ra, a * (r / (a + b + eps)), rtol=0, atol=1e-4)
    assert torch.allclose(rb, b * (r / (a + b + eps)), rtol=0, atol=1e-4)
    ra, rb = O.add2_relevance(a.double(), b.double(), r.double(), 0.0)
    assert torch.allclose(ra + rb, r.double(), rtol=0, atol=1e-9)


def test_identity_and_uniform_rules():
    r = torch.randn(1, 4, 32, generator=_g(17))
    assert torch.equal(O.rms_norm_identity_relevance(r), r)            # reference test_normalize: relevance passes through
    assert torch.allclose(O.mul2_relevance(r, 2) * 2, r) and torch.equal(O.mul2_relevance(r, 1), r)
    # GxI identity rule times the input gives back f(x) * g: the relevance of a point-wise non-linearity is conserved
    x = torch.randn(64, generator=_g(18)).double()
    g = torch.randn(64, generator=_g(19)).double()
    fx = torch.nn.functional.silu(x)
    assert torch.allclose(O.identity_rule_implicit_grad(fx, x, g) * x, fx * g, rtol=1e-6, atol=1e-9)
    assert torch.allclose(O.divide_gradient_grad(g, 4.0) * 4.0, g)


def test_uniform_epsilon_rule_is_half_of_the_epsilon_rule_per_operand():
    a = torch.randn(2, 6, 8, generator=_g(20)).double()
    b = torch.randn(2, 8, 3, generator=_g(21)).double()
    r = torch.randn(2, 6, 3, generator=_g(22)).double()
    ua, ub = O.uniform_epsilon_matmul_relevance(a, b, r, 1e-12)
    ma, mb = O.matmul_relevance(a, b, r, 1e-12)     # s = R / (2 O): the functional form of the same rule
    assert torch.allclose(ua, ma, rtol=1e-9, atol=1e-9) and torch.allclose(ub, mb, rtol=1e-9, atol=1e-9)
