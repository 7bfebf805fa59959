"""GPU: the drop-in rule API (`lxt_b200.explicit.functional`, `.explicit.rules`, `.efficient.rules`) against the
CPU oracle on the reference's own test shapes (tests/test_rules.py, tests/test_functional.py) and the BASELINE
"2-layer 128-d MLP".  Inputs are the golden inputs rounded to bf16 (the tensor-core operands are bf16), the oracle
is evaluated in fp32 on exactly those values.  Tolerances: element-wise fp32 kernels 1e-5; rules whose
normalised relevance s = R/(z+eps) is a bf16 GEMM operand 2.5e-3 rel-L2 (bound of one bf16 rounding)."""
import functools

import pytest
import torch

from helpers import load_npz, rel_l2
from oracle import attnlrp_oracle as O

pytestmark = pytest.mark.gpu
T = torch.from_numpy


def q(t):
    """bf16-representable fp32"""
    return t.bfloat16().float()


@pytest.fixture(scope="module")
def G():
    return {k: T(v) for k, v in load_npz("rules.npz").items()}


def _grad(fn, *xs, seed):
    xs = [x.cuda().requires_grad_() for x in xs]
    y = fn(*xs)
    y.backward(seed.cuda().to(y.dtype))
    return y.detach().cpu(), [x.grad.float().cpu() for x in xs]


def test_linear_epsilon_reference_test_shapes(G):
    import lxt_b200.explicit.functional as lf
    import lxt_b200.explicit.rules as rules
    for x, W, b, R, eps in ((G["le_x"], G["le_W"], G["le_b"], G["le_R"], 1e-6), (G["lin_x"], G["lin_W"], None, G["lin_R"], 1e-9)):
        x, W, R = q(x), q(W), q(R)
        b = None if b is None else q(b)
        exp = O.linear_epsilon_relevance(x, W, b, R, eps)
        Wc, bc = W.cuda(), None if b is None else b.cuda()
        y, (got,) = _grad(lambda t: lf.linear_epsilon(t, Wc, bc, eps), x, seed=R)
        assert rel_l2(y, torch.nn.functional.linear(x, W, b)) < 1e-5
        assert rel_l2(got, exp) < 2.5e-3
        # tests/test_rules.py:9-24 — EpsilonRule(F.linear partial) == linear_epsilon
        rule = rules.EpsilonRule(functools.partial(torch.nn.functional.linear, weight=Wc, bias=bc), epsilon=eps)
        _, (got2,) = _grad(rule, x, seed=R)
        assert rel_l2(got2, exp) < 1e-4       # generic VJP path: fp32 element-wise kernels + autograd.grad
        lin = torch.nn.Linear(W.shape[1], W.shape[0], bias=b is not None).cuda()
        lin.weight.data.copy_(Wc)
        if b is not None:
            lin.bias.data.copy_(bc)


def test_two_layer_128d_mlp_epsilon_rule(G):
    # BASELINE.json configs[0]
    import lxt_b200.explicit.functional as lf
    x, W1, b1, W2, b2, R = (q(G[k]) for k in ("mlp_x", "mlp_W1", "mlp_b1", "mlp_W2", "mlp_b2", "mlp_R"))
    h1 = q(torch.nn.functional.linear(x, W1, b1))  # the kernel's layer-2 input is the bf16-rounded layer-1 output
    R1 = O.linear_epsilon_relevance(h1, W2, b2, R, 1e-6)
    exp = O.linear_epsilon_relevance(x, W1, b1, R1, 1e-6)
    W1c, b1c, W2c, b2c = W1.cuda(), b1.cuda(), W2.cuda(), b2.cuda()
    _, (got,) = _grad(lambda t: lf.linear_epsilon(lf.linear_epsilon(t, W1c, b1c, 1e-6), W2c, b2c, 1e-6), x, seed=R)
    assert rel_l2(got, exp) < 5e-3   # two chained bf16-operand rules


def test_fused_linear_eps_kernel_large(G):
    from lxt_b200 import ops
    g = torch.Generator().manual_seed(3)
    Tn, N, K = 1500, 1024, 768
    x = q(torch.rand(Tn, K, generator=g) + 0.5)
    W = q(torch.rand(N, K, generator=g) + 0.5)
    b = torch.rand(N, generator=g)
    R = torch.randn(Tn, N, generator=g)
    exp = O.linear_epsilon_relevance(x, W, b, R, 1e-6)
    got = ops.linear_eps_bwd(x.cuda(), W.cuda(), b.cuda(), R.cuda(), 1e-6).cpu()
    assert rel_l2(got, exp) < 2.5e-3


def test_matmul_softmax_add_mul_rms(G):
    import lxt_b200.explicit.functional as lf
    import lxt_b200.explicit.rules as rules
    a, b, R = q(G["mm_a"]), q(G["mm_b"]), q(G["mm_R"])
    ea, eb = O.matmul_relevance(a, b, R, 1e-9)
    _, (ga, gb) = _grad(lambda p, r: lf.matmul(p, r, False, 1e-9), a, b, seed=R)
    assert rel_l2(ga, ea) < 2.5e-3 and rel_l2(gb, eb) < 2.5e-3
    x, R = G["sm_x"], G["sm_R"]
    _, (gs,) = _grad(lambda t: lf.softmax(t, -1), x, seed=R)
    assert rel_l2(gs, O.softmax_relevance(x, R)) < 1e-5
    a, b, R = G["add_a"], G["add_b"], G["add_R"]
    _, (ra, rb) = _grad(lambda p, r: lf.add2(p, r, False, 1e-8), a, b, seed=R)
    ea, eb = O.add2_relevance(a, b, R, 1e-8)
    assert rel_l2(ra, ea) < 1e-5 and rel_l2(rb, eb) < 1e-5
    _, (ma, mb) = _grad(lf.mul2, a, b, seed=R)
    assert torch.allclose(ma, R / 2) and torch.allclose(mb, R / 2)
    x, w, R = G["rms_x"], G["rms_w"], G["rms_R"]
    y, (gx,) = _grad(lambda t: lf.rms_norm_identity(t, w.cuda(), 1e-6), x, seed=R)
    assert torch.allclose(y, G["rms_y"], atol=1e-5) and torch.equal(gx, R)

    class MM(torch.nn.Module):
        def forward(self, p, r):
            return torch.matmul(p, r)

    a, b, R = G["ue_a"], G["ue_b"], G["ue_R"]
    _, (ua, ub) = _grad(rules.UniformEpsilonRule(MM(), epsilon=1e-6), a, b, seed=R)
    ea, eb = O.uniform_epsilon_matmul_relevance(a, b, R, 1e-6)
    assert rel_l2(ua, ea) < 1e-4 and rel_l2(ub, eb) < 1e-4


def test_efficient_rules(G):
    from lxt_b200.efficient import rules as R
    F = torch.nn.functional
    x, g = G["id_x"], G["id_gout"]
    for name, fn in (("silu", F.silu), ("gelu", torch.nn.GELU()), ("gelu_tanh", torch.nn.GELU(approximate="tanh")),
                     ("silu", lambda t: F.silu(t))):  # last: generic-callable path
        y, (gx,) = _grad(lambda t: R.identity_rule_implicit(fn, t), x, seed=g)
        assert rel_l2(gx, G[f"id_{name}_g"]) < 1e-5
    _, (gd,) = _grad(lambda t: R.divide_gradient(t * 1.0, 4), x, seed=g)
    assert torch.allclose(gd, G["div4_g"])


def test_matmul_and_softmax_rules_at_attention_shape():
    """lf.matmul / lf.softmax on the [B,H,S,S] attention products (reference lxt/explicit/models/llama.py:379-391 uses exactly
    these): one strided-batched launch per contraction instead of a Python loop over B*H slices; fp32 operands are held to the
    reference test's own tolerance (tests/test_functional.py:29-54, atol 1e-4 on O(1) data), bf16 operands to a bf16 bar."""
    import lxt_b200.explicit.functional as lf
    from lxt_b200 import ops
    g = torch.Generator().manual_seed(11)
    B, H, S, D = 2, 8, 512, 64
    # positive operands keep 2 O + eps away from zero: the rule divides by it, and a near-singular element would dominate any
    # norm (the reference's own test avoids this with 10 x 5 outputs)
    qh = torch.rand(B, H, S, D, generator=g) * 0.5 + 0.1
    kh = torch.rand(B, H, D, S, generator=g) * 0.5 + 0.1
    R = torch.randn(B, H, S, S, generator=g)
    ea, eb = O.matmul_relevance(qh, kh, R, 1e-8)
    n0 = ops.launch_count()
    _, (ga, gb) = _grad(lambda p, r: lf.matmul(p, r, False, 1e-8), qh, kh, seed=R)
    launches = ops.launch_count() - n0
    assert launches <= 3 * 3 + 2 * 3 + 3 + 8, f"{launches} launches: the batch must not be looped over in Python"
    assert rel_l2(ga, ea) < 1e-4 and rel_l2(gb, eb) < 1e-4
    # bf16 operands: one bf16 product per contraction
    qb, kb, Rb = (t.to(torch.bfloat16) for t in (qh, kh, R))
    ea16, eb16 = O.matmul_relevance(qb.float(), kb.float(), Rb.float(), 1e-8)
    _, (ga16, gb16) = _grad(lambda p, r: lf.matmul(p, r, False, 1e-8), qb, kb, seed=Rb)
    assert rel_l2(ga16, ea16) < 2e-2 and rel_l2(gb16, eb16) < 2e-2
    # P V with the soft-max rule in front: softmax forward kernel + Deep-Taylor backward at [B,H,S,S]
    x = torch.randn(B, H, S, S, generator=g)
    Rp = torch.randn(B, H, S, S, generator=g)
    y, (gx,) = _grad(lambda t: lf.softmax(t, -1), x, seed=Rp)
    assert rel_l2(y, torch.softmax(x, -1)) < 1e-6
    assert rel_l2(gx, O.softmax_relevance(x, Rp)) < 1e-5
    y2, (gx2,) = _grad(lambda t: lf.softmax(t, 1, None, 2.0), x[:, :, :64, :96].contiguous(), seed=Rp[:, :, :64, :96].contiguous())
    xs = x[:, :, :64, :96] / 2.0
    assert rel_l2(y2, torch.softmax(xs, 1)) < 1e-6
    exp2 = (xs.transpose(1, -1) * (Rp[:, :, :64, :96].transpose(1, -1) - torch.softmax(xs, 1).transpose(1, -1) *
                                   Rp[:, :, :64, :96].transpose(1, -1).sum(-1, keepdim=True))).transpose(1, -1)
    assert rel_l2(gx2, exp2) < 1e-5


def test_reference_matmul_test_shapes_fp32_tolerance():
    """the reference's tests/test_functional.py:29-54 (`test_matmul`): a [2,10,32] x [2,32,5], eps 1e-9, closed form via einsum.
    The reference holds its own fp32 einsum restatement to atol 1e-4; here the products come from two-term bf16 splits (|dO| ~ 1e-6),
    and the rule divides by 2 O + 1e-9, so the elements where O ~ 0 amplify that: the forward is held to the reference's atol, the
    relevances to 1e-3 rel-L2 (the path's stated tolerance) against a float64 closed form."""
    import lxt_b200.explicit.functional as lf
    g = torch.Generator().manual_seed(5)
    a, b, R = torch.randn(2, 10, 32, generator=g), torch.randn(2, 32, 5, generator=g), torch.randn(2, 10, 5, generator=g)
    ad, bd, Rd = a.double(), b.double(), R.double()
    out = torch.einsum("bij,bjk->bik", ad, bd)
    s = Rd / (2 * out + 1e-9)
    exp_a = torch.einsum("bik,bjk->bij", s, bd) * ad
    exp_b = torch.einsum("bij,bik->bjk", ad, s) * bd
    y, (ga, gb) = _grad(lambda p, r: lf.matmul(p, r, False, 1e-9), a, b, seed=R)
    assert torch.allclose(y.double(), out, atol=1e-4)
    assert rel_l2(ga, exp_a) < 1e-3 and rel_l2(gb, exp_b) < 1e-3


def test_mean_layer_norm_normalize_rules_like_the_reference_tests():
    """lf.mean / lf.layer_norm / lf.normalize (reference tests/test_functional.py:109-178): closed forms, the reference's shapes"""
    import lxt_b200.explicit.functional as lf
    g = torch.Generator().manual_seed(9)
    x, R = torch.randn(1, 8, 32, generator=g) + 2.0, torch.randn(1, 8, generator=g)
    for keep in (False, True):
        Rk = R.unsqueeze(-1) if keep else R
        y, (gx,) = _grad(lambda t: lf.mean(t, -1, keep, 1e-9), x, seed=Rk)
        assert torch.allclose(y, x.mean(-1, keepdim=keep), atol=1e-5)
        exp = x * (R / (x.sum(-1) + 1e-9)).unsqueeze(-1)
        assert rel_l2(gx, exp) < 1e-4
    y, (gx,) = _grad(lambda t: lf.mean(t, 1, False, 1e-9), x, seed=torch.randn(1, 32, generator=torch.Generator().manual_seed(1)))
    assert torch.allclose(y, x.mean(1), atol=1e-5)
    # layer_norm: detached-std LayerNorm under the epsilon rule == the composition of the primitive rules (reference cross-check)
    x = torch.randn(1, 2, 8, generator=g)
    w, b, R = torch.randn(8, generator=g), torch.randn(8, generator=g), torch.randn(1, 2, 8, generator=g)
    y1, (g1,) = _grad(lambda t: lf.layer_norm(t, w.cuda(), b.cuda(), 1e-5), x, seed=R)
    ref = torch.nn.functional.layer_norm(x, (8,), w, b, 1e-5)
    assert torch.allclose(y1, ref, atol=1e-5)
    xd = x.double().requires_grad_()
    mu = xd.mean(-1, keepdim=True)
    yd = (xd - mu) / ((xd - mu).pow(2).mean(-1, keepdim=True) + 1e-5).sqrt().detach() * w.double() + b.double()
    (gd,) = torch.autograd.grad(yd, xd, R.double() / (yd.detach() + 1e-6))
    assert rel_l2(g1, gd * x.double()) < 1e-4
    y2, (g2,) = _grad(lambda t: lf._layer_norm_slower(t, w.cuda(), b.cuda(), 1e-5), x, seed=R)
    assert torch.allclose(y2, ref, atol=1e-4)
    cos = torch.nn.functional.cosine_similarity(g1.flatten(), g2.flatten(), dim=0)
    assert cos > 0.99                       # the reference's own bar for this pair (atol 1e-1 + cosine > 0.99)
    # normalize: identity rule
    x, R = torch.randn(1, 4, 32, generator=g), torch.randn(1, 4, 32, generator=g)
    y, (gx,) = _grad(lambda t: lf.normalize(t, 2.0, -1), x, seed=R)
    assert torch.allclose(y, torch.nn.functional.normalize(x, 2.0, -1), atol=1e-6) and torch.equal(gx, R)


def test_taylor_decomposition_rule_without_bias():
    """TaylorDecompositionRule at ref = 0 on a linear map equals the epsilon rule without bias (AttnLRP Eq. 4-5)"""
    import lxt_b200.explicit.rules as rules
    g = torch.Generator().manual_seed(4)
    x, W, R = torch.rand(6, 16, generator=g) + 0.5, torch.rand(8, 16, generator=g) + 0.5, torch.randn(6, 8, generator=g)
    lin = torch.nn.Linear(16, 8, bias=False).cuda()
    lin.weight.data.copy_(W)
    lin.weight.requires_grad_(False)
    rule = rules.TaylorDecompositionRule(lin, ref=(torch.zeros(6, 16, device="cuda"),), bias=False)
    _, (gx,) = _grad(rule, x, seed=R)
    exp = O.linear_epsilon_relevance(x, W, None, R, 1e-6)
    assert rel_l2(gx, exp) < 1e-4
    import pytest as _pt
    with _pt.raises(NotImplementedError):
        rules.TaylorDecompositionRule(lin, ref=(torch.zeros(6, 16, device="cuda"),), bias=True)(x.cuda().requires_grad_())


def test_taylor_rule_on_a_linear_whose_class_forward_is_patched():
    """monkey_patch routes nn.Linear through the B200 GEMM process-wide; the torch.func jvp / vjp of TaylorDecompositionRule must
    still work on such a module (the patched forward steps aside for transformed tensors)"""
    import warnings
    import lxt_b200.explicit.rules as rules
    from lxt_b200.efficient import patches as P
    saved = torch.nn.Linear.forward, torch.nn.Linear.__dict__.get("original_forward")
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            P.patch_method(P.linear_forward, torch.nn.Linear, keep_original=True)
        g = torch.Generator().manual_seed(4)
        x, W, R = torch.rand(6, 16, generator=g) + 0.5, torch.rand(8, 16, generator=g) + 0.5, torch.randn(6, 8, generator=g)
        lin = torch.nn.Linear(16, 8, bias=False).cuda()
        lin.weight.data.copy_(W)
        lin.weight.requires_grad_(False)
        y, (gx0,) = _grad(lin, x, seed=R)                              # plain call: the kernel path
        assert rel_l2(y, x @ W.t()) < 1e-5 and rel_l2(gx0, R @ W) < 1e-5
        rule = rules.TaylorDecompositionRule(lin, ref=(torch.zeros(6, 16, device="cuda"),), bias=False)
        _, (gx,) = _grad(rule, x, seed=R)
        assert rel_l2(gx, O.linear_epsilon_relevance(x, W, None, R, 1e-6)) < 1e-4
    finally:
        torch.nn.Linear.forward = saved[0]
        if saved[1] is None and "original_forward" in torch.nn.Linear.__dict__:
            del torch.nn.Linear.original_forward


def test_flash_attention_relevance_space_equals_the_rule_chain():
    """lf.flash_attention (extension: the explicit rule chain matmul -> mul2 -> add2(mask) -> softmax -> matmul of reference
    lxt/explicit/models/llama.py:378-391 without any [B,H,S,S] tensor) against (a) that chain evaluated rule by rule with the CPU
    oracle in float64 and (b) the Gradient x Input closed form under a causal mask with grouped heads."""
    import lxt_b200.explicit.functional as lf
    g = torch.Generator().manual_seed(23)
    B, H, S, D = 2, 4, 256, 64
    qh = torch.rand(B, H, S, D, generator=g) * 0.5 + 0.1       # positive: keeps O and the scores away from the poles of the rules
    kh = torch.rand(B, H, S, D, generator=g) * 0.5 + 0.1
    vh = torch.rand(B, H, S, D, generator=g) * 0.5 + 0.1
    R = torch.randn(B, H, S, D, generator=g)
    scale, eps = D ** -0.5, 1e-9
    q64, k64, v64, R64 = (t.double() for t in (qh, kh, vh, R))
    A = q64 @ k64.transpose(-1, -2)
    x = A * scale
    P = torch.softmax(x, -1)
    Rp, Rv = O.matmul_relevance(P, v64, R64, eps)
    Rx = O.softmax_relevance(x, Rp)
    Rq, Rkt = O.matmul_relevance(q64, k64.transpose(-1, -2), Rx, eps)      # mul2 by a constant hands the relevance through
    y, (gq, gk, gv) = _grad(lambda a, b, c: lf.flash_attention(a, b, c, None, False, 0, eps), qh, kh, vh, seed=R)
    assert rel_l2(y, (P @ v64).float()) < 1e-5
    for got, want in ((gq, Rq), (gk, Rkt.transpose(-1, -2)), (gv, Rv)):
        assert rel_l2(got, want.float()) < 1e-3
    # conservation of the uniform rule: R_Q + R_K + R_V = R_O summed (the soft-max rule itself is not conservative: it sheds the
    # bias term), checked on the part that is: R_V carries half
    assert abs(float(gv.double().sum()) - 0.5 * float(R64.sum())) < 1e-3 * float(R64.abs().sum())

    # (b) causal, grouped heads, bf16 operands: closed form Q*dQ/4, K*dK/4, V*dV/2 with dO = R/(O + eps/2) by float64 autograd
    B, H, Hkv, S, D = 1, 4, 2, 4096, 128
    qb = (torch.randn(B, H, S, D, generator=g) * 0.5).bfloat16()
    kb = (torch.randn(B, Hkv, S, D, generator=g) * 0.5).bfloat16()
    vb = (torch.rand(B, Hkv, S, D, generator=g) + 0.2).bfloat16()
    Rb = torch.randn(B, H, S, D, generator=g).bfloat16()
    qd, kd, vd = (t.double().requires_grad_() for t in (qb, kb, vb))
    rep = H // Hkv
    sc = (qd @ kd.repeat_interleave(rep, 1).transpose(-1, -2)) * D ** -0.5
    sc = sc.masked_fill(torch.triu(torch.ones(S, S, dtype=torch.bool), 1), float("-inf"))
    od = torch.softmax(sc, -1) @ vd.repeat_interleave(rep, 1)
    od.backward((Rb.double() / (od.detach() + 0.5e-6)))
    want = (qd.detach() * qd.grad / 4, kd.detach() * kd.grad / 4, vd.detach() * vd.grad / 2)
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    y, got = _grad(lambda a, b, c: lf.flash_attention(a, b, c), qb, kb, vb, seed=Rb)
    peak = torch.cuda.max_memory_allocated() - base
    assert peak < B * H * S * S * 2, f"peak {peak} B: a [B,H,S,S] tensor was materialised"
    assert rel_l2(y, od.detach().float()) < 1e-2
    for gg, ww in zip(got, want):
        assert rel_l2(gg, ww.float()) < 2e-2
