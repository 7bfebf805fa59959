"""GPU: every C-ABI kernel against a plain PyTorch fp32 reference of the same op on the same bf16 inputs.
Tolerances (rel-L2): fp32-output kernels 1e-5; bf16-output kernels 3e-3 (one bf16 rounding is 1.6e-3 rel-L2)."""
import math

import pytest
import torch

from helpers import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from lxt_b200 import ops as o
    return o


def rnd(*shape, scale=1.0, dtype=torch.bfloat16, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(dtype)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 520, 200), (1000, 1024, 4096), (64, 2048, 512), (4224, 768, 384)])
@pytest.mark.parametrize("layout", [0, 1])
def test_gemm_plain_and_fused_epilogue(ops, M, N, K, layout):
    a = rnd(M, K, seed=1)
    b = rnd(N, K, seed=2) if layout == 0 else rnd(K, N, seed=2)
    ref = a.float() @ (b.float().T if layout == 0 else b.float())
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    ops.gemm(a, b, out, b_layout=layout)
    assert rel_l2(out.float(), ref) < 3e-3
    # fp32 out + residual + row/col scale + bias + bf16 shadow
    res, rs, cs, bias = rnd(M, N, dtype=torch.float32, seed=3), rnd(M, dtype=torch.float32, seed=4), rnd(N, dtype=torch.float32, seed=5), rnd(N, dtype=torch.float32, seed=6)
    out32 = torch.empty(M, N, dtype=torch.float32, device="cuda")
    sh = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    ops.gemm(a, b, out32, b_layout=layout, resid=res, rowscale=rs, colscale=cs, bias=bias, shadow=sh, alpha=0.5)
    ref2 = res + 0.5 * ref * rs[:, None] * cs[None, :] + bias[None, :]
    assert rel_l2(out32, ref2) < 1e-5
    assert rel_l2(sh.float(), ref2) < 3e-3


def test_gemm_strided_views_and_inplace_accumulate(ops):
    T, I, d = 384, 256, 128
    gu = rnd(T, 2 * I, seed=7)
    w = rnd(I, d, seed=8)
    acc = rnd(T, d, dtype=torch.float32, seed=9)
    ref = acc + gu[:, I:].float() @ w.float()
    ops.linear_dgrad(gu[:, I:], w, acc, resid=acc)  # A is a strided view, output accumulates in place
    assert rel_l2(acc, ref) < 1e-5


def test_gemm_argument_errors_are_loud(ops):
    from lxt_b200._capi import LrpError
    a, b = rnd(16, 12), rnd(16, 12)
    with pytest.raises(LrpError):
        ops.gemm(a, b, torch.empty(16, 16, dtype=torch.bfloat16, device="cuda"), b_layout=0)  # K % 8 != 0
    with pytest.raises(LrpError):
        ops.gemm(rnd(16, 16), rnd(16, 8), torch.empty(16, 16, dtype=torch.bfloat16, device="cuda"), b_layout=0)


def _attn_ref(q, k, v, d_o, scale, causal, window):
    """fp32 torch reference; q [B,S,H,D], k/v [B,S,Hkv,D]"""
    B, S, H, D = q.shape
    G = H // k.shape[2]
    qf, kf, vf = (t.float().transpose(1, 2).requires_grad_() for t in (q, k, v))
    kr, vr = kf.repeat_interleave(G, 1), vf.repeat_interleave(G, 1)
    sc = qf @ kr.transpose(-1, -2) * scale
    i = torch.arange(S, device=q.device)
    m = torch.zeros(S, S, dtype=torch.bool, device=q.device)
    if causal:
        m |= i[None, :] > i[:, None]
    if window:
        m |= (i[:, None] - i[None, :]) >= window
    sc = sc.masked_fill(m, float("-inf"))
    p = sc.softmax(-1)
    o = (p @ vr).transpose(1, 2)
    o.backward(d_o.float())
    return o.detach(), torch.logsumexp(sc, -1).detach(), qf.grad.transpose(1, 2), kf.grad.transpose(1, 2), vf.grad.transpose(1, 2)


@pytest.mark.parametrize("B,S,H,Hkv,D,causal,window", [(1, 128, 1, 1, 128, True, 0), (2, 300, 4, 2, 128, True, 0),
                                                       (2, 300, 4, 2, 64, True, 0), (2, 197, 4, 4, 64, False, 0),
                                                       (1, 520, 4, 1, 128, True, 200), (1, 1024, 8, 2, 128, True, 0)])
def test_flash_attnlrp_fwd_bwd(ops, B, S, H, Hkv, D, causal, window):
    qkv = rnd(B, S, (H + 2 * Hkv) * D, scale=1.0, seed=11)  # packed buffer: kernels read strided views, no copies
    q = qkv[:, :, : H * D].view(B, S, H, D)
    k = qkv[:, :, H * D: (H + Hkv) * D].view(B, S, Hkv, D)
    v = qkv[:, :, (H + Hkv) * D:].view(B, S, Hkv, D)
    d_o = rnd(B, S, H, D, seed=12)
    scale = 1 / math.sqrt(D)
    o, lse = ops.attn_fwd(q, k, v, scale, causal=causal, window=window)
    dq, dk, dv = ops.attn_bwd(q, k, v, o, d_o, lse, scale, causal=causal, window=window)
    ro, rlse, rdq, rdk, rdv = _attn_ref(q, k, v, d_o, scale, causal, window)
    assert rel_l2(o.float(), ro) < 5e-3
    assert rel_l2(lse, rlse) < 1e-5
    # AttnLRP: uniform rule = dQ/4, dK/4, dV/2 (lxt/efficient/patches.py:193-203)
    assert rel_l2(dq.float(), rdq / 4) < 8e-3
    assert rel_l2(dk.float(), rdk / 4) < 8e-3
    assert rel_l2(dv.float(), rdv / 2) < 8e-3
    # CP-LRP: q,k detached
    dq0, dk0, dv1 = ops.attn_bwd(q, k, v, o, d_o, lse, scale, causal=causal, window=window, q_div=0.0, k_div=0.0, v_div=1.0)
    assert float(dq0.abs().max()) == 0.0 and float(dk0.abs().max()) == 0.0
    assert rel_l2(dv1.float(), rdv) < 8e-3


def test_rmsnorm_fwd_bwd(ops):
    T, d = 777, 1024
    for xdt in (torch.float32, torch.bfloat16):
        x, w, g = rnd(T, d, dtype=xdt, seed=21), (1 + 0.1 * rnd(d, dtype=torch.float32, seed=22)).bfloat16(), rnd(T, d, seed=23)
        for w_off in (0.0, 1.0):
            y, rstd = ops.rmsnorm_fwd(x, w, 1e-5, w_offset=w_off)
            xf = x.float()
            r = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5)
            assert rel_l2(rstd, r.squeeze(-1)) < 1e-6
            assert rel_l2(y.float(), xf * r * (w.float() + w_off)) < 3e-3
            gx = ops.rmsnorm_bwd(g, w, rstd, w_offset=w_off, out_dtype=torch.float32)
            assert rel_l2(gx, g.float() * (w.float() + w_off) * r) < 1e-6


def test_layernorm_detached_std(ops):
    T, d = 394, 1024
    for dt, tol in ((torch.float32, 1e-5), (torch.bfloat16, 4e-3)):
        x, w, b, g = rnd(T, d, dtype=dt, seed=31), rnd(d, dtype=dt, seed=32), rnd(d, dtype=dt, seed=33), rnd(T, d, dtype=dt, seed=34)
        y, mean, rstd = ops.layernorm_fwd(x, w, b, 1e-6)
        xf = x.float().requires_grad_()
        mu = xf.mean(-1, keepdim=True)
        std = ((xf - mu) ** 2).mean(-1, keepdim=True).add(1e-6).sqrt()
        yr = (xf - mu) / std.detach() * w.float() + b.float()  # lxt/efficient/patches.py:126-142
        yr.backward(g.float())
        assert rel_l2(y.float(), yr.detach()) < tol
        assert rel_l2(ops.layernorm_bwd(g, w, rstd).float(), xf.grad) < tol


def test_rope_forward_and_transpose(ops):
    from oracle.attnlrp_oracle import rope_tables, _rotate_half, _rotate_half_T
    B, S, Hn, D = 2, 200, 6, 128
    x = rnd(B * S, Hn * D + 256, seed=41)
    cos, sin = rope_tables(S, D, 10000.0)
    c, s = cos[:, : D // 2].cuda().contiguous(), sin[:, : D // 2].cuda().contiguous()
    ref_in = x[:, : Hn * D].float().view(B, S, Hn, D).transpose(1, 2)
    ref = ref_in * cos.cuda() + _rotate_half(ref_in) * sin.cuda()
    y = x.clone()
    ops.rope_inplace(y, Hn, D, c, s, S)
    assert rel_l2(y[:, : Hn * D].float().view(B, S, Hn, D).transpose(1, 2), ref) < 3e-3
    assert torch.equal(y[:, Hn * D:], x[:, Hn * D:])  # columns beyond the rotated heads are untouched
    g = x.clone()
    ops.rope_inplace(g, Hn, D, c, s, S, inverse=True)
    refT = ref_in * cos.cuda() + _rotate_half_T(ref_in * sin.cuda())
    assert rel_l2(g[:, : Hn * D].float().view(B, S, Hn, D).transpose(1, 2), refT) < 3e-3


def test_gated_act_rules(ops):
    T, I = 500, 1536
    gu, ga = rnd(T, 2 * I, scale=2.0, seed=51), rnd(T, I, seed=52)
    gate, up = gu[:, :I].float(), gu[:, I:].float()
    s = torch.nn.functional.silu(gate)
    assert rel_l2(ops.gated_act_fwd(gu).float(), s * up) < 3e-3
    ggu = ops.gated_act_bwd(ga, gu).float()
    gh = ga.float() / 2                                   # divide_gradient (rules.py:125-127)
    assert rel_l2(ggu[:, I:], gh * s) < 3e-3              # product rule, up branch
    assert rel_l2(ggu[:, :I], gh * up * (s / (gate + 1e-10))) < 3e-3   # identity rule on SiLU (rules.py:88-100)


def test_path_ends(ops):
    V, d, T = 1000, 512, 300
    emb = rnd(V, d, seed=61)
    ids = torch.randint(0, V, (T,), device="cuda")
    h = ops.embed_gather(ids, emb)
    assert torch.equal(h, emb[ids].float())
    logits = rnd(7, 5003, dtype=torch.float32, seed=62)
    idx, val = ops.argmax_rows(logits)
    assert torch.equal(idx.long(), logits.argmax(-1)) and torch.equal(val, logits.max(-1).values)
    g = rnd(T, d, dtype=torch.float32, seed=63)
    assert rel_l2(ops.gxi_reduce(h, g), (h * g).sum(-1)) < 1e-6
    assert torch.equal(ops.cast_bf16(g), g.bfloat16())


@pytest.mark.parametrize("cp", [False, True])
def test_fused_gated_backward_epilogue_equals_unfused(ops, cp):
    """down-proj LRP dgrad with the gated-MLP rules in its epilogue == dgrad followed by lrp_gated_act_bwd"""
    T, d, I = 640, 512, 1280
    gy, w, gu = rnd(T, d, seed=71), rnd(d, I, scale=0.05, seed=72), rnd(T, 2 * I, scale=2.0, seed=73)
    ga = torch.empty(T, I, dtype=torch.bfloat16, device="cuda")
    ops.linear_dgrad(gy, w, ga)
    ref = ops.gated_act_bwd(ga, gu, cp=cp)
    got = ops.linear_dgrad_gated_bwd(gy, w, gu, torch.empty_like(gu), cp=cp)
    # the unfused path rounds g_a to bf16 before the point-wise rule, the fused one does not
    assert rel_l2(got.float(), ref.float()) < 4e-3
    ga32 = gy.float() @ w.float()
    gate, up = gu[:, :I].float(), gu[:, I:].float()
    s = torch.nn.functional.silu(gate)
    exp = torch.cat([torch.zeros_like(gate), ga32 * s], 1) if cp else torch.cat([(s / (gate + 1e-10)) * (ga32 / 2 * up), ga32 / 2 * s], 1)
    assert rel_l2(got.float(), exp) < 3e-3
    # interleaved (gate | up) layout + GELU-tanh: the form the engine uses
    for act in (0, 1):
        blk = lambda t: torch.stack([t[:, :I].reshape(T, I // 32, 32), t[:, I:].reshape(T, I // 32, 32)], 2).reshape(T, 2 * I).contiguous()
        gu_il = blk(gu)
        got_il = ops.linear_dgrad_gated_bwd(gy, w, gu_il, torch.empty_like(gu), act=act, cp=cp, layout=1)
        ref_il = ops.gated_act_bwd(ga, gu_il, act, cp=cp, layout=1)
        assert rel_l2(got_il.float(), ref_il.float()) < 4e-3
        if act == 0:
            assert rel_l2(got_il.float(), blk(exp)) < 3e-3


@pytest.mark.parametrize("shape", [(300, 512, 256), (4096, 28672 // 4, 1024), (257, 1152, 192)])
@pytest.mark.parametrize("act", [0, 1])
def test_gate_up_gemm_with_fused_activation_epilogue(shape, act):
    """gate|up forward with a = act(gate) * up leaving the GEMM epilogue (interleaved 32-row blocks): gu and a must be
    bit-identical to the unfused pair (plain GEMM, then lrp_gated_act_fwd on the stored bf16 values)."""
    from lxt_b200 import ops
    T, I, K = shape
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(T, K, generator=g, device="cuda").to(torch.bfloat16)
    wg = (torch.randn(I, K, generator=g, device="cuda") * 0.05).to(torch.bfloat16)
    wu = (torch.randn(I, K, generator=g, device="cuda") * 0.05).to(torch.bfloat16)
    w_il = torch.stack([wg.view(I // 32, 32, K), wu.view(I // 32, 32, K)], 1).reshape(2 * I, K).contiguous()
    gu_f = torch.empty(T, 2 * I, dtype=torch.bfloat16, device="cuda")
    a_f = torch.empty(T, I, dtype=torch.bfloat16, device="cuda")
    ops.linear_fwd(x, w_il, gu_f, act_out=a_f, act=act)
    gu_u = torch.empty_like(gu_f)
    ops.linear_fwd(x, w_il, gu_u)
    a_u = ops.gated_act_fwd(gu_u, act, layout=1)
    assert torch.equal(gu_f, gu_u)
    # the fused epilogue evaluates the activation with ex2 / rcp / tanh approximations (fp32 relative error ~1e-6), the point-wise
    # kernel with IEEE division: the bf16 results agree except where that difference crosses a rounding boundary (1 bf16 ulp)
    da = (a_f.float() - a_u.float()).abs()
    # (absolute floor: in GELU-tanh's far negative tail 1 + tanh(u) cancels in the fp32 of the point-wise kernel and of torch,
    #  the x * sigmoid(2u) form of the epilogue does not; the values there are ~1e-6 of the typical magnitude)
    assert bool((da <= 0.0079 * a_u.float().abs() + 2e-6 * float(a_u.float().abs().max())).all())
    assert float((da > 0).float().mean()) < 0.02
    # and against torch on the de-interleaved halves
    gate = (x.float() @ wg.float().T).to(torch.bfloat16).float()
    up = (x.float() @ wu.float().T).to(torch.bfloat16).float()
    ref = (torch.nn.functional.silu(gate) if act == 0 else torch.nn.functional.gelu(gate, approximate="tanh")) * up
    err = float((a_f.float() - ref).norm() / ref.norm())
    assert err < 6e-3
    # interleaved backward layout == halves layout after de-interleaving
    ga = torch.randn(T, I, generator=g, device="cuda").to(torch.bfloat16)
    ggu_il = ops.gated_act_bwd(ga, gu_u, act, layout=1)
    blk = gu_u.view(T, I // 32, 2, 32)
    gu_h = torch.cat([blk[:, :, 0].reshape(T, I), blk[:, :, 1].reshape(T, I)], 1).contiguous()
    ggu_h = ops.gated_act_bwd(ga, gu_h, act, layout=0)
    b2 = ggu_il.view(T, I // 32, 2, 32)
    assert torch.equal(torch.cat([b2[:, :, 0].reshape(T, I), b2[:, :, 1].reshape(T, I)], 1), ggu_h)


@pytest.mark.parametrize("B,S,H,Hkv,D,causal,window", [(3, 300, 4, 2, 128, True, 0), (3, 520, 4, 2, 64, True, 0), (2, 300, 2, 1, 256, True, 0),
                                                       (3, 333, 4, 4, 128, False, 0), (2, 520, 4, 1, 64, True, 200), (2, 2100, 8, 2, 128, True, 0)])
def test_flash_attnlrp_key_padding_ranges(ops, B, S, H, Hkv, D, causal, window):
    """batches of prompts of different lengths: one valid-key range [lo, hi) per sequence (left / right padding / both) in every
    attention kernel (persistent forward D=128, first-generation forward D=64/256, pipelined backward, two-pass backward D=256) vs an
    fp32 torch reference with the corresponding boolean mask.  Query rows that see no key give o = 0, lse = -inf, zero gradients."""
    qkv = rnd(B, S, (H + 2 * Hkv) * D, scale=1.0, seed=31)
    q = qkv[:, :, : H * D].view(B, S, H, D)
    k = qkv[:, :, H * D: (H + Hkv) * D].view(B, S, Hkv, D)
    v = qkv[:, :, (H + Hkv) * D:].view(B, S, Hkv, D)
    d_o = rnd(B, S, H, D, seed=32)
    scale = 1 / math.sqrt(D)
    rng = torch.tensor([[0, S], [S // 3 + 5, S], [7, S - S // 4]][:B], dtype=torch.int32, device="cuda")   # none, left, both
    o, lse = ops.attn_fwd(q, k, v, scale, causal=causal, window=window, kv_range=rng)
    dq, dk, dv = ops.attn_bwd(q, k, v, o, d_o, lse, scale, causal=causal, window=window, kv_range=rng)
    # reference with the same mask
    G = H // Hkv
    qf, kf, vf = (t.float().transpose(1, 2).requires_grad_() for t in (q, k, v))
    kr, vr = kf.repeat_interleave(G, 1), vf.repeat_interleave(G, 1)
    sc = qf @ kr.transpose(-1, -2) * scale
    i = torch.arange(S, device="cuda")
    m = torch.zeros(B, 1, S, S, dtype=torch.bool, device="cuda")
    if causal:
        m |= (i[None, :] > i[:, None])[None, None]
    if window:
        m |= ((i[:, None] - i[None, :]) >= window)[None, None]
    m |= ((i[None, :] < rng[:, 0, None]) | (i[None, :] >= rng[:, 1, None]))[:, None, None, :]
    dead = m.all(-1, keepdim=True)                                           # query rows that see nothing
    p = sc.masked_fill(m, float("-inf")).masked_fill(dead, 0.0).softmax(-1).masked_fill(dead, 0.0)
    oref = (p @ vr).transpose(1, 2)
    oref.backward(d_o.float())
    live = (~dead.squeeze(-1)).transpose(1, 2).unsqueeze(-1).expand(B, S, H, 1)   # [B,S,H,1]
    assert rel_l2(o.float(), oref.detach()) < 5e-3
    assert bool((o.float().abs().amax(-1, keepdim=True)[~live] == 0).all())
    lse_ref = torch.logsumexp(sc.masked_fill(m, float("-inf")), -1)
    ok = torch.isfinite(lse_ref)
    assert rel_l2(lse[ok], lse_ref[ok]) < 1e-5 and bool((lse[~ok] == float("-inf")).all())
    assert rel_l2(dq.float(), qf.grad.transpose(1, 2) / 4) < 8e-3
    assert rel_l2(dk.float(), kf.grad.transpose(1, 2) / 4) < 8e-3
    assert rel_l2(dv.float(), vf.grad.transpose(1, 2) / 2) < 8e-3
