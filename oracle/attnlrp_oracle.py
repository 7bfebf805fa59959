"""CPU oracle for the AttnLRP hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may import this
module, and only as the checker / the timed CPU baseline.  The product (`lxt_b200`) never imports it and has no
CPU fallback.

What it is: a plain-torch (CPU, no autograd, explicit backward formulas) restatement of the reference algorithm
of rachtibat/LRP-eXplains-Transformers (`lxt` 2.1) for the path SURVEY.md §8 names.  Every function cites the
reference file:line it follows.  Parity status: PINNED — `tests/golden/make_golden.py` imports the real reference
from /root/reference (lxt.efficient.monkey_patch on a HuggingFace Llama, lxt.explicit.functional / rules) and
stores its outputs on seeded inputs; `tests/test_oracle_vs_golden.py` checks this file against those vectors.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

# ------------------------------------------------------------------------------------------------------
# relevance-space rules  (lxt/explicit/functional.py, lxt/explicit/rules.py)
# ------------------------------------------------------------------------------------------------------


def stabilize(z: torch.Tensor, eps: float) -> torch.Tensor:
    """`_stabilize`: plain `z + eps`, no sign handling (lxt/explicit/functional.py:266-273)."""
    return z + eps


def linear_epsilon_relevance(x, W, b, R_out, eps: float = 1e-6):
    """epsilon-LRP for nn.functional.linear (lxt/explicit/functional.py:353-364):
    z = x W^T + b ; R_in = x * ((R_out / (z + eps)) W)."""
    z = F.linear(x, W, b)
    s = R_out / stabilize(z, eps)
    return torch.matmul(s, W) * x


def epsilon_rule_relevance(x, W, b, R_out, eps: float = 1e-8):
    """`EpsilonRule` wrapped around F.linear (lxt/explicit/rules.py:187-222): R/(out+eps) -> VJP -> * input.
    For a linear module the VJP is `s W`, i.e. the same closed form with the rule's own default eps."""
    return linear_epsilon_relevance(x, W, b, R_out, eps)


def matmul_relevance(a, b, R_out, eps: float = 1e-8):
    """epsilon + uniform rule for torch.matmul (lxt/explicit/functional.py:393-408):
    s = R/(2 O + eps); R_a = (s b^T) * a ; R_b = (a^T s) * b."""
    O = torch.matmul(a, b)
    s = R_out / stabilize(O * 2, eps)
    return torch.matmul(s, b.transpose(-1, -2)) * a, torch.matmul(a.transpose(-1, -2), s) * b


def softmax_relevance(x, R_out, dim: int = -1, temperature: float = 1.0):
    """Deep-Taylor softmax rule, Prop 3.1 (lxt/explicit/functional.py:293-322):
    R_in = x * (R_out - p * sum(R_out)), -inf inputs treated as 0."""
    xs = x / temperature
    p = F.softmax(xs, dim=dim)
    xs = torch.where(torch.isneginf(xs), torch.zeros_like(xs), xs)
    return xs * (R_out - p * R_out.sum(dim, keepdim=True))


def add2_relevance(a, b, R_out, eps: float = 1e-8):
    """epsilon rule for a + b (lxt/explicit/functional.py:439-459)."""
    s = R_out / stabilize(a + b, eps)
    return s * a, s * b


def mul2_relevance(R_out, n_requiring_grad: int = 2):
    """uniform rule for a * b (lxt/explicit/functional.py:524-536): each input gets R / (#inputs requiring grad)."""
    return R_out / n_requiring_grad


def rms_norm_identity_relevance(R_out):
    """identity rule on RMSNorm (lxt/explicit/functional.py:481-495)."""
    return R_out


def uniform_epsilon_matmul_relevance(a, b, R_out, eps: float = 1e-6):
    """`UniformEpsilonRule` around a two-input matmul module (lxt/explicit/rules.py:253-282):
    s = R/(out+eps)/n_inputs ; R_i = VJP_i(s) * input_i."""
    O = torch.matmul(a, b)
    s = R_out / stabilize(O, eps) / 2
    return torch.matmul(s, b.transpose(-1, -2)) * a, torch.matmul(a.transpose(-1, -2), s) * b


# ------------------------------------------------------------------------------------------------------
# Gradient x Input rules  (lxt/efficient/rules.py)
# ------------------------------------------------------------------------------------------------------


def identity_rule_implicit_grad(fn_out, x, g):
    """identity rule in GxI space (lxt/efficient/rules.py:88-100): g_in = g_out * f(x)/(x + 1e-10)."""
    return (fn_out / (x + 1e-10)) * g


def divide_gradient_grad(g, factor):
    """uniform rule in GxI space (lxt/efficient/rules.py:125-127)."""
    return g / factor


# ------------------------------------------------------------------------------------------------------
# model-level restatement: Llama forward + AttnLRP (GxI) backward with explicit formulas
#   reference: lxt/efficient/models/llama.py:9-14 (attnLRP map) applied to transformers modeling_llama;
#   workload: examples/quantized_llama.py:35-47 (embed -> forward -> max logit at last position -> backward ->
#   (emb * emb.grad).float().sum(-1)).
# ------------------------------------------------------------------------------------------------------


def rope_tables(S: int, D: int, theta: float, dtype=torch.float32):
    """cos/sin [S, D] as HF builds them (transformers modeling_llama.py: LlamaRotaryEmbedding.forward)."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, D, 2, dtype=torch.int64).float() / D))
    pos = torch.arange(S, dtype=torch.float32)
    freqs = torch.outer(pos, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def _rotate_half(x):
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def _rotate_half_T(y):
    """transpose of rotate_half (its backward)"""
    y1, y2 = y[..., : y.shape[-1] // 2], y[..., y.shape[-1] // 2:]
    return torch.cat((y2, -y1), dim=-1)


def _rmsnorm_fwd(h, w, eps):
    """lxt/efficient/patches.py:111-123 (`rms_norm_forward`)."""
    dt = h.dtype
    hf = h.to(torch.float32)
    rstd = torch.rsqrt(hf.pow(2).mean(-1, keepdim=True) + eps)
    return w * (hf * rstd).to(dt), rstd


def _rmsnorm_bwd(g, w, rstd):
    """variance detached: g_x = (g * w) * rstd (fp32 inside, result in g's dtype)."""
    dt = g.dtype
    return ((g * w).to(torch.float32) * rstd).to(dt)


def llama_attnlrp(weights: Dict, ids: torch.Tensor, cfg: Dict, dtype=torch.float32, return_aux: bool = False,
                  rule: str = "attnlrp"):
    """One AttnLRP attribution per prompt, entirely with explicit formulas (no autograd).

    weights: {'emb','norm','lm_head', 'layers':[{'wq','wk','wv','wo','wg','wu','wd','ln1','ln2'}]} (any float dtype)
    ids: int64 [B,S].  cfg: d, H, Hkv, D, eps, theta.  Returns relevance fp32 [B,S] (and aux dict).
    rule="cp" restates the CP-LRP map (lxt/efficient/models/llama.py:16-21; patches.py:228-280): q,k detached
    before attention, MLP gate detached and no uniform split on the product.
    """
    cp = rule == "cp"
    H, Hkv, D, eps = cfg["H"], cfg["Hkv"], cfg["D"], cfg["eps"]
    G = H // Hkv
    B, S = ids.shape
    scale = 1.0 / math.sqrt(D)
    W = lambda t: t.to(dtype)
    cos, sin = rope_tables(S, D, cfg["theta"], dtype)
    mask = torch.full((S, S), float("-inf")).triu(1)

    emb = W(weights["emb"])[ids]  # [B,S,d]
    h = emb
    stash: List[Dict] = []
    for lw in weights["layers"]:
        st = {}
        xn, st["rstd1"] = _rmsnorm_fwd(h, W(lw["ln1"]), eps)
        q = (xn @ W(lw["wq"]).T).view(B, S, H, D).transpose(1, 2)
        k = (xn @ W(lw["wk"]).T).view(B, S, Hkv, D).transpose(1, 2)
        v = (xn @ W(lw["wv"]).T).view(B, S, Hkv, D).transpose(1, 2)
        q = q * cos + _rotate_half(q) * sin  # transformers modeling_llama.py:146-168
        k = k * cos + _rotate_half(k) * sin
        kr = k.repeat_interleave(G, dim=1)
        vr = v.repeat_interleave(G, dim=1)
        sc = (q @ kr.transpose(-1, -2)) * scale + mask
        P = F.softmax(sc, dim=-1, dtype=torch.float32).to(dtype)  # modeling_llama.py:199-222 (eager path)
        o = (P @ vr).transpose(1, 2).reshape(B, S, H * D)
        st.update(q=q, kr=kr, vr=vr, P=P)
        h = h + o @ W(lw["wo"]).T
        xn2, st["rstd2"] = _rmsnorm_fwd(h, W(lw["ln2"]), eps)
        gate = xn2 @ W(lw["wg"]).T
        up = xn2 @ W(lw["wu"]).T
        s = F.silu(gate)
        st.update(gate=gate, up=up, s=s)
        h = h + (s * up) @ W(lw["wd"]).T  # lxt/efficient/patches.py:145-157 (`gated_mlp_forward`)
        st["h_out"] = h
        stash.append(st)
    hN, rstdN = _rmsnorm_fwd(h, W(weights["norm"]), eps)
    logits = hN[:, -1, :] @ W(weights["lm_head"]).T  # only the last position is read (quantized_llama.py:40)
    idx = logits.float().argmax(-1)

    # ---------------- LRP backward in Gradient x Input space ----------------
    g_h = torch.zeros_like(h)
    g_hN_last = W(weights["lm_head"])[idx]  # d(logit_max)/d(hN_last)
    g_h[:, -1, :] = _rmsnorm_bwd(g_hN_last, W(weights["norm"]), rstdN[:, -1, :])
    trace = []
    for lw, st in zip(reversed(weights["layers"]), reversed(stash)):
        # latent relevance of this layer's output: output * grad (docs/source/latent-feature-attribution-efficient.rst:49-90)
        trace.append((st["h_out"] * g_h).float().sum(-1))
        # gated MLP: uniform rule on the product, identity rule on SiLU (patches.py:145-157, rules.py:88-127)
        if cp:
            g_xn2 = ((g_h @ W(lw["wd"])) * st["s"]) @ W(lw["wu"])
        else:
            g_a = divide_gradient_grad(g_h @ W(lw["wd"]), 2)
            g_s = g_a * st["up"]
            g_up = g_a * st["s"]
            g_gate = identity_rule_implicit_grad(st["s"], st["gate"], g_s)
            g_xn2 = g_gate @ W(lw["wg"]) + g_up @ W(lw["wu"])
        g_h = g_h + _rmsnorm_bwd(g_xn2, W(lw["ln2"]), st["rstd2"])
        # attention: ordinary softmax-attention backward, then dQ/4, dK/4, dV/2 (patches.py:193-203)
        g_o = (g_h @ W(lw["wo"])).view(B, S, H, D).transpose(1, 2)
        P, q, kr, vr = st["P"], st["q"], st["kr"], st["vr"]
        dV = P.transpose(-1, -2) @ g_o
        dP = (g_o @ vr.transpose(-1, -2)).to(torch.float32)
        Pf = P.to(torch.float32)
        dS = (Pf * (dP - (dP * Pf).sum(-1, keepdim=True))).to(dtype) * scale
        dQ = dS @ kr
        dK = dS.transpose(-1, -2) @ q
        dK = dK.view(B, Hkv, G, S, D).sum(2)
        dV = dV.view(B, Hkv, G, S, D).sum(2)
        dQ, dK, dV = (dQ * 0, dK * 0, dV) if cp else (dQ / 4, dK / 4, dV / 2)
        dQ = dQ * cos + _rotate_half_T(dQ * sin)
        dK = dK * cos + _rotate_half_T(dK * sin)
        g_xn = (dQ.transpose(1, 2).reshape(B, S, H * D) @ W(lw["wq"])
                + dK.transpose(1, 2).reshape(B, S, Hkv * D) @ W(lw["wk"])
                + dV.transpose(1, 2).reshape(B, S, Hkv * D) @ W(lw["wv"]))
        g_h = g_h + _rmsnorm_bwd(g_xn, W(lw["ln1"]), st["rstd1"])
    rel = (emb * g_h).float().sum(-1)  # quantized_llama.py:47
    if return_aux:
        return rel, {"idx": idx, "logits": logits.float(), "g_emb": g_h, "layer_relevance": torch.stack(trace[::-1])}
    return rel


def _act(x, kind):
    return F.silu(x) if kind == "silu" else F.gelu(x, approximate="tanh")


def decoder_attnlrp(weights: Dict, ids: torch.Tensor, cfg: Dict, dtype=torch.float32, return_aux: bool = False):
    """Generalised decoder restatement (explicit formulas, no autograd) covering Llama and Gemma-3 text models under the
    reference's attnLRP maps (lxt/efficient/models/llama.py:9-14, gemma3.py:11-19).  Options in `cfg` beyond d,H,Hkv,D,eps:
      norm_offset  0 (Llama `w*x_hat`) or 1 (Gemma `(1+w)*x_hat`, gemma3.py:11-12 + HF Gemma3RMSNorm.forward)
      act          "silu" | "gelu_tanh"           qk_norm     per-head RMSNorm on q,k before RoPE (weights 'qn','kn' [D])
      post_norms   Gemma layer layout: h += post_norm(branch(pre_norm(h)))  (weights 'ln_post_attn','ln_pre_ff','ln_post_ff')
      windows      per-layer sliding window (0 = global), thetas: per-layer RoPE base
      attn_scale   soft-max scale (default D^-0.5), emb_scale: embedding multiplier (Gemma sqrt(d))
    """
    H, Hkv, D, eps = cfg["H"], cfg["Hkv"], cfg["D"], cfg["eps"]
    off, act = float(cfg.get("norm_offset", 0)), cfg.get("act", "silu")
    qk_norm, post = bool(cfg.get("qk_norm", False)), bool(cfg.get("post_norms", False))
    L = len(weights["layers"])
    windows = cfg.get("windows") or [0] * L
    thetas = cfg.get("thetas") or [cfg["theta"]] * L
    scale = cfg.get("attn_scale") or 1.0 / math.sqrt(D)
    emb_scale = cfg.get("emb_scale", 1.0)
    G = H // Hkv
    B, S = ids.shape
    W = lambda t: t.to(dtype)

    def norm_f(x, w):  # rms norm with detached variance; returns (y, rstd)
        xf = x.to(torch.float32)
        r = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
        return ((xf * r) * (off + w.to(torch.float32))).to(x.dtype), r

    def norm_b(g, w, r):  # identity rule: g * (off + w) * rstd
        return (g.to(torch.float32) * (off + w.to(torch.float32)) * r).to(g.dtype)

    i = torch.arange(S)
    emb = (W(weights["emb"])[ids] * torch.tensor(emb_scale, dtype=dtype)) if emb_scale != 1.0 else W(weights["emb"])[ids]
    h = emb
    stash = []
    for l, lw in enumerate(weights["layers"]):
        st = {}
        cos, sin = rope_tables(S, D, thetas[l], dtype)
        mask = torch.zeros(S, S)
        bad = i[None, :] > i[:, None]
        if windows[l]:
            bad = bad | ((i[:, None] - i[None, :]) >= windows[l])
        mask = mask.masked_fill(bad, float("-inf"))
        xn, st["r1"] = norm_f(h, W(lw["ln1"]))
        q = (xn @ W(lw["wq"]).T).view(B, S, H, D).transpose(1, 2)
        k = (xn @ W(lw["wk"]).T).view(B, S, Hkv, D).transpose(1, 2)
        v = (xn @ W(lw["wv"]).T).view(B, S, Hkv, D).transpose(1, 2)
        if "bq" in lw:
            q = q + W(lw["bq"]).view(1, H, 1, D); k = k + W(lw["bk"]).view(1, Hkv, 1, D); v = v + W(lw["bv"]).view(1, Hkv, 1, D)
        if qk_norm:
            q, st["rq"] = norm_f(q, W(lw["qn"]))
            k, st["rk"] = norm_f(k, W(lw["kn"]))
        q = q * cos + _rotate_half(q) * sin
        k = k * cos + _rotate_half(k) * sin
        kr, vr = k.repeat_interleave(G, 1), v.repeat_interleave(G, 1)
        P = F.softmax((q @ kr.transpose(-1, -2)) * scale + mask, dim=-1, dtype=torch.float32).to(dtype)
        o = (P @ vr).transpose(1, 2).reshape(B, S, H * D)
        st.update(q=q, kr=kr, vr=vr, P=P, cos=cos, sin=sin)
        a_out = o @ W(lw["wo"]).T
        if post:
            a_out, st["rpa"] = norm_f(a_out, W(lw["ln_post_attn"]))
        h = h + a_out
        xn2, st["r2"] = norm_f(h, W(lw["ln_pre_ff"] if post else lw["ln2"]))
        gate, up = xn2 @ W(lw["wg"]).T, xn2 @ W(lw["wu"]).T
        sg = _act(gate, act)
        st.update(gate=gate, up=up, s=sg)
        m_out = (sg * up) @ W(lw["wd"]).T
        if post:
            m_out, st["rpf"] = norm_f(m_out, W(lw["ln_post_ff"]))
        h = h + m_out
        stash.append(st)
    hN, rN = norm_f(h, W(weights["norm"]))
    logits = hN[:, -1, :] @ W(weights["lm_head"]).T
    idx = logits.float().argmax(-1)

    g_h = torch.zeros_like(h)
    g_h[:, -1, :] = norm_b(W(weights["lm_head"])[idx], W(weights["norm"]), rN[:, -1, :])
    for lw, st in zip(reversed(weights["layers"]), reversed(stash)):
        g_m = norm_b(g_h, W(lw["ln_post_ff"]), st["rpf"]) if post else g_h
        g_a = divide_gradient_grad(g_m @ W(lw["wd"]), 2)
        g_gate = identity_rule_implicit_grad(st["s"], st["gate"], g_a * st["up"])
        g_xn2 = g_gate @ W(lw["wg"]) + (g_a * st["s"]) @ W(lw["wu"])
        g_h = g_h + norm_b(g_xn2, W(lw["ln_pre_ff"] if post else lw["ln2"]), st["r2"])
        g_ao = norm_b(g_h, W(lw["ln_post_attn"]), st["rpa"]) if post else g_h
        g_o = (g_ao @ W(lw["wo"])).view(B, S, H, D).transpose(1, 2)
        P, q, kr, vr, cos, sin = st["P"], st["q"], st["kr"], st["vr"], st["cos"], st["sin"]
        dV = P.transpose(-1, -2) @ g_o
        dP = (g_o @ vr.transpose(-1, -2)).to(torch.float32)
        Pf = P.to(torch.float32)
        dS = (Pf * (dP - (dP * Pf).sum(-1, keepdim=True))).to(dtype) * scale
        dQ = (dS @ kr) / 4
        dK = (dS.transpose(-1, -2) @ q).view(B, Hkv, G, S, D).sum(2) / 4
        dV = dV.view(B, Hkv, G, S, D).sum(2) / 2
        dQ = dQ * cos + _rotate_half_T(dQ * sin)
        dK = dK * cos + _rotate_half_T(dK * sin)
        if qk_norm:
            dQ = norm_b(dQ, W(lw["qn"]), st["rq"])
            dK = norm_b(dK, W(lw["kn"]), st["rk"])
        g_xn = (dQ.transpose(1, 2).reshape(B, S, H * D) @ W(lw["wq"]) + dK.transpose(1, 2).reshape(B, S, Hkv * D) @ W(lw["wk"])
                + dV.transpose(1, 2).reshape(B, S, Hkv * D) @ W(lw["wv"]))
        g_h = g_h + norm_b(g_xn, W(lw["ln1"]), st["r1"])
    rel = (emb * g_h).float().sum(-1)
    if return_aux:
        return rel, {"idx": idx, "logits": logits.float(), "g_emb": g_h}
    return rel


def random_llama_weights(cfg: Dict, seed: int = 0, std: float = 0.02, dtype=torch.bfloat16) -> Dict:
    """HF-style random init (normal(0, 0.02), norm weights 1) for the synthetic workloads."""
    g = torch.Generator().manual_seed(seed)
    d, I, H, Hkv, D, V = cfg["d"], cfg["I"], cfg["H"], cfg["Hkv"], cfg["D"], cfg["V"]
    rn = lambda *s: (torch.randn(*s, generator=g) * std).to(dtype)
    layers = []
    for _ in range(cfg["L"]):
        layers.append(dict(wq=rn(H * D, d), wk=rn(Hkv * D, d), wv=rn(Hkv * D, d), wo=rn(d, H * D), wg=rn(I, d),
                           wu=rn(I, d), wd=rn(d, I), ln1=torch.ones(d, dtype=dtype), ln2=torch.ones(d, dtype=dtype)))
    return dict(emb=rn(V, d), norm=torch.ones(d, dtype=dtype), lm_head=rn(V, d), layers=layers)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-300))
