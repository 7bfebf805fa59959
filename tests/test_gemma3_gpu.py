"""GPU: Gemma-3 text model through `lxt_b200.efficient.monkey_patch(modeling_gemma3)` — sliding-window + global layers,
per-head q/k RMSNorm, `(1+w)` RMSNorm with the identity rule, GELU-tanh gated MLP — against the golden relevance of the
real reference (tests/golden/make_golden.py --gemma), at head_dim 64 and at Gemma's production head_dim 256 (BASELINE
configs[4]: forward with a 320-column TMEM budget, backward as the two-pass dV / dK + dQ kernels of attn_bwd_v2.cu)."""
import warnings

import numpy as np
import pytest
import torch

from helpers import bf16_from_bits, load_npz, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,head_dim", [("gemma3_tiny.npz", 64), ("gemma3_tiny_d256.npz", 256)])
def test_patched_gemma3_matches_reference(name, head_dim):
    from transformers import Gemma3ForCausalLM, Gemma3TextConfig
    from transformers.models.gemma3 import modeling_gemma3
    from lxt_b200.efficient import monkey_patch
    from lxt_b200 import ops
    z = load_npz(name)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")   # the attention registry / nn.Linear may already be patched by another family
        monkey_patch(modeling_gemma3, verbose=True)
    cfg = Gemma3TextConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=6, num_attention_heads=2,
                           num_key_value_heads=1, head_dim=head_dim, vocab_size=384, sliding_window=48, max_position_embeddings=512,
                           query_pre_attn_scalar=head_dim, rms_norm_eps=1e-6, tie_word_embeddings=True)
    cfg._attn_implementation = "sdpa"
    assert list(z["layer_types"]) == cfg.layer_types and "full_attention" in cfg.layer_types
    from helpers import build_hf
    model = build_hf(Gemma3ForCausalLM, cfg)
    model.load_state_dict({k[3:]: bf16_from_bits(v) for k, v in z.items() if k.startswith("sd_")}, strict=False)
    model = model.cuda().eval()
    for p in model.parameters():
        p.requires_grad_(False)
    ids = torch.from_numpy(z["ids"]).cuda()
    n0 = ops.launch_count()
    emb = model.get_input_embeddings()(ids).detach().requires_grad_()
    logits = model(inputs_embeds=emb, use_cache=False).logits
    mx, mi = logits[:, -1, :].max(-1)
    mx.sum().backward()
    rel = (emb * emb.grad).float().sum(-1).detach().cpu()
    assert ops.launch_count() - n0 > 6 * 20
    assert np.array_equal(mi.cpu().numpy(), z["idx"])
    err = rel_l2(rel, z["rel_fp32"])
    cos = float(torch.nn.functional.cosine_similarity(rel.flatten(), torch.from_numpy(z["rel_fp32"]).flatten(), dim=0))
    print(f"Gemma-3 tiny head_dim {head_dim} (sliding window 48 + global layer): rel-L2 vs reference fp32 = {err:.3e}, cos = {cos:.5f}")
    assert err < 2e-2 and cos > 0.9995   # bf16 HF module graph, 6 layers with 4 norms each


def test_unsupported_head_dim_is_reported_not_faked():
    from lxt_b200 import ops
    from lxt_b200._capi import LrpError
    q = torch.zeros(1, 128, 1, 96, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(LrpError, match="head_dim"):
        ops.attn_fwd(q, q, q, 1.0)


def test_gemma3_4b_attention_shape_long_context():
    """BASELINE configs[4] attention shape: S = 8192, 8 query / 4 kv heads, head_dim 256, sliding window 1024 and global:
    no [B,H,S,S] tensor exists; checked through size-independent properties (row-stochastic soft-max => o is a convex
    combination of v rows; windowed == global when the window covers the sequence; dV column sums)."""
    from lxt_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    B, S, H, Hkv, D = 1, 8192, 8, 4, 256
    qkv = torch.randn(B, S, (H + 2 * Hkv) * D, generator=g, device="cuda").to(torch.bfloat16)
    q = qkv[:, :, : H * D].view(B, S, H, D)
    k = qkv[:, :, H * D: (H + Hkv) * D].view(B, S, Hkv, D)
    v = qkv[:, :, (H + Hkv) * D:].view(B, S, Hkv, D)
    o_w, lse_w = ops.attn_fwd(q, k, v, D ** -0.5, causal=True, window=1024)
    o_g, lse_g = ops.attn_fwd(q, k, v, D ** -0.5, causal=True, window=0)
    o_c, lse_c = ops.attn_fwd(q, k, v, D ** -0.5, causal=True, window=S)      # window >= S is the global mask
    assert torch.isfinite(o_w).all() and torch.isfinite(lse_w).all()
    assert torch.equal(o_c, o_g) and torch.equal(lse_c, lse_g)
    assert torch.equal(o_w[:, :1024], o_g[:, :1024])                           # first `window` rows see the same keys
    assert float(o_w.float().abs().max()) <= float(v.float().abs().max()) + 1e-2
    # uniform d_o => dV[j] = sum_i P[i,j] d_o: the column sums of P over the queries; total mass = number of queries
    d_o = torch.ones(B, S, H, D, dtype=torch.bfloat16, device="cuda")
    dq, dk, dv = ops.attn_bwd(q, k, v, o_g, d_o, lse_g, D ** -0.5, causal=True, window=0, q_div=1.0, k_div=1.0, v_div=1.0)
    mass = dv.float()[..., 0].sum(dim=1)                                      # [B, Hkv]: sum_j sum_{i,h in group} P[i,j]
    assert rel_l2(mass, torch.full_like(mass, float(S * (H // Hkv)))) < 5e-3
    assert torch.isfinite(dq).all() and torch.isfinite(dk).all()
