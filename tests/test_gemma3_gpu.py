"""GPU: Gemma-3 text model through `lxt_b200.efficient.monkey_patch(modeling_gemma3)` — sliding-window + global layers,
per-head q/k RMSNorm, `(1+w)` RMSNorm with the identity rule, GELU-tanh gated MLP — against the golden relevance of the
real reference (tests/golden/make_golden.py --gemma).  head_dim 64 (BASELINE configs[4] at head_dim 256 needs the
round-2 two-pass attention backward; `lrp_attn_*` reports D=256 as unsupported)."""
import warnings

import numpy as np
import pytest
import torch

from helpers import bf16_from_bits, load_npz, rel_l2

pytestmark = pytest.mark.gpu


def test_patched_gemma3_matches_reference():
    from transformers import Gemma3ForCausalLM, Gemma3TextConfig
    from transformers.models.gemma3 import modeling_gemma3
    from lxt_b200.efficient import monkey_patch
    from lxt_b200 import ops
    z = load_npz("gemma3_tiny.npz")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")   # the attention registry / nn.Linear may already be patched by another family
        monkey_patch(modeling_gemma3, verbose=True)
    cfg = Gemma3TextConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=6, num_attention_heads=2,
                           num_key_value_heads=1, head_dim=64, vocab_size=384, sliding_window=48, max_position_embeddings=512,
                           query_pre_attn_scalar=64, rms_norm_eps=1e-6, tie_word_embeddings=True)
    cfg._attn_implementation = "sdpa"
    assert list(z["layer_types"]) == cfg.layer_types and "full_attention" in cfg.layer_types
    model = Gemma3ForCausalLM(cfg).to(torch.bfloat16)
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
    print(f"Gemma-3 tiny (sliding window 48 + global layer): rel-L2 vs reference fp32 = {err:.3e}, cos = {cos:.5f}")
    assert err < 2e-2 and cos > 0.9995   # bf16 HF module graph, 6 layers with 4 norms each


def test_head_dim_256_is_reported_not_faked():
    from lxt_b200 import ops
    from lxt_b200._capi import LrpError
    q = torch.zeros(1, 128, 1, 256, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(LrpError, match="head_dim"):
        ops.attn_fwd(q, q, q, 1.0)
