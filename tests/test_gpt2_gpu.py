"""GPU: GPT-2 through `lxt_b200.efficient.monkey_patch(modeling_gpt2)` (LayerNorm with the identity rule, plain GELU MLP
with the identity rule on the activation, flash AttnLRP) against the golden relevance of the real reference
(lxt/efficient/models/gpt2.py:11-32; tests/golden/make_golden.py --gpt2)."""
import warnings

import numpy as np
import pytest
import torch

from helpers import bf16_from_bits, load_npz, rel_l2

pytestmark = pytest.mark.gpu


def test_patched_gpt2_matches_reference():
    from transformers import GPT2Config, GPT2LMHeadModel
    from transformers.models.gpt2 import modeling_gpt2
    from lxt_b200.efficient import monkey_patch
    from lxt_b200 import ops
    z = load_npz("gpt2_tiny.npz")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        monkey_patch(modeling_gpt2, verbose=True)
    cfg = GPT2Config(n_embd=128, n_head=2, n_layer=2, vocab_size=384, n_positions=256, resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
    cfg._attn_implementation = "sdpa"
    model = GPT2LMHeadModel(cfg).to(torch.bfloat16)
    model.load_state_dict({k[3:]: bf16_from_bits(v) for k, v in z.items() if k.startswith("sd_")}, strict=False)
    model = model.cuda().eval()
    for p in model.parameters():
        p.requires_grad_(False)
    ids = torch.from_numpy(z["ids"]).cuda()
    n0 = ops.launch_count()
    emb = model.transformer.wte(ids).detach().requires_grad_()
    logits = model(inputs_embeds=emb, use_cache=False).logits
    mx, mi = logits[:, -1, :].max(-1)
    mx.sum().backward()
    rel = (emb * emb.grad).float().sum(-1).detach().cpu()
    assert ops.launch_count() - n0 >= 2 * 8
    assert np.array_equal(mi.cpu().numpy(), z["idx"])
    err = rel_l2(rel, z["rel_fp32"])
    print(f"GPT-2 tiny: rel-L2 vs reference fp32 = {err:.3e}")
    assert err < 2e-2
