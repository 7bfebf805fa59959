"""GPU: the drop-in boundary beyond the plain causal case — what the reference gets for free by delegating to HF
(lxt/efficient/patches.py:193-203) and what its examples rely on (examples/quantized_llama.py:21-27):
  * batches of prompts of different lengths (HF `attention_mask`, left or right padding) vs a golden of the REAL reference;
  * `model.train()` + `gradient_checkpointing_enable()` (torch.utils.checkpoint re-runs the patched forwards);
  * fp32 HF models = validation precision through the same kernels: <= 1e-3 against the reference's fp32 goldens
    (llama, gpt2, qwen2, qwen3, gemma3), which shows the bf16 drop-in figures (up to 2e-2) are the bf16 module graph's rounding.
"""
import importlib
import warnings

import numpy as np
import pytest
import torch

from helpers import bf16_from_bits, load_llama_golden, load_npz, rel_l2

pytestmark = pytest.mark.gpu

TOL = 1e-3


def _patch(modeling):
    from lxt_b200.efficient import monkey_patch
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        monkey_patch(modeling)


def _llama(name, dtype, impl="sdpa"):
    from test_monkey_patch_gpu import _hf_model
    from transformers.models.llama import modeling_llama
    _patch(modeling_llama)
    cfg, w, ids, z = load_llama_golden(name)
    return _hf_model(cfg, w, impl).to(dtype), ids.cuda(), z


def _attribute(model, ids, attention_mask=None, last=None):
    emb = model.get_input_embeddings()(ids).detach().requires_grad_()
    logits = model(inputs_embeds=emb, attention_mask=attention_mask, use_cache=False).logits
    B = ids.shape[0]
    pos = torch.full((B,), ids.shape[1] - 1, device=ids.device) if last is None else last
    mx, mi = torch.max(logits[torch.arange(B, device=ids.device), pos, :], dim=-1)
    mx.sum().backward()
    return (emb * emb.grad).float().sum(-1).detach(), mi


@pytest.mark.parametrize("side", ["left", "right"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_padded_batch_matches_reference_golden(side, dtype):
    model, ids, _ = _llama("llama_tiny_d64.npz", dtype)
    z = load_npz("llama_tiny_d64_padded.npz")
    mask = torch.from_numpy(z[f"mask_{side}"]).cuda()
    last = None if side == "left" else torch.from_numpy(z["lens"]).cuda() - 1
    rel, mi = _attribute(model, ids, attention_mask=mask, last=last)
    rel = (rel * mask).cpu()
    err = rel_l2(rel, z[f"rel_{side}"])
    print(f"padded batch ({side}, {dtype}): rel-L2 vs reference fp32 = {err:.2e}")
    if dtype == torch.float32:
        assert np.array_equal(mi.cpu().numpy(), z[f"idx_{side}"])
        assert err <= TOL
    else:
        assert err <= 1.5e-2
    # and the defining property: a padded prompt's relevance equals the relevance of the same prompt run alone
    n = int(z["lens"][1])
    sl = slice(ids.shape[1] - n, None) if side == "left" else slice(0, n)
    alone, _ = _attribute(model, ids[1:2, sl])
    e2 = rel_l2(rel[1, sl], alone[0].cpu())
    print(f"   padded vs alone: {e2:.2e}")
    assert e2 <= (2e-4 if dtype == torch.float32 else 1.5e-2)


def test_gradient_checkpointing_train_mode_equals_plain_run():
    """examples/quantized_llama.py:21-27: model.train() + gradient_checkpointing_enable(); the patched Functions must survive
    torch.utils.checkpoint's second forward, and Dropout must be the identity (dropout_forward)."""
    model, ids, z = _llama("llama_tiny_d64.npz", torch.bfloat16)
    model.config.attention_dropout = 0.1           # would change the result if the patched attention did not force dropout 0
    rel0, _ = _attribute(model, ids)
    model.train()
    model.gradient_checkpointing_enable()
    rel1, _ = _attribute(model, ids)
    model.gradient_checkpointing_disable()
    model.eval()
    e = rel_l2(rel1.cpu(), rel0.cpu())
    print(f"checkpointed train-mode run vs plain eval run: {e:.2e}; vs reference fp32 {rel_l2(rel1.cpu(), z['rel_fp32_sdpa']):.2e}")
    assert e < 1e-4      # same kernels twice; fp32 dQ atomics may reorder the last bits


@pytest.mark.parametrize("name,impl", [("llama_tiny_d64.npz", "sdpa"), ("llama_tiny_d128.npz", "eager")])
def test_fp32_hf_llama_validation_precision(name, impl):
    from lxt_b200 import ops
    model, ids, z = _llama(name, torch.float32, impl)
    n0 = ops.launch_count()
    rel, mi = _attribute(model, ids)
    assert ops.launch_count() - n0 > 40, "the B200 kernels did not run"
    assert np.array_equal(mi.cpu().numpy(), z["idx_fp32_sdpa"])
    err = rel_l2(rel.cpu(), z["rel_fp32_sdpa"])
    print(f"{name}: fp32 HF model through the drop-in path, rel-L2 vs reference fp32 = {err:.2e}")
    assert err <= TOL


@pytest.mark.parametrize("family", ["qwen2", "qwen3", "gpt2", "gemma3"])
def test_fp32_hf_families_validation_precision(family):
    import transformers
    z = load_npz({"gemma3": "gemma3_tiny.npz"}.get(family, f"{family}_tiny.npz"))
    if family == "gemma3":
        from transformers import Gemma3ForCausalLM, Gemma3TextConfig
        from transformers.models.gemma3 import modeling_gemma3
        _patch(modeling_gemma3)
        cfg = Gemma3TextConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=6, num_attention_heads=2,
                               num_key_value_heads=1, head_dim=64, vocab_size=384, sliding_window=48, max_position_embeddings=512,
                               query_pre_attn_scalar=64, rms_norm_eps=1e-6, tie_word_embeddings=True)
        assert list(z["layer_types"]) == cfg.layer_types
        Model = Gemma3ForCausalLM
    elif family == "gpt2":
        from transformers import GPT2Config, GPT2LMHeadModel
        from transformers.models.gpt2 import modeling_gpt2
        _patch(modeling_gpt2)
        cfg = GPT2Config(n_embd=128, n_head=2, n_layer=2, vocab_size=384, n_positions=256, resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
        Model = GPT2LMHeadModel
    else:
        modeling = importlib.import_module(f"transformers.models.{family}.modeling_{family}")
        _patch(modeling)
        kw = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=1,
                  vocab_size=384, max_position_embeddings=512, rms_norm_eps=1e-6, tie_word_embeddings=False)
        if family == "qwen3":
            kw["head_dim"] = 64
        cfg = getattr(transformers, f"{family.capitalize()}Config")(**kw)
        Model = getattr(transformers, f"{family.capitalize()}ForCausalLM")
    cfg._attn_implementation = "sdpa"
    model = Model(cfg)
    model.load_state_dict({k[3:]: bf16_from_bits(v).float() for k, v in z.items() if k.startswith("sd_")}, strict=False)
    model = model.float().cuda().eval()
    for p_ in model.parameters():
        p_.requires_grad_(False)
    rel, mi = _attribute(model, torch.from_numpy(z["ids"]).cuda())
    assert np.array_equal(mi.cpu().numpy(), z["idx"])
    err = rel_l2(rel.cpu(), z["rel_fp32"])
    print(f"{family}: fp32 HF model through the drop-in path, rel-L2 vs reference fp32 = {err:.2e}")
    assert err <= TOL
