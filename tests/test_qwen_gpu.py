"""GPU: Qwen2 (q/k/v projection biases) and Qwen3 (per-head q/k RMSNorm) through the drop-in default maps, against the
golden relevance of the real reference (the other decoder families re-use the same kernels)."""
import warnings

import numpy as np
import pytest
import torch

from helpers import bf16_from_bits, load_npz, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("family", ["qwen2", "qwen3"])
def test_patched_qwen_matches_reference(family):
    import importlib
    import transformers
    from lxt_b200.efficient import monkey_patch
    from lxt_b200 import ops
    z = load_npz(f"{family}_tiny.npz")
    modeling = importlib.import_module(f"transformers.models.{family}.modeling_{family}")
    Cfg = getattr(transformers, f"{family.capitalize()}Config")
    Model = getattr(transformers, f"{family.capitalize()}ForCausalLM")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        monkey_patch(modeling, verbose=True)
    kw = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=1,
              vocab_size=384, max_position_embeddings=512, rms_norm_eps=1e-6, tie_word_embeddings=False)
    if family == "qwen3":
        kw["head_dim"] = 64
    cfg = Cfg(**kw)
    cfg._attn_implementation = "sdpa"
    from helpers import build_hf
    model = build_hf(Model, cfg)
    model.load_state_dict({k[3:]: bf16_from_bits(v) for k, v in z.items() if k.startswith("sd_")}, strict=True)
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
    assert ops.launch_count() - n0 > 2 * 20
    assert np.array_equal(mi.cpu().numpy(), z["idx"])
    err = rel_l2(rel, z["rel_fp32"])
    print(f"{family} tiny: rel-L2 vs reference fp32 = {err:.3e}")
    assert err < 2e-2
    # the fused engine built from the same HF model (Qwen2: q/k/v biases in the packed-QKV GEMM epilogue; Qwen3: per-head
    # q/k RMSNorm kernel)
    from lxt_b200.engine import LlamaAttnLRPEngine
    eng = LlamaAttnLRPEngine.from_hf(model, micro_batch=2)
    r_eng, aux = eng.attribute_device(ids, return_aux=True)
    assert np.array_equal(aux["idx"].cpu().numpy(), z["idx"])
    e2 = rel_l2(r_eng.cpu(), z["rel_fp32"])
    print(f"{family} tiny: fused engine rel-L2 vs reference fp32 = {e2:.3e}")
    assert e2 < 6e-3
