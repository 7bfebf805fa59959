"""Generate golden vectors by running the REAL reference (rachtibat/LRP-eXplains-Transformers, `lxt` 2.1).

Run in the build container (the reference is not available on the GPU box):

    PYTHONPATH=/root/reference python tests/golden/make_golden.py

Writes small .npz fixtures next to this file.  Nothing here is imported by the product or by the GPU tests; the
tests only read the .npz files.  Fixtures:

  rules.npz        lxt.explicit.functional / lxt.explicit.rules / lxt.efficient.rules on seeded inputs
                   (shapes of the reference's own tests/test_rules.py, tests/test_functional.py + the BASELINE
                   "2-layer 128-d MLP" epsilon-rule case)
  llama_tiny_*.npz lxt.efficient.monkey_patch(modeling_llama) on seeded tiny HuggingFace Llama models:
                   token ids, weights (bf16 bit patterns), relevance from an fp32 run and from a bf16 run.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import lxt.explicit.functional as lf  # noqa: E402  (the reference)
import lxt.explicit.rules as lrules  # noqa: E402
import lxt.efficient.rules as erules  # noqa: E402
from lxt.efficient import monkey_patch  # noqa: E402
from transformers import LlamaConfig, LlamaForCausalLM  # noqa: E402
from transformers.models.llama import modeling_llama  # noqa: E402

from oracle.attnlrp_oracle import random_llama_weights  # noqa: E402  (weight generator only)


def bf16_bits(t: torch.Tensor) -> np.ndarray:
    return t.to(torch.bfloat16).view(torch.int16).numpy().astype(np.uint16)


def gen_rules():
    out = {}
    g = torch.Generator().manual_seed(1234)
    rn = lambda *s: torch.randn(*s, generator=g)

    # tests/test_rules.py:9-24 : EpsilonRule(partial(F.linear, W, b)) vs lf.linear_epsilon
    x, W, b, R = rn(1, 5), rn(5, 5), rn(5), rn(1, 5)
    xx = x.clone().requires_grad_()
    y = lf.linear_epsilon(xx, W, b, 1e-6)
    y.backward(R)
    out.update(le_x=x, le_W=W, le_b=b, le_R=R, le_y=y.detach(), le_Rin=xx.grad.clone())
    xx = x.clone().requires_grad_()
    import functools
    rule = lrules.EpsilonRule(functools.partial(torch.nn.functional.linear, weight=W, bias=b), epsilon=1e-6)
    y2 = rule(xx)
    y2.backward(R)
    out.update(le_rule_Rin=xx.grad.clone())

    # tests/test_functional.py:57-76 shapes
    x, W, R = rn(16, 10), rn(5, 10), rn(16, 5)
    xx = x.clone().requires_grad_()
    lf.linear_epsilon(xx, W, None, 1e-9).backward(R)
    out.update(lin_x=x, lin_W=W, lin_R=R, lin_Rin=xx.grad.clone())

    # BASELINE config[0]: 2-layer 128-d MLP under the epsilon rule, CPU fp32
    x, W1, b1, W2, b2, R = rn(8, 128), rn(128, 128) * 0.1, rn(128) * 0.1, rn(128, 128) * 0.1, rn(128) * 0.1, rn(8, 128)
    xx = x.clone().requires_grad_()
    h1 = lf.linear_epsilon(xx, W1, b1, 1e-6)
    h2 = lf.linear_epsilon(h1, W2, b2, 1e-6)
    h2.backward(R)
    out.update(mlp_x=x, mlp_W1=W1, mlp_b1=b1, mlp_W2=W2, mlp_b2=b2, mlp_R=R, mlp_h1=h1.detach(), mlp_Rin=xx.grad.clone())

    # matmul (test_functional.py:29-54)
    a, bb, R = rn(2, 10, 32), rn(2, 32, 5), rn(2, 10, 5)
    aa, b2_ = a.clone().requires_grad_(), bb.clone().requires_grad_()
    lf.matmul(aa, b2_, False, 1e-9).backward(R)
    out.update(mm_a=a, mm_b=bb, mm_R=R, mm_Ra=aa.grad.clone(), mm_Rb=b2_.grad.clone())

    # softmax (test_functional.py:8-26, called with inplace as keyword)
    x, R = rn(16, 10, 32), rn(16, 10, 32)
    xx = x.clone().requires_grad_()
    lf.softmax(xx, -1, None, 1.0, inplace=False).backward(R)
    out.update(sm_x=x, sm_R=R, sm_Rin=xx.grad.clone())

    # add2 / mul2 / rms_norm_identity
    a, bb, R = rn(16, 10, 32), rn(16, 10, 32), rn(16, 10, 32)
    aa, b2_ = a.clone().requires_grad_(), bb.clone().requires_grad_()
    lf.add2(aa, b2_, False, 1e-8).backward(R)
    out.update(add_a=a, add_b=bb, add_R=R, add_Ra=aa.grad.clone(), add_Rb=b2_.grad.clone())
    aa, b2_ = a.clone().requires_grad_(), bb.clone().requires_grad_()
    lf.mul2(aa, b2_).backward(R)
    out.update(mul_Ra=aa.grad.clone(), mul_Rb=b2_.grad.clone())
    x, w, R = rn(1, 4, 32), rn(32), rn(1, 4, 32)
    xx = x.clone().requires_grad_()
    yn = lf.rms_norm_identity(xx, w, 1e-6)
    yn.backward(R)
    out.update(rms_x=x, rms_w=w, rms_R=R, rms_y=yn.detach(), rms_Rin=xx.grad.clone())

    # UniformEpsilonRule on a matmul module (rules.py:253-282)
    class MM(torch.nn.Module):
        def forward(self, p, q):
            return torch.matmul(p, q)
    a, bb, R = rn(2, 6, 16), rn(2, 16, 8), rn(2, 6, 8)
    aa, b2_ = a.clone().requires_grad_(), bb.clone().requires_grad_()
    lrules.UniformEpsilonRule(MM(), epsilon=1e-6)(aa, b2_).backward(R)
    out.update(ue_a=a, ue_b=bb, ue_R=R, ue_Ra=aa.grad.clone(), ue_Rb=b2_.grad.clone())

    # efficient rules: identity_rule_implicit(silu / gelu), divide_gradient
    x, gout = rn(4, 64), rn(4, 64)
    for name, fn in (("silu", torch.nn.functional.silu), ("gelu", torch.nn.functional.gelu),
                     ("gelu_tanh", lambda t: torch.nn.functional.gelu(t, approximate="tanh"))):
        xx = x.clone().requires_grad_()
        erules.identity_rule_implicit(fn, xx).backward(gout)
        out[f"id_{name}_g"] = xx.grad.clone()
    xx = x.clone().requires_grad_()
    erules.divide_gradient(xx * 1.0, 4).backward(gout)
    out.update(id_x=x, id_gout=gout, div4_g=xx.grad.clone())
    np.savez_compressed(os.path.join(HERE, "rules.npz"), **{k: v.numpy() for k, v in out.items()})
    print("rules.npz:", len(out), "arrays")


_patched = False


def gen_llama(name, cfg, B, S, seed):
    global _patched
    if not _patched:
        monkey_patch(modeling_llama, verbose=True)
        _patched = True
    w = random_llama_weights(cfg, seed=seed, std=0.02, dtype=torch.bfloat16)
    # make the norm weights non-trivial so that the w*rstd path is exercised
    g = torch.Generator().manual_seed(seed + 100)
    for lw in w["layers"]:
        lw["ln1"] = (1 + 0.1 * torch.randn(cfg["d"], generator=g)).to(torch.bfloat16)
        lw["ln2"] = (1 + 0.1 * torch.randn(cfg["d"], generator=g)).to(torch.bfloat16)
    w["norm"] = (1 + 0.1 * torch.randn(cfg["d"], generator=g)).to(torch.bfloat16)
    ids = torch.randint(0, cfg["V"], (B, S), generator=torch.Generator().manual_seed(seed + 1))
    hf_cfg = LlamaConfig(hidden_size=cfg["d"], intermediate_size=cfg["I"], num_hidden_layers=cfg["L"],
                         num_attention_heads=cfg["H"], num_key_value_heads=cfg["Hkv"], head_dim=cfg["D"],
                         vocab_size=cfg["V"], rms_norm_eps=cfg["eps"],
                         rope_parameters={"rope_type": "default", "rope_theta": cfg["theta"]},
                         max_position_embeddings=max(512, S), attention_bias=False, tie_word_embeddings=False)
    res = {}
    for dt, tag in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
        for impl in ("sdpa", "eager"):
            hf_cfg._attn_implementation = impl
            model = LlamaForCausalLM(hf_cfg).to(dt).eval()
            sd = {"model.embed_tokens.weight": w["emb"], "model.norm.weight": w["norm"], "lm_head.weight": w["lm_head"]}
            for i, lw in enumerate(w["layers"]):
                p = f"model.layers.{i}."
                sd.update({p + "self_attn.q_proj.weight": lw["wq"], p + "self_attn.k_proj.weight": lw["wk"],
                           p + "self_attn.v_proj.weight": lw["wv"], p + "self_attn.o_proj.weight": lw["wo"],
                           p + "mlp.gate_proj.weight": lw["wg"], p + "mlp.up_proj.weight": lw["wu"],
                           p + "mlp.down_proj.weight": lw["wd"], p + "input_layernorm.weight": lw["ln1"],
                           p + "post_attention_layernorm.weight": lw["ln2"]})
            missing = model.load_state_dict({k: v.to(dt) for k, v in sd.items()}, strict=True)
            for p_ in model.parameters():
                p_.requires_grad_(False)
            # examples/quantized_llama.py:35-47
            emb = model.get_input_embeddings()(ids)
            emb = emb.detach().requires_grad_()
            # latent relevance trace exactly as docs/source/latent-feature-attribution-efficient.rst:49-90 does it:
            # forward hooks keep each decoder layer's output and retain its grad; relevance = output * output.grad
            kept, hooks = [], []

            def keep(mod, inp, out):
                t = out[0] if isinstance(out, tuple) else out
                t.retain_grad()
                kept.append(t)

            for layer in model.model.layers:
                hooks.append(layer.register_forward_hook(keep))
            logits = model(inputs_embeds=emb, use_cache=False).logits
            max_logits, max_idx = torch.max(logits[:, -1, :], dim=-1)
            max_logits.sum().backward()
            for hk in hooks:
                hk.remove()
            rel = (emb * emb.grad).float().sum(-1)
            res[f"trace_{tag}_{impl}"] = torch.stack([(t * t.grad).float().sum(-1) for t in kept]).detach().numpy()
            res[f"rel_{tag}_{impl}"] = rel.detach().numpy()
            res[f"idx_{tag}_{impl}"] = max_idx.numpy()
            if tag == "fp32" and impl == "sdpa":
                res[f"gemb_{tag}_{impl}"] = emb.grad.float().numpy()
    print(name, "sdpa-vs-eager fp32 rel diff",
          float(np.linalg.norm(res["rel_fp32_sdpa"] - res["rel_fp32_eager"]) / np.linalg.norm(res["rel_fp32_eager"])),
          " bf16-vs-fp32", float(np.linalg.norm(res["rel_bf16_sdpa"] - res["rel_fp32_sdpa"]) / np.linalg.norm(res["rel_fp32_sdpa"])))
    save = dict(ids=ids.numpy(), cfg_keys=np.array(list(cfg.keys())), cfg_vals=np.array([float(v) for v in cfg.values()]))
    save.update({"w_emb": bf16_bits(w["emb"]), "w_norm": bf16_bits(w["norm"]), "w_lm_head": bf16_bits(w["lm_head"])})
    for i, lw in enumerate(w["layers"]):
        for k, v in lw.items():
            save[f"w_l{i}_{k}"] = bf16_bits(v)
    save.update(res)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **save)


def gen_llama_padded():
    """A batch of two prompts of DIFFERENT lengths through the real reference (lxt.efficient.monkey_patch(modeling_llama), HF
    `attention_mask`): left padding with the logit read at the last position, right padding with the logit read at each
    prompt's own last token.  Weights / ids of llama_tiny_d64.npz; stores the relevance of the fp32 run (sdpa)."""
    global _patched
    if not _patched:
        monkey_patch(modeling_llama, verbose=True)
        _patched = True
    sys.path.insert(0, os.path.dirname(HERE))
    from helpers import load_llama_golden
    cfg, w, ids, _ = load_llama_golden("llama_tiny_d64.npz")
    B, S = ids.shape
    lens = [S, 97]
    hf_cfg = LlamaConfig(hidden_size=cfg["d"], intermediate_size=cfg["I"], num_hidden_layers=cfg["L"],
                         num_attention_heads=cfg["H"], num_key_value_heads=cfg["Hkv"], head_dim=cfg["D"],
                         vocab_size=cfg["V"], rms_norm_eps=cfg["eps"],
                         rope_parameters={"rope_type": "default", "rope_theta": cfg["theta"]},
                         max_position_embeddings=512, attention_bias=False, tie_word_embeddings=False)
    hf_cfg._attn_implementation = "sdpa"
    model = LlamaForCausalLM(hf_cfg).to(torch.float32).eval()
    sd = {"model.embed_tokens.weight": w["emb"], "model.norm.weight": w["norm"], "lm_head.weight": w["lm_head"]}
    for i, lw in enumerate(w["layers"]):
        p = f"model.layers.{i}."
        sd.update({p + "self_attn.q_proj.weight": lw["wq"], p + "self_attn.k_proj.weight": lw["wk"],
                   p + "self_attn.v_proj.weight": lw["wv"], p + "self_attn.o_proj.weight": lw["wo"],
                   p + "mlp.gate_proj.weight": lw["wg"], p + "mlp.up_proj.weight": lw["wu"],
                   p + "mlp.down_proj.weight": lw["wd"], p + "input_layernorm.weight": lw["ln1"],
                   p + "post_attention_layernorm.weight": lw["ln2"]})
    model.load_state_dict({k: v.float() for k, v in sd.items()}, strict=True)
    for p_ in model.parameters():
        p_.requires_grad_(False)
    save = dict(lens=np.array(lens))
    for side in ("left", "right"):
        mask = torch.zeros(B, S, dtype=torch.long)
        for b, n in enumerate(lens):
            if side == "left":
                mask[b, S - n:] = 1
            else:
                mask[b, :n] = 1
        emb = model.get_input_embeddings()(ids).detach().requires_grad_()
        logits = model(inputs_embeds=emb, attention_mask=mask, use_cache=False).logits
        last = torch.tensor([S - 1] * B) if side == "left" else torch.tensor([n - 1 for n in lens])
        mx, mi = torch.max(logits[torch.arange(B), last, :], dim=-1)
        mx.sum().backward()
        rel = (emb * emb.grad).float().sum(-1) * mask      # relevance of padding tokens is not defined: zeroed
        save[f"mask_{side}"] = mask.numpy()
        save[f"rel_{side}"] = rel.detach().numpy()
        save[f"idx_{side}"] = mi.numpy()
        print("padded", side, "relevance norm", float(rel.norm()), "idx", mi.tolist())
    np.savez_compressed(os.path.join(HERE, "llama_tiny_d64_padded.npz"), **save)


def gen_vit():
    """torchvision ViT, lxt.efficient.monkey_patch(vision_transformer) (cp_LRP map), examples/vit_torch.py:84-91."""
    from torchvision.models import vision_transformer
    monkey_patch(vision_transformer, verbose=True)
    torch.manual_seed(5)
    model = vision_transformer.VisionTransformer(image_size=64, patch_size=16, num_layers=2, num_heads=2, hidden_dim=128,
                                                 mlp_dim=256, num_classes=16).eval()
    torch.nn.init.normal_(model.heads.head.weight, std=0.02)  # torchvision zero-fills the head (all-zero relevance)
    for p_ in model.parameters():
        p_.requires_grad_(False)
    x = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(6)).requires_grad_()
    y = model(x)
    cls = y.argmax(-1)
    y[torch.arange(2), cls].sum().backward()
    heat = (x * x.grad).sum(1)
    save = {"x": x.detach().numpy(), "heat": heat.detach().numpy(), "cls": cls.numpy(), "logits": y.detach().numpy()}
    for k, v in model.state_dict().items():
        save["sd_" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "vit_tiny.npz"), **save)
    print("vit_tiny.npz heat norm", float(heat.norm()))


def build_vit_l16(seed=5):
    """torchvision vit_l_16 (BASELINE configs[3]) with seeded random weights and the head re-initialised (torchvision zero-fills it).
    The 304 M weights are NOT stored: the GPU test rebuilds them from the same seed (same torch build, CPU generator)."""
    from torchvision.models import vision_transformer
    torch.manual_seed(seed)
    model = vision_transformer.vit_l_16(weights=None).eval()
    torch.nn.init.normal_(model.heads.head.weight, std=0.02)
    for p_ in model.parameters():
        p_.requires_grad_(False)
    return model


def vit_weight_fingerprint(model):
    sd = model.state_dict()
    keys = ["conv_proj.weight", "encoder.layers.encoder_layer_11.mlp.0.weight", "heads.head.weight"]
    return np.array([float(sd[k].double().abs().sum()) for k in keys])


def gen_vit_l16():
    """ViT-L/16 at full size, lxt.efficient.monkey_patch(vision_transformer) (cp_LRP map), examples/vit_torch.py:84-91:
    pixel relevance of the arg-max class from the reference's fp32 run and from its bf16 run."""
    from torchvision.models import vision_transformer
    monkey_patch(vision_transformer, verbose=True)
    model = build_vit_l16()
    x0 = torch.randn(1, 3, 224, 224, generator=torch.Generator().manual_seed(6))
    x = x0.clone().requires_grad_()
    y = model(x)
    cls = y.argmax(-1)
    y[torch.arange(1), cls].sum().backward()
    heat = (x * x.grad).sum(1)
    fp = vit_weight_fingerprint(model)
    m16 = model.to(torch.bfloat16)
    x16 = x0.to(torch.bfloat16).requires_grad_()
    y16 = m16(x16)
    y16[torch.arange(1), cls].sum().backward()
    heat16 = (x16 * x16.grad).float().sum(1)
    np.savez_compressed(os.path.join(HERE, "vit_l16.npz"), x=x0.numpy(), heat=heat.detach().numpy(), heat_bf16=heat16.detach().numpy(),
                        cls=cls.numpy(), logits=y.detach().numpy(), fingerprint=fp)
    e = float((heat16.detach().double() - heat.detach().double()).norm() / heat.detach().double().norm())
    print("vit_l16.npz heat norm", float(heat.norm()), "reference bf16 vs fp32 rel-L2", e)


def gen_llama_cp():
    """CP-LRP variant (lxt/efficient/models/llama.py:16-21) — run in a fresh process: patches are process-global."""
    from lxt.efficient.models.llama import cp_LRP
    monkey_patch(modeling_llama, cp_LRP, verbose=True)
    cfg = dict(d=256, I=512, H=4, Hkv=2, D=64, L=2, V=256, eps=1e-5, theta=10000.0)
    w = random_llama_weights(cfg, seed=0, std=0.02, dtype=torch.bfloat16)
    ids = torch.randint(0, cfg["V"], (2, 160), generator=torch.Generator().manual_seed(1))
    hf_cfg = LlamaConfig(hidden_size=cfg["d"], intermediate_size=cfg["I"], num_hidden_layers=cfg["L"],
                         num_attention_heads=cfg["H"], num_key_value_heads=cfg["Hkv"], head_dim=cfg["D"],
                         vocab_size=cfg["V"], rms_norm_eps=cfg["eps"],
                         rope_parameters={"rope_type": "default", "rope_theta": cfg["theta"]},
                         max_position_embeddings=512, attention_bias=False, tie_word_embeddings=False)
    hf_cfg._attn_implementation = "sdpa"
    model = LlamaForCausalLM(hf_cfg).float().eval()
    sd = {"model.embed_tokens.weight": w["emb"], "model.norm.weight": w["norm"], "lm_head.weight": w["lm_head"]}
    for i, lw in enumerate(w["layers"]):
        p = f"model.layers.{i}."
        sd.update({p + "self_attn.q_proj.weight": lw["wq"], p + "self_attn.k_proj.weight": lw["wk"],
                   p + "self_attn.v_proj.weight": lw["wv"], p + "self_attn.o_proj.weight": lw["wo"],
                   p + "mlp.gate_proj.weight": lw["wg"], p + "mlp.up_proj.weight": lw["wu"],
                   p + "mlp.down_proj.weight": lw["wd"], p + "input_layernorm.weight": lw["ln1"],
                   p + "post_attention_layernorm.weight": lw["ln2"]})
    model.load_state_dict({k: v.float() for k, v in sd.items()}, strict=True)
    for p_ in model.parameters():
        p_.requires_grad_(False)
    emb = model.get_input_embeddings()(ids).detach().requires_grad_()
    logits = model(inputs_embeds=emb, use_cache=False).logits
    max_logits, max_idx = torch.max(logits[:, -1, :], dim=-1)
    max_logits.sum().backward()
    rel = (emb * emb.grad).float().sum(-1)
    np.savez_compressed(os.path.join(HERE, "llama_tiny_cp.npz"), ids=ids.numpy(), rel_fp32=rel.detach().numpy(),
                        idx=max_idx.numpy(), seed=np.array([0]))
    print("llama_tiny_cp.npz rel norm", float(rel.norm()))


def gen_gemma(head_dim=64, name="gemma3_tiny"):
    """Gemma-3 text model (sliding-window + global layers, q/k-norm, (1+w) RMSNorm, GELU-tanh gated MLP) under
    lxt.efficient.monkey_patch(modeling_gemma3) — lxt/efficient/models/gemma3.py:11-19.  head_dim 64 (the B200 attention
    tiles cover 64/128; Gemma's production head_dim 256 is a round-2 item)."""
    from transformers import Gemma3TextConfig, Gemma3ForCausalLM
    from transformers.models.gemma3 import modeling_gemma3
    monkey_patch(modeling_gemma3, verbose=True)
    cfg = Gemma3TextConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=6, num_attention_heads=2,
                           num_key_value_heads=1, head_dim=head_dim, vocab_size=384, sliding_window=48, max_position_embeddings=512,
                           query_pre_attn_scalar=head_dim, rms_norm_eps=1e-6, tie_word_embeddings=True)
    cfg._attn_implementation = "sdpa"
    torch.manual_seed(11)
    model = Gemma3ForCausalLM(cfg).float().eval()
    g = torch.Generator().manual_seed(12)
    sd = model.state_dict()
    for k_, v_ in sd.items():
        if "norm" in k_:
            v_.copy_(0.1 * torch.randn(v_.shape, generator=g))
        v_.copy_(v_.to(torch.bfloat16).float())   # bf16-representable weights
    for p_ in model.parameters():
        p_.requires_grad_(False)
    ids = torch.randint(0, cfg.vocab_size, (2, 160), generator=torch.Generator().manual_seed(13))
    emb = model.get_input_embeddings()(ids).detach().requires_grad_()
    logits = model(inputs_embeds=emb, use_cache=False).logits
    max_logits, max_idx = torch.max(logits[:, -1, :], dim=-1)
    max_logits.sum().backward()
    rel = (emb * emb.grad).float().sum(-1)
    save = {"ids": ids.numpy(), "rel_fp32": rel.detach().numpy(), "idx": max_idx.numpy(), "layer_types": np.array(cfg.layer_types)}
    for k_, v_ in model.state_dict().items():
        save["sd_" + k_] = bf16_bits(v_)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **save)
    print(f"{name}.npz", cfg.layer_types, "rel norm", float(rel.norm()))


def gen_gpt2():
    """GPT-2 (LayerNorm + plain GELU MLP + Conv1D projections) under lxt.efficient.monkey_patch(modeling_gpt2) —
    lxt/efficient/models/gpt2.py:11-32 (`mlp_forward`, `layer_norm_forward`, `patch_attention`)."""
    from transformers import GPT2Config, GPT2LMHeadModel
    from transformers.models.gpt2 import modeling_gpt2
    monkey_patch(modeling_gpt2, verbose=True)
    cfg = GPT2Config(n_embd=128, n_head=2, n_layer=2, vocab_size=384, n_positions=256, resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
    cfg._attn_implementation = "sdpa"
    torch.manual_seed(21)
    model = GPT2LMHeadModel(cfg).float().eval()
    sd = model.state_dict()
    g = torch.Generator().manual_seed(22)
    for k_, v_ in sd.items():
        if v_.dtype.is_floating_point:
            if "ln_" in k_ and k_.endswith("weight"):
                v_.copy_(1 + 0.1 * torch.randn(v_.shape, generator=g))
            v_.copy_(v_.to(torch.bfloat16).float())
    for p_ in model.parameters():
        p_.requires_grad_(False)
    ids = torch.randint(0, cfg.vocab_size, (2, 144), generator=torch.Generator().manual_seed(23))
    emb = model.transformer.wte(ids).detach().requires_grad_()
    logits = model(inputs_embeds=emb, use_cache=False).logits
    max_logits, max_idx = torch.max(logits[:, -1, :], dim=-1)
    max_logits.sum().backward()
    rel = (emb * emb.grad).float().sum(-1)
    save = {"ids": ids.numpy(), "rel_fp32": rel.detach().numpy(), "idx": max_idx.numpy()}
    for k_, v_ in model.state_dict().items():
        if v_.dtype.is_floating_point:
            save["sd_" + k_] = bf16_bits(v_)
    np.savez_compressed(os.path.join(HERE, "gpt2_tiny.npz"), **save)
    print("gpt2_tiny.npz rel norm", float(rel.norm()))


def gen_qwen():
    """Qwen2 (q/k/v biases) and Qwen3 (per-head q/k RMSNorm) under the reference's default maps
    (lxt/efficient/models/qwen2.py, qwen3.py) — same rule set as Llama."""
    import warnings
    from transformers import Qwen2Config, Qwen2ForCausalLM, Qwen3Config, Qwen3ForCausalLM
    from transformers.models.qwen2 import modeling_qwen2
    from transformers.models.qwen3 import modeling_qwen3
    for name, Cfg, Model, modeling in (("qwen2_tiny", Qwen2Config, Qwen2ForCausalLM, modeling_qwen2),
                                       ("qwen3_tiny", Qwen3Config, Qwen3ForCausalLM, modeling_qwen3)):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            monkey_patch(modeling, verbose=True)
        kw = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=1,
                  vocab_size=384, max_position_embeddings=512, rms_norm_eps=1e-6, tie_word_embeddings=False)
        if Cfg is Qwen3Config:
            kw["head_dim"] = 64
        cfg = Cfg(**kw)
        cfg._attn_implementation = "sdpa"
        torch.manual_seed(31)
        model = Model(cfg).float().eval()
        g = torch.Generator().manual_seed(32)
        for k_, v_ in model.state_dict().items():
            if "norm" in k_:
                v_.copy_(1 + 0.1 * torch.randn(v_.shape, generator=g))
            if k_.endswith("bias"):
                v_.copy_(0.05 * torch.randn(v_.shape, generator=g))
            v_.copy_(v_.to(torch.bfloat16).float())
        for p_ in model.parameters():
            p_.requires_grad_(False)
        ids = torch.randint(0, cfg.vocab_size, (2, 144), generator=torch.Generator().manual_seed(33))
        emb = model.get_input_embeddings()(ids).detach().requires_grad_()
        logits = model(inputs_embeds=emb, use_cache=False).logits
        max_logits, max_idx = torch.max(logits[:, -1, :], dim=-1)
        max_logits.sum().backward()
        rel = (emb * emb.grad).float().sum(-1)
        save = {"ids": ids.numpy(), "rel_fp32": rel.detach().numpy(), "idx": max_idx.numpy()}
        for k_, v_ in model.state_dict().items():
            save["sd_" + k_] = bf16_bits(v_)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **save)
        print(f"{name}.npz rel norm", float(rel.norm()), "biases:", sum(k_.endswith("bias") for k_ in model.state_dict()))


if __name__ == "__main__":
    torch.manual_seed(0)
    if len(sys.argv) > 1 and sys.argv[1] == "--qwen":
        gen_qwen()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "--gpt2":
        gen_gpt2()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "--gemma":
        gen_gemma()
        gen_gemma(256, "gemma3_tiny_d256")   # Gemma's production head_dim
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "--cp":
        gen_llama_cp()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "--vit":
        gen_vit()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "--padded":
        gen_llama_padded()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "--vit-l16":
        gen_vit_l16()
        sys.exit(0)
    gen_rules()
    gen_llama("llama_tiny_d64", dict(d=256, I=512, H=4, Hkv=2, D=64, L=2, V=256, eps=1e-5, theta=10000.0), B=2, S=160, seed=0)
    gen_llama("llama_tiny_d128", dict(d=256, I=384, H=2, Hkv=1, D=128, L=2, V=320, eps=1e-5, theta=500000.0), B=1, S=130, seed=7)
