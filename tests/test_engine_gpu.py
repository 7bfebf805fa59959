"""GPU parity of the engine (all CUDA kernels through the C ABI) against the golden vectors of the real
reference and against the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest
import torch

from helpers import load_llama_golden, rel_l2

pytestmark = pytest.mark.gpu


def _engine(cfg, w, **kw):
    from lxt_b200.engine import LlamaAttnLRPEngine, LlamaDims
    dims = LlamaDims(d=cfg["d"], I=cfg["I"], H=cfg["H"], Hkv=cfg["Hkv"], D=cfg["D"], L=cfg["L"], V=cfg["V"], eps=cfg["eps"],
                     theta=cfg["theta"])
    return LlamaAttnLRPEngine.from_weights(dims, w, device="cuda", **kw)


@pytest.mark.parametrize("name", ["llama_tiny_d64.npz", "llama_tiny_d128.npz"])
@pytest.mark.parametrize("store", ["all", "sqrt"])
def test_engine_matches_reference_golden(name, store):
    cfg, w, ids, z = load_llama_golden(name)
    eng = _engine(cfg, w, micro_batch=4, store=store)
    rel, aux = eng.attribute_device(ids.cuda(), return_aux=True)
    rel = rel.float().cpu()
    assert np.array_equal(aux["idx"].cpu().numpy(), z["idx_fp32_sdpa"])
    ref_gap = rel_l2(z["rel_bf16_sdpa"], z["rel_fp32_sdpa"])  # the reference's own bf16-vs-fp32 distance
    err = rel_l2(rel, z["rel_fp32_sdpa"])
    print(f"{name} store={store}: rel-L2 vs reference fp32 = {err:.3e} (reference bf16 vs fp32 = {ref_gap:.3e})")
    # bf16 activation storage: a bf16 pipeline cannot reach 1e-3 against an fp32 run (one bf16 rounding alone
    # is ~1.6e-3 rel-L2); the bar is "the same distance from the fp32 reference as the reference's own bf16
    # run" (2.0e-3..2.4e-3 on these fixtures), stated as an absolute 4e-3.  Kernel-level tests hold 1e-3.
    assert err < 4e-3
    cos = torch.nn.functional.cosine_similarity(rel.flatten().double(), torch.from_numpy(z["rel_fp32_sdpa"]).flatten().double(), dim=0)
    assert cos > 0.9999


def test_engine_public_api_host_roundtrip():
    cfg, w, ids, z = load_llama_golden("llama_tiny_d64.npz")
    eng = _engine(cfg, w, micro_batch=1)
    out = eng.attribute(ids.pin_memory())
    assert out.shape == ids.shape and out.dtype == torch.float32 and not out.is_cuda
    assert rel_l2(out, z["rel_fp32_sdpa"]) < 5e-3


def test_engine_matches_oracle_on_fresh_seed():
    from oracle import attnlrp_oracle as O
    cfg = dict(d=512, I=1024, H=8, Hkv=2, D=64, L=3, V=1000, eps=1e-5, theta=10000.0)
    w = O.random_llama_weights(cfg, seed=11)
    ids = torch.randint(0, cfg["V"], (3, 200), generator=torch.Generator().manual_seed(5))
    ref, aux = O.llama_attnlrp(w, ids, cfg, dtype=torch.float32, return_aux=True)
    eng = _engine(cfg, w, micro_batch=3)
    rel, a2 = eng.attribute_device(ids.cuda(), return_aux=True)
    assert torch.equal(a2["idx"].cpu().long(), aux["idx"])
    assert rel_l2(rel.cpu(), ref) < 5e-3


def test_engine_cp_lrp_matches_reference_golden():
    from helpers import load_npz
    from oracle import attnlrp_oracle as O
    z = load_npz("llama_tiny_cp.npz")
    cfg = dict(d=256, I=512, H=4, Hkv=2, D=64, L=2, V=256, eps=1e-5, theta=10000.0)
    w = O.random_llama_weights(cfg, seed=0)
    eng = _engine(cfg, w, micro_batch=2, rule="cp")
    rel, aux = eng.attribute_device(torch.from_numpy(z["ids"]).cuda(), return_aux=True)
    assert np.array_equal(aux["idx"].cpu().numpy(), z["idx"])
    err = rel_l2(rel.cpu(), z["rel_fp32"])
    print(f"CP-LRP engine rel-L2 vs reference fp32 = {err:.3e}")
    assert err < 4e-3


def test_engine_tinyllama_dims_seq512():
    """BASELINE configs[1] at reduced depth: TinyLlama-1.1B widths (d=2048, I=5632, 32/4 heads, D=64), S=512, B=1."""
    from oracle import attnlrp_oracle as O
    cfg = dict(d=2048, I=5632, H=32, Hkv=4, D=64, L=2, V=4096, eps=1e-5, theta=10000.0)
    w = O.random_llama_weights(cfg, seed=21)
    ids = torch.randint(0, cfg["V"], (1, 512), generator=torch.Generator().manual_seed(22))
    ref, aux = O.llama_attnlrp(w, ids, cfg, dtype=torch.float32, return_aux=True)
    ref_bf16, aux_bf16 = O.llama_attnlrp(w, ids, cfg, dtype=torch.bfloat16, return_aux=True)  # reference-style bf16 run
    rel, a2 = _engine(cfg, w, micro_batch=1).attribute_device(ids.cuda(), return_aux=True)
    assert torch.equal(a2["idx"].cpu().long(), aux["idx"])
    err, gerr = rel_l2(rel.cpu(), ref), rel_l2(a2["g_emb"].cpu(), aux["g_emb"])
    err_b, gerr_b = rel_l2(ref_bf16, ref), rel_l2(aux_bf16["g_emb"].float(), aux["g_emb"])
    print(f"TinyLlama-width rel-L2 vs oracle fp32: engine relevance {err:.3e} g_emb {gerr:.3e}; "
          f"bf16 oracle run relevance {err_b:.3e} g_emb {gerr_b:.3e}")
    # At d=2048 even the forward logits of any bf16 pipeline are ~1e-2 from fp32.  The bar is self-calibrating: the
    # engine (fp32 residual/gradient streams) must be at least as close to the fp32 oracle as a reference-style
    # bf16 run of the same model, with an absolute ceiling.
    assert err <= 1.1 * err_b and gerr <= 1.1 * gerr_b
    assert err < 1.5e-2 and gerr < 2e-2


def test_engine_latent_relevance_trace():
    """per-layer latent relevance as a by-product of the backward sweep, vs the reference's hooks"""
    cfg, w, ids, z = load_llama_golden("llama_tiny_d128.npz")
    rel, aux = _engine(cfg, w, micro_batch=1).attribute_device(ids.cuda(), trace=True)
    tr = aux["layer_relevance"].cpu()
    assert tuple(tr.shape) == (cfg["L"],) + tuple(ids.shape)
    err = rel_l2(tr, z["trace_fp32_sdpa"])
    print(f"latent trace rel-L2 vs reference hooks = {err:.3e}")
    assert err < 6e-3


def test_engine_cuda_graph_replay_matches_eager():
    cfg, w, ids, z = load_llama_golden("llama_tiny_d64.npz")
    eager = _engine(cfg, w, micro_batch=2).attribute(ids.pin_memory()).clone()
    eng = _engine(cfg, w, micro_batch=2, cuda_graph=True)
    for _ in range(3):  # first call captures, the next ones replay
        got = eng.attribute(ids.pin_memory()).clone()
        assert rel_l2(got, eager) < 1e-5   # dQ uses fp32 atomics: order-dependent in the last bits only
    ids2 = torch.roll(ids, 1, 0)           # different content, same shape: replay must pick up the new ids
    assert rel_l2(eng.attribute(ids2.pin_memory()), torch.roll(eager, 1, 0)) < 1e-5


def test_engine_cuda_graph_survives_partial_micro_batches_and_other_shapes():
    """a trailing partial micro-batch (N % micro_batch != 0) or any other [B,S] between two graphed calls must not
    invalidate the captured graph's workspace."""
    cfg, w, ids, z = load_llama_golden("llama_tiny_d64.npz")
    ids3 = torch.cat([ids, ids[:1]], 0)                       # N = 3, micro_batch = 2 -> one graphed batch + one eager tail
    eager = _engine(cfg, w, micro_batch=2).attribute(ids3.pin_memory()).clone()
    eng = _engine(cfg, w, micro_batch=2, cuda_graph=True)
    for _ in range(3):
        got = eng.attribute(ids3.pin_memory()).clone()
        assert rel_l2(got, eager) < 1e-5
        eng.attribute_device(ids[:1, :96].cuda())             # a different shape through the shared workspace slot
        torch.cuda.empty_cache()
    assert rel_l2(eng.attribute(ids3.pin_memory()), eager) < 1e-5


def test_engine_from_hf_reproduces_scaled_rope():
    """rope_scaling (llama3 / linear) must be honoured by from_hf — engine vs the patched HF model itself."""
    from transformers import LlamaConfig, LlamaForCausalLM
    from transformers.models.llama import modeling_llama
    from lxt_b200.efficient import monkey_patch
    from lxt_b200.engine import LlamaAttnLRPEngine
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        monkey_patch(modeling_llama)
    for rope in ({"rope_type": "llama3", "rope_theta": 10000.0, "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
                  "original_max_position_embeddings": 64},
                 {"rope_type": "linear", "rope_theta": 10000.0, "factor": 4.0}):
        torch.manual_seed(0)
        hf = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                         head_dim=64, vocab_size=256, rope_parameters=rope, max_position_embeddings=512, tie_word_embeddings=False)
        hf._attn_implementation = "sdpa"
        from helpers import build_hf
        model = build_hf(LlamaForCausalLM, hf).cuda().eval()
        for p_ in model.parameters():
            p_.requires_grad_(False)
        ids = torch.randint(0, 256, (2, 200), generator=torch.Generator().manual_seed(3)).cuda()
        emb = model.get_input_embeddings()(ids).detach().requires_grad_()
        logits = model(inputs_embeds=emb, use_cache=False).logits
        mx, mi = torch.max(logits[:, -1, :], dim=-1)
        mx.sum().backward()
        rel_hf = (emb * emb.grad).float().sum(-1).detach().cpu()
        eng = LlamaAttnLRPEngine.from_hf(model, micro_batch=2)
        rel, aux = eng.attribute_device(ids, return_aux=True)
        plain = LlamaAttnLRPEngine.from_hf(model, micro_batch=2, rope=None).attribute_device(ids).cpu()   # default tables: must differ
        e, e_plain = rel_l2(rel.cpu(), rel_hf), rel_l2(plain, rel_hf)
        print(f"rope {rope['rope_type']}: engine vs patched HF {e:.2e}; with default RoPE tables {e_plain:.2e}")
        assert torch.equal(aux["idx"].cpu().long(), mi.cpu())
        assert e < 1.5e-2 and e_plain > 3 * e


@pytest.mark.parametrize("name", ["gemma3_tiny.npz", "gemma3_tiny_d256.npz"])
def test_engine_gemma3_matches_reference_golden(name):
    """The fused engine on the Gemma-3 layer layout ((1+w) norms, pre/post norms around both branches, per-head q/k-norm,
    GELU-tanh gate, sliding-window + global layers with their own RoPE base, scaled embedding, tied lm_head) vs the golden
    relevance of the real reference (lxt.efficient.monkey_patch(modeling_gemma3)) and vs the CPU oracle."""
    from helpers import load_gemma_golden
    from lxt_b200.engine import LlamaAttnLRPEngine, LlamaDims
    from oracle import attnlrp_oracle as O
    cfg, w, ids, z = load_gemma_golden(name)
    keys = ("d", "I", "H", "Hkv", "D", "L", "V", "eps", "theta", "norm_offset", "act", "qk_norm", "post_norms", "windows", "thetas",
            "attn_scale", "emb_scale")
    eng = LlamaAttnLRPEngine.from_weights(LlamaDims(**{k: cfg[k] for k in keys}), w, device="cuda", micro_batch=2)
    rel, aux = eng.attribute_device(ids.cuda(), return_aux=True)
    assert np.array_equal(aux["idx"].cpu().numpy(), z["idx"])
    err = rel_l2(rel.cpu(), z["rel_fp32"])
    ora = O.decoder_attnlrp(w, ids, cfg, dtype=torch.float32)
    print(f"{name}: Gemma-3 engine rel-L2 vs reference fp32 = {err:.3e}, vs oracle = {rel_l2(rel.cpu(), ora):.3e}")
    assert err < 6e-3
