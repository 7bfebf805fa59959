"""GPU: the validation-precision mode (`precision="high"`): fp32 activations, every GEMM as a two-term bf16 split on the
SAME tcgen05 kernels, fp32 instantiations of the same point-wise templates, fp32 CUDA-core attention.

What it proves: the stated tolerance of the path is 1e-3 rel-L2 against the reference; the bf16 production mode cannot meet
it because one bf16 storage rounding is already 1.6e-3.  With storage rounding removed and nothing else changed (same
kernels, same launch order, same rules) the engine must sit at <= 1e-3 from the real reference's fp32 golden on EVERY
fixture — then the bf16 mode's 2e-3..6e-3 is demonstrably rounding, not a defect.  Tolerance: 1e-3 (BASELINE.json north_star).
"""
import numpy as np
import pytest
import torch

from helpers import bf16_from_bits, load_gemma_golden, load_llama_golden, load_npz, rel_l2

pytestmark = pytest.mark.gpu

TOL = 1e-3   # the tolerance north_star states


def _llama_engine(cfg, w, **kw):
    from lxt_b200.engine import LlamaAttnLRPEngine, LlamaDims
    dims = LlamaDims(d=cfg["d"], I=cfg["I"], H=cfg["H"], Hkv=cfg["Hkv"], D=cfg["D"], L=cfg["L"], V=cfg["V"], eps=cfg["eps"],
                     theta=cfg["theta"])
    return LlamaAttnLRPEngine.from_weights(dims, w, device="cuda", **kw)


def test_split_gemm_is_fp32_accurate():
    """the two-term bf16 split through the tcgen05 GEMM vs an fp64 matmul: NT (forward) and NN (LRP dgrad) layouts, fused epilogue"""
    from lxt_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    T, K, N = 520, 1024, 768
    x = torch.randn(T, K, generator=g, device="cuda")
    w = (torch.randn(N, K, generator=g, device="cuda") * 0.05).to(torch.bfloat16)
    resid = torch.randn(T, N, generator=g, device="cuda")
    rs = torch.rand(T, generator=g, device="cuda") + 0.5
    cs = torch.rand(N, generator=g, device="cuda") + 0.5
    out = torch.empty(T, N, device="cuda")
    ops.linear_fwd(x, w, out, resid=resid, rowscale=rs, colscale=cs)
    ref = resid.double() + (x.double() @ w.double().T) * rs.double()[:, None] * cs.double()[None, :]
    e_nt = rel_l2(out.cpu(), ref.cpu())
    gy = torch.randn(T, N, generator=g, device="cuda")
    gx = torch.empty(T, K, device="cuda")
    ops.linear_dgrad(gy, w, gx)
    e_nn = rel_l2(gx.cpu(), (gy.double() @ w.double()).cpu())
    # the bf16 GEMM on the same operands, for scale
    y16 = torch.empty(T, N, device="cuda")
    ops.linear_fwd(x.to(torch.bfloat16), w, y16)
    e16 = rel_l2(y16.cpu(), (x.double() @ w.double().T).cpu())
    print(f"split GEMM rel-L2 vs fp64: NT {e_nt:.2e}, NN {e_nn:.2e}  (one-term bf16 GEMM: {e16:.2e})")
    assert e_nt < 2e-5 and e_nn < 2e-5 and e16 > 5e-4


@pytest.mark.parametrize("shape", [(2, 300, 4, 2, 64, True, 0), (1, 520, 4, 1, 128, True, 200), (2, 197, 4, 4, 64, False, 0),
                                   (1, 260, 2, 1, 256, True, 0)])
def test_fp32_attention_kernels_vs_torch_fp64(shape):
    from lxt_b200 import ops
    B, S, H, Hkv, D, causal, window = shape
    g = torch.Generator(device="cuda").manual_seed(3)
    q = torch.randn(B, S, H, D, generator=g, device="cuda")
    k = torch.randn(B, S, Hkv, D, generator=g, device="cuda")
    v = torch.randn(B, S, Hkv, D, generator=g, device="cuda")
    d_o = torch.randn(B, S, H, D, generator=g, device="cuda")
    scale = D ** -0.5
    o, lse = ops.attn_fwd(q, k, v, scale, causal=causal, window=window)
    dq, dk, dv = ops.attn_bwd(q, k, v, o, d_o, lse, scale, causal=causal, window=window)
    qd, kd, vd = (t.double().requires_grad_() for t in (q, k, v))
    G = H // Hkv
    kr, vr = kd.repeat_interleave(G, 2), vd.repeat_interleave(G, 2)
    sc = torch.einsum("bshd,bthd->bhst", qd, kr) * scale
    i = torch.arange(S, device="cuda")
    bad = torch.zeros(S, S, dtype=torch.bool, device="cuda")
    if causal:
        bad |= i[None, :] > i[:, None]
    if window:
        bad |= (i[:, None] - i[None, :]) >= window
    p = torch.softmax(sc.masked_fill(bad, float("-inf")), -1)
    oref = torch.einsum("bhst,bthd->bshd", p, vr)
    oref.backward(d_o.double())
    errs = dict(o=rel_l2(o.cpu(), oref.detach().cpu()), dq=rel_l2(dq.cpu(), (qd.grad / 4).cpu()), dk=rel_l2(dk.cpu(), (kd.grad / 4).cpu()),
                dv=rel_l2(dv.cpu(), (vd.grad / 2).cpu()))
    print(shape, {k_: f"{v_:.1e}" for k_, v_ in errs.items()})
    assert max(errs.values()) < 2e-5


@pytest.mark.parametrize("name", ["llama_tiny_d64.npz", "llama_tiny_d128.npz"])
def test_engine_high_precision_meets_1e3_on_llama_goldens(name):
    cfg, w, ids, z = load_llama_golden(name)
    rel, aux = _llama_engine(cfg, w, micro_batch=2, precision="high").attribute_device(ids.cuda(), return_aux=True)
    rel16 = _llama_engine(cfg, w, micro_batch=2).attribute_device(ids.cuda()).cpu()
    assert np.array_equal(aux["idx"].cpu().numpy(), z["idx_fp32_sdpa"])
    err = rel_l2(rel.cpu(), z["rel_fp32_sdpa"])
    gerr = rel_l2(aux["g_emb"].cpu(), z["gemb_fp32_sdpa"])
    e16, ref16 = rel_l2(rel16, z["rel_fp32_sdpa"]), rel_l2(z["rel_bf16_sdpa"], z["rel_fp32_sdpa"])
    print(f"{name}: validation mode rel-L2 vs reference fp32 = {err:.2e} (g_emb {gerr:.2e}); bf16 mode {e16:.2e}; reference's own bf16 run {ref16:.2e}")
    assert err <= TOL and gerr <= TOL
    # self-calibrated bf16 bar: the same order of distance from fp32 as the reference's own bf16 run.  Two bf16 pipelines round at
    # different points, so their distances to fp32 are two draws of the same noise: measured ratio 0.57 (d128) .. 1.27 (d64).
    assert e16 <= 1.5 * ref16
    # and the bf16 engine against the reference's bf16 run (two bf16 pipelines: their distance is bounded by the sum)
    assert rel_l2(rel16, z["rel_bf16_sdpa"]) <= e16 + ref16


def test_engine_high_precision_cp_lrp_golden():
    from oracle import attnlrp_oracle as O
    z = load_npz("llama_tiny_cp.npz")
    cfg = dict(d=256, I=512, H=4, Hkv=2, D=64, L=2, V=256, eps=1e-5, theta=10000.0)
    w = O.random_llama_weights(cfg, seed=0)
    rel = _llama_engine(cfg, w, micro_batch=2, rule="cp", precision="high").attribute_device(torch.from_numpy(z["ids"]).cuda())
    err = rel_l2(rel.cpu(), z["rel_fp32"])
    print(f"CP-LRP validation mode rel-L2 vs reference fp32 = {err:.2e}")
    assert err <= TOL


@pytest.mark.parametrize("name", ["gemma3_tiny.npz", "gemma3_tiny_d256.npz"])
def test_engine_high_precision_gemma3_goldens(name):
    from lxt_b200.engine import LlamaAttnLRPEngine, LlamaDims
    cfg, w, ids, z = load_gemma_golden(name)
    keys = ("d", "I", "H", "Hkv", "D", "L", "V", "eps", "theta", "norm_offset", "act", "qk_norm", "post_norms", "windows", "thetas",
            "attn_scale", "emb_scale")
    eng = LlamaAttnLRPEngine.from_weights(LlamaDims(**{k: cfg[k] for k in keys}), w, device="cuda", micro_batch=2, precision="high")
    rel, aux = eng.attribute_device(ids.cuda(), return_aux=True)
    assert np.array_equal(aux["idx"].cpu().numpy(), z["idx"])
    err = rel_l2(rel.cpu(), z["rel_fp32"])
    print(f"{name}: Gemma-3 validation mode rel-L2 vs reference fp32 = {err:.2e}")
    assert err <= TOL


@pytest.mark.parametrize("family", ["qwen2", "qwen3"])
def test_engine_high_precision_qwen_goldens(family):
    import transformers
    from lxt_b200.engine import LlamaAttnLRPEngine
    z = load_npz(f"{family}_tiny.npz")
    kw = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=1,
              vocab_size=384, max_position_embeddings=512, rms_norm_eps=1e-6, tie_word_embeddings=False)
    if family == "qwen3":
        kw["head_dim"] = 64
    from helpers import build_hf
    model = build_hf(getattr(transformers, f"{family.capitalize()}ForCausalLM"), getattr(transformers, f"{family.capitalize()}Config")(**kw))
    model.load_state_dict({k[3:]: bf16_from_bits(v) for k, v in z.items() if k.startswith("sd_")}, strict=True)
    eng = LlamaAttnLRPEngine.from_hf(model, micro_batch=2, precision="high")
    rel, aux = eng.attribute_device(torch.from_numpy(z["ids"]).cuda(), return_aux=True)
    assert np.array_equal(aux["idx"].cpu().numpy(), z["idx"])
    err = rel_l2(rel.cpu(), z["rel_fp32"])
    print(f"{family}: validation mode rel-L2 vs reference fp32 = {err:.2e}")
    assert err <= TOL
