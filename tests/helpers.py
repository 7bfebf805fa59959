"""shared test helpers (CPU side)"""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_npz(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def bf16_from_bits(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(a.astype(np.int16)).view(torch.bfloat16)


def load_llama_golden(name):
    """-> (cfg dict, weights dict (bf16 tensors), ids int64 tensor, raw npz dict)"""
    z = load_npz(name)
    cfg = {k: (float(v) if k in ("eps", "theta") else int(v)) for k, v in zip(z["cfg_keys"], z["cfg_vals"])}
    w = {"emb": bf16_from_bits(z["w_emb"]), "norm": bf16_from_bits(z["w_norm"]), "lm_head": bf16_from_bits(z["w_lm_head"]),
         "layers": []}
    for i in range(cfg["L"]):
        w["layers"].append({k: bf16_from_bits(z[f"w_l{i}_{k}"]) for k in ("wq", "wk", "wv", "wo", "wg", "wu", "wd", "ln1", "ln2")})
    return cfg, w, torch.from_numpy(z["ids"]), z


def rel_l2(a, b) -> float:
    a = torch.as_tensor(a).double().flatten()
    b = torch.as_tensor(b).double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-300))


def load_gemma_golden(name):
    """-> (cfg for oracle.decoder_attnlrp / engine, weights dict (bf16), ids, raw npz) from tests/golden/gemma3_tiny*.npz"""
    z = load_npz(name)
    sd = {k[3:]: bf16_from_bits(v) for k, v in z.items() if k.startswith("sd_")}
    L = len(z["layer_types"])
    D = sd["model.layers.0.self_attn.q_norm.weight"].shape[0]
    d = sd["model.embed_tokens.weight"].shape[1]
    H = sd["model.layers.0.self_attn.q_proj.weight"].shape[0] // D
    Hkv = sd["model.layers.0.self_attn.k_proj.weight"].shape[0] // D
    cfg = dict(d=d, I=sd["model.layers.0.mlp.gate_proj.weight"].shape[0], H=H, Hkv=Hkv, D=D, L=L, V=sd["model.embed_tokens.weight"].shape[0],
               eps=1e-6, theta=10000.0, norm_offset=1, act="gelu_tanh", qk_norm=True, post_norms=True,
               windows=[48 if t == "sliding_attention" else 0 for t in z["layer_types"]],
               thetas=[10000.0 if t == "sliding_attention" else 1000000.0 for t in z["layer_types"]],
               attn_scale=float(D) ** -0.5, emb_scale=float(d) ** 0.5)
    w = {"emb": sd["model.embed_tokens.weight"], "norm": sd["model.norm.weight"], "lm_head": sd["model.embed_tokens.weight"], "layers": []}
    for i in range(L):
        p = f"model.layers.{i}."
        w["layers"].append(dict(wq=sd[p + "self_attn.q_proj.weight"], wk=sd[p + "self_attn.k_proj.weight"], wv=sd[p + "self_attn.v_proj.weight"],
                                wo=sd[p + "self_attn.o_proj.weight"], qn=sd[p + "self_attn.q_norm.weight"], kn=sd[p + "self_attn.k_norm.weight"],
                                wg=sd[p + "mlp.gate_proj.weight"], wu=sd[p + "mlp.up_proj.weight"], wd=sd[p + "mlp.down_proj.weight"],
                                ln1=sd[p + "input_layernorm.weight"], ln_post_attn=sd[p + "post_attention_layernorm.weight"],
                                ln_pre_ff=sd[p + "pre_feedforward_layernorm.weight"], ln_post_ff=sd[p + "post_feedforward_layernorm.weight"]))
    return cfg, w, torch.from_numpy(z["ids"]), z


def build_vit_l16(seed=5):
    """same construction as tests/golden/make_golden.py::build_vit_l16 (seeded torchvision vit_l_16, head re-initialised)"""
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


def build_hf(Model, cfg, dtype=torch.bfloat16):
    """Construct an HF model the way `from_pretrained(torch_dtype=...)` does: parameters in `dtype`, RoPE `inv_freq` buffers in
    fp32.  (`Model(cfg).to(torch.bfloat16)` also rounds the non-persistent inv_freq buffer to bf16, and HF computes RoPE from that
    buffer: at position 2048 the angle of the fastest dimension is then off by radians — a property of that model object, which
    both the reference and this repo reproduce faithfully, but not what the fp32 goldens / oracle compute.)"""
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        return Model(cfg)
    finally:
        torch.set_default_dtype(old)


# ---- BASELINE.json configurations at full width (tests/test_baseline_configs_gpu.py, tests/golden/make_baseline_oracle.py) ----
BASELINE_CASES = ("llama3_8b_l2", "tinyllama_22l", "gemma3_4b_l6")


def baseline_case(name: str, S_override: int = 0):
    """-> (tag, cfg, weights (bf16, seeded), ids [1,S], kind) with kind 'llama' (oracle.llama_attnlrp) or 'decoder' (oracle.decoder_attnlrp)"""
    from oracle import attnlrp_oracle as O
    if name == "llama3_8b_l2":
        cfg = dict(d=4096, I=14336, H=32, Hkv=8, D=128, L=2, V=128256, eps=1e-5, theta=500000.0)
        S, seed = S_override or 2048, 31
        w = O.random_llama_weights(cfg, seed=seed)
        ids = torch.randint(0, cfg["V"], (1, S), generator=torch.Generator().manual_seed(seed + 1))
        return "llama3-8b dims, L=2", cfg, w, ids, "llama"
    if name == "tinyllama_22l":
        cfg = dict(d=2048, I=5632, H=32, Hkv=4, D=64, L=22, V=32000, eps=1e-5, theta=10000.0)
        S, seed = S_override or 512, 41
        w = O.random_llama_weights(cfg, seed=seed)
        ids = torch.randint(0, cfg["V"], (1, S), generator=torch.Generator().manual_seed(seed + 1))
        return "tinyllama-1.1b dims, all 22 layers", cfg, w, ids, "llama"
    if name == "gemma3_4b_l6":
        L, pattern, S = 6, 6, S_override or 8192
        d, I, H, Hkv, D, V = 2560, 10240, 8, 4, 256, 262208
        glob = [(l + 1) % pattern == 0 for l in range(L)]
        cfg = dict(d=d, I=I, H=H, Hkv=Hkv, D=D, L=L, V=V, eps=1e-6, theta=10000.0, norm_offset=1.0, act="gelu_tanh", qk_norm=True,
                   post_norms=True, windows=[0 if g else 1024 for g in glob], thetas=[1000000.0 if g else 10000.0 for g in glob],
                   attn_scale=float(D) ** -0.5, emb_scale=float(d) ** 0.5)
        g = torch.Generator().manual_seed(51)
        rn = lambda *s: (torch.randn(*s, generator=g) * 0.02).to(torch.bfloat16)
        nw = lambda n: (torch.randn(n, generator=g) * 0.05).to(torch.bfloat16)    # (1 + w) norms: non-trivial weights
        layers = [dict(wq=rn(H * D, d), wk=rn(Hkv * D, d), wv=rn(Hkv * D, d), wo=rn(d, H * D), wg=rn(I, d), wu=rn(I, d), wd=rn(d, I),
                       ln1=nw(d), qn=nw(D), kn=nw(D), ln_post_attn=nw(d), ln_pre_ff=nw(d), ln_post_ff=nw(d)) for _ in range(L)]
        emb = rn(V, d)
        w = dict(emb=emb, norm=nw(d), lm_head=emb, layers=layers)
        ids = torch.randint(0, V, (1, S), generator=torch.Generator().manual_seed(52))
        return "gemma3-4b dims, 6 layers (5 sliding-window + 1 global)", cfg, w, ids, "decoder"
    raise KeyError(name)


def baseline_oracle(name: str, S_override: int = 0, want_bf16: bool = True):
    """fp32 (and reference-style bf16) CPU oracle of a BASELINE case: minutes of host time at full width.
    -> dict(ref, idx, g_rows, row_stride[, ref16, idx16, g16_rows]) of numpy arrays; g_rows = every row_stride-th row of g_emb"""
    from oracle import attnlrp_oracle as O
    tag, cfg, w, ids, kind = baseline_case(name, S_override)
    fn = O.llama_attnlrp if kind == "llama" else O.decoder_attnlrp
    torch.set_num_threads(os.cpu_count() or 8)
    S = ids.shape[1]
    stride = max(1, S // 32)
    ref, aux = fn(w, ids, cfg, dtype=torch.float32, return_aux=True)
    out = dict(ref=ref.numpy(), idx=aux["idx"].numpy(), g_rows=aux["g_emb"][:, ::stride].float().numpy(), row_stride=np.array([stride]),
               S=np.array([S]))
    if want_bf16 and kind == "llama":
        ref16, aux16 = fn(w, ids, cfg, dtype=torch.bfloat16, return_aux=True)
        out.update(ref16=ref16.float().numpy(), idx16=aux16["idx"].numpy(), g16_rows=aux16["g_emb"][:, ::stride].float().numpy())
    return out
