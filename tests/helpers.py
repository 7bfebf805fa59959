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
