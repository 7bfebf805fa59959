"""GPU: 4-bit NormalFloat weight storage.  The CUDA quantiser / de-quantiser are pinned BIT-EXACTLY against
oracle/nf4_oracle.py; the engine on NF4 weights is pinned against the fp32 AttnLRP oracle evaluated on the de-quantised weights
(bitsandbytes is absent: parity unpinned against bnb itself)."""
import numpy as np
import pytest
import torch

from helpers import load_llama_golden, rel_l2

pytestmark = pytest.mark.gpu


def test_cuda_nf4_matches_the_oracle_bit_exactly():
    from lxt_b200 import ops
    from oracle import nf4_oracle as Q
    g = torch.Generator().manual_seed(1)
    for shape, std in (((256, 512), 0.02), ((1024, 4096), 0.02), ((64, 64), 3.0)):
        w = (torch.randn(*shape, generator=g) * std).to(torch.bfloat16)
        w[0, :64] = 0                                      # an all-zero block
        packed, absmax = ops.quant_nf4(w.cuda())
        p_ref, a_ref = Q.quantize_nf4(w)
        assert torch.equal(absmax.cpu(), a_ref)
        assert torch.equal(packed.cpu(), p_ref), f"{int((packed.cpu() != p_ref).sum())} code bytes differ"
        out = ops.dequant_nf4(packed, absmax, torch.empty(shape, dtype=torch.bfloat16, device="cuda"))
        assert torch.equal(out.cpu(), Q.dequantize_nf4(p_ref, a_ref, shape))


def test_engine_on_nf4_weights_matches_oracle_on_dequantised_weights():
    from lxt_b200.engine import LlamaAttnLRPEngine, LlamaDims
    from oracle import attnlrp_oracle as O
    from oracle import nf4_oracle as Q
    cfg = dict(d=512, I=1024, H=8, Hkv=2, D=64, L=3, V=1000, eps=1e-5, theta=10000.0)
    w = O.random_llama_weights(cfg, seed=11)
    ids = torch.randint(0, cfg["V"], (2, 200), generator=torch.Generator().manual_seed(5))
    dims = LlamaDims(**{k: cfg[k] for k in ("d", "I", "H", "Hkv", "D", "L", "V", "eps", "theta")})
    eng = LlamaAttnLRPEngine.from_weights(dims, w, device="cuda", micro_batch=2, quant="nf4")
    full = LlamaAttnLRPEngine.from_weights(dims, w, device="cuda", micro_batch=2)
    scratch = sum(t.numel() * t.element_size() for t in eng._wscratch.values())     # one layer's worth of bf16, shared by all layers
    assert eng.weight_bytes() - scratch < 0.30 * full.weight_bytes()              # 4.5 bits per weight instead of 16
    wq = dict(emb=w["emb"], norm=w["norm"], lm_head=w["lm_head"], layers=[])
    for lw in w["layers"]:
        e = dict(lw)
        for k in ("wq", "wk", "wv", "wo", "wg", "wu", "wd"):
            e[k] = Q.dequantize_nf4(*Q.quantize_nf4(lw[k]), lw[k].shape)
        wq["layers"].append(e)
    ref, aux = O.llama_attnlrp(wq, ids, cfg, dtype=torch.float32, return_aux=True)
    rel, a2 = eng.attribute_device(ids.cuda(), return_aux=True)
    assert torch.equal(a2["idx"].cpu().long(), aux["idx"])
    err = rel_l2(rel.cpu(), ref)
    # and the validation mode on the same 4-bit weights: de-quantised weights are bf16 values, so the 1e-3 tolerance applies
    hp = LlamaAttnLRPEngine.from_weights(dims, w, device="cuda", micro_batch=2, quant="nf4", precision="high")
    e_hp = rel_l2(hp.attribute_device(ids.cuda()).cpu(), ref)
    moved = rel_l2(full.attribute_device(ids.cuda()).cpu(), ref)     # how far 4-bit weights move the relevance at all
    print(f"NF4 engine vs oracle on de-quantised weights: bf16 {err:.2e}, validation mode {e_hp:.2e}; bf16-weight engine vs the same: {moved:.2e}")
    assert err < 6e-3 and e_hp <= 1e-3 and moved > 5 * err
