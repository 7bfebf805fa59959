"""BASELINE configs[1]: TinyLlama-1.1B dims, random init, seq 512, one B200.  Three ways to get the same relevance:
(a) drop-in API  lxt_b200.efficient.monkey_patch(modeling_llama) on the HuggingFace model (unchanged user code),
(b) the engine launch-by-launch, (c) the engine replaying one CUDA graph.  Prints attributions/s for each."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lrp-explains-transformers_b200"))

import torch  # noqa: E402
from transformers import LlamaConfig, LlamaForCausalLM  # noqa: E402
from transformers.models.llama import modeling_llama  # noqa: E402

from lxt_b200.efficient import monkey_patch  # noqa: E402
from lxt_b200.engine import LlamaAttnLRPEngine  # noqa: E402


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n, out


def main(S=512, B=1):
    cfg = LlamaConfig(hidden_size=2048, intermediate_size=5632, num_hidden_layers=22, num_attention_heads=32,
                      num_key_value_heads=4, head_dim=64, vocab_size=32000, rms_norm_eps=1e-5, max_position_embeddings=2048,
                      rope_parameters={"rope_type": "default", "rope_theta": 10000.0}, tie_word_embeddings=False)
    cfg._attn_implementation = "sdpa"
    torch.manual_seed(0)
    torch.set_default_dtype(torch.bfloat16)   # bf16 parameters, fp32 RoPE tables (like from_pretrained(torch_dtype=bfloat16))
    try:
        model = LlamaForCausalLM(cfg).cuda().eval()
    finally:
        torch.set_default_dtype(torch.float32)
    for p in model.parameters():
        p.requires_grad_(False)
    ids = torch.randint(0, cfg.vocab_size, (B, S), generator=torch.Generator().manual_seed(1)).cuda()
    eng = LlamaAttnLRPEngine.from_hf(model, micro_batch=B)
    eng_g = LlamaAttnLRPEngine.from_hf(model, micro_batch=B, cuda_graph=True)
    monkey_patch(modeling_llama)

    def hf_path():
        emb = model.get_input_embeddings()(ids).detach().requires_grad_()
        logits = model(inputs_embeds=emb, use_cache=False).logits
        logits[:, -1, :].max(-1).values.sum().backward()
        return (emb * emb.grad).float().sum(-1)

    t_hf, r_hf = timed(hf_path)
    t_en, r_en = timed(lambda: eng.attribute_device(ids))
    t_gr, r_gr = timed(lambda: eng_g.attribute_device_graphed(ids))
    rl2 = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    print(f"TinyLlama-1.1B dims, S={S}, B={B}: monkey_patch(HF) {B / t_hf:.1f} attr/s | engine {B / t_en:.1f} attr/s | "
          f"engine + CUDA graph {B / t_gr:.1f} attr/s ; rel-L2 HF-path vs engine {rl2(r_hf, r_en):.2e}, graph vs engine {rl2(r_gr, r_en):.2e}")


if __name__ == "__main__":
    main()
