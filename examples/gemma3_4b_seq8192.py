"""BASELINE configs[4]: Gemma-3-4B text dims (34 layers, d=2560, I=10240, 8/4 heads, head_dim 256, sliding window 1024 with
one global layer in six), random init, ONE 8192-token prompt, one B200, through the drop-in API
`lxt_b200.efficient.monkey_patch(modeling_gemma3)` — unchanged HuggingFace model code, flash AttnLRP (no [B,H,S,S] tensor),
logits restricted to the last position by the standard `logits_to_keep=1` argument."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lrp-explains-transformers_b200"))

import torch  # noqa: E402
from transformers import Gemma3ForCausalLM, Gemma3TextConfig  # noqa: E402
from transformers.models.gemma3 import modeling_gemma3  # noqa: E402

from lxt_b200 import ops  # noqa: E402
from lxt_b200.efficient import monkey_patch  # noqa: E402


def main(S=8192, layers=34):
    cfg = Gemma3TextConfig(hidden_size=2560, intermediate_size=10240, num_hidden_layers=layers, num_attention_heads=8,
                           num_key_value_heads=4, head_dim=256, vocab_size=262208, sliding_window=1024,
                           max_position_embeddings=131072, query_pre_attn_scalar=256, rms_norm_eps=1e-6, tie_word_embeddings=True)
    cfg._attn_implementation = "sdpa"
    torch.manual_seed(0)
    # bf16 parameters, fp32 RoPE tables — as `from_pretrained(torch_dtype=torch.bfloat16)` builds it (a later `.to(torch.bfloat16)` would
    # also round HF's non-persistent inv_freq buffer, which ruins RoPE at long context)
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device("cuda"):
            model = Gemma3ForCausalLM(cfg).eval()
    finally:
        torch.set_default_dtype(torch.float32)
    for p in model.parameters():
        p.requires_grad_(False)
    monkey_patch(modeling_gemma3)
    ids = torch.randint(0, cfg.vocab_size, (1, S), generator=torch.Generator().manual_seed(1)).cuda()

    def once():
        emb = model.get_input_embeddings()(ids).detach().requires_grad_()
        logits = model(inputs_embeds=emb, use_cache=False, logits_to_keep=1).logits
        logits[:, -1, :].max(-1).values.sum().backward()
        return (emb * emb.grad).float().sum(-1)

    once()
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    n0, t0 = ops.launch_count(), time.perf_counter()
    rel = once()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # the fused engine on the same weights (no autograd, fused epilogues, last-position lm_head)
    from lxt_b200.engine import LlamaAttnLRPEngine
    eng = LlamaAttnLRPEngine.from_hf(model, micro_batch=1)
    r_eng = eng.attribute_device(ids)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    r_eng = eng.attribute_device(ids)
    torch.cuda.synchronize()
    dt_eng = time.perf_counter() - t1
    rl2 = float((r_eng.double() - rel.double()).norm() / rel.double().norm())
    print(f"fused engine: {dt_eng * 1e3:.0f} ms per attribution = {1 / dt_eng:.2f} attributions/s; rel-L2 engine vs drop-in path {rl2:.2e}")
    print(f"Gemma-3-4B dims, {layers} layers ({cfg.layer_types.count('full_attention')} global), S={S}, B=1: {dt * 1e3:.0f} ms per attribution "
          f"= {1 / dt:.2f} attributions/s ({ops.launch_count() - n0} B200 kernel launches); relevance {tuple(rel.shape)}, "
          f"finite={bool(torch.isfinite(rel).all())}, peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")


if __name__ == "__main__":
    main()
