"""A/B tool (GPU): which part of the patched-HF bf16 path sets its distance from the fp32 oracle?

    python tests/tools/ab_hf_path_precision.py none|mlp|rms|lin|attn|mlp,rms,lin,attn

Each named patch uses the B200 kernels, the others a pure-torch autograd restatement of the reference patch (stock
kernels).  Round-1 result at TinyLlama width, 4 layers, S=512 (rel-L2 of relevance / g_emb vs the fp32 oracle):
  none 2.28e-2 / 4.01e-2   all ours 2.13e-2 / 4.01e-2   (any single patch: 2.1-2.3e-2)   engine 6.2e-3 / 9.4e-3
i.e. the distance is a property of the bf16 HuggingFace module graph (bf16 residual/gradient streams, stock bf16
reductions), not of the kernels; the engine (fp32 streams) is ~3.5x closer to fp32.
"""
import sys, torch
from functools import partial
sys.path.insert(0,'.'); sys.path.insert(0,'lrp-explains-transformers_b200'); sys.path.insert(0,'tests')
from oracle import attnlrp_oracle as O
from test_monkey_patch_gpu import _hf_model
from transformers.models.llama import modeling_llama
from transformers.models.llama.modeling_llama import LlamaMLP, LlamaRMSNorm
from lxt_b200.efficient import monkey_patch, patches as P, rules as R
variant = sys.argv[1]
cfg = dict(d=2048, I=5632, H=32, Hkv=4, D=64, L=4, V=32000, eps=1e-5, theta=10000.0)
w = O.random_llama_weights(cfg, seed=0)
ids = torch.randint(0, cfg["V"], (1, 512), generator=torch.Generator().manual_seed(1))
torch.set_num_threads(32)
r32, a32 = O.llama_attnlrp(w, ids, cfg, dtype=torch.float32, return_aux=True)

# torch re-statements of the reference's patches (autograd over stock kernels), for A/B only
class DivFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, f): ctx.f = f; return x
    @staticmethod
    def backward(ctx, g): return g / ctx.f, None
class IdFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fn, x):
        y = fn(x); ctx.save_for_backward(y / (x + 1e-10)); return y
    @staticmethod
    def backward(ctx, g): return None, ctx.saved_tensors[0] * g
def t_rms(self, h):
    dt = h.dtype; hf = h.float(); var = hf.pow(2).mean(-1, keepdim=True)
    return self.weight * (hf * torch.rsqrt(var + self.variance_epsilon).detach()).to(dt)
def t_mlp(self, x):
    g = IdFn.apply(self.act_fn, self.gate_proj(x))
    return self.down_proj(DivFn.apply(g * self.up_proj(x), 2))
def t_attn_wrap(fn):
    def f(module, q, k, v, *a, **kw):
        return fn(module, DivFn.apply(q, 4), DivFn.apply(k, 4), DivFn.apply(v, 2), *a, **kw)
    return f
def t_patch_attn(module):
    module.eager_attention_forward = t_attn_wrap(module.eager_attention_forward)
    for k_, v_ in list(module.ALL_ATTENTION_FUNCTIONS.items()):
        module.ALL_ATTENTION_FUNCTIONS[k_] = t_attn_wrap(v_)
    return True

ours = dict(mlp=partial(P.patch_method, P.gated_mlp_forward), rms=partial(P.patch_method, P.rms_norm_forward),
            lin=partial(P.patch_method, P.linear_forward, keep_original=True), attn=P.patch_attention)
ref = dict(mlp=partial(P.patch_method, t_mlp), rms=partial(P.patch_method, t_rms), lin=lambda t: True, attn=t_patch_attn)
use = {k: (ours[k] if k in variant.split(",") else ref[k]) for k in ours}
pm = {LlamaMLP: use["mlp"], LlamaRMSNorm: use["rms"], torch.nn.Linear: use["lin"], modeling_llama: use["attn"]}
monkey_patch(modeling_llama, pm)
model = _hf_model(cfg, w, "sdpa")
emb = model.get_input_embeddings()(ids.cuda()).detach().requires_grad_()
logits = model(inputs_embeds=emb, use_cache=False).logits
mx, mi = logits[:, -1, :].max(-1)
mx.sum().backward()
r_hf = (emb * emb.grad).float().sum(-1).detach().cpu()
print(f"variant[{variant}] vs fp32:", O.rel_l2(r_hf, r32), "g_emb", O.rel_l2(emb.grad.float().cpu(), a32["g_emb"]))
