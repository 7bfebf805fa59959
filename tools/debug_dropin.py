"""debug (GPU): patched-HF path vs the fused engine on the same weights, with individual patches swapped for torch restatements.
    python tools/debug_dropin.py <variant> [d I H Hkv D L V S]"""
import sys, time, torch
from functools import partial
sys.path.insert(0, '.'); sys.path.insert(0, 'lrp-explains-transformers_b200'); sys.path.insert(0, 'tests')
from oracle import attnlrp_oracle as O
from test_monkey_patch_gpu import _hf_model
from transformers.models.llama import modeling_llama
from transformers.models.llama.modeling_llama import LlamaMLP, LlamaRMSNorm
from lxt_b200.efficient import monkey_patch, patches as P
from lxt_b200.engine import LlamaAttnLRPEngine, LlamaDims
variant = sys.argv[1]
nums = [int(x) for x in sys.argv[2:]] if len(sys.argv) > 2 else [4096, 14336, 32, 8, 128, 1, 128256, 2048]
d, I, H, Hkv, D, L, V, S = nums
cfg = dict(d=d, I=I, H=H, Hkv=Hkv, D=D, L=L, V=V, eps=1e-5, theta=500000.0)
w = O.random_llama_weights(cfg, seed=31)
ids = torch.randint(0, V, (1, S), generator=torch.Generator().manual_seed(32))

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
t0 = time.time()
model = _hf_model(cfg, w, "sdpa", max_pos=max(4096, S))
emb = model.get_input_embeddings()(ids.cuda()).detach().requires_grad_()
logits = model(inputs_embeds=emb, use_cache=False).logits
mx, mi = logits[:, -1, :].max(-1)
mx.sum().backward()
r_hf = (emb * emb.grad).float().sum(-1).detach().cpu()
g_hf = emb.grad.float().cpu()
del model, logits
torch.cuda.empty_cache()
dims = LlamaDims(**{k: cfg[k] for k in ("d", "I", "H", "Hkv", "D", "L", "V", "eps", "theta")})
eng = LlamaAttnLRPEngine.from_weights(dims, w, device="cuda", micro_batch=1)
r_e, aux = eng.attribute_device(ids.cuda(), return_aux=True)
print(f"variant[{variant}] dims {nums}: patched-HF vs engine rel {O.rel_l2(r_hf, r_e.cpu()):.3e}  g_emb {O.rel_l2(g_hf, aux['g_emb'].cpu().view_as(g_hf)):.3e}"
      f"  idx hf {mi.tolist()} engine {aux['idx'].tolist()}  max logit hf {float(mx):.4f} engine {float(aux['logits'].max()):.4f}  ({time.time()-t0:.0f}s)", flush=True)
