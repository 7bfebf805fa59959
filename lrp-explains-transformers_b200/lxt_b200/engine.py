"""Layer-by-layer AttnLRP executor for Llama-family decoders (the headline hot path).

One attribution = embed -> forward -> arg-max logit at the last position -> LRP backward in Gradient x Input
space -> (emb * g_emb).sum(-1)   (reference workload: examples/quantized_llama.py:35-47 with
lxt.efficient.monkey_patch(modeling_llama), lxt/efficient/models/llama.py:9-14).

The engine does not use autograd: it launches the hand-written sm_100a kernels of liblrp_b200.so in the order the
reference's autograd graph would execute its stock kernels, with the point-wise LRP rules fused into GEMM
epilogues:

  forward  per layer : rmsnorm -> [QKV GEMM] -> RoPE -> flash-attn fwd -> [O GEMM + residual]
                       -> rmsnorm -> [gate|up GEMM] -> act*up -> [down GEMM + residual]
  backward per layer : [down dgrad] -> (÷2, identity rule on SiLU, product rule) -> [gate|up dgrad + w*rstd +
                       residual] -> [O dgrad] -> flash AttnLRP bwd (dQ/4,dK/4,dV/2) -> RoPE^T
                       -> [QKV dgrad + w*rstd + residual]

Data layout in HBM (T = micro_batch * S tokens, all row-major):
  residual stream h, gradient stream g_h : fp32 [T, d]     (+ bf16 shadow of g_h = next GEMM's A operand)
  activation store per layer (kept for the backward, nothing is recomputed when it fits — B200 has 180 GB):
      qkv  bf16 [T, (H+2Hkv) D]  (post-RoPE q,k and v packed; attention reads strided views, no copies)
      o    bf16 [T, H D]         lse fp32 [B, H, S]
      gu   bf16 [T, 2 I]         (gate | up)
      rstd1, rstd2 fp32 [T]
  weights bf16 in nn.Linear layout [out, in]: wqkv = [wq; wk; wv], wgu = [wg; wu]; the backward consumes the
  same storage through MN-major UMMA descriptors (no transposed copies).
If the store for all layers does not fit, `store="sqrt"` keeps only the layer-boundary residual stream at
segment starts (sqrt(L) segments) and recomputes the forward inside a segment before its backward.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

from . import ops


@dataclass
class LlamaDims:
    """Decoder description.  The defaults are the Llama layout; the optional fields describe Gemma-3 / Qwen3 variants
    (reference maps lxt/efficient/models/{llama,gemma3,qwen3}.py share one rule set)."""
    d: int
    I: int
    H: int
    Hkv: int
    D: int
    L: int
    V: int
    eps: float = 1e-5
    theta: float = 10000.0
    norm_offset: float = 0.0          # RMSNorm weight offset: 0 = `w * x_hat` (Llama), 1 = `(1 + w) * x_hat` (Gemma)
    act: str = "silu"                 # gated-MLP activation: "silu" | "gelu_tanh"
    qk_norm: bool = False             # per-head RMSNorm on q and k before RoPE (weights 'qn', 'kn')
    post_norms: bool = False          # Gemma layer layout: h += post_norm(branch(pre_norm(h)))
    windows: Optional[Sequence[int]] = None   # per-layer sliding window (0 = global attention)
    thetas: Optional[Sequence[float]] = None  # per-layer RoPE base
    attn_scale: Optional[float] = None        # soft-max scale (default D^-0.5; Gemma: query_pre_attn_scalar^-0.5)
    emb_scale: float = 1.0                    # embedding multiplier (Gemma: sqrt(d))

    @property
    def qkv_width(self) -> int:
        return (self.H + 2 * self.Hkv) * self.D

    def window(self, l: int) -> int:
        return int(self.windows[l]) if self.windows else 0

    def layer_theta(self, l: int) -> float:
        return float(self.thetas[l]) if self.thetas else self.theta

    @property
    def act_code(self) -> int:
        return {"silu": ops.ACT_SILU, "gelu_tanh": ops.ACT_GELU_TANH}[self.act]


DecoderDims = LlamaDims


def gemma3_dims(d, I, H, Hkv, D, L, V, sliding_window=1024, pattern=6, eps=1e-6) -> "LlamaDims":
    """Gemma-3 text layout: `pattern-1` sliding-window layers then one global layer, local / global RoPE bases 1e4 / 1e6."""
    glob = [(l + 1) % pattern == 0 for l in range(L)]
    return LlamaDims(d=d, I=I, H=H, Hkv=Hkv, D=D, L=L, V=V, eps=eps, norm_offset=1.0, act="gelu_tanh", qk_norm=True, post_norms=True,
                     windows=[0 if g else sliding_window for g in glob], thetas=[1000000.0 if g else 10000.0 for g in glob],
                     attn_scale=float(D) ** -0.5, emb_scale=float(d) ** 0.5)


LLAMA3_8B = LlamaDims(d=4096, I=14336, H=32, Hkv=8, D=128, L=32, V=128256, eps=1e-5, theta=500000.0)
TINYLLAMA_1B = LlamaDims(d=2048, I=5632, H=32, Hkv=4, D=64, L=22, V=32000, eps=1e-5, theta=10000.0)
GEMMA3_4B = gemma3_dims(d=2560, I=10240, H=8, Hkv=4, D=256, L=34, V=262208)


class _LayerStore:
    __slots__ = ("qkv", "o", "lse", "gu", "rstd1", "rstd2", "rstd_qk", "rstd_pa", "rstd_pf")


class LlamaAttnLRPEngine:
    """B200-native AttnLRP engine.  Construct with `from_weights`, `from_hf` or `random_init`."""

    def __init__(self, dims: LlamaDims, device: torch.device, weights: Dict, micro_batch: int = 8, store: str = "all",
                 rule: str = "attnlrp", cuda_graph: bool = False, precision: str = "bf16", rope: Optional[Sequence] = None,
                 quant: Optional[str] = None):
        if device.type != "cuda":
            raise RuntimeError("LlamaAttnLRPEngine runs on a CUDA (B200) device only; there is no CPU path")
        ops._capi.require_device()
        self.dims, self.device, self.micro_batch, self.store_policy = dims, device, micro_batch, store
        if rule not in ("attnlrp", "cp"):
            raise ValueError("rule must be 'attnlrp' (lxt attnLRP map) or 'cp' (lxt cp_LRP map)")
        self.cp = rule == "cp"  # CP-LRP: q,k and the MLP gate detached (lxt/efficient/models/llama.py:16-21)
        # precision="high" is the VALIDATION mode: every activation / gradient tensor is stored in fp32, every GEMM runs as a
        # two-term bf16 split (x = hi + lo) on the same tcgen05 kernel, the point-wise kernels are the fp32 instantiations of
        # the same templates and attention runs on the fp32 CUDA-core kernel (attn_f32.cu).  It exists to show that the bf16
        # mode's distance to the reference's fp32 run is storage rounding, not a defect; it is ~5-10x slower.
        if precision not in ("bf16", "high"):
            raise ValueError("precision must be 'bf16' (production) or 'high' (fp32-activation validation mode)")
        self.hp = precision == "high"
        self.adt = torch.float32 if self.hp else torch.bfloat16
        import os
        # fusing the gated-MLP backward rules into the down-dgrad epilogue was measured SLOWER on B200 (16.2 vs 16.9
        # attributions/s: the exp/div epilogue outlasts the K=4096 mainloop), so it is opt-in
        # gate|up layout: the weight rows (hence the columns of gu / g_gu) are interleaved in blocks of 32 (32 gate, 32 up) when I
        # allows it, so that one 64-column slab of the gate|up GEMM accumulator holds matching gate and up values and the
        # epilogue can emit a = act(gate) * up itself (no separate point-wise pass over [T, 2I]).  LRP_FUSE_ACT=0 keeps the
        # separate kernel (A/B); the validation mode always uses the separate fp32 kernel on the same layout.
        self.gu_layout = 1 if dims.I % 32 == 0 else 0
        self.fuse_act = self.gu_layout == 1 and not self.hp and os.environ.get("LRP_FUSE_ACT", "1") == "1"
        self.fuse_delta = os.environ.get("LRP_FUSE_DELTA", "1") == "1" and dims.D in (32, 64, 128, 256) and (dims.H * dims.D) % 256 == 0
        # the backward counterpart: (g_a/2, identity rule on the activation, product rule) in the down-dgrad epilogue.  Round 1's
        # version (IEEE division, activation chosen per element) was slower than the separate kernel; the lean epilogue
        # (approximate ex2 / rcp, activation as a template parameter) is not: LRP_FUSE_GATED=0 restores the separate kernel.
        self.fuse_gated = os.environ.get("LRP_FUSE_GATED", "1") == "1" and not self.hp and dims.I % 32 == 0
        bf = lambda t: t.to(device=device, dtype=torch.bfloat16).contiguous()
        f32 = lambda t: t.to(device=device, dtype=torch.float32).contiguous()
        self.emb = bf(weights["emb"])
        self.lm_head = bf(weights["lm_head"])
        self.norm_w = bf(weights["norm"])
        self.layers: List[Dict[str, torch.Tensor]] = []
        for lw in weights["layers"]:
            wqkv = torch.cat([lw["wq"], lw["wk"], lw["wv"]], dim=0)
            wgu = self._pack_gu(lw["wg"], lw["wu"])
            ln2 = lw["ln_pre_ff"] if dims.post_norms else lw["ln2"]
            layer = dict(wqkv=bf(wqkv), wo=bf(lw["wo"]), wgu=bf(wgu), wd=bf(lw["wd"]), ln1=bf(lw["ln1"]), ln2=bf(ln2),
                         ln1_f=f32(bf(lw["ln1"])) + dims.norm_offset, ln2_f=f32(bf(ln2)) + dims.norm_offset)
            if "bq" in lw:   # Qwen2-style projection biases: one fp32 vector for the packed QKV GEMM epilogue
                layer["bqkv"] = f32(torch.cat([lw["bq"], lw["bk"], lw["bv"]], dim=0))
            if dims.qk_norm:
                layer.update(qn=bf(lw["qn"]), kn=bf(lw["kn"]))
            if dims.post_norms:
                layer.update(ln_post_attn=bf(lw["ln_post_attn"]), ln_post_ff=bf(lw["ln_post_ff"]))
            self.layers.append(layer)
        # quant="nf4": the four projection matrices of every layer rest in HBM as 4-bit NormalFloat codes (+ fp32 absmax per 64
        # values) — 8B parameters in 4.5 GB instead of 16 GB, as the reference's examples do through bitsandbytes
        # (examples/quantized_llama.py:13-19) — and are expanded to bf16 into ONE per-layer scratch right before the layer's GEMMs
        # (forward and backward), so the tcgen05 kernels and the LRP rules are untouched.  Embedding / lm_head stay bf16.
        if quant not in (None, "nf4"):
            raise ValueError("quant must be None or 'nf4'")
        self.quant = quant
        self._wq_names = ("wqkv", "wo", "wgu", "wd")
        if quant == "nf4":
            self._wscratch = {n: torch.empty_like(self.layers[0][n]) for n in self._wq_names}
            self._wscratch_layer = None
            for layer in self.layers:
                for n in self._wq_names:
                    layer[n] = ops.quant_nf4(layer[n]) + (tuple(layer[n].shape),)
            torch.cuda.empty_cache()
        self._ws_key = None
        self._ws = None
        self._rope_cache = {}
        # RoPE per layer as (inv_freq fp32 [D/2], attention_scaling): the default is theta^(-2i/D) (transformers
        # modeling_llama.py LlamaRotaryEmbedding.compute_default_rope_parameters); `from_hf` passes the model's own tables so
        # that scaled RoPE variants (llama3, linear, yarn: static inv_freq + a cos/sin factor) are reproduced exactly.
        if rope is None:
            rope = []
            for l in range(dims.L):
                th = dims.layer_theta(l)
                rope.append((1.0 / (th ** (torch.arange(0, dims.D, 2, dtype=torch.int64).float() / dims.D)), 1.0))
        if len(rope) != dims.L:
            raise ValueError("rope: one (inv_freq, attention_scaling) pair per layer expected")
        self._rope_spec, self._rope_id, uniq = [], [], {}
        for inv_freq, att in rope:
            inv_freq = torch.as_tensor(inv_freq, dtype=torch.float32).detach().cpu().contiguous()
            if inv_freq.numel() != dims.D // 2:
                raise ValueError(f"rope: inv_freq must have head_dim/2 = {dims.D // 2} entries")
            key = (inv_freq.numpy().tobytes(), float(att))
            if key not in uniq:
                uniq[key] = len(self._rope_spec)
                self._rope_spec.append((inv_freq, float(att)))
            self._rope_id.append(uniq[key])
        # cuda_graph=True captures the whole attribution (≈20 launches per layer) of a given [B,S] once and replays it:
        # small / latency-bound shapes (e.g. TinyLlama, B=1, S=512) stop being bound by host launch overhead.
        self.cuda_graph = cuda_graph
        self._graphs = {}
        self._graph_ws = {}   # [B,S] -> workspace owned by the captured graph of that shape

    def _pack_gu(self, wg, wu):
        if self.gu_layout == 0:
            return torch.cat([wg, wu], dim=0)
        I, d = wg.shape
        return torch.stack([wg.reshape(I // 32, 32, d), wu.reshape(I // 32, 32, d)], dim=1).reshape(2 * I, d)

    def export_weights(self) -> Dict:
        """the weights in the `from_weights` format (packed projections split / de-interleaved again)"""
        m = self.dims
        out = dict(emb=self.emb, norm=self.norm_w, lm_head=self.lm_head, layers=[])
        for l, lw in enumerate(self.layers):
            if self.quant is not None:   # de-quantised copies
                lw = dict(lw)
                for n in self._wq_names:
                    packed, absmax, shape = lw[n]
                    lw[n] = ops.dequant_nf4(packed, absmax, torch.empty(shape, dtype=torch.bfloat16, device=self.device))
            wq, wk, wv = lw["wqkv"].split([m.H * m.D, m.Hkv * m.D, m.Hkv * m.D], 0)
            if self.gu_layout == 0:
                wg, wu = lw["wgu"].split([m.I, m.I], 0)
            else:
                blk = lw["wgu"].view(m.I // 32, 2, 32, m.d)
                wg, wu = blk[:, 0].reshape(m.I, m.d), blk[:, 1].reshape(m.I, m.d)
            e = dict(wq=wq, wk=wk, wv=wv, wo=lw["wo"], wg=wg, wu=wu, wd=lw["wd"], ln1=lw["ln1"])
            if m.post_norms:
                e.update(ln_pre_ff=lw["ln2"], ln_post_attn=lw["ln_post_attn"], ln_post_ff=lw["ln_post_ff"])
            else:
                e["ln2"] = lw["ln2"]
            if m.qk_norm:
                e.update(qn=lw["qn"], kn=lw["kn"])
            if "bqkv" in lw:
                bq, bk, bv = lw["bqkv"].split([m.H * m.D, m.Hkv * m.D, m.Hkv * m.D], 0)
                e.update(bq=bq, bk=bk, bv=bv)
            out["layers"].append(e)
        return out

    # ------------------------------------------------------------------ constructors
    @classmethod
    def from_weights(cls, dims: LlamaDims, weights: Dict, device="cuda", **kw):
        return cls(dims, torch.device(device), weights, **kw)

    @classmethod
    def from_hf(cls, model, device="cuda", **kw):
        """Build from a HuggingFace causal LM of the Llama, Qwen3 or Gemma-3 (text) family; weights are copied to bf16."""
        c = model.config
        mt = getattr(c, "model_type", "llama")
        D = getattr(c, "head_dim", None) or c.hidden_size // c.num_attention_heads
        rp = getattr(c, "rope_parameters", None) or {}
        theta = rp.get("rope_theta") or getattr(c, "rope_theta", None) or 10000.0
        extra = {}
        if mt.startswith("gemma3"):
            lt = list(c.layer_types)
            th = lambda kind: float((rp.get(kind) or {}).get("rope_theta", 10000.0 if kind == "sliding_attention" else 1000000.0))
            extra = dict(norm_offset=1.0, act="gelu_tanh", qk_norm=True, post_norms=True,
                         windows=[int(c.sliding_window) if t == "sliding_attention" else 0 for t in lt],
                         thetas=[th(t) for t in lt], attn_scale=float(c.query_pre_attn_scalar) ** -0.5,
                         emb_scale=float(c.hidden_size) ** 0.5)
        elif mt == "qwen3":
            extra = dict(qk_norm=True)
        elif mt not in ("llama", "qwen2"):
            raise ValueError(f"LlamaAttnLRPEngine.from_hf: unsupported model_type {mt!r} (use the drop-in monkey_patch API)")
        dims = LlamaDims(d=c.hidden_size, I=c.intermediate_size, H=c.num_attention_heads, Hkv=c.num_key_value_heads, D=D,
                         L=c.num_hidden_layers, V=c.vocab_size, eps=c.rms_norm_eps, theta=float(theta), **extra)
        kw.setdefault("rope", cls._rope_from_hf(model, dims))
        sd = model.state_dict()
        has_bias = any(k.endswith("self_attn.q_proj.bias") for k in sd)
        if any(k.endswith(("o_proj.bias", "gate_proj.bias", "up_proj.bias", "down_proj.bias")) for k in sd):
            raise ValueError("only q/k/v projection biases are supported by the fused engine (use the drop-in monkey_patch API)")
        w = dict(emb=sd["model.embed_tokens.weight"], norm=sd["model.norm.weight"],
                 lm_head=sd.get("lm_head.weight", sd["model.embed_tokens.weight"]), layers=[])
        for i in range(dims.L):
            p = f"model.layers.{i}."
            lw = dict(wq=sd[p + "self_attn.q_proj.weight"], wk=sd[p + "self_attn.k_proj.weight"],
                      wv=sd[p + "self_attn.v_proj.weight"], wo=sd[p + "self_attn.o_proj.weight"],
                      wg=sd[p + "mlp.gate_proj.weight"], wu=sd[p + "mlp.up_proj.weight"],
                      wd=sd[p + "mlp.down_proj.weight"], ln1=sd[p + "input_layernorm.weight"])
            if has_bias:
                lw.update(bq=sd[p + "self_attn.q_proj.bias"], bk=sd[p + "self_attn.k_proj.bias"], bv=sd[p + "self_attn.v_proj.bias"])
            if dims.qk_norm:
                lw.update(qn=sd[p + "self_attn.q_norm.weight"], kn=sd[p + "self_attn.k_norm.weight"])
            if dims.post_norms:
                lw.update(ln_post_attn=sd[p + "post_attention_layernorm.weight"], ln_pre_ff=sd[p + "pre_feedforward_layernorm.weight"],
                          ln_post_ff=sd[p + "post_feedforward_layernorm.weight"])
            else:
                lw.update(ln2=sd[p + "post_attention_layernorm.weight"])
            w["layers"].append(lw)
        return cls(dims, torch.device(device), w, **kw)

    @staticmethod
    def _rope_from_hf(model, dims: LlamaDims):
        """Per-layer (inv_freq, attention_scaling) exactly as the HF model's rotary module holds them (rope_type default, linear,
        llama3, yarn: `ROPE_INIT_FUNCTIONS`, transformers modeling_rope_utils.py).  Types whose tables depend on the sequence
        length at run time (dynamic NTK, longrope) are refused: the engine would silently compute something else."""
        c = model.config
        rot = getattr(getattr(model, "model", model), "rotary_emb", None)
        if rot is None:
            raise ValueError("from_hf: the model has no `model.rotary_emb` module to take the RoPE tables from")
        kinds = getattr(rot, "rope_type", "default")
        layer_types = list(getattr(c, "layer_types", None) or [None] * dims.L)
        out = []
        for l in range(dims.L):
            lt = layer_types[l] if isinstance(kinds, dict) else None
            kind = kinds[lt] if isinstance(kinds, dict) else kinds
            if kind in ("dynamic", "longrope"):
                raise ValueError(f"from_hf: rope_type {kind!r} changes its tables with the sequence length; not supported by the fused "
                                 "engine (use the drop-in monkey_patch API)")
            inv = getattr(rot, f"{lt}_inv_freq" if lt is not None else "inv_freq")
            att = getattr(rot, f"{lt}_attention_scaling" if lt is not None else "attention_scaling", 1.0)
            out.append((inv.detach().float().cpu(), float(att)))
        return out

    @classmethod
    def random_init(cls, dims: LlamaDims, device="cuda", seed: int = 0, std: float = 0.02, **kw):
        """HF-style random init generated directly on the device (synthetic benchmark weights)."""
        dev = torch.device(device)
        g = torch.Generator(device=dev).manual_seed(seed)
        rn = lambda *s: (torch.randn(*s, generator=g, device=dev, dtype=torch.float32) * std).to(torch.bfloat16)
        ones = lambda n: torch.ones(n, device=dev, dtype=torch.bfloat16)
        nw = (lambda n: torch.zeros(n, device=dev, dtype=torch.bfloat16)) if dims.norm_offset else ones  # (1 + 0) = 1
        layers = []
        for _ in range(dims.L):
            lw = dict(wq=rn(dims.H * dims.D, dims.d), wk=rn(dims.Hkv * dims.D, dims.d), wv=rn(dims.Hkv * dims.D, dims.d),
                      wo=rn(dims.d, dims.H * dims.D), wg=rn(dims.I, dims.d), wu=rn(dims.I, dims.d), wd=rn(dims.d, dims.I), ln1=nw(dims.d))
            if dims.qk_norm:
                lw.update(qn=nw(dims.D), kn=nw(dims.D))
            if dims.post_norms:
                lw.update(ln_post_attn=nw(dims.d), ln_pre_ff=nw(dims.d), ln_post_ff=nw(dims.d))
            else:
                lw.update(ln2=nw(dims.d))
            layers.append(lw)
        emb = rn(dims.V, dims.d)
        w = dict(emb=emb, norm=nw(dims.d), lm_head=emb if dims.post_norms else rn(dims.V, dims.d), layers=layers)
        return cls(dims, dev, w, **kw)

    # ------------------------------------------------------------------ workspace
    def store_bytes_per_token_layer(self) -> int:
        m = self.dims
        return 2 * (m.qkv_width + m.H * m.D + 2 * m.I) + 4 * m.H + 8

    def _workspace(self, B: int, S: int):
        key = (B, S)
        if key in self._graph_ws:       # a captured CUDA graph replays on these buffers: they are never freed or re-keyed
            return self._graph_ws[key]
        if self._ws_key == key:
            return self._ws
        self._ws = None  # release the previous (un-graphed) workspace before allocating the new one
        m, dev = self.dims, self.device
        T = B * S
        e = lambda *s, dt=self.adt: torch.empty(*s, dtype=dt, device=dev)
        ws = {}
        n_store = m.L if self.store_policy == "all" else self._segment_len()
        stores = []
        for _ in range(n_store):
            st = _LayerStore()
            st.qkv, st.o, st.gu = e(T, m.qkv_width), e(T, m.H * m.D), e(T, 2 * m.I)
            st.lse = e(B, m.H, S, dt=torch.float32)
            st.rstd1, st.rstd2 = e(T, dt=torch.float32), e(T, dt=torch.float32)
            st.rstd_qk = e(T, m.H + m.Hkv, dt=torch.float32) if m.qk_norm else None
            st.rstd_pa = e(T, dt=torch.float32) if m.post_norms else None
            st.rstd_pf = e(T, dt=torch.float32) if m.post_norms else None
            stores.append(st)
        ws["stores"] = stores
        if self.store_policy != "all":
            ws["h_ckpt"] = [e(T, m.d, dt=torch.float32) for _ in range(math.ceil(m.L / self._segment_len()))]
        ws["h"] = e(T, m.d, dt=torch.float32)
        ws["g_h"] = e(T, m.d, dt=torch.float32)
        ws["g_hb"] = ws["g_h"] if self.hp else e(T, m.d)   # GEMM A operand: bf16 shadow (fp32 stream itself in validation mode)
        ws["xn"] = e(T, m.d)
        if m.post_norms:
            ws["y"] = e(T, m.d)      # branch output before its post-norm (forward) / gradient after it (backward)
        ws["a"] = e(T, m.I)          # act(gate)*up in forward, g_a in backward
        ws["g_gu"] = e(T, 2 * m.I)
        ws["g_o"] = e(T, m.H * m.D)
        ws["g_qkv"] = e(T, m.qkv_width)
        # zero once: every attention backward leaves it zero again (LRP_ATTN_ACC_ZERO)
        ws["dq_acc"] = None if self.hp else torch.zeros(ops.attn_bwd_workspace_floats(B, S, m.H, m.D), dtype=torch.float32, device=dev)
        ws["delta"] = e(B, m.H, S, dt=torch.float32)
        ws["logits"] = e(B, m.V, dt=torch.float32)
        ws["h_last"] = e(B, m.d, dt=torch.float32)
        ws["xn_last"] = e(B, m.d)
        ws["rstd_last"] = e(B, dt=torch.float32)
        ws["idx"] = e(B, dt=torch.int32)
        ws["last_rows"] = (torch.arange(B, device=dev, dtype=torch.int64) + 1) * S - 1
        self._ws_key, self._ws = key, ws
        return ws

    def _segment_len(self) -> int:
        return max(1, int(math.ceil(math.sqrt(self.dims.L))))

    def _rope(self, S: int, l: int = 0):
        key = (S, self._rope_id[l])
        if key not in self._rope_cache:
            inv_freq, att = self._rope_spec[self._rope_id[l]]
            fr = torch.outer(torch.arange(S, dtype=torch.float32), inv_freq)
            self._rope_cache[key] = ((fr.cos() * att).to(self.device).contiguous(), (fr.sin() * att).to(self.device).contiguous())
        return self._rope_cache[key]

    # ------------------------------------------------------------------ one layer
    def _materialize(self, lw, l: int):
        """NF4 storage: expand layer l's four projection matrices into the shared bf16 scratch (4 launches; skipped when the scratch
        already holds this layer, e.g. backward right after the forward of the same layer in the sqrt schedule)"""
        if self.quant is None:
            return lw
        if self._wscratch_layer != l:
            for n in self._wq_names:
                packed, absmax, _ = lw[n]
                ops.dequant_nf4(packed, absmax, self._wscratch[n])
            self._wscratch_layer = l
        out = dict(lw)
        out.update(self._wscratch)
        return out

    def weight_bytes(self) -> int:
        """bytes of HBM held by the decoder-layer weights (packed codes + absmax for NF4, incl. the one-layer bf16 scratch)"""
        tot = 0
        for lw in self.layers:
            for v in lw.values():
                for t in (v if isinstance(v, tuple) else (v,)):
                    if isinstance(t, torch.Tensor):
                        tot += t.numel() * t.element_size()
        if self.quant is not None:
            tot += sum(t.numel() * t.element_size() for t in self._wscratch.values())
        return tot

    def _layer_fwd(self, lw, st: _LayerStore, h, ws, B, S, l: int = 0, h_copy=None):
        """h_copy: optional bf16 [T,d] buffer that receives a copy of this layer's OUTPUT residual stream, written by the last
        residual epilogue of the layer (the latent-relevance trace needs the layer outputs in the backward sweep)"""
        lw = self._materialize(lw, l)
        m = self.dims
        T = B * S
        cos, sin = self._rope(S, l)
        scale = m.attn_scale or 1.0 / math.sqrt(m.D)
        off, win = m.norm_offset, m.window(l)
        lib, C = ops._capi.lib(), ops._capi
        f = int(self.hp)   # activation dtype flag of the typed entry points: 0 = bf16 (production), 1 = fp32 (validation mode)
        C.check(lib.lrp_rmsnorm_fwd_t(h.data_ptr(), 1, lw["ln1"].data_ptr(), off, m.eps, ws["xn"].data_ptr(), f, st.rstd1.data_ptr(),
                                      T, m.d, ops._stream()), "rmsnorm_fwd")
        ops.linear_fwd(ws["xn"], lw["wqkv"], st.qkv, bias=lw.get("bqkv"))
        if m.qk_norm:
            C.check(lib.lrp_headnorm_inplace_t(st.qkv.data_ptr(), f, m.qkv_width, m.H, m.Hkv, m.D, lw["qn"].data_ptr(), lw["kn"].data_ptr(),
                                               off, m.eps, st.rstd_qk.data_ptr(), T, 0, ops._stream()), "headnorm_fwd")
        ops.rope_inplace(st.qkv, m.H + m.Hkv, m.D, cos, sin, S)
        q, k, v = self._qkv_views(st.qkv, B, S)
        if self.hp:
            C.check(lib.lrp_attn_fwd_f32(q.data_ptr(), k.data_ptr(), v.data_ptr(), m.qkv_width, m.qkv_width, m.qkv_width,
                                         st.o.data_ptr(), st.lse.data_ptr(), None, B, S, m.H, m.Hkv, m.D, scale, 1, win, ops._stream()),
                    "attn_fwd_f32")
        else:
            C.check(lib.lrp_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), m.qkv_width, m.qkv_width, m.qkv_width,
                                     st.o.data_ptr(), st.lse.data_ptr(), B, S, m.H, m.Hkv, m.D, scale, 1, win, ops._stream()), "attn_fwd")
        if m.post_norms:
            ops.linear_fwd(st.o, lw["wo"], ws["y"])
            C.check(lib.lrp_rmsnorm_fwd_residual_t(ws["y"].data_ptr(), f, lw["ln_post_attn"].data_ptr(), off, m.eps, h.data_ptr(),
                                                   st.rstd_pa.data_ptr(), T, m.d, ops._stream()), "rmsnorm_fwd_residual")
        else:
            ops.linear_fwd(st.o, lw["wo"], h, resid=h)
        C.check(lib.lrp_rmsnorm_fwd_t(h.data_ptr(), 1, lw["ln2"].data_ptr(), off, m.eps, ws["xn"].data_ptr(), f, st.rstd2.data_ptr(),
                                      T, m.d, ops._stream()), "rmsnorm_fwd")
        if self.fuse_act:   # a = act(gate) * up leaves the gate|up GEMM's epilogue together with gu
            ops.linear_fwd(ws["xn"], lw["wgu"], st.gu, act_out=ws["a"], act=m.act_code)
        else:
            ops.linear_fwd(ws["xn"], lw["wgu"], st.gu)
            C.check(lib.lrp_gated_act_fwd_t(st.gu.data_ptr(), ws["a"].data_ptr(), f, self.gu_layout, T, m.I, m.act_code, ops._stream()),
                    "gated_act_fwd")
        if m.post_norms:
            ops.linear_fwd(ws["a"], lw["wd"], ws["y"])
            C.check(lib.lrp_rmsnorm_fwd_residual_t(ws["y"].data_ptr(), f, lw["ln_post_ff"].data_ptr(), off, m.eps, h.data_ptr(),
                                                   st.rstd_pf.data_ptr(), T, m.d, ops._stream()), "rmsnorm_fwd_residual")
            if h_copy is not None:
                h_copy.copy_(h)
        elif h_copy is not None and not self.hp:
            ops.linear_fwd(ws["a"], lw["wd"], h, resid=h, shadow=h_copy)   # bf16 copy of the layer output from the GEMM epilogue
        else:
            ops.linear_fwd(ws["a"], lw["wd"], h, resid=h)
            if h_copy is not None:
                h_copy.copy_(h)

    def _layer_bwd(self, lw, st: _LayerStore, ws, B, S, l: int = 0):
        lw = self._materialize(lw, l)
        m = self.dims
        T = B * S
        cos, sin = self._rope(S, l)
        scale = m.attn_scale or 1.0 / math.sqrt(m.D)
        off, win = m.norm_offset, m.window(l)
        lib, C = ops._capi.lib(), ops._capi
        g_h, g_hb = ws["g_h"], ws["g_hb"]
        f = int(self.hp)
        shadow = None if self.hp else g_hb
        # ---- gated MLP
        src = g_hb
        if m.post_norms:   # identity rule through the post-feed-forward norm: g * (off + w) * rstd of the branch output
            src = ops.rmsnorm_bwd(g_hb, lw["ln_post_ff"], st.rstd_pf, w_offset=off, out=ws["y"])
        if self.fuse_gated:
            # down dgrad with (÷2, identity rule on SiLU, product rule) fused into its epilogue: g_a never touches HBM
            ops.linear_dgrad_gated_bwd(src, lw["wd"], st.gu, ws["g_gu"], m.act_code, self.cp, layout=self.gu_layout)
        else:
            ops.linear_dgrad(src, lw["wd"], ws["a"])                               # g_a [T, I]
            C.check(lib.lrp_gated_act_bwd_t(ws["a"].data_ptr(), st.gu.data_ptr(), ws["g_gu"].data_ptr(), f, self.gu_layout, T, m.I,
                                            m.act_code, int(self.cp), ops._stream()), "gated_act_bwd")
        ops.linear_dgrad(ws["g_gu"], lw["wgu"], g_h, resid=g_h, rowscale=st.rstd2, colscale=lw["ln2_f"], shadow=shadow)
        # ---- attention
        src = g_hb
        if m.post_norms:
            src = ops.rmsnorm_bwd(g_hb, lw["ln_post_attn"], st.rstd_pa, w_offset=off, out=ws["y"])
        # g_o [T, H D]; its epilogue also emits delta = sum_d o * g_o per (b, h, s) for the attention backward (no separate pass)
        fuse_delta = not self.hp and self.fuse_delta
        if fuse_delta:
            ops.linear_dgrad(src, lw["wo"], ws["g_o"], delta=(st.o, ws["delta"], m.D, S))
        else:
            ops.linear_dgrad(src, lw["wo"], ws["g_o"])
        q, k, v = self._qkv_views(st.qkv, B, S)
        dq, dk, dv = self._qkv_views(ws["g_qkv"], B, S)
        divs = (0.0, 0.0, 1.0) if self.cp else (4.0, 4.0, 2.0)
        if self.hp:
            C.check(lib.lrp_attn_bwd_f32(q.data_ptr(), k.data_ptr(), v.data_ptr(), m.qkv_width, m.qkv_width, m.qkv_width,
                                         st.o.data_ptr(), ws["g_o"].data_ptr(), st.lse.data_ptr(), dq.data_ptr(), dk.data_ptr(),
                                         dv.data_ptr(), m.qkv_width, m.qkv_width, m.qkv_width, ws["delta"].data_ptr(), None, B, S, m.H,
                                         m.Hkv, m.D, scale, 1, win, *divs, ops._stream()), "attn_bwd_f32")
        else:
            flags = (1 if fuse_delta else 0) | 2      # LRP_ATTN_DELTA_READY | LRP_ATTN_ACC_ZERO
            C.check(lib.lrp_attn_bwd_varlen(q.data_ptr(), k.data_ptr(), v.data_ptr(), m.qkv_width, m.qkv_width, m.qkv_width,
                                            st.o.data_ptr(), ws["g_o"].data_ptr(), st.lse.data_ptr(), dq.data_ptr(), dk.data_ptr(),
                                            dv.data_ptr(), m.qkv_width, m.qkv_width, m.qkv_width, ws["dq_acc"].data_ptr(),
                                            ws["delta"].data_ptr(), None, flags, B, S, m.H, m.Hkv, m.D, scale, 1, win, *divs,
                                            ops._stream()), "attn_bwd")
        ops.rope_inplace(ws["g_qkv"], m.H + m.Hkv, m.D, cos, sin, S, inverse=True)
        if m.qk_norm:
            C.check(lib.lrp_headnorm_inplace_t(ws["g_qkv"].data_ptr(), f, m.qkv_width, m.H, m.Hkv, m.D, lw["qn"].data_ptr(),
                                               lw["kn"].data_ptr(), off, m.eps, st.rstd_qk.data_ptr(), T, 1, ops._stream()), "headnorm_bwd")
        ops.linear_dgrad(ws["g_qkv"], lw["wqkv"], g_h, resid=g_h, rowscale=st.rstd1, colscale=lw["ln1_f"], shadow=shadow)

    def _qkv_views(self, buf, B, S):
        m = self.dims
        x = buf.view(B, S, m.qkv_width)
        q = x[:, :, : m.H * m.D]
        k = x[:, :, m.H * m.D: (m.H + m.Hkv) * m.D]
        v = x[:, :, (m.H + m.Hkv) * m.D:]
        return q, k, v

    # ------------------------------------------------------------------ one micro-batch, device resident
    @torch.no_grad()
    def attribute_device(self, ids: torch.Tensor, return_aux: bool = False, trace: bool = False):
        """ids int64 [B,S] on the device -> relevance fp32 [B,S] on the device.
        trace=True additionally returns the latent relevance of every decoder layer's output, `[L,B,S]`
        (= `output * output.grad` summed over features, what docs/source/latent-feature-attribution-efficient.rst
        :49-90 obtains with forward hooks + retain_grad), emitted here as one reduction kernel per layer."""
        if trace and self.store_policy != "all":
            raise ValueError("trace=True needs store='all'")
        m = self.dims
        B, S = ids.shape
        T = B * S
        ws = self._workspace(B, S)
        lib, C = ops._capi.lib(), ops._capi
        h, g_h, g_hb = ws["h"], ws["g_h"], ws["g_hb"]
        flat = ids.reshape(-1).contiguous()
        C.check(lib.lrp_embed_gather(flat.data_ptr(), self.emb.data_ptr(), m.emb_scale, h.data_ptr(), T, m.d, ops._stream()), "embed")

        seg = self._segment_len()
        if self.store_policy == "all":
            # latent trace: the layer outputs are kept as the bf16 shadow the down-projection's residual epilogue writes anyway
            # (no extra pass over the residual stream, half the bytes of an fp32 clone; fp32 in validation mode)
            h_outs = [torch.empty((T, m.d), dtype=torch.float32 if self.hp else torch.bfloat16, device=self.device)
                      for _ in range(m.L)] if trace else None
            for l, lw in enumerate(self.layers):
                self._layer_fwd(lw, ws["stores"][l], h, ws, B, S, l, h_copy=h_outs[l] if trace else None)
        else:
            for l, lw in enumerate(self.layers):
                if l % seg == 0:
                    ws["h_ckpt"][l // seg].copy_(h)
                self._layer_fwd(lw, ws["stores"][l % seg], h, ws, B, S, l)

        # ---- head: only the last position is read (examples/quantized_llama.py:40)
        C.check(lib.lrp_gather_rows_f32(h.data_ptr(), ws["last_rows"].data_ptr(), ws["h_last"].data_ptr(), B, m.d, ops._stream()),
                "gather_rows")
        xn_last, rstd_last = ops.rmsnorm_fwd(ws["h_last"], self.norm_w, m.eps, w_offset=m.norm_offset, out_dtype=self.adt,
                                             out=ws["xn_last"], rstd=ws["rstd_last"])
        ops.linear_fwd(xn_last, self.lm_head, ws["logits"])
        C.check(lib.lrp_argmax_rows(ws["logits"].data_ptr(), ws["idx"].data_ptr(), None, B, m.V, ops._stream()), "argmax_rows")
        idx = ws["idx"]
        # seed: d(max logit)/d(xn_last) = lm_head[idx], through the final norm with the identity rule, scattered to the last position
        # of every prompt (zero elsewhere) together with the bf16 shadow: one launch
        C.check(lib.lrp_seed_gradient(self.lm_head.data_ptr(), idx.data_ptr(), self.norm_w.data_ptr(), m.norm_offset,
                                      rstd_last.data_ptr(), S, g_h.data_ptr(), None if self.hp else g_hb.data_ptr(), T, m.d,
                                      ops._stream()), "seed_gradient")

        if self.store_policy == "all":
            layer_rel = [None] * m.L
            for l in range(m.L - 1, -1, -1):
                if trace:
                    layer_rel[l] = ops.gxi_reduce(h_outs[l], g_h).view(B, S)
                    h_outs[l] = None
                self._layer_bwd(self.layers[l], ws["stores"][l], ws, B, S, l)
        else:
            nseg = math.ceil(m.L / seg)
            for sgi in range(nseg - 1, -1, -1):
                lo, hi = sgi * seg, min(m.L, (sgi + 1) * seg)
                # recompute this segment's forward from its checkpoint (uniformly, also for the last segment: its
                # layer stores were overwritten unless L is a multiple of the segment length)
                h.copy_(ws["h_ckpt"][sgi])
                for l in range(lo, hi):
                    self._layer_fwd(self.layers[l], ws["stores"][l % seg], h, ws, B, S, l)
                for l in range(hi - 1, lo - 1, -1):
                    self._layer_bwd(self.layers[l], ws["stores"][l % seg], ws, B, S, l)

        # ---- Gradient x Input at the embedding
        C.check(lib.lrp_embed_gather(flat.data_ptr(), self.emb.data_ptr(), m.emb_scale, h.data_ptr(), T, m.d, ops._stream()), "embed")
        rel = ops.gxi_reduce(h, g_h).view(B, S)
        if return_aux or trace:
            aux = {"idx": idx, "logits": ws["logits"], "g_emb": g_h.view(B, S, m.d)}
            if trace:
                aux["layer_relevance"] = torch.stack(layer_rel)
            return rel, aux
        return rel

    @torch.no_grad()
    def attribute_device_graphed(self, ids: torch.Tensor) -> torch.Tensor:
        """Same as `attribute_device(ids)` but replayed from a CUDA graph captured on first use of this [B,S]
        (static input/output buffers; the returned tensor is overwritten by the next call of the same shape)."""
        key = tuple(ids.shape)
        ent = self._graphs.get(key)
        if ent is None:
            static_ids = ids.clone()
            # the graph bakes raw pointers into its workspace: give this shape a private one that `_workspace` never frees
            # (a later call of another shape, e.g. the trailing partial micro-batch of `attribute`, allocates its own)
            self._ws_key, self._ws = None, None          # drop the shared slot, allocate a fresh workspace ...
            ws_private = self._workspace(*key)
            self._ws_key, self._ws = None, None          # ... and move it out of the shared slot into the graph's ownership
            self._graph_ws[key] = ws_private
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):  # warm-up outside capture: lazy attribute setup, workspace allocation, rope tables
                    self.attribute_device(static_ids)
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                static_out = self.attribute_device(static_ids)
            ent = self._graphs[key] = (g, static_ids, static_out)
        g, static_ids, static_out = ent
        static_ids.copy_(ids, non_blocking=True)
        g.replay()
        return static_out

    # ------------------------------------------------------------------ public API: host in, host out
    @torch.no_grad()
    def attribute(self, input_ids: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """input_ids: int64 [N,S] on the HOST (pinned memory recommended) or the device.
        Returns token relevance fp32 [N,S] on the host.  Prompts are processed `micro_batch` at a time."""
        N, S = input_ids.shape
        if out is None:
            out = torch.empty((N, S), dtype=torch.float32, pin_memory=True)
        for i in range(0, N, self.micro_batch):
            ids = input_ids[i:i + self.micro_batch].to(self.device, non_blocking=True)
            full = ids.shape[0] == self.micro_batch
            rel = self.attribute_device_graphed(ids) if (self.cuda_graph and full) else self.attribute_device(ids)
            out[i:i + self.micro_batch].copy_(rel, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out
