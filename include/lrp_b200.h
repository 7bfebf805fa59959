/*
 * lrp_b200.h — C ABI of the B200-native AttnLRP hot path (liblrp_b200.so).
 *
 * The reference (rachtibat/LRP-eXplains-Transformers, `lxt` 2.1) has no FFI: its hot path is Python
 * autograd glue over stock PyTorch kernels.  This header is the boundary a maintainer would bind instead
 * (ctypes stub shown in INTEGRATION.md).  Every entry point cites the reference code it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers owned by the caller unless stated;
 *   - bf16 tensors are `uint16_t`-sized IEEE bfloat16, fp32 tensors are `float`;
 *   - row-major, innermost dimension contiguous, leading dimensions given in ELEMENTS;
 *   - `stream` is a `cudaStream_t` passed as `void*` (0 = legacy default stream); kernels are enqueued,
 *     never synchronised;
 *   - return value 0 on success, negative `LRP_ERR_*` otherwise; `lrp_last_error()` gives the
 *     thread-local message.  Nothing throws, nothing allocates device memory (workspaces are passed in);
 *   - re-entrant from any host thread (the autograd engine calls backward rules from its own thread).
 */
#ifndef LRP_B200_H_
#define LRP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LRP_OK 0
#define LRP_ERR_ARG (-1)      /* invalid argument / unsupported shape */
#define LRP_ERR_CUDA (-2)     /* CUDA runtime or driver error */
#define LRP_ERR_NO_DEVICE (-3) /* no sm_100 device visible */

#define LRP_ACT_SILU 0
#define LRP_ACT_GELU_TANH 1
#define LRP_ACT_GELU_ERF 2

/* library ABI version (major*1000 + minor).  Library plumbing: the reference has no FFI (its version is lxt/__init__.py / setup.py:5) */
int lrp_version(void);
/* thread-local description of the last error returned on this thread ("" if none); the reference reports failures as Python
 * warnings / exceptions (lxt/efficient/core.py:39-44), the Python binding turns a non-zero return code into LrpError */
const char* lrp_last_error(void);
/* number of GPU kernels (and memset nodes) this library has enqueued in this process so far (bench.py `gpu_launches`; no
 * reference counterpart: the reference launches stock torch kernels from autograd, lxt/efficient/rules.py:69-127) */
int64_t lrp_launch_count(void);
/* 0 if an sm_100 device is present and usable, LRP_ERR_NO_DEVICE otherwise (the reference runs wherever torch runs:
 * examples/quantized_llama.py:13-22 `device_map="cuda"`; this library has no CPU fallback by design) */
int lrp_check_device(void);

/* ------------------------------------------------------------------------------------------------
 * Fused GEMM epilogue:  out = resid + alpha * acc * rowscale[m] * colscale[n] + bias[n]
 * NULL pointers disable the corresponding term.  `out` is bf16 or fp32 (out_is_f32); `shadow_bf16`
 * optionally receives a bf16 copy of the same values (the next GEMM's A operand).
 * ---------------------------------------------------------------------------------------------- */
typedef struct lrp_epilogue {
  void* out;
  int32_t out_is_f32;
  void* shadow_bf16;
  const float* resid_f32; /* may alias `out` when out_is_f32 (in-place accumulate) */
  const float* rowscale;  /* [M] fp32 */
  const float* colscale;  /* [N] fp32 */
  const float* bias;      /* [N] fp32 */
  float alpha;
  int64_t ldc;            /* leading dimension of out / shadow / resid, in elements */
  /* Optional fused gated-MLP LRP backward (replaces `out`, which may then be NULL): the accumulator is g_a, the
   * gradient w.r.t. act(gate)*up, and the epilogue applies divide_gradient(2) + the identity rule on the activation +
   * the product rule (lxt/efficient/patches.py:145-157, rules.py:88-127) exactly like lrp_gated_act_bwd:
   *   gated_out[m, n] = g_gate,  gated_out[m, I+n] = g_up,  reading gate/up from gated_gu[m, n] / [m, I+n]. */
  const void* gated_gu;   /* bf16 [M, 2I] (gate | up), I = N of this GEMM */
  void* gated_out;        /* bf16 [M, 2I] */
  int32_t gated_act;      /* LRP_ACT_* */
  int32_t gated_cp;       /* 1 = CP-LRP variant */
  /* Optional fused gated-MLP FORWARD (lxt/efficient/patches.py:145-157): when `act_out` is set, this GEMM is the packed gate|up
   * projection with its weight rows interleaved in blocks of 32 (32 gate rows, then the 32 matching up rows); `out` (bf16
   * [M, 2I], same interleaved column order) is written as usual and act_out[m, i] = act(gate[m, i]) * up[m, i] (bf16 [M, I],
   * gated_act selects the activation) — bit-identical to lrp_gated_act_fwd on the stored bf16 values. */
  void* act_out;
  /* Optional fused prologue of the attention backward: when `delta_o` is set, this GEMM is the O-projection dgrad (bf16 output
   * [B*S, H*D] = dO) and the epilogue also emits delta[b,h,s] = sum_d o[b,s,h,d] * dO[b,s,h,d] (fp32 [B,H,S]), the row term of the
   * soft-max backward that lrp_attn_bwd otherwise computes in a separate pass (lxt/efficient/patches.py:193-203 -> SDPA backward).
   * `delta_o` (bf16) has the output's layout; delta_head_dim = D, delta_seq = S. */
  const void* delta_o;
  float* delta_out;
  int32_t delta_head_dim;
  int32_t delta_seq;
  int32_t gated_layout;   /* layout of gated_gu / gated_out: 0 = (gate | up) halves, 1 = blocks of 32 interleaved (see act_out) */
} lrp_epilogue_t;

/* Generic tcgen05 GEMM, A [M,K] bf16.  b_layout 0: B is [N,K] (NT);  b_layout 1: B is [K,N] (NN).
 * tile_n = 0 lets the library choose (one-CTA 128 x 128 / 128 x 256 tiles, or 256 x 256 tiles on CTA pairs with
 * tcgen05 cta_group::2 when they fill the machine); 128 / 256 force a one-CTA tile, 2 forces the CTA-pair kernel. */
int lrp_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, int b_layout, int M, int N, int K,
                  const lrp_epilogue_t* epi, int tile_n, void* stream);

/* Strided-batched form (one launch for `batch` independent problems; 3-D tensor maps): the two contractions of the matmul rule
 * R_A = (s B^T) * A, R_B = (A^T s) * B of lxt/explicit/functional.py:393-408 over all [B*H] slices, and its forward A B.
 * a_layout 0: A_i is [M,K] row-major;  a_layout 1: the caller holds A_i^T, i.e. a [K,M] row-major tensor (consumed through an
 * MN-major descriptor, no transposed copy).  b_layout as in lrp_gemm_bf16.  Problem i reads A + i*stride_a, B + i*stride_b and
 * writes out (resid / shadow) + i*stride_c (strides in elements); rowscale / colscale / bias are shared by all problems. */
int lrp_gemm_bf16_batched(const void* A, int64_t lda, int64_t stride_a, int a_layout, const void* B, int64_t ldb, int64_t stride_b,
                          int b_layout, int batch, int M, int N, int K, const lrp_epilogue_t* epi, int64_t stride_c, void* stream);

/* nn.Linear forward  y[T,N] = x[T,K] W[N,K]^T (+bias) with the fused epilogue above.
 * Replaces: every `nn.Linear` on the path (transformers modeling_llama.py:183,262-264,288), left unpatched
 * by the reference and therefore executed by cuBLAS there. */
int lrp_linear_fwd(const void* x, int64_t ldx, const void* W, int64_t ldw, int T, int N, int K,
                   const lrp_epilogue_t* epi, void* stream);

/* GxI-space ε-LRP backward of nn.Linear  g_x[T,K] = epilogue(g_y[T,N] W[N,K]); the epilogue carries the
 * rules that follow it in the reference graph (RMSNorm identity rule `g*w*rstd`: lxt/efficient/patches.py
 * :111-123; residual accumulation; divide_gradient: lxt/efficient/rules.py:103-127).  W is read in its
 * stored [N,K] layout (no transposed copy). */
int lrp_linear_dgrad_fused(const void* gy, int64_t ldg, const void* W, int64_t ldw, int T, int N, int K,
                           const lrp_epilogue_t* epi, void* stream);

/* Relevance-space ε-LRP rule of nn.Linear (lxt/explicit/functional.py:325-364, `linear_epsilon_fn`):
 *     z = x W^T + b ;  s = R_out / (z + eps) ;  R_in = x ⊙ (s W)
 * One launch: a persistent kernel whose first tile phase forms z and s (s kept in `s_ws`, [T,N] bf16,
 * L2-resident between phases) and whose second phase contracts s with W and multiplies by x.
 * x [T,K] bf16, W [N,K] bf16, bias [N] fp32 or NULL, r_out [T,N] (bf16 or fp32), r_in [T,K] (same dtype
 * as r_out).  `flags_ws` is an int32 scratch of lrp_linear_eps_flags_count(T) entries, zero on entry. */
int lrp_linear_eps_bwd(const void* x, const void* W, const float* bias, const void* r_out, int r_is_f32,
                       void* r_in, void* s_ws, int32_t* flags_ws, int T, int N, int K, float eps, void* stream);
/* number of int32 flags lrp_linear_eps_bwd needs in `flags` (zeroed by the caller) for T rows: the per-op workspace query of
 * SURVEY.md 8(b); replaces the implicit autograd-saved tensors of lxt/explicit/functional.py:340-346 */
int64_t lrp_linear_eps_flags_count(int T);

/* ------------------------------------------------------------------------------------------------
 * RMSNorm with the identity rule (lxt/efficient/patches.py:111-123 `rms_norm_forward`,
 * lxt/efficient/models/gemma3.py:11-12 `gemma3_norm`; relevance form lxt/explicit/functional.py:463-495).
 *   fwd: y = (x * rsqrt(mean(x^2)+eps)) [cast to bf16] * (w + w_offset)   rstd saved (fp32 [T])
 *   bwd (GxI): g_x = g_y * (w + w_offset) * rstd       (the variance path is detached)
 * x_is_f32 selects the dtype of x / g_x; y / g_y are bf16.  w is bf16 [d]; w_offset 0 (Llama) or 1 (Gemma).
 * ---------------------------------------------------------------------------------------------- */
int lrp_rmsnorm_fwd(const void* x, int x_is_f32, const void* w, float w_offset, float eps, void* y, float* rstd,
                    int T, int d, void* stream);
/* backward of the line above with the variance detached (identity rule): gx = gy * (w + w_offset) * rstd,
 * lxt/efficient/patches.py:111-123 (autograd of `hidden_states * rsqrt(variance.detach() + eps)`) */
int lrp_rmsnorm_bwd(const void* gy, const void* w, float w_offset, const float* rstd, void* gx, int gx_is_f32,
                    int accumulate, int T, int d, void* stream);

/* Gemma layer layout `h = residual + post_norm(branch)`: h[t,:] += rmsnorm(y[t,:]) * (w + w_offset), rstd of y saved
 * (transformers 5.5 modeling_gemma3.py:425,431 under the norm patch lxt/efficient/models/gemma3.py:11-19; the backward of the norm is
 * lrp_rmsnorm_bwd). y,w bf16, h fp32. */
int lrp_rmsnorm_fwd_residual(const void* y, const void* w, float w_offset, float eps, float* h, float* rstd, int T, int d,
                             void* stream);
/* Per-head RMSNorm over D of the q and k slices of a packed [T, ld] bf16 buffer, in place (Gemma-3 / Qwen3 q_norm, k_norm:
 * transformers 5.5 modeling_gemma3.py:361-362, modeling_qwen3.py:263-264, patched by lxt/efficient/models/gemma3.py:11-12 / qwen3.py):
 * heads [0, n_q_heads) use wq, the next n_k_heads use wk.  backward = 0: normalise and save rstd [T, n_q+n_k];
 * backward = 1: identity rule g <- g * (w + w_offset) * rstd with the saved rstd. */
int lrp_headnorm_inplace(void* qk, int64_t ld, int n_q_heads, int n_k_heads, int D, const void* wq, const void* wk,
                         float w_offset, float eps, float* rstd, int T, int backward, void* stream);

/* LayerNorm with detached std (lxt/efficient/patches.py:126-142 `layer_norm_forward`); x,w,b,y all bf16 or
 * all fp32 (is_f32).
 *   fwd: y = (x-mean)/sqrt(var+eps) * w + b ; saves mean,rstd.  bwd: g_x = (g_y*w*rstd) - mean_d(g_y*w*rstd) */
int lrp_layernorm_fwd(const void* x, const void* w, const void* b, float eps, void* y, float* mean, float* rstd,
                      int T, int d, int is_f32, void* stream);
/* backward with the std detached (formula above), lxt/efficient/patches.py:126-142 */
int lrp_layernorm_bwd(const void* gy, const void* w, const float* rstd, void* gx, int T, int d, int is_f32,
                      void* stream);

/* Rotary embedding applied in place to the q and k slices of a packed qkv buffer [T, ld] (bf16):
 * heads are contiguous D-wide slices; rotate_half convention (transformers modeling_llama.py:146-168).
 * `inverse`=1 applies the transposed rotation (the backward of RoPE, which is linear in q,k).
 * positions: token t has position (t % S).  cos/sin tables: fp32 [S, D/2]. */
int lrp_rope_inplace(void* qk, int64_t ld, int n_heads_total, int D, const float* cos_t, const float* sin_t,
                     int T, int S, int inverse, void* stream);

/* Gated MLP point-wise part with the identity rule on the activation and the uniform rule on the product
 * (lxt/efficient/patches.py:145-157 `gated_mlp_forward`, lxt/efficient/rules.py:69-127).
 *   fwd: a = act(gate) * up                          gu = [T, 2I] bf16 (gate | up), a = [T, I] bf16
 *   bwd: g_up = (g_a/2) * act(gate);  g_gate = (g_a/2) * up * act(gate)/(gate + 1e-10)
 *        (fp32 arithmetic on the bf16 inputs, one rounding per output)
 *   cp_variant = 1 is CP-LRP (`cp_gated_mlp_forward`, patches.py:272-280): g_gate = 0, g_up = g_a * act(gate). */
int lrp_gated_act_fwd(const void* gu, void* a, int T, int I, int act, void* stream);
/* LRP backward of the gated product: identity rule on act(gate) (lxt/efficient/rules.py:88-100), uniform rule /2 on the product
 * (rules.py:125-127 via patches.py:154); cp_variant = CP-LRP gate detached (patches.py:228-246) */
int lrp_gated_act_bwd(const void* ga, const void* gu, void* ggu, int T, int I, int act, int cp_variant, void* stream);

/* Identity rule on a plain element-wise non-linearity (lxt/efficient/rules.py:88-100,
 * lxt/efficient/patches.py:159-169 `mlp_forward`, :206-211 `non_linear_forward`):
 *   fwd: y = act(x);   bwd: g_x = g_y * act(x)/(x + 1e-10) */
int lrp_act_identity_fwd(const void* x, void* y, int64_t n, int act, int is_f32, void* stream);
/* gx = gy * act(x) / (x + 1e-10): lxt/efficient/rules.py:96-100 */
int lrp_act_identity_bwd(const void* gy, const void* x, void* gx, int64_t n, int act, int is_f32, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Flash AttnLRP (lxt/efficient/patches.py:171-203: SDPA with dQ/4, dK/4, dV/2; softmax is propagated as
 * its ordinary backward = Deep-Taylor rule in GxI space, lxt/explicit/functional.py:276-322).
 * No [B,H,S,S] tensor is materialised.  q [B,S,H,D], k/v [B,S,Hkv,D] given as strided views
 * (row strides in elements: token stride `ld*`, head stride D), o [B,S,H,D] contiguous, lse fp32 [B,H,S].
 * causal: 0/1; window: 0 = none, else sliding window size (keys j with i-j < window).
 * bwd writes dq,dk,dv with the same strides as q,k,v, already scaled by q_div,k_div,v_div (4,4,2 for
 * AttnLRP; 1,1,1 gives the plain gradient; q_div=k_div=0 is CP-LRP: dq=dk=0 is written).
 * `dq_acc_ws` fp32 [B,S,H,D] scratch (zero-filled by the call).
 * ---------------------------------------------------------------------------------------------- */
int lrp_attn_fwd(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv, void* o,
                 float* lse, int B, int S, int H, int Hkv, int D, float scale, int causal, int window,
                 void* stream);
/* LRP backward: plain flash-attention backward with dQ/q_div, dK/k_div, dV/v_div = (4,4,2) for AttnLRP
 * (lxt/efficient/patches.py:193-203), (0,0,1) for CP-LRP (patches.py:249-258; a divisor 0 means "detached": zeros) */
int lrp_attn_bwd(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                 const void* o, const void* d_o, const float* lse, void* dq, void* dk, void* dv, int64_t lddq,
                 int64_t lddk, int64_t lddv, float* dq_acc_ws, float* delta_ws, int B, int S, int H, int Hkv, int D,
                 float scale, int causal, int window, float q_div, float k_div, float v_div, void* stream);

/* Batches of prompts of different lengths (the HF `attention_mask` of a padded batch, which the reference passes through to the
 * wrapped HF attention function untouched: lxt/efficient/patches.py:193-203): `kv_range` is a DEVICE int32 [B,2] array, keys
 * outside [kv_range[2b], kv_range[2b+1]) are masked for every query of sequence b (left or right padding); NULL = no padding.
 * Query rows that see no key produce o = 0, lse = -inf and receive / contribute zero gradient. */
int lrp_attn_fwd_varlen(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv, void* o,
                        float* lse, const int32_t* kv_range, int B, int S, int H, int Hkv, int D, float scale, int causal,
                        int window, void* stream);
/* backward of lrp_attn_fwd_varlen (same rule and workspaces as lrp_attn_bwd; lxt/efficient/patches.py:193-203) */
#define LRP_ATTN_DELTA_READY 1 /* delta_ws already holds sum_d o*dO (written by the O-dgrad GEMM epilogue, lrp_epilogue_t.delta_o) */
#define LRP_ATTN_ACC_ZERO 2    /* dq_acc_ws is zero on entry and is left zero on return (skips the per-call zero-fill) */
int lrp_attn_bwd_varlen(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv,
                        const void* o, const void* d_o, const float* lse, void* dq, void* dk, void* dv, int64_t lddq,
                        int64_t lddk, int64_t lddv, float* dq_acc_ws, float* delta_ws, const int32_t* kv_range, int flags, int B,
                        int S, int H, int Hkv, int D, float scale, int causal, int window, float q_div, float k_div, float v_div,
                        void* stream);

/* ------------------------------------------------------------------------------------------------
 * Ends of the path (examples/quantized_llama.py:35-47)
 * ---------------------------------------------------------------------------------------------- */
/* h[t,:] = float(emb[ids[t],:]) * scale   (ids int64 on device; emb bf16 [V,d]; h fp32 [T,d]) */
int lrp_embed_gather(const int64_t* ids, const void* emb, float scale, float* h, int T, int d, void* stream);
/* out[i,:] = src[rows[i],:] (fp32): the hidden states of the last position of every prompt, `logits[0, -1, :]` reads only those
 * (examples/quantized_llama.py:40) */
int lrp_gather_rows_f32(const float* src, const int64_t* rows, float* out, int n_rows, int d, void* stream);
/* seed of the LRP backward sweep, `max_logits.backward()` (examples/quantized_llama.py:40-44): g_h[t,:] = lm_head[idx[b],:] *
 * (norm_w + w_offset) * rstd_last[b] at the last token t = b*S + S-1 of each prompt (the final RMSNorm's identity rule,
 * lxt/efficient/patches.py:111-123), 0 elsewhere; g_hb (bf16 copy, may be NULL) is the first dgrad's A operand. */
int lrp_seed_gradient(const void* lm_head, const int32_t* idx, const void* norm_w, float w_offset, const float* rstd_last, int S,
                      float* g_h, void* g_hb, int T, int d, void* stream);
/* argmax over logits[b,:] (fp32 [B,V]) -> idx[b] (int32), val[b] */
/* arg-max logit per prompt: `output_logits[0, -1, :].max()` examples/quantized_llama.py:40 */
int lrp_argmax_rows(const float* logits, int32_t* idx, float* val, int B, int V, void* stream);
/* relevance[t] = sum_d x[t,d] * g[t,d]  (x,g fp32) — `(emb * emb.grad).float().sum(-1)` */
/* rel[t] = sum_d x[t,d] * g[t,d]: `(input_embeds * input_embeds.grad).float().sum(-1)` examples/quantized_llama.py:47 */
int lrp_gxi_reduce(const float* x, const float* g, float* rel, int T, int d, void* stream);
/* bf16 variant of the same reduction (x, g bf16) */
/* same for bf16 x / g (the dtype the reference's example runs in, examples/quantized_llama.py:13-19,47) */
int lrp_gxi_reduce_bf16(const void* x, const void* g, float* rel, int T, int d, void* stream);
/* latent relevance of a layer output (docs/source/latent-feature-attribution-efficient.rst:49-90: `output * output.grad` summed over
 * features): x = the bf16 copy of the layer output that the forward residual GEMM epilogue writes as its shadow, g = the fp32
 * gradient stream at that layer in the backward sweep (same reduction as examples/quantized_llama.py:47 applied per layer) */
int lrp_gxi_reduce_mixed(const void* x_bf16, const float* g, float* rel, int T, int d, void* stream);
/* out_bf16 = bf16(in_f32), n elements */
/* fp32 -> bf16 rounding of an activation / gradient stream (the reference's `.to(input_dtype)` casts,
 * lxt/efficient/patches.py:118-123) */
int lrp_cast_f32_to_bf16(const float* in, void* out, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Element-wise rule kernels on arbitrary tensors (the `lxt.explicit.functional` / `lxt.efficient.rules` API).
 * All tensors of one call share a dtype: is_f32 = 1 fp32, 0 bf16; n elements, contiguous.
 * ---------------------------------------------------------------------------------------------- */
/* out = r / (alpha*z + eps): the `relevance_norm` step of every epsilon rule (lxt/explicit/functional.py:359,399;
 * lxt/explicit/rules.py:215,271); `_stabilize` is a plain `+ eps` (functional.py:266-273). */
int lrp_eps_div(const void* r, const void* z, void* out, int64_t n, float alpha, float eps, int is_f32, void* stream);
/* out = a*b*scale: the `.mul_(inputs)` step of the epsilon rules (functional.py:361, rules.py:220). */
int lrp_mul(const void* a, const void* b, void* out, int64_t n, float scale, int is_f32, void* stream);
/* zennit's Gamma rule (third-party `zennit`, unpinned in the reference's setup.py:18, absent here: restated from its published
 * algorithm, PARITY UNPINNED) as the reference runs it in Gradient x Input space (lxt/efficient/zennit_patches.py:33-62,
 * examples/vit_torch.py:59-62).  xcat [rows, 2K] = [max(x,0) | min(x,0)]: the clamped inputs of the four modified passes
 * (zennit_patches.py:42-44) side by side, so the passes become GEMMs with a doubled contraction. */
int lrp_gamma_split(const void* x, void* xcat, int64_t rows, int K, int is_f32, void* stream);
/* scat [rows, 2N] = [ [y>0] g*y/stab(zp) | [y<0] g*y/stab(zn) ]: `grad_output * output` (zennit_patches.py:38) normalised by the
 * positive / negative modified pre-activations with zennit's signed stabiliser; the branch follows the sign of the unmodified output
 * (the rule's gradient_mapper, applied at zennit_patches.py:50). */
int lrp_gamma_s(const void* g, const void* y, const void* zp, const void* zn, void* scat, int64_t rows, int N, float eps, int is_f32,
                void* stream);
/* out = x * (x>0 ? g1 : g2) / stabilize(x, 1e-10): the rule's reducer sum(input_i * gradient_i) (zennit_patches.py:57) followed by
 * the division that returns to the gradient domain (zennit_patches.py:59-60). */
int lrp_gamma_combine(const void* x, const void* g1, const void* g2, void* out, int64_t n, int is_f32, void* stream);
/* out = x*factor: divide_gradient backward (lxt/efficient/rules.py:125-127), mul2 uniform rule (functional.py:524-536). */
int lrp_scale(const void* x, void* out, int64_t n, float factor, int is_f32, void* stream);
/* gx = gy * y/(x+1e-10): identity rule in GxI space for an arbitrary y = f(x) (lxt/efficient/rules.py:88-100). */
int lrp_identity_rule_bwd(const void* gy, const void* x, const void* y, void* gx, int64_t n, int is_f32, void* stream);
/* Deep-Taylor softmax rule over the last dim (functional.py:308-322): out = x~*(r - p*sum(r)), -inf -> 0. */
int lrp_softmax_dt_bwd(const void* x, const void* p, const void* r, void* out, int64_t rows, int cols, int is_f32,
                       void* stream);
/* soft-max forward over the last dim, out = softmax(x / temperature), fp32 arithmetic (lxt/explicit/functional.py:293-306
 * `softmax_fn.forward`; its Deep-Taylor backward is lrp_softmax_dt_bwd). */
int lrp_softmax_fwd(const void* x, void* out, int64_t rows, int cols, float temperature, int is_f32, void* stream);
/* epsilon rule of a+b (functional.py:439-459): ra = a*r/(a+b+eps), rb = b*r/(a+b+eps). */
int lrp_add2_bwd(const void* a, const void* b, const void* r, void* ra, void* rb, int64_t n, float eps, int is_f32,
                 void* stream);

/* ------------------------------------------------------------------------------------------------
 * Validation-precision mode (fp32 activations): the same rule kernels with the activation dtype selectable, the two-term
 * bf16 split that feeds fp32 activations to the bf16 tcgen05 GEMM (out = hi W + lo W, fp32 accumulation), and an fp32
 * CUDA-core flash AttnLRP.  Purpose: show that the bf16 path's distance to the reference's fp32 run is storage rounding
 * (the reference itself runs in whatever dtype the HF model has; fp32: tests of lxt/explicit, bf16: examples/quantized_llama.py:13-19).
 * `*_is_f32` = 1 selects fp32, 0 bf16, for the named tensor; weights stay bf16.
 * ---------------------------------------------------------------------------------------------- */
/* lrp_rmsnorm_fwd with a selectable output dtype (lxt/efficient/patches.py:111-123) */
int lrp_rmsnorm_fwd_t(const void* x, int x_is_f32, const void* w, float w_offset, float eps, void* y, int y_is_f32, float* rstd,
                      int T, int d, void* stream);
/* lrp_rmsnorm_bwd with a selectable input dtype (identity rule, lxt/efficient/patches.py:111-123) */
int lrp_rmsnorm_bwd_t(const void* gy, int gy_is_f32, const void* w, float w_offset, const float* rstd, void* gx, int gx_is_f32,
                      int accumulate, int T, int d, void* stream);
/* lrp_rmsnorm_fwd_residual with a selectable branch dtype (lxt/efficient/models/gemma3.py:11-19) */
int lrp_rmsnorm_fwd_residual_t(const void* y, int y_is_f32, const void* w, float w_offset, float eps, float* h, float* rstd, int T,
                               int d, void* stream);
/* lrp_headnorm_inplace on a bf16 or fp32 packed buffer (lxt/efficient/models/gemma3.py:11-12) */
int lrp_headnorm_inplace_t(void* qk, int is_f32, int64_t ld, int n_q_heads, int n_k_heads, int D, const void* wq, const void* wk,
                           float w_offset, float eps, float* rstd, int T, int backward, void* stream);
/* lrp_rope_inplace on a bf16 or fp32 packed buffer (RoPE is left unpatched by the reference: transformers modeling_llama.py:146-168,
 * SURVEY 0 table; explicit form lxt/explicit/models/llama.py:226-260) */
int lrp_rope_inplace_t(void* qk, int is_f32, int64_t ld, int n_heads_total, int D, const float* cos_t, const float* sin_t, int T,
                       int S, int inverse, void* stream);
/* lrp_gated_act_fwd on bf16 or fp32 tensors (lxt/efficient/patches.py:145-157); layout 0: gu = (gate | up) halves, layout 1:
 * blocks of 32 interleaved (32 gate columns, the 32 matching up columns, ...) as the fused gate|up GEMM epilogue writes them */
int lrp_gated_act_fwd_t(const void* gu, void* a, int is_f32, int layout, int T, int I, int act, void* stream);
/* lrp_gated_act_bwd on bf16 or fp32 tensors, same layouts for gu and ggu (lxt/efficient/rules.py:88-127) */
int lrp_gated_act_bwd_t(const void* ga, const void* gu, void* ggu, int is_f32, int layout, int T, int I, int act, int cp_variant,
                        void* stream);
/* x (fp32, n elements) -> hi = bf16(x), lo = bf16(x - hi): operands of the split GEMM that replaces the fp32 `F.linear` of an fp32
 * model (transformers modeling_llama.py:183,262-264,288 executed in fp32, as in the reference's own tests/test_functional.py:57-76) */
int lrp_split_bf16x2(const float* x, void* hi, void* lo, int64_t n, void* stream);
/* fp32 flash AttnLRP forward (same contract as lrp_attn_fwd_varlen, all tensors fp32; lxt/efficient/patches.py:193-203) */
int lrp_attn_fwd_f32(const float* q, const float* k, const float* v, int64_t ldq, int64_t ldk, int64_t ldv, float* o, float* lse,
                     const int32_t* kv_range, int B, int S, int H, int Hkv, int D, float scale, int causal, int window, void* stream);
/* fp32 flash AttnLRP backward (same contract as lrp_attn_bwd_varlen; delta_ws fp32 [B,H,S]; lxt/efficient/patches.py:193-203) */
int lrp_attn_bwd_f32(const float* q, const float* k, const float* v, int64_t ldq, int64_t ldk, int64_t ldv, const float* o,
                     const float* d_o, const float* lse, float* dq, float* dk, float* dv, int64_t lddq, int64_t lddk,
                     int64_t lddv, float* delta_ws, const int32_t* kv_range, int B, int S, int H, int Hkv, int D, float scale,
                     int causal, int window, float q_div, float k_div, float v_div, void* stream);
/* workspace query of lrp_attn_bwd (SURVEY.md 8(b) "a *_workspace_bytes query per op"): bytes of `dq_acc_ws` and `delta_ws` for this
 * shape; replaces the tensors autograd saves / allocates in the reference's SDPA backward (lxt/efficient/patches.py:193-203) */
int lrp_attn_bwd_workspace_bytes(int B, int S, int H, int D, int64_t* dq_acc_bytes, int64_t* delta_bytes);

/* ------------------------------------------------------------------------------------------------
 * 4-bit NormalFloat (NF4) weight storage: weights rest in HBM as 4-bit codes + one fp32 absmax per block and
 * are expanded to bf16 into a scratch right before the tcgen05 GEMMs of a layer.  Replaces bitsandbytes' Linear4bit storage that
 * every reference example uses (examples/quantized_llama.py:13-19; wrapped by lxt/explicit/models/llama.py:91-92).  Format: 16 NF4
 * code points, two codes per byte (first value in the high nibble), block size a multiple of 8 dividing n.  Parity is pinned
 * against oracle/nf4_oracle.py (bit-exact), not against bitsandbytes (absent).
 * ---------------------------------------------------------------------------------------------- */
/* quantise n bf16 weights (examples/quantized_llama.py:13-19 `BitsAndBytesConfig(load_in_4bit=True)`) */
int lrp_quant_nf4(const void* w_bf16, void* packed, float* absmax, int64_t n, int blocksize, void* stream);
/* expand to bf16: what bitsandbytes does inside Linear4bit.forward / backward (lxt/explicit/models/llama.py:91-92 call sites) */
int lrp_dequant_nf4(const void* packed, const float* absmax, void* out_bf16, int64_t n, int blocksize, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LRP_B200_H_ */
