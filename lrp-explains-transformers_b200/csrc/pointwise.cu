// HBM-bound kernels of the AttnLRP path: norms with the identity rule, RoPE, the gated-MLP point-wise rules,
// the ends of the path (embedding gather, arg-max seed, Gradient x Input reduction).
// All are single-pass, 16-byte vectorised, coalesced along the contiguous feature dimension; grids are sized
// in rows (>> 148 SMs x resident CTAs at the benchmark shapes).
#include "ptx_sm100.cuh"
#include "lrp_internal.h"

namespace lrp {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int THREADS>
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < THREADS / 32) ? red[l] : 0.f;
  t = warp_sum(t);
  __syncthreads();
  return t;
}

__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ void load8(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                            pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}
__device__ __forceinline__ void store8(float* p, const float (&f)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}

// ------------------------------------------------------------------------------------------------
// RMSNorm, identity rule
// ------------------------------------------------------------------------------------------------
constexpr int NORM_THREADS = 256;

// y = bf16( (x * rstd) * (w_offset + w) ), all arithmetic in fp32 with one final rounding.
//   Llama: w_offset = 0 (patches.py:111-123; the reference's intermediate bf16 downcast before `weight *` is a
//   storage artefact of its bf16 tensors, not part of the rule);  Gemma: w_offset = 1 (gemma3.py:11-12).
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(NORM_THREADS) rmsnorm_fwd_kernel(const TIn* __restrict__ x,
                                                                   const __nv_bfloat16* __restrict__ w,
                                                                   float w_offset, float eps,
                                                                   TOut* __restrict__ y,
                                                                   float* __restrict__ rstd_out, int d) {
  __shared__ float red[NORM_THREADS / 32];
  const int64_t row = blockIdx.x;
  const TIn* xr = x + row * d;
  float ss = 0.f;
  for (int i = threadIdx.x * 8; i < d; i += NORM_THREADS * 8) {
    float f[8];
    load8(xr + i, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
  }
  ss = block_sum<NORM_THREADS>(ss, red);
  const float rstd = rsqrtf(ss / float(d) + eps);
  if (threadIdx.x == 0 && rstd_out != nullptr) rstd_out[row] = rstd;
  TOut* yr = y + row * d;
  for (int i = threadIdx.x * 8; i < d; i += NORM_THREADS * 8) {
    float f[8], wf[8], o[8];
    load8(xr + i, f);
    load8(w + i, wf);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j] = (f[j] * rstd) * (w_offset + wf[j]);
    }
    store8(yr + i, o);
  }
}

// GxI backward: g_x = g_y * (w + w_offset) * rstd  (variance path detached = identity rule)
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(NORM_THREADS) rmsnorm_bwd_kernel(const TIn* __restrict__ gy,
                                                                   const __nv_bfloat16* __restrict__ w,
                                                                   float w_offset,
                                                                   const float* __restrict__ rstd,
                                                                   TOut* __restrict__ gx, int accumulate, int d) {
  const int64_t row = blockIdx.x;
  const float r = rstd[row];
  for (int i = threadIdx.x * 8; i < d; i += NORM_THREADS * 8) {
    float g[8], wf[8], o[8];
    load8(gy + row * d + i, g);
    load8(w + i, wf);
    if (accumulate) load8(gx + row * d + i, o);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = g[j] * (wf[j] + w_offset) * r;
      o[j] = accumulate ? o[j] + v : v;
    }
    store8(gx + row * d + i, o);
  }
}

// h[t,:] += (y[t,:] * rsqrt(mean(y^2)+eps)) * (w_offset + w)   — Gemma's post-branch RMSNorm fused with the residual add
// (HF Gemma3DecoderLayer: hidden = residual + post_norm(branch)); rstd of the branch output saved for the backward.
template <typename TY>
__global__ void __launch_bounds__(NORM_THREADS) rmsnorm_fwd_residual_kernel(const TY* __restrict__ y,
                                                                            const __nv_bfloat16* __restrict__ w,
                                                                            float w_offset, float eps, float* __restrict__ h,
                                                                            float* __restrict__ rstd_out, int d) {
  __shared__ float red[NORM_THREADS / 32];
  const int64_t row = blockIdx.x;
  const TY* yr = y + row * d;
  float ss = 0.f;
  for (int i = threadIdx.x * 8; i < d; i += NORM_THREADS * 8) {
    float f[8];
    load8(yr + i, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
  }
  ss = block_sum<NORM_THREADS>(ss, red);
  const float rstd = rsqrtf(ss / float(d) + eps);
  if (threadIdx.x == 0) rstd_out[row] = rstd;
  for (int i = threadIdx.x * 8; i < d; i += NORM_THREADS * 8) {
    float f[8], wf[8], o[8];
    load8(yr + i, f);
    load8(w + i, wf);
    load8(h + row * d + i, o);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] += (f[j] * rstd) * (w_offset + wf[j]);
    store8(h + row * d + i, o);
  }
}

// Per-head RMSNorm over D on the q and k slices of a packed qkv buffer, in place (Gemma-3 / Qwen3 q_norm, k_norm), one
// warp per (token, head).  backward = 0: x <- x * rstd * (off + w), rstd saved [T, n_heads];
// backward = 1 (identity rule): g <- g * (off + w) * rstd with the saved rstd.
template <typename TQ>
__global__ void __launch_bounds__(256) headnorm_kernel(TQ* __restrict__ qk, int64_t ld, int n_q, int n_heads, int D,
                                                       const __nv_bfloat16* __restrict__ wq, const __nv_bfloat16* __restrict__ wk,
                                                       float w_offset, float eps, float* __restrict__ rstd, int64_t T, int backward) {
  const int64_t row = blockIdx.x * int64_t(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= T * n_heads) return;
  const int lane = threadIdx.x & 31;
  const int64_t t = row / n_heads;
  const int hd = int(row - t * n_heads);
  TQ* p = qk + t * ld + int64_t(hd) * D;
  const __nv_bfloat16* w = hd < n_q ? wq : wk;
  float r;
  if (!backward) {
    float ss = 0.f;
    for (int i = lane * 8; i < D; i += 256) {
      float f[8];
      load8(p + i, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
    }
    ss = warp_sum(ss);
    r = rsqrtf(ss / float(D) + eps);
    if (lane == 0) rstd[row] = r;
  } else {
    r = rstd[row];
  }
  for (int i = lane * 8; i < D; i += 256) {
    float f[8], wf[8];
    load8(p + i, f);
    load8(w + i, wf);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = f[j] * r * (w_offset + wf[j]);
    store8(p + i, f);
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm with detached std (ViT path), bf16 or fp32 rows
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(NORM_THREADS) layernorm_fwd_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                                     const T* __restrict__ b, float eps,
                                                                     T* __restrict__ y, float* __restrict__ mean_out,
                                                                     float* __restrict__ rstd_out, int d) {
  __shared__ float red[NORM_THREADS / 32];
  const int64_t row = blockIdx.x;
  const T* xr = x + row * d;
  float s = 0.f;
  for (int i = threadIdx.x * 8; i < d; i += NORM_THREADS * 8) {
    float f[8];
    load8(xr + i, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += f[j];
  }
  const float mean = block_sum<NORM_THREADS>(s, red) / float(d);
  float ss = 0.f;
  for (int i = threadIdx.x * 8; i < d; i += NORM_THREADS * 8) {
    float f[8];
    load8(xr + i, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) ss += (f[j] - mean) * (f[j] - mean);
  }
  const float var = block_sum<NORM_THREADS>(ss, red) / float(d);
  const float rstd = 1.f / sqrtf(var + eps);
  if (threadIdx.x == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  for (int i = threadIdx.x * 8; i < d; i += NORM_THREADS * 8) {
    float f[8], wf[8], bf[8], o[8];
    load8(xr + i, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (f[j] - mean) * rstd;
    if (w != nullptr) {
      load8(w + i, wf);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] *= wf[j];
    }
    if (b != nullptr) {
      load8(b + i, bf);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += bf[j];
    }
    store8(y + row * d + i, o);
  }
}

// y = (x - mean(x)) * rstd * w + b with rstd detached:  g_x = u - mean_d(u),  u = g_y * w * rstd
template <typename T>
__global__ void __launch_bounds__(NORM_THREADS) layernorm_bwd_kernel(const T* __restrict__ gy, const T* __restrict__ w,
                                                                     const float* __restrict__ rstd,
                                                                     T* __restrict__ gx, int d) {
  __shared__ float red[NORM_THREADS / 32];
  const int64_t row = blockIdx.x;
  const float r = rstd[row];
  float s = 0.f;
  for (int i = threadIdx.x * 8; i < d; i += NORM_THREADS * 8) {
    float g[8], wf[8];
    load8(gy + row * d + i, g);
    if (w != nullptr) load8(w + i, wf);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += g[j] * (w != nullptr ? wf[j] : 1.f) * r;
  }
  const float mu = block_sum<NORM_THREADS>(s, red) / float(d);
  for (int i = threadIdx.x * 8; i < d; i += NORM_THREADS * 8) {
    float g[8], wf[8], o[8];
    load8(gy + row * d + i, g);
    if (w != nullptr) load8(w + i, wf);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = g[j] * (w != nullptr ? wf[j] : 1.f) * r - mu;
    store8(gx + row * d + i, o);
  }
}

// ------------------------------------------------------------------------------------------------
// RoPE in place on packed heads (rotate_half convention)
// ------------------------------------------------------------------------------------------------
// one thread handles 8 consecutive "pair" indices i..i+7 of one head: x1 = x[i], x2 = x[i + D/2]
template <typename TQ>
__global__ void __launch_bounds__(256) rope_kernel(TQ* __restrict__ qk, int64_t ld, int n_heads, int D,
                                                   const float* __restrict__ cos_t, const float* __restrict__ sin_t,
                                                   int64_t T, int S, int inverse) {
  const int half = D >> 1;
  const int chunks_per_head = half >> 3;
  const int64_t total = T * int64_t(n_heads) * chunks_per_head;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(idx % chunks_per_head);
    const int64_t r = idx / chunks_per_head;
    const int h = int(r % n_heads);
    const int64_t t = r / n_heads;
    const int pos = int(t % S);
    TQ* p = qk + t * ld + int64_t(h) * D + c * 8;
    float x1[8], x2[8], cs[8], sn[8], y1[8], y2[8];
    load8(p, x1);
    load8(p + half, x2);
    load8(cos_t + int64_t(pos) * half + c * 8, cs);
    load8(sin_t + int64_t(pos) * half + c * 8, sn);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s = inverse ? -sn[j] : sn[j];
      y1[j] = x1[j] * cs[j] - x2[j] * s;
      y2[j] = x2[j] * cs[j] + x1[j] * s;
    }
    store8(p, y1);
    store8(p + half, y2);
  }
}

// ------------------------------------------------------------------------------------------------
// activations + identity / uniform rules
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_eval(float x, int act) {
  if (act == LRP_ACT_SILU) return x / (1.f + __expf(-x));
  if (act == LRP_ACT_GELU_TANH) {
    const float k = 0.7978845608028654f;
    return 0.5f * x * (1.f + tanhf(k * (x + 0.044715f * x * x * x)));
  }
  return 0.5f * x * (1.f + erff(x * 0.7071067811865476f));
}

// a = bf16( act(gate) * up )
// column of gate element c / distance to the matching up element in a [T, 2I] gate|up row:
//   layout 0: halves (gate | up);  layout 1: blocks of 32 interleaved (32 gate, 32 up, ...) as the fused GEMM epilogue writes them
__device__ __forceinline__ int gu_col(int c, int layout) { return layout ? ((c >> 5) << 6) + (c & 31) : c; }
__device__ __forceinline__ int gu_up(int I, int layout) { return layout ? 32 : I; }

template <typename TA>
__global__ void __launch_bounds__(256) gated_act_fwd_kernel(const TA* __restrict__ gu,
                                                            TA* __restrict__ a, int64_t T, int I, int act, int layout) {
  const int chunks = I >> 3;
  const int64_t total = T * chunks;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t t = idx / chunks;
    const int c = int(idx - t * chunks) * 8;
    float g[8], u[8], o[8];
    load8(gu + t * 2 * I + gu_col(c, layout), g);
    load8(gu + t * 2 * I + gu_col(c, layout) + gu_up(I, layout), u);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = act_eval(g[j], act) * u[j];
    store8(a + t * I + c, o);
  }
}

// g_half = g_a / 2                           (divide_gradient, rules.py:125-127)
// g_up   = g_half * s                        s = act(gate)
// g_gate = (s / (gate + 1e-10)) * (g_half * up)                      (identity rule, rules.py:88-100)
// fp32 arithmetic on the bf16 inputs, one rounding on each output.
template <typename TA>
__global__ void __launch_bounds__(256) gated_act_bwd_kernel(const TA* __restrict__ ga,
                                                            const TA* __restrict__ gu,
                                                            TA* __restrict__ ggu, int64_t T, int I, int act,
                                                            int cp_variant, int layout) {
  const int chunks = I >> 3;
  const int64_t total = T * chunks;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t t = idx / chunks;
    const int c = int(idx - t * chunks) * 8;
    float g[8], u[8], d[8], og[8], ou[8];
    const int64_t gc = t * 2 * I + gu_col(c, layout);
    const int uo = gu_up(I, layout);
    load8(gu + gc, g);
    load8(gu + gc + uo, u);
    load8(ga + t * I + c, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s = act_eval(g[j], act);
      if (cp_variant) {  // CP-LRP (patches.py:272-280): gate detached, no uniform split
        ou[j] = d[j] * s;
        og[j] = 0.f;
      } else {
        const float gh = d[j] * 0.5f;
        ou[j] = gh * s;
        og[j] = (s / (g[j] + 1e-10f)) * (gh * u[j]);
      }
    }
    store8(ggu + gc, og);
    store8(ggu + gc + uo, ou);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) act_identity_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t n8,
                                                               int act) {
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < n8; idx += int64_t(gridDim.x) * blockDim.x) {
    float f[8], o[8];
    load8(x + idx * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = act_eval(f[j], act);
    store8(y + idx * 8, o);
  }
}
template <typename T>
__global__ void __launch_bounds__(256) act_identity_bwd_kernel(const T* __restrict__ gy, const T* __restrict__ x,
                                                               T* __restrict__ gx, int64_t n8, int act) {
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < n8; idx += int64_t(gridDim.x) * blockDim.x) {
    float f[8], g[8], o[8];
    load8(x + idx * 8, f);
    load8(gy + idx * 8, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j] = (act_eval(f[j], act) / (f[j] + 1e-10f)) * g[j];
    }
    store8(gx + idx * 8, o);
  }
}

// ------------------------------------------------------------------------------------------------
// ends of the path
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) embed_gather_kernel(const int64_t* __restrict__ ids,
                                                           const __nv_bfloat16* __restrict__ emb, float scale,
                                                           float* __restrict__ h, int d) {
  const int64_t t = blockIdx.x;
  const int64_t id = ids[t];
  for (int i = threadIdx.x * 8; i < d; i += blockDim.x * 8) {
    float f[8];
    load8(emb + id * d + i, f);
    if (scale != 1.f) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = f[j] * scale;
    }
    store8(h + t * d + i, f);
  }
}

// out[i, :] = src[rows[i], :]  (fp32 rows; the last-position hidden states that feed the final norm / lm_head)
__global__ void __launch_bounds__(256) gather_rows_f32_kernel(const float* __restrict__ src, const int64_t* __restrict__ rows,
                                                              float* __restrict__ out, int d) {
  const int64_t r = rows[blockIdx.x];
  for (int i = threadIdx.x * 4; i < d; i += blockDim.x * 4)
    *reinterpret_cast<float4*>(out + int64_t(blockIdx.x) * d + i) = *reinterpret_cast<const float4*>(src + r * d + i);
}

// Seed of the backward sweep (examples/quantized_llama.py:40-44: `max_logits.backward()` of the arg-max logit at the last position):
// d logit[idx_b] / d h is lm_head[idx_b] pulled through the final RMSNorm with the identity rule (g * (w + off) * rstd) at the last
// token of prompt b and zero everywhere else.  One pass writes the fp32 gradient stream and its bf16 shadow.
__global__ void __launch_bounds__(256) seed_gradient_kernel(const __nv_bfloat16* __restrict__ lm_head, const int32_t* __restrict__ idx,
                                                            const __nv_bfloat16* __restrict__ norm_w, float w_offset,
                                                            const float* __restrict__ rstd_last, int S, float* __restrict__ g_h,
                                                            __nv_bfloat16* __restrict__ g_hb, int d) {
  const int64_t t = blockIdx.x;
  const bool last = (t % S) == S - 1;
  const int64_t b = t / S;
  const __nv_bfloat16* wrow = last ? lm_head + int64_t(idx[b]) * d : nullptr;
  const float r = last ? rstd_last[b] : 0.f;
  for (int i = threadIdx.x * 8; i < d; i += blockDim.x * 8) {
    float o[8];
    if (last) {
      float lw[8], nw[8];
      load8(wrow + i, lw);
      load8(norm_w + i, nw);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = lw[j] * (nw[j] + w_offset) * r;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;
    }
    store8(g_h + t * d + i, o);
    if (g_hb != nullptr) store8(g_hb + t * d + i, o);
  }
}

__global__ void __launch_bounds__(1024) argmax_rows_kernel(const float* __restrict__ logits, int32_t* __restrict__ idx,
                                                           float* __restrict__ val, int V) {
  __shared__ float sv[32];
  __shared__ int si[32];
  const float* row = logits + int64_t(blockIdx.x) * V;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    const float v = row[i];
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    best = threadIdx.x < (blockDim.x >> 5) ? sv[threadIdx.x] : -INFINITY;
    bi = threadIdx.x < (blockDim.x >> 5) ? si[threadIdx.x] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (threadIdx.x == 0) { idx[blockIdx.x] = bi; if (val) val[blockIdx.x] = best; }
  }
}

template <typename T, typename TG = T>
__global__ void __launch_bounds__(256) gxi_reduce_kernel(const T* __restrict__ x, const TG* __restrict__ g,
                                                         float* __restrict__ rel, int64_t Tn, int d) {
  // one warp per row
  const int64_t row = blockIdx.x * int64_t(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= Tn) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int i = lane * 8; i < d; i += 32 * 8) {
    float a[8], b[8];
    load8(x + row * d + i, a);
    load8(g + row * d + i, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s += a[j] * b[j];
    }
  }
  s = warp_sum(s);
  if (lane == 0) rel[row] = s;
}

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                            int64_t n8) {
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < n8; idx += int64_t(gridDim.x) * blockDim.x) {
    float f[8];
    load8(in + idx * 8, f);
    store8(out + idx * 8, f);
  }
}

// x (fp32) = hi + lo with hi = bf16(x), lo = bf16(x - hi): the two-term bf16 split that lets the bf16 tcgen05 GEMM
// contract an fp32 activation with ~16 mantissa bits (validation-precision mode: out = hi W + lo W, fp32 accumulation)
__global__ void __launch_bounds__(256) split_bf16x2_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ hi,
                                                           __nv_bfloat16* __restrict__ lo, int64_t n8) {
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < n8; idx += int64_t(gridDim.x) * blockDim.x) {
    float f[8], h[8], l[8];
    load8(in + idx * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      h[j] = round_bf16(f[j]);
      l[j] = f[j] - h[j];
    }
    store8(hi + idx * 8, h);
    store8(lo + idx * 8, l);
  }
}

// ------------------------------------------------------------------------------------------------
// generic element-wise rule kernels (explicit / efficient rule API on arbitrary tensors); scalar tail handled
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ldf(const float* p, int64_t i) { return p[i]; }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p, int64_t i) { return __bfloat162float(p[i]); }
__device__ __forceinline__ void stf(float* p, int64_t i, float v) { p[i] = v; }
__device__ __forceinline__ void stf(__nv_bfloat16* p, int64_t i, float v) { p[i] = __float2bfloat16_rn(v); }

// out = r / (alpha * z + eps)      (_stabilize: plain "+ eps", lxt/explicit/functional.py:266-273)
template <typename T>
__global__ void __launch_bounds__(256) eps_div_kernel(const T* __restrict__ r, const T* __restrict__ z, T* __restrict__ out,
                                                      int64_t n, float alpha, float eps) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    stf(out, i, ldf(r, i) / (alpha * ldf(z, i) + eps));
}
// out = a * b * scale
template <typename T>
__global__ void __launch_bounds__(256) mul_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out,
                                                  int64_t n, float scale) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    stf(out, i, ldf(a, i) * ldf(b, i) * scale);
}
// out = x * factor
template <typename T>
__global__ void __launch_bounds__(256) scale_kernel(const T* __restrict__ x, T* __restrict__ out, int64_t n, float factor) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    stf(out, i, ldf(x, i) * factor);
}
// ---- Gamma rule in Gradient x Input space (zennit Gamma through lxt/efficient/zennit_patches.py:26-62) ----
// xcat[t, 0:K] = max(x, 0), xcat[t, K:2K] = min(x, 0): the two clamped inputs of the rule side by side, so that the four modified
// forward passes become two GEMMs with a doubled contraction
template <typename T>
__global__ void __launch_bounds__(256) gamma_split_kernel(const T* __restrict__ x, T* __restrict__ xcat, int64_t rows, int K) {
  const int64_t n = rows * K;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t t = i / K;
    const int k = int(i - t * K);
    const float v = ldf(x, i);
    stf(xcat, t * 2 * K + k, fmaxf(v, 0.f));
    stf(xcat, t * 2 * K + K + k, fminf(v, 0.f));
  }
}
// scat[t, 0:N] = [y > 0] g y / stab(zp), scat[t, N:2N] = [y < 0] g y / stab(zn): relevance g*y of the layer output normalised by the
// positive / negative modified pre-activation; the branch follows the sign of the UNMODIFIED output y (the rule's fifth, unmodified
// pass), stab(z) = z + ((z == 0) + sign(z)) eps is zennit's signed stabiliser
template <typename T>
__global__ void __launch_bounds__(256) gamma_s_kernel(const T* __restrict__ g, const T* __restrict__ y, const T* __restrict__ zp,
                                                      const T* __restrict__ zn, T* __restrict__ scat, int64_t rows, int N, float eps) {
  const int64_t n = rows * N;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t t = i / N;
    const int c = int(i - t * N);
    const float yv = ldf(y, i);
    const float r = ldf(g, i) * yv;
    const float p = ldf(zp, i), q = ldf(zn, i);
    stf(scat, t * 2 * N + c, yv > 0.f ? r / (p + (p < 0.f ? -eps : eps)) : 0.f);
    stf(scat, t * 2 * N + N + c, yv < 0.f ? r / (q + (q < 0.f ? -eps : eps)) : 0.f);
  }
}
// g_x = x (x > 0 ? g1 : g2) / (x + sign(x) 1e-10), 0 at x = 0: Sigma_i input_i * gradient_i of the rule, then the division that takes
// the relevance back to a gradient (zennit_patches.py:59-60)
template <typename T>
__global__ void __launch_bounds__(256) gamma_combine_kernel(const T* __restrict__ x, const T* __restrict__ g1, const T* __restrict__ g2,
                                                            T* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const float v = ldf(x, i);
    float o = 0.f;
    if (v > 0.f) o = v * ldf(g1, i) / (v + 1e-10f);
    else if (v < 0.f) o = v * ldf(g2, i) / (v - 1e-10f);
    stf(out, i, o);
  }
}
// gx = gy * (y / (x + 1e-10))     identity rule in GxI space for an arbitrary f with y = f(x)  (rules.py:88-100)
template <typename T>
__global__ void __launch_bounds__(256) identity_rule_bwd_kernel(const T* __restrict__ gy, const T* __restrict__ x,
                                                                const T* __restrict__ y, T* __restrict__ gx, int64_t n) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    stf(gx, i, ldf(gy, i) * (ldf(y, i) / (ldf(x, i) + 1e-10f)));
}
// Deep-Taylor softmax rule (functional.py:308-322): out = x~ * (r - p * sum_row(r)), x~ = x with -inf -> 0; one warp per row
template <typename T>
__global__ void __launch_bounds__(256) softmax_dt_bwd_kernel(const T* __restrict__ x, const T* __restrict__ p,
                                                             const T* __restrict__ r, T* __restrict__ out, int64_t rows, int cols) {
  const int64_t row = blockIdx.x * int64_t(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s += ldf(r, row * cols + c);
  s = warp_sum(s);
  for (int c = lane; c < cols; c += 32) {
    const int64_t i = row * cols + c;
    float xv = ldf(x, i);
    if (xv == -INFINITY) xv = 0.f;
    stf(out, i, xv * (ldf(r, i) - ldf(p, i) * s));
  }
}
// soft-max forward over the last dimension, p = softmax(x / temperature) (functional.py:293-306); fp32 arithmetic, one warp per
// row, -inf inputs give 0, a row of only -inf gives zeros
template <typename T>
__global__ void __launch_bounds__(256) softmax_fwd_kernel(const T* __restrict__ x, T* __restrict__ out, int64_t rows, int cols,
                                                          float inv_temp) {
  const int64_t row = blockIdx.x * int64_t(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int64_t base = row * cols;
  float m = -INFINITY;
  for (int c = lane; c < cols; c += 32) m = fmaxf(m, ldf(x, base + c) * inv_temp);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (m == -INFINITY) {
    for (int c = lane; c < cols; c += 32) stf(out, base + c, 0.f);
    return;
  }
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s += expf(ldf(x, base + c) * inv_temp - m);
  s = warp_sum(s);
  const float inv = 1.f / s;
  for (int c = lane; c < cols; c += 32) stf(out, base + c, expf(ldf(x, base + c) * inv_temp - m) * inv);
}
// epsilon rule for a + b (functional.py:439-459): s = r / (a + b + eps); ra = s*a; rb = s*b
template <typename T>
__global__ void __launch_bounds__(256) add2_bwd_kernel(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ r,
                                                       T* __restrict__ ra, T* __restrict__ rb, int64_t n, float eps) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const float av = ldf(a, i), bv = ldf(b, i);
    const float sv = ldf(r, i) / (av + bv + eps);
    stf(ra, i, sv * av);
    stf(rb, i, sv * bv);
  }
}

static inline int grid_for(int64_t work_items, int threads) {
  int64_t g = (work_items + threads - 1) / threads;
  const int64_t cap = int64_t(sm_count()) * 16;
  return int(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace lrp

using namespace lrp;
typedef __nv_bfloat16 bf16;

extern "C" {

int lrp_rmsnorm_fwd_t(const void* x, int x_is_f32, const void* w, float w_offset, float eps, void* y, int y_is_f32, float* rstd,
                      int T, int d, void* stream) {
  if (T <= 0 || d <= 0 || (d % 8) != 0) return set_error(LRP_ERR_ARG, "rmsnorm_fwd: d must be a positive multiple of 8");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (x_is_f32 && y_is_f32)
    rmsnorm_fwd_kernel<float, float><<<T, NORM_THREADS, 0, st>>>((const float*)x, (const bf16*)w, w_offset, eps, (float*)y, rstd, d);
  else if (x_is_f32)
    rmsnorm_fwd_kernel<float, bf16><<<T, NORM_THREADS, 0, st>>>((const float*)x, (const bf16*)w, w_offset, eps, (bf16*)y, rstd, d);
  else if (y_is_f32)
    rmsnorm_fwd_kernel<bf16, float><<<T, NORM_THREADS, 0, st>>>((const bf16*)x, (const bf16*)w, w_offset, eps, (float*)y, rstd, d);
  else
    rmsnorm_fwd_kernel<bf16, bf16><<<T, NORM_THREADS, 0, st>>>((const bf16*)x, (const bf16*)w, w_offset, eps, (bf16*)y, rstd, d);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

int lrp_rmsnorm_fwd(const void* x, int x_is_f32, const void* w, float w_offset, float eps, void* y, float* rstd,
                    int T, int d, void* stream) {
  return lrp_rmsnorm_fwd_t(x, x_is_f32, w, w_offset, eps, y, 0, rstd, T, d, stream);
}

int lrp_rmsnorm_bwd_t(const void* gy, int gy_is_f32, const void* w, float w_offset, const float* rstd, void* gx, int gx_is_f32,
                      int accumulate, int T, int d, void* stream) {
  if (T <= 0 || d <= 0 || (d % 8) != 0) return set_error(LRP_ERR_ARG, "rmsnorm_bwd: d must be a positive multiple of 8");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (gy_is_f32 && gx_is_f32)
    rmsnorm_bwd_kernel<float, float><<<T, NORM_THREADS, 0, st>>>((const float*)gy, (const bf16*)w, w_offset, rstd, (float*)gx, accumulate, d);
  else if (gy_is_f32)
    rmsnorm_bwd_kernel<float, bf16><<<T, NORM_THREADS, 0, st>>>((const float*)gy, (const bf16*)w, w_offset, rstd, (bf16*)gx, accumulate, d);
  else if (gx_is_f32)
    rmsnorm_bwd_kernel<bf16, float><<<T, NORM_THREADS, 0, st>>>((const bf16*)gy, (const bf16*)w, w_offset, rstd, (float*)gx, accumulate, d);
  else
    rmsnorm_bwd_kernel<bf16, bf16><<<T, NORM_THREADS, 0, st>>>((const bf16*)gy, (const bf16*)w, w_offset, rstd, (bf16*)gx, accumulate, d);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

int lrp_rmsnorm_bwd(const void* gy, const void* w, float w_offset, const float* rstd, void* gx, int gx_is_f32,
                    int accumulate, int T, int d, void* stream) {
  return lrp_rmsnorm_bwd_t(gy, 0, w, w_offset, rstd, gx, gx_is_f32, accumulate, T, d, stream);
}

int lrp_rmsnorm_fwd_residual_t(const void* y, int y_is_f32, const void* w, float w_offset, float eps, float* h, float* rstd, int T,
                               int d, void* stream) {
  if (T <= 0 || d <= 0 || (d % 8) != 0) return set_error(LRP_ERR_ARG, "rmsnorm_fwd_residual: d must be a positive multiple of 8");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (y_is_f32)
    rmsnorm_fwd_residual_kernel<float><<<T, NORM_THREADS, 0, st>>>((const float*)y, (const bf16*)w, w_offset, eps, h, rstd, d);
  else
    rmsnorm_fwd_residual_kernel<bf16><<<T, NORM_THREADS, 0, st>>>((const bf16*)y, (const bf16*)w, w_offset, eps, h, rstd, d);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

int lrp_rmsnorm_fwd_residual(const void* y, const void* w, float w_offset, float eps, float* h, float* rstd, int T, int d,
                             void* stream) {
  return lrp_rmsnorm_fwd_residual_t(y, 0, w, w_offset, eps, h, rstd, T, d, stream);
}

int lrp_headnorm_inplace_t(void* qk, int is_f32, int64_t ld, int n_q_heads, int n_k_heads, int D, const void* wq, const void* wk,
                           float w_offset, float eps, float* rstd, int T, int backward, void* stream) {
  if (T <= 0 || n_q_heads < 0 || n_k_heads < 0 || n_q_heads + n_k_heads <= 0 || D <= 0 || (D % 8) != 0 || (ld % 8) != 0)
    return set_error(LRP_ERR_ARG, "headnorm: D and ld must be multiples of 8");
  const int64_t rows = int64_t(T) * (n_q_heads + n_k_heads);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (is_f32)
    headnorm_kernel<float><<<unsigned((rows + 7) / 8), 256, 0, st>>>((float*)qk, ld, n_q_heads, n_q_heads + n_k_heads, D, (const bf16*)wq,
                                                                    (const bf16*)wk, w_offset, eps, rstd, T, backward);
  else
    headnorm_kernel<bf16><<<unsigned((rows + 7) / 8), 256, 0, st>>>((bf16*)qk, ld, n_q_heads, n_q_heads + n_k_heads, D, (const bf16*)wq,
                                                                   (const bf16*)wk, w_offset, eps, rstd, T, backward);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

int lrp_headnorm_inplace(void* qk, int64_t ld, int n_q_heads, int n_k_heads, int D, const void* wq, const void* wk,
                         float w_offset, float eps, float* rstd, int T, int backward, void* stream) {
  return lrp_headnorm_inplace_t(qk, 0, ld, n_q_heads, n_k_heads, D, wq, wk, w_offset, eps, rstd, T, backward, stream);
}

int lrp_layernorm_fwd(const void* x, const void* w, const void* b, float eps, void* y, float* mean, float* rstd, int T,
                      int d, int is_f32, void* stream) {
  if (T <= 0 || d <= 0 || (d % 8) != 0) return set_error(LRP_ERR_ARG, "layernorm_fwd: d must be a positive multiple of 8");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (is_f32)
    layernorm_fwd_kernel<float><<<T, NORM_THREADS, 0, st>>>((const float*)x, (const float*)w, (const float*)b, eps, (float*)y, mean, rstd, d);
  else
    layernorm_fwd_kernel<bf16><<<T, NORM_THREADS, 0, st>>>((const bf16*)x, (const bf16*)w, (const bf16*)b, eps, (bf16*)y, mean, rstd, d);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

int lrp_layernorm_bwd(const void* gy, const void* w, const float* rstd, void* gx, int T, int d, int is_f32,
                      void* stream) {
  if (T <= 0 || d <= 0 || (d % 8) != 0) return set_error(LRP_ERR_ARG, "layernorm_bwd: d must be a positive multiple of 8");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (is_f32)
    layernorm_bwd_kernel<float><<<T, NORM_THREADS, 0, st>>>((const float*)gy, (const float*)w, rstd, (float*)gx, d);
  else
    layernorm_bwd_kernel<bf16><<<T, NORM_THREADS, 0, st>>>((const bf16*)gy, (const bf16*)w, rstd, (bf16*)gx, d);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

int lrp_rope_inplace_t(void* qk, int is_f32, int64_t ld, int n_heads_total, int D, const float* cos_t, const float* sin_t, int T,
                       int S, int inverse, void* stream) {
  if (T <= 0 || n_heads_total <= 0 || D <= 0 || (D % 16) != 0 || (ld % 8) != 0 || S <= 0)
    return set_error(LRP_ERR_ARG, "rope: D must be a multiple of 16 and ld a multiple of 8");
  const int64_t total = int64_t(T) * n_heads_total * (D / 16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (is_f32)
    rope_kernel<float><<<grid_for(total, 256), 256, 0, st>>>((float*)qk, ld, n_heads_total, D, cos_t, sin_t, T, S, inverse);
  else
    rope_kernel<bf16><<<grid_for(total, 256), 256, 0, st>>>((bf16*)qk, ld, n_heads_total, D, cos_t, sin_t, T, S, inverse);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

int lrp_rope_inplace(void* qk, int64_t ld, int n_heads_total, int D, const float* cos_t, const float* sin_t, int T,
                     int S, int inverse, void* stream) {
  return lrp_rope_inplace_t(qk, 0, ld, n_heads_total, D, cos_t, sin_t, T, S, inverse, stream);
}

int lrp_gated_act_fwd_t(const void* gu, void* a, int is_f32, int layout, int T, int I, int act, void* stream) {
  if (T <= 0 || I <= 0 || (I % 8) != 0) return set_error(LRP_ERR_ARG, "gated_act_fwd: I must be a positive multiple of 8");
  if (act < 0 || act > 2) return set_error(LRP_ERR_ARG, "gated_act_fwd: unknown activation");
  if (layout != 0 && (layout != 1 || (I % 32) != 0)) return set_error(LRP_ERR_ARG, "gated_act_fwd: the interleaved layout needs I % 32 == 0");
  const int64_t total = int64_t(T) * (I / 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (is_f32) gated_act_fwd_kernel<float><<<grid_for(total, 256), 256, 0, st>>>((const float*)gu, (float*)a, T, I, act, layout);
  else gated_act_fwd_kernel<bf16><<<grid_for(total, 256), 256, 0, st>>>((const bf16*)gu, (bf16*)a, T, I, act, layout);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

int lrp_gated_act_fwd(const void* gu, void* a, int T, int I, int act, void* stream) {
  return lrp_gated_act_fwd_t(gu, a, 0, 0, T, I, act, stream);
}

int lrp_gated_act_bwd_t(const void* ga, const void* gu, void* ggu, int is_f32, int layout, int T, int I, int act, int cp_variant,
                        void* stream) {
  if (T <= 0 || I <= 0 || (I % 8) != 0) return set_error(LRP_ERR_ARG, "gated_act_bwd: I must be a positive multiple of 8");
  if (act < 0 || act > 2) return set_error(LRP_ERR_ARG, "gated_act_bwd: unknown activation");
  if (layout != 0 && (layout != 1 || (I % 32) != 0)) return set_error(LRP_ERR_ARG, "gated_act_bwd: the interleaved layout needs I % 32 == 0");
  const int64_t total = int64_t(T) * (I / 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (is_f32)
    gated_act_bwd_kernel<float><<<grid_for(total, 256), 256, 0, st>>>((const float*)ga, (const float*)gu, (float*)ggu, T, I, act, cp_variant, layout);
  else
    gated_act_bwd_kernel<bf16><<<grid_for(total, 256), 256, 0, st>>>((const bf16*)ga, (const bf16*)gu, (bf16*)ggu, T, I, act, cp_variant, layout);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

int lrp_gated_act_bwd(const void* ga, const void* gu, void* ggu, int T, int I, int act, int cp_variant, void* stream) {
  return lrp_gated_act_bwd_t(ga, gu, ggu, 0, 0, T, I, act, cp_variant, stream);
}

int lrp_split_bf16x2(const float* x, void* hi, void* lo, int64_t n, void* stream) {
  if (n <= 0 || (n % 8) != 0) return set_error(LRP_ERR_ARG, "split_bf16x2: n must be a positive multiple of 8");
  split_bf16x2_kernel<<<grid_for(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, (bf16*)hi, (bf16*)lo, n / 8);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

int lrp_act_identity_fwd(const void* x, void* y, int64_t n, int act, int is_f32, void* stream) {
  if (n <= 0 || (n % 8) != 0) return set_error(LRP_ERR_ARG, "act_identity_fwd: n must be a positive multiple of 8");
  if (act < 0 || act > 2) return set_error(LRP_ERR_ARG, "act_identity_fwd: unknown activation");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (is_f32)
    act_identity_fwd_kernel<float><<<grid_for(n / 8, 256), 256, 0, st>>>((const float*)x, (float*)y, n / 8, act);
  else
    act_identity_fwd_kernel<bf16><<<grid_for(n / 8, 256), 256, 0, st>>>((const bf16*)x, (bf16*)y, n / 8, act);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

int lrp_act_identity_bwd(const void* gy, const void* x, void* gx, int64_t n, int act, int is_f32, void* stream) {
  if (n <= 0 || (n % 8) != 0) return set_error(LRP_ERR_ARG, "act_identity_bwd: n must be a positive multiple of 8");
  if (act < 0 || act > 2) return set_error(LRP_ERR_ARG, "act_identity_bwd: unknown activation");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (is_f32)
    act_identity_bwd_kernel<float><<<grid_for(n / 8, 256), 256, 0, st>>>((const float*)gy, (const float*)x, (float*)gx, n / 8, act);
  else
    act_identity_bwd_kernel<bf16><<<grid_for(n / 8, 256), 256, 0, st>>>((const bf16*)gy, (const bf16*)x, (bf16*)gx, n / 8, act);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

int lrp_embed_gather(const int64_t* ids, const void* emb, float scale, float* h, int T, int d, void* stream) {
  if (T <= 0 || d <= 0 || (d % 8) != 0) return set_error(LRP_ERR_ARG, "embed_gather: d must be a positive multiple of 8");
  embed_gather_kernel<<<T, 256, 0, static_cast<cudaStream_t>(stream)>>>(ids, (const bf16*)emb, scale, h, d);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

int lrp_gather_rows_f32(const float* src, const int64_t* rows, float* out, int n_rows, int d, void* stream) {
  if (n_rows <= 0 || d <= 0 || (d % 4) != 0) return set_error(LRP_ERR_ARG, "gather_rows: d must be a positive multiple of 4");
  gather_rows_f32_kernel<<<n_rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, rows, out, d);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

int lrp_seed_gradient(const void* lm_head, const int32_t* idx, const void* norm_w, float w_offset, const float* rstd_last, int S,
                      float* g_h, void* g_hb, int T, int d, void* stream) {
  if (T <= 0 || S <= 0 || (T % S) != 0 || d <= 0 || (d % 8) != 0)
    return set_error(LRP_ERR_ARG, "seed_gradient: T must be a multiple of S and d a multiple of 8");
  seed_gradient_kernel<<<T, 256, 0, static_cast<cudaStream_t>(stream)>>>((const bf16*)lm_head, idx, (const bf16*)norm_w, w_offset,
                                                                        rstd_last, S, g_h, (bf16*)g_hb, d);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

int lrp_argmax_rows(const float* logits, int32_t* idx, float* val, int B, int V, void* stream) {
  if (B <= 0 || V <= 0) return set_error(LRP_ERR_ARG, "argmax_rows: empty input");
  argmax_rows_kernel<<<B, 1024, 0, static_cast<cudaStream_t>(stream)>>>(logits, idx, val, V);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

int lrp_gxi_reduce(const float* x, const float* g, float* rel, int T, int d, void* stream) {
  if (T <= 0 || d <= 0 || (d % 8) != 0) return set_error(LRP_ERR_ARG, "gxi_reduce: d must be a positive multiple of 8");
  gxi_reduce_kernel<float><<<(T + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, g, rel, T, d);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

int lrp_gxi_reduce_bf16(const void* x, const void* g, float* rel, int T, int d, void* stream) {
  if (T <= 0 || d <= 0 || (d % 8) != 0) return set_error(LRP_ERR_ARG, "gxi_reduce: d must be a positive multiple of 8");
  gxi_reduce_kernel<bf16><<<(T + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>((const bf16*)x, (const bf16*)g, rel, T, d);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

int lrp_gxi_reduce_mixed(const void* x_bf16, const float* g, float* rel, int T, int d, void* stream) {
  if (T <= 0 || d <= 0 || (d % 8) != 0) return set_error(LRP_ERR_ARG, "gxi_reduce: d must be a positive multiple of 8");
  gxi_reduce_kernel<bf16, float><<<(T + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>((const bf16*)x_bf16, g, rel, T, d);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

int lrp_cast_f32_to_bf16(const float* in, void* out, int64_t n, void* stream) {
  if (n <= 0 || (n % 8) != 0) return set_error(LRP_ERR_ARG, "cast: n must be a positive multiple of 8");
  cast_f32_bf16_kernel<<<grid_for(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, (bf16*)out, n / 8);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

}  // extern "C"

template <typename F32, typename BF>
static int ew_launch(int64_t n, int is_f32, void* stream, const char* what, F32 f32, BF bf) {
  if (n <= 0) return set_error(LRP_ERR_ARG, what);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int g = grid_for(n, 256);
  if (is_f32) f32(g, st); else bf(g, st);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

extern "C" {

int lrp_eps_div(const void* r, const void* z, void* out, int64_t n, float alpha, float eps, int is_f32, void* stream) {
  return ew_launch(n, is_f32, stream, "eps_div: empty tensor",
                   [&](int g, cudaStream_t st) { eps_div_kernel<float><<<g, 256, 0, st>>>((const float*)r, (const float*)z, (float*)out, n, alpha, eps); },
                   [&](int g, cudaStream_t st) { eps_div_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)r, (const bf16*)z, (bf16*)out, n, alpha, eps); });
}
int lrp_mul(const void* a, const void* b, void* out, int64_t n, float scale, int is_f32, void* stream) {
  return ew_launch(n, is_f32, stream, "mul: empty tensor",
                   [&](int g, cudaStream_t st) { mul_kernel<float><<<g, 256, 0, st>>>((const float*)a, (const float*)b, (float*)out, n, scale); },
                   [&](int g, cudaStream_t st) { mul_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)a, (const bf16*)b, (bf16*)out, n, scale); });
}
int lrp_scale(const void* x, void* out, int64_t n, float factor, int is_f32, void* stream) {
  return ew_launch(n, is_f32, stream, "scale: empty tensor",
                   [&](int g, cudaStream_t st) { scale_kernel<float><<<g, 256, 0, st>>>((const float*)x, (float*)out, n, factor); },
                   [&](int g, cudaStream_t st) { scale_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)x, (bf16*)out, n, factor); });
}
int lrp_gamma_split(const void* x, void* xcat, int64_t rows, int K, int is_f32, void* stream) {
  if (K <= 0) return set_error(LRP_ERR_ARG, "gamma_split: bad width");
  return ew_launch(rows * K, is_f32, stream, "gamma_split: empty tensor",
                   [&](int g, cudaStream_t st) { gamma_split_kernel<float><<<g, 256, 0, st>>>((const float*)x, (float*)xcat, rows, K); },
                   [&](int g, cudaStream_t st) { gamma_split_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)x, (bf16*)xcat, rows, K); });
}
int lrp_gamma_s(const void* g, const void* y, const void* zp, const void* zn, void* scat, int64_t rows, int N, float eps, int is_f32,
                void* stream) {
  if (N <= 0) return set_error(LRP_ERR_ARG, "gamma_s: bad width");
  return ew_launch(rows * N, is_f32, stream, "gamma_s: empty tensor",
                   [&](int gr, cudaStream_t st) { gamma_s_kernel<float><<<gr, 256, 0, st>>>((const float*)g, (const float*)y, (const float*)zp, (const float*)zn, (float*)scat, rows, N, eps); },
                   [&](int gr, cudaStream_t st) { gamma_s_kernel<bf16><<<gr, 256, 0, st>>>((const bf16*)g, (const bf16*)y, (const bf16*)zp, (const bf16*)zn, (bf16*)scat, rows, N, eps); });
}
int lrp_gamma_combine(const void* x, const void* g1, const void* g2, void* out, int64_t n, int is_f32, void* stream) {
  return ew_launch(n, is_f32, stream, "gamma_combine: empty tensor",
                   [&](int g, cudaStream_t st) { gamma_combine_kernel<float><<<g, 256, 0, st>>>((const float*)x, (const float*)g1, (const float*)g2, (float*)out, n); },
                   [&](int g, cudaStream_t st) { gamma_combine_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)x, (const bf16*)g1, (const bf16*)g2, (bf16*)out, n); });
}
int lrp_identity_rule_bwd(const void* gy, const void* x, const void* y, void* gx, int64_t n, int is_f32, void* stream) {
  return ew_launch(n, is_f32, stream, "identity_rule_bwd: empty tensor",
                   [&](int g, cudaStream_t st) { identity_rule_bwd_kernel<float><<<g, 256, 0, st>>>((const float*)gy, (const float*)x, (const float*)y, (float*)gx, n); },
                   [&](int g, cudaStream_t st) { identity_rule_bwd_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)gy, (const bf16*)x, (const bf16*)y, (bf16*)gx, n); });
}
int lrp_softmax_dt_bwd(const void* x, const void* p, const void* r, void* out, int64_t rows, int cols, int is_f32, void* stream) {
  if (rows <= 0 || cols <= 0) return set_error(LRP_ERR_ARG, "softmax_dt_bwd: empty tensor");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned g = unsigned((rows + 7) / 8);
  if (is_f32) softmax_dt_bwd_kernel<float><<<g, 256, 0, st>>>((const float*)x, (const float*)p, (const float*)r, (float*)out, rows, cols);
  else softmax_dt_bwd_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)x, (const bf16*)p, (const bf16*)r, (bf16*)out, rows, cols);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}
int lrp_softmax_fwd(const void* x, void* out, int64_t rows, int cols, float temperature, int is_f32, void* stream) {
  if (rows <= 0 || cols <= 0 || !(temperature != 0.f)) return set_error(LRP_ERR_ARG, "softmax_fwd: empty tensor or zero temperature");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned g = unsigned((rows + 7) / 8);
  if (is_f32) softmax_fwd_kernel<float><<<g, 256, 0, st>>>((const float*)x, (float*)out, rows, cols, 1.f / temperature);
  else softmax_fwd_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)x, (bf16*)out, rows, cols, 1.f / temperature);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}
int lrp_add2_bwd(const void* a, const void* b, const void* r, void* ra, void* rb, int64_t n, float eps, int is_f32, void* stream) {
  return ew_launch(n, is_f32, stream, "add2_bwd: empty tensor",
                   [&](int g, cudaStream_t st) { add2_bwd_kernel<float><<<g, 256, 0, st>>>((const float*)a, (const float*)b, (const float*)r, (float*)ra, (float*)rb, n, eps); },
                   [&](int g, cudaStream_t st) { add2_bwd_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)a, (const bf16*)b, (const bf16*)r, (bf16*)ra, (bf16*)rb, n, eps); });
}

}  // extern "C"
