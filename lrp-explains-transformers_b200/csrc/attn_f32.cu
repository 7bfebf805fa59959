// Validation-precision flash AttnLRP: fp32 inputs, fp32 CUDA-core arithmetic, no tensor cores, no [B,H,S,S] tensor.
//
// Purpose: the validation mode of the engine / drop-in path (every activation fp32, GEMMs as two-term bf16 splits on the
// tcgen05 kernels) needs an attention step whose error is fp32 rounding only, so that a model-level result can be held
// against the reference's fp32 run at <= 1e-3 (the bf16 tcgen05 kernels of attn_fwd_ws.cu / attn_lrp.cu round P and dS
// to bf16 once, which alone is 1.6e-3).  Same rule as the production kernels (reference lxt/efficient/patches.py:193-203):
// plain soft-max attention forward; backward = ordinary attention backward with dQ/q_div, dK/k_div, dV/v_div (0 = detached).
// Not a performance kernel: one warp per query row (forward, dQ) / per key row (dK, dV), lanes split head_dim.
#include <math.h>
#include "ptx_sm100.cuh"
#include "lrp_internal.h"

namespace lrp {

namespace {
constexpr int F32_WARPS = 8;

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ bool visible(int qpos, int kpos, int causal, int window) {
  if (causal && kpos > qpos) return false;
  if (window > 0 && qpos - kpos >= window) return false;
  return true;
}

// E = head_dim / 32 elements per lane
template <int E>
__global__ void __launch_bounds__(F32_WARPS * 32)
attn_f32_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, int64_t ldq,
                    int64_t ldk, int64_t ldv, float* __restrict__ o, float* __restrict__ lse, int B, int S, int H, int Hkv,
                    float scale, int causal, int window, const int* __restrict__ kv_range) {
  constexpr int D = E * 32;
  const int64_t row = blockIdx.x * int64_t(F32_WARPS) + (threadIdx.x >> 5);
  if (row >= int64_t(B) * H * S) return;
  const int lane = threadIdx.x & 31;
  const int qi = int(row % S), h = int((row / S) % H), b = int(row / (int64_t(S) * H));
  const int hk = h / (H / Hkv);
  const float* qr = q + (int64_t(b) * S + qi) * ldq + h * D;
  float qv[E], acc[E];
#pragma unroll
  for (int e = 0; e < E; ++e) { qv[e] = qr[lane + 32 * e]; acc[e] = 0.f; }
  const int kvlo = kv_range ? max(0, kv_range[2 * b]) : 0, kvhi = kv_range ? min(S, kv_range[2 * b + 1]) : S;
  const int j_lo = max(kvlo, window > 0 ? max(0, qi - window + 1) : 0);
  const int j_hi = min(kvhi - 1, causal ? qi : S - 1);
  float m = -INFINITY, l = 0.f;
  for (int j = j_lo; j <= j_hi; ++j) {
    const float* kr = k + (int64_t(b) * S + j) * ldk + hk * D;
    const float* vr = v + (int64_t(b) * S + j) * ldv + hk * D;
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) s = fmaf(qv[e], kr[lane + 32 * e], s);
    s = wsum(s) * scale;
    const float m_new = fmaxf(m, s);
    const float corr = expf(m - m_new);   // exp(-inf) = 0 on the first key
    const float p = expf(s - m_new);
    l = l * corr + p;
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = fmaf(p, vr[lane + 32 * e], acc[e] * corr);
    m = m_new;
  }
  float* orow = o + ((int64_t(b) * S + qi) * H + h) * D;
  const float inv = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e) orow[lane + 32 * e] = acc[e] * inv;
  if (lane == 0) lse[(int64_t(b) * H + h) * S + qi] = l > 0.f ? m + logf(l) : -INFINITY;
}

// dQ (one warp per query row) and delta = sum(o * dO) saved for the key-major pass
template <int E>
__global__ void __launch_bounds__(F32_WARPS * 32)
attn_f32_dq_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, int64_t ldq,
                   int64_t ldk, int64_t ldv, const float* __restrict__ o, const float* __restrict__ d_o,
                   const float* __restrict__ lse, float* __restrict__ delta, float* __restrict__ dq, int64_t lddq, int B, int S,
                   int H, int Hkv, float scale, int causal, int window, float inv_q_div, const int* __restrict__ kv_range) {
  constexpr int D = E * 32;
  const int64_t row = blockIdx.x * int64_t(F32_WARPS) + (threadIdx.x >> 5);
  if (row >= int64_t(B) * H * S) return;
  const int lane = threadIdx.x & 31;
  const int qi = int(row % S), h = int((row / S) % H), b = int(row / (int64_t(S) * H));
  const int hk = h / (H / Hkv);
  const float* qr = q + (int64_t(b) * S + qi) * ldq + h * D;
  const float* orow = o + ((int64_t(b) * S + qi) * H + h) * D;
  const float* dor = d_o + ((int64_t(b) * S + qi) * H + h) * D;
  float qv[E], dov[E], acc[E];
  float dl = 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    qv[e] = qr[lane + 32 * e];
    dov[e] = dor[lane + 32 * e];
    dl = fmaf(orow[lane + 32 * e], dov[e], dl);
    acc[e] = 0.f;
  }
  dl = wsum(dl);
  const float ls = lse[(int64_t(b) * H + h) * S + qi];
  if (lane == 0) delta[(int64_t(b) * H + h) * S + qi] = dl;
  const int kvlo = kv_range ? max(0, kv_range[2 * b]) : 0, kvhi = kv_range ? min(S, kv_range[2 * b + 1]) : S;
  const int j_lo = max(kvlo, window > 0 ? max(0, qi - window + 1) : 0);
  const int j_hi = min(kvhi - 1, causal ? qi : S - 1);
  if (ls != -INFINITY) {
    for (int j = j_lo; j <= j_hi; ++j) {
      const float* kr = k + (int64_t(b) * S + j) * ldk + hk * D;
      const float* vr = v + (int64_t(b) * S + j) * ldv + hk * D;
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        s = fmaf(qv[e], kr[lane + 32 * e], s);
        dp = fmaf(dov[e], vr[lane + 32 * e], dp);
      }
      s = wsum(s) * scale;
      dp = wsum(dp);
      const float ds = expf(s - ls) * (dp - dl) * scale;
#pragma unroll
      for (int e = 0; e < E; ++e) acc[e] = fmaf(ds, kr[lane + 32 * e], acc[e]);
    }
  }
  float* dqr = dq + (int64_t(b) * S + qi) * lddq + h * D;
#pragma unroll
  for (int e = 0; e < E; ++e) dqr[lane + 32 * e] = acc[e] * inv_q_div;
}

// dK, dV: one warp per key row (b, hk, j); loops over the G query heads of the group and the query rows that see key j
template <int E>
__global__ void __launch_bounds__(F32_WARPS * 32)
attn_f32_dkv_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, int64_t ldq,
                    int64_t ldk, int64_t ldv, const float* __restrict__ d_o, const float* __restrict__ lse,
                    const float* __restrict__ delta, float* __restrict__ dk, float* __restrict__ dv, int64_t lddk, int64_t lddv,
                    int B, int S, int H, int Hkv, float scale, int causal, int window, float inv_k_div, float inv_v_div,
                    const int* __restrict__ kv_range) {
  constexpr int D = E * 32;
  const int64_t row = blockIdx.x * int64_t(F32_WARPS) + (threadIdx.x >> 5);
  if (row >= int64_t(B) * Hkv * S) return;
  const int lane = threadIdx.x & 31;
  const int j = int(row % S), hk = int((row / S) % Hkv), b = int(row / (int64_t(S) * Hkv));
  const int G = H / Hkv;
  const float* kr = k + (int64_t(b) * S + j) * ldk + hk * D;
  const float* vr = v + (int64_t(b) * S + j) * ldv + hk * D;
  float kv[E], vv[E], ak[E], av[E];
#pragma unroll
  for (int e = 0; e < E; ++e) { kv[e] = kr[lane + 32 * e]; vv[e] = vr[lane + 32 * e]; ak[e] = 0.f; av[e] = 0.f; }
  const int kvlo = kv_range ? max(0, kv_range[2 * b]) : 0, kvhi = kv_range ? min(S, kv_range[2 * b + 1]) : S;
  const int i_lo = causal ? j : 0;
  const int i_hi = (j < kvlo || j >= kvhi) ? -1 : (window > 0 ? min(S - 1, j + window - 1) : S - 1);   // a padded key is attended by nobody
  for (int g = 0; g < G; ++g) {
    const int h = hk * G + g;
    for (int i = i_lo; i <= i_hi; ++i) {
      if (!visible(i, j, causal, window)) continue;
      const float ls = lse[(int64_t(b) * H + h) * S + i];
      if (ls == -INFINITY) continue;
      const float* qr = q + (int64_t(b) * S + i) * ldq + h * D;
      const float* dor = d_o + ((int64_t(b) * S + i) * H + h) * D;
      float qv[E], dov[E];
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        qv[e] = qr[lane + 32 * e];
        dov[e] = dor[lane + 32 * e];
        s = fmaf(qv[e], kv[e], s);
        dp = fmaf(dov[e], vv[e], dp);
      }
      s = wsum(s) * scale;
      dp = wsum(dp);
      const float p = expf(s - ls);
      const float ds = p * (dp - delta[(int64_t(b) * H + h) * S + i]) * scale;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        av[e] = fmaf(p, dov[e], av[e]);
        ak[e] = fmaf(ds, qv[e], ak[e]);
      }
    }
  }
  float* dkr = dk + (int64_t(b) * S + j) * lddk + hk * D;
  float* dvr = dv + (int64_t(b) * S + j) * lddv + hk * D;
#pragma unroll
  for (int e = 0; e < E; ++e) { dkr[lane + 32 * e] = ak[e] * inv_k_div; dvr[lane + 32 * e] = av[e] * inv_v_div; }
}

static int check_f32(int B, int S, int H, int Hkv, int D) {
  if (B <= 0 || S <= 0 || H <= 0 || Hkv <= 0 || H % Hkv != 0) return set_error(LRP_ERR_ARG, "attn_f32: bad shape");
  if (D != 32 && D != 64 && D != 128 && D != 256) return set_error(LRP_ERR_ARG, "attn_f32: head_dim must be 32, 64, 128 or 256");
  return LRP_OK;
}
}  // namespace

}  // namespace lrp

using namespace lrp;

extern "C" {

int lrp_attn_fwd_f32(const float* q, const float* k, const float* v, int64_t ldq, int64_t ldk, int64_t ldv, float* o, float* lse,
                     const int32_t* kv_range, int B, int S, int H, int Hkv, int D, float scale, int causal, int window, void* stream) {
  if (int e = check_f32(B, S, H, Hkv, D)) return e;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t rows = int64_t(B) * H * S;
  const unsigned grid = unsigned((rows + F32_WARPS - 1) / F32_WARPS);
#define LRP_F32_FWD(E) attn_f32_fwd_kernel<E><<<grid, F32_WARPS * 32, 0, st>>>(q, k, v, ldq, ldk, ldv, o, lse, B, S, H, Hkv, scale, causal, window, kv_range)
  if (D == 32) LRP_F32_FWD(1); else if (D == 64) LRP_F32_FWD(2); else if (D == 128) LRP_F32_FWD(4); else LRP_F32_FWD(8);
#undef LRP_F32_FWD
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

int lrp_attn_bwd_f32(const float* q, const float* k, const float* v, int64_t ldq, int64_t ldk, int64_t ldv, const float* o,
                     const float* d_o, const float* lse, float* dq, float* dk, float* dv, int64_t lddq, int64_t lddk,
                     int64_t lddv, float* delta_ws, const int32_t* kv_range, int B, int S, int H, int Hkv, int D, float scale, int causal,
                     int window, float q_div, float k_div, float v_div, void* stream) {
  if (int e = check_f32(B, S, H, Hkv, D)) return e;
  if (delta_ws == nullptr) return set_error(LRP_ERR_ARG, "attn_bwd_f32: missing delta workspace");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float iq = q_div > 0.f ? 1.f / q_div : 0.f, ik = k_div > 0.f ? 1.f / k_div : 0.f, iv = v_div > 0.f ? 1.f / v_div : 0.f;
  const int64_t rows = int64_t(B) * H * S, krows = int64_t(B) * Hkv * S;
  const unsigned g1 = unsigned((rows + F32_WARPS - 1) / F32_WARPS), g2 = unsigned((krows + F32_WARPS - 1) / F32_WARPS);
#define LRP_F32_BWD(E)                                                                                                          \
  do {                                                                                                                          \
    attn_f32_dq_kernel<E><<<g1, F32_WARPS * 32, 0, st>>>(q, k, v, ldq, ldk, ldv, o, d_o, lse, delta_ws, dq, lddq, B, S, H, Hkv,  \
                                                          scale, causal, window, iq, kv_range);                                         \
    attn_f32_dkv_kernel<E><<<g2, F32_WARPS * 32, 0, st>>>(q, k, v, ldq, ldk, ldv, d_o, lse, delta_ws, dk, dv, lddk, lddv, B, S,  \
                                                           H, Hkv, scale, causal, window, ik, iv, kv_range);                    \
  } while (0)
  if (D == 32) LRP_F32_BWD(1); else if (D == 64) LRP_F32_BWD(2); else if (D == 128) LRP_F32_BWD(4); else LRP_F32_BWD(8);
#undef LRP_F32_BWD
  LRP_CHECK_LAUNCH();
  note_launch();
  return LRP_OK;
}

}  // extern "C"
