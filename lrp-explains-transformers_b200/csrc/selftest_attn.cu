// Native self-test + micro-benchmark of the flash AttnLRP kernels through the C ABI.
// Reference: naive fp32 soft-max attention + analytic backward on the same bf16 inputs.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../include/lrp_b200.h"

#define CK(x)                                                                       \
  do {                                                                              \
    cudaError_t e = (x);                                                            \
    if (e != cudaSuccess) {                                                         \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                      \
    }                                                                               \
  } while (0)

typedef __nv_bfloat16 bf16;

__global__ void fill_bf16(bf16* p, size_t n, uint32_t seed, float scale) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  uint32_t x = uint32_t(i) * 2654435761u + seed;
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  float u = (x & 0xffffff) / float(0x1000000) - 0.5f;
  p[i] = __float2bfloat16(u * scale);
}

__device__ bool masked(int qi, int kj, int causal, int window) {
  if (causal && kj > qi) return true;
  if (window > 0 && qi - kj >= window) return true;
  return false;
}

// one thread per (b, h, q): forward + per-row backward contributions
__global__ void ref_attn(const bf16* q, const bf16* k, const bf16* v, const bf16* d_o, int64_t ldq, int64_t ldk,
                         int64_t ldv, int B, int S, int H, int Hkv, int D, float scale, int causal, int window,
                         float* o_ref, float* lse_ref, float* dq_ref, float* dk_ref, float* dv_ref, int do_bwd) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H * S) return;
  int qi = idx % S, h = (idx / S) % H, b = idx / (S * H);
  int hk = h / (H / Hkv);
  const bf16* qr = q + (int64_t(b) * S + qi) * ldq + h * D;
  float m = -INFINITY;
  for (int j = 0; j < S; ++j) {
    if (masked(qi, j, causal, window)) continue;
    const bf16* kr = k + (int64_t(b) * S + j) * ldk + hk * D;
    float s = 0;
    for (int d = 0; d < D; ++d) s += __bfloat162float(qr[d]) * __bfloat162float(kr[d]);
    m = fmaxf(m, s * scale);
  }
  float l = 0;
  float oacc[256];
  for (int d = 0; d < D; ++d) oacc[d] = 0;
  for (int j = 0; j < S; ++j) {
    if (masked(qi, j, causal, window)) continue;
    const bf16* kr = k + (int64_t(b) * S + j) * ldk + hk * D;
    const bf16* vr = v + (int64_t(b) * S + j) * ldv + hk * D;
    float s = 0;
    for (int d = 0; d < D; ++d) s += __bfloat162float(qr[d]) * __bfloat162float(kr[d]);
    float p = expf(s * scale - m);
    l += p;
    for (int d = 0; d < D; ++d) oacc[d] += p * __bfloat162float(vr[d]);
  }
  float* orow = o_ref + ((int64_t(b) * S + qi) * H + h) * D;
  for (int d = 0; d < D; ++d) orow[d] = oacc[d] / l;
  float lse = m + logf(l);
  lse_ref[(int64_t(b) * H + h) * S + qi] = lse;
  if (!do_bwd) return;
  const bf16* dor = d_o + ((int64_t(b) * S + qi) * H + h) * D;
  float delta = 0;
  for (int d = 0; d < D; ++d) delta += __bfloat162float(dor[d]) * __bfloat162float(__float2bfloat16(orow[d]));
  float* dqr = dq_ref + ((int64_t(b) * S + qi) * H + h) * D;
  for (int j = 0; j < S; ++j) {
    if (masked(qi, j, causal, window)) continue;
    const bf16* kr = k + (int64_t(b) * S + j) * ldk + hk * D;
    const bf16* vr = v + (int64_t(b) * S + j) * ldv + hk * D;
    float s = 0, dp = 0;
    for (int d = 0; d < D; ++d) {
      s += __bfloat162float(qr[d]) * __bfloat162float(kr[d]);
      dp += __bfloat162float(dor[d]) * __bfloat162float(vr[d]);
    }
    float p = expf(s * scale - lse);
    float ds = p * (dp - delta) * scale;
    float* dkr = dk_ref + ((int64_t(b) * S + j) * Hkv + hk) * D;
    float* dvr = dv_ref + ((int64_t(b) * S + j) * Hkv + hk) * D;
    for (int d = 0; d < D; ++d) {
      dqr[d] += ds * __bfloat162float(kr[d]);
      atomicAdd(&dkr[d], ds * __bfloat162float(qr[d]));
      atomicAdd(&dvr[d], p * __bfloat162float(dor[d]));
    }
  }
}

static int g_fail = 0;

static double rel_l2(const std::vector<float>& got, const std::vector<float>& ref) {
  double num = 0, den = 0;
  for (size_t i = 0; i < got.size(); ++i) {
    double d = double(got[i]) - ref[i];
    if (!(got[i] == got[i])) d = 1e30;
    num += d * d;
    den += double(ref[i]) * ref[i];
  }
  return sqrt(num / (den + 1e-30));
}

static std::vector<float> fetch_bf16_strided(const bf16* dev, int64_t rows, int width, int64_t ld) {
  std::vector<bf16> h(size_t(rows - 1) * ld + width);
  CK(cudaMemcpy(h.data(), dev, h.size() * 2, cudaMemcpyDeviceToHost));
  std::vector<float> out(size_t(rows) * width);
  for (int64_t r = 0; r < rows; ++r)
    for (int c = 0; c < width; ++c) out[r * width + c] = __bfloat162float(h[r * ld + c]);
  return out;
}

static void run_case(int B, int S, int H, int Hkv, int D, int causal, int window, bool packed) {
  const int64_t T = int64_t(B) * S;
  const int wq = H * D, wk = Hkv * D;
  const int64_t ld = packed ? (wq + 2 * wk) : 0;
  bf16 *qkv, *q, *k, *v, *o, *d_o, *dqkv, *dq, *dk, *dv;
  int64_t ldq, ldk, ldv;
  if (packed) {
    CK(cudaMalloc(&qkv, T * ld * 2));
    CK(cudaMalloc(&dqkv, T * ld * 2));
    CK(cudaMemset(dqkv, 0xff, T * ld * 2));
    fill_bf16<<<(T * ld + 255) / 256, 256>>>(qkv, T * ld, 3u, 2.f);
    q = qkv; k = qkv + wq; v = qkv + wq + wk; ldq = ldk = ldv = ld;
    dq = dqkv; dk = dqkv + wq; dv = dqkv + wq + wk;
  } else {
    CK(cudaMalloc(&q, T * wq * 2)); CK(cudaMalloc(&k, T * wk * 2)); CK(cudaMalloc(&v, T * wk * 2));
    CK(cudaMalloc(&dq, T * wq * 2)); CK(cudaMalloc(&dk, T * wk * 2)); CK(cudaMalloc(&dv, T * wk * 2));
    fill_bf16<<<(T * wq + 255) / 256, 256>>>(q, T * wq, 3u, 2.f);
    fill_bf16<<<(T * wk + 255) / 256, 256>>>(k, T * wk, 5u, 2.f);
    fill_bf16<<<(T * wk + 255) / 256, 256>>>(v, T * wk, 7u, 2.f);
    ldq = wq; ldk = ldv = wk;
  }
  CK(cudaMalloc(&o, T * wq * 2));
  CK(cudaMalloc(&d_o, T * wq * 2));
  fill_bf16<<<(T * wq + 255) / 256, 256>>>(d_o, T * wq, 11u, 1.f);
  float *lse, *dq_acc, *delta, *o_ref, *lse_ref, *dq_ref, *dk_ref, *dv_ref;
  CK(cudaMalloc(&lse, size_t(B) * H * S * 4)); CK(cudaMalloc(&delta, size_t(B) * H * S * 4));
  { int64_t nb = 0; lrp_attn_bwd_workspace_bytes(B, S, H, D, &nb, nullptr); CK(cudaMalloc(&dq_acc, nb)); }
  CK(cudaMalloc(&o_ref, T * wq * 4)); CK(cudaMalloc(&lse_ref, size_t(B) * H * S * 4));
  CK(cudaMalloc(&dq_ref, T * wq * 4)); CK(cudaMalloc(&dk_ref, T * wk * 4)); CK(cudaMalloc(&dv_ref, T * wk * 4));
  CK(cudaMemset(dq_ref, 0, T * wq * 4)); CK(cudaMemset(dk_ref, 0, T * wk * 4)); CK(cudaMemset(dv_ref, 0, T * wk * 4));
  const float scale = 1.f / sqrtf(float(D));
  int rc = lrp_attn_fwd(q, k, v, ldq, ldk, ldv, o, lse, B, S, H, Hkv, D, scale, causal, window, 0);
  if (rc) { printf("FAIL fwd rc=%d %s\n", rc, lrp_last_error()); g_fail++; return; }
  cudaError_t ce = cudaDeviceSynchronize();
  if (ce != cudaSuccess) { printf("FAIL fwd exec: %s\n", cudaGetErrorString(ce)); exit(3); }
  rc = lrp_attn_bwd(q, k, v, ldq, ldk, ldv, o, d_o, lse, dq, dk, dv, ldq, ldk, ldv, dq_acc, delta, B, S, H, Hkv, D, scale,
                    causal, window, 4.f, 4.f, 2.f, 0);
  if (rc) { printf("FAIL bwd rc=%d %s\n", rc, lrp_last_error()); g_fail++; return; }
  ce = cudaDeviceSynchronize();
  if (ce != cudaSuccess) { printf("FAIL bwd exec: %s\n", cudaGetErrorString(ce)); exit(3); }
  ref_attn<<<(B * H * S + 63) / 64, 64>>>(q, k, v, d_o, ldq, ldk, ldv, B, S, H, Hkv, D, scale, causal, window, o_ref,
                                          lse_ref, dq_ref, dk_ref, dv_ref, 1);
  CK(cudaDeviceSynchronize());
  std::vector<float> r_o(T * wq), r_lse(size_t(B) * H * S), r_dq(T * wq), r_dk(T * wk), r_dv(T * wk), g_lse(size_t(B) * H * S);
  CK(cudaMemcpy(r_o.data(), o_ref, r_o.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(r_lse.data(), lse_ref, r_lse.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(g_lse.data(), lse, g_lse.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(r_dq.data(), dq_ref, r_dq.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(r_dk.data(), dk_ref, r_dk.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(r_dv.data(), dv_ref, r_dv.size() * 4, cudaMemcpyDeviceToHost));
  for (auto& x : r_dq) x *= 0.25f;
  for (auto& x : r_dk) x *= 0.25f;
  for (auto& x : r_dv) x *= 0.5f;
  auto g_o = fetch_bf16_strided(o, T, wq, wq);
  auto g_dq = fetch_bf16_strided(dq, T, wq, ldq);
  auto g_dk = fetch_bf16_strided(dk, T, wk, ldk);
  auto g_dv = fetch_bf16_strided(dv, T, wk, ldv);
  double eo = rel_l2(g_o, r_o), el = rel_l2(g_lse, r_lse), eq = rel_l2(g_dq, r_dq), ek = rel_l2(g_dk, r_dk),
         ev = rel_l2(g_dv, r_dv);
  bool ok = eo < 8e-3 && el < 1e-4 && eq < 1.5e-2 && ek < 1.5e-2 && ev < 1.5e-2;
  printf("%s B=%d S=%d H=%d Hkv=%d D=%d causal=%d window=%d packed=%d  o=%.2e lse=%.2e dq=%.2e dk=%.2e dv=%.2e\n",
         ok ? "ok  " : "FAIL", B, S, H, Hkv, D, causal, window, int(packed), eo, el, eq, ek, ev);
  if (!ok) g_fail++;
  if (packed) { cudaFree(qkv); cudaFree(dqkv); } else { cudaFree(q); cudaFree(k); cudaFree(v); cudaFree(dq); cudaFree(dk); cudaFree(dv); }
  cudaFree(o); cudaFree(d_o); cudaFree(lse); cudaFree(delta); cudaFree(dq_acc); cudaFree(o_ref); cudaFree(lse_ref);
  cudaFree(dq_ref); cudaFree(dk_ref); cudaFree(dv_ref);
}

static void perf_case(int B, int S, int H, int Hkv, int D) {
  const int64_t T = int64_t(B) * S;
  const int64_t ld = int64_t(H + 2 * Hkv) * D;
  bf16 *qkv, *dqkv, *o, *d_o;
  float *lse, *delta, *dq_acc;
  CK(cudaMalloc(&qkv, T * ld * 2)); CK(cudaMalloc(&dqkv, T * ld * 2));
  CK(cudaMalloc(&o, T * H * D * 2)); CK(cudaMalloc(&d_o, T * H * D * 2));
  CK(cudaMalloc(&lse, size_t(B) * H * S * 4)); CK(cudaMalloc(&delta, size_t(B) * H * S * 4));
  { int64_t nb = 0; lrp_attn_bwd_workspace_bytes(B, S, H, D, &nb, nullptr); CK(cudaMalloc(&dq_acc, nb)); }
  fill_bf16<<<(T * ld + 255) / 256, 256>>>(qkv, T * ld, 3u, 2.f);
  fill_bf16<<<(T * H * D + 255) / 256, 256>>>(d_o, T * H * D, 11u, 1.f);
  const float scale = 1.f / sqrtf(float(D));
  bf16 *q = qkv, *k = qkv + H * D, *v = qkv + (H + Hkv) * D;
  cudaEvent_t e0, e1, e2;
  cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  // warm-up (also leaves the dQ workspace zero: every backward clears it behind itself), then 5 timed calls each
  const int REPS = 5;
  auto bwd = [&](int flags) {
    lrp_attn_bwd_varlen(q, k, v, ld, ld, ld, o, d_o, lse, dqkv, dqkv + H * D, dqkv + (H + Hkv) * D, ld, ld, ld, dq_acc, delta, nullptr,
                        flags, B, S, H, Hkv, D, scale, 1, 0, 4.f, 4.f, 2.f, 0);
  };
  lrp_attn_fwd(q, k, v, ld, ld, ld, o, lse, B, S, H, Hkv, D, scale, 1, 0, 0);
  bwd(0);
  CK(cudaDeviceSynchronize());
  float f = 0, bw = 0, bws = 0;
  cudaEventRecord(e0);
  for (int rep = 0; rep < REPS; ++rep) lrp_attn_fwd(q, k, v, ld, ld, ld, o, lse, B, S, H, Hkv, D, scale, 1, 0, 0);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  cudaEventElapsedTime(&f, e0, e1);
  cudaEventRecord(e0);
  for (int rep = 0; rep < REPS; ++rep) bwd(0);                    // workspace content unknown: zero-filled by the delta kernel
  cudaEventRecord(e1);
  for (int rep = 0; rep < REPS; ++rep) bwd(LRP_ATTN_ACC_ZERO);    // steady state of a caller that keeps the workspace (engine, ops cache)
  cudaEventRecord(e2);
  CK(cudaDeviceSynchronize());
  cudaEventElapsedTime(&bw, e0, e1);
  cudaEventElapsedTime(&bws, e1, e2);
  f /= REPS; bw /= REPS; bws /= REPS;
  const double flops_f = 4.0 * B * H * double(S) * S * D / 2;
  printf("perf B=%d S=%d H=%d Hkv=%d D=%d causal: fwd %.3f ms (%.0f TFLOP/s)  bwd %.3f ms (%.0f TFLOP/s)  bwd, kept workspace %.3f ms (%.0f TFLOP/s)\n",
         B, S, H, Hkv, D, f, flops_f / f * 1e-9, bw, 2.5 * flops_f / bw * 1e-9, bws, 2.5 * flops_f / bws * 1e-9);
}

int main(int argc, char** argv) {
  bool perf = argc > 1 && !strcmp(argv[1], "--perf");
  if (lrp_check_device() != 0) { printf("no device: %s\n", lrp_last_error()); return 1; }
  run_case(1, 128, 1, 1, 128, 0, 0, false);
  run_case(1, 128, 1, 1, 128, 1, 0, false);
  run_case(1, 256, 2, 1, 128, 1, 0, false);
  run_case(2, 300, 4, 2, 128, 1, 0, true);
  run_case(1, 384, 2, 2, 128, 1, 0, false);     // odd number of query tiles: the second tile of the last pair is absent
  run_case(2, 1000, 8, 2, 128, 1, 0, true);     // ragged tail, several work items per CTA
  run_case(1, 2048, 40, 8, 128, 1, 0, true);    // more work items than SMs: persistent loop, phases across items
  run_case(2, 333, 4, 4, 128, 0, 0, false);     // full (non-causal) attention
  run_case(2, 300, 4, 2, 64, 1, 0, true);
  run_case(2, 197, 4, 4, 64, 0, 0, false);
  run_case(1, 520, 4, 1, 128, 1, 200, true);
  run_case(1, 640, 2, 2, 64, 1, 130, false);
  run_case(1, 128, 1, 1, 256, 1, 0, false);
  run_case(2, 300, 4, 2, 256, 1, 0, true);
  run_case(1, 520, 2, 1, 256, 1, 200, true);
  run_case(2, 197, 2, 2, 256, 0, 0, false);
  printf(g_fail ? "SELFTEST FAILED (%d)\n" : "SELFTEST PASSED\n", g_fail);
  if (perf) {
    perf_case(4, 2048, 32, 8, 128);
    perf_case(8, 2048, 32, 8, 128);
    perf_case(1, 512, 32, 4, 64);
    perf_case(1, 8192, 8, 4, 256);
  }
  return g_fail ? 1 : 0;
}
