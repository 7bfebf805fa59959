// Native self-test + micro-benchmark of the tcgen05 GEMM through the C ABI (no Python, no torch).
// Correctness: against a naive fp32-accumulate CUDA kernel on the same bf16 inputs.
// Usage: selftest_gemm [--perf]
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

__global__ void fill_bf16(__nv_bfloat16* p, size_t n, uint32_t seed, float scale, float offset = 0.f) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  uint32_t x = uint32_t(i) * 2654435761u + seed;
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  float u = (x & 0xffffff) / float(0x1000000) - 0.5f;
  p[i] = __float2bfloat16(u * scale + offset);
}
__global__ void fill_f32(float* p, size_t n, uint32_t seed, float scale, float offset) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  uint32_t x = uint32_t(i) * 2654435761u + seed;
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  float u = (x & 0xffffff) / float(0x1000000) - 0.5f;
  p[i] = u * scale + offset;
}

// ref[m,n] = resid + alpha*acc*rs[m]*cs[n] + bias[n]
__global__ void ref_gemm(const __nv_bfloat16* A, const __nv_bfloat16* B, int b_layout, int M, int N, int K,
                         const float* resid, const float* rs, const float* cs, const float* bias, float alpha,
                         float* out) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    float a = __bfloat162float(A[size_t(m) * K + k]);
    float b = b_layout == 0 ? __bfloat162float(B[size_t(n) * K + k]) : __bfloat162float(B[size_t(k) * N + n]);
    acc += a * b;
  }
  float v = acc * alpha;
  if (rs) v *= rs[m];
  if (cs) v *= cs[n];
  if (bias) v += bias[n];
  if (resid) v += resid[size_t(m) * N + n];
  out[size_t(m) * N + n] = v;
}

static int g_fail = 0;

static void check_case(int M, int N, int K, int b_layout, int tile_n, int mode) {
  __nv_bfloat16 *A, *B, *Cb, *Sh;
  float *Cf, *Ref, *Res, *rs, *cs, *bias;
  CK(cudaMalloc(&A, size_t(M) * K * 2));
  CK(cudaMalloc(&B, size_t(N) * K * 2));
  CK(cudaMalloc(&Cb, size_t(M) * N * 2));
  CK(cudaMalloc(&Sh, size_t(M) * N * 2));
  CK(cudaMalloc(&Cf, size_t(M) * N * 4));
  CK(cudaMalloc(&Ref, size_t(M) * N * 4));
  CK(cudaMalloc(&Res, size_t(M) * N * 4));
  CK(cudaMalloc(&rs, size_t(M) * 4));
  CK(cudaMalloc(&cs, size_t(N) * 4));
  CK(cudaMalloc(&bias, size_t(N) * 4));
  fill_bf16<<<(size_t(M) * K + 255) / 256, 256>>>(A, size_t(M) * K, 1u, 2.f);
  fill_bf16<<<(size_t(N) * K + 255) / 256, 256>>>(B, size_t(N) * K, 77u, 2.f);
  fill_f32<<<(size_t(M) * N + 255) / 256, 256>>>(Res, size_t(M) * N, 5u, 4.f, 0.f);
  fill_f32<<<(M + 255) / 256, 256>>>(rs, M, 9u, 1.f, 1.f);
  fill_f32<<<(N + 255) / 256, 256>>>(cs, N, 11u, 1.f, 1.f);
  fill_f32<<<(N + 255) / 256, 256>>>(bias, N, 13u, 2.f, 0.f);
  CK(cudaMemset(Cb, 0xff, size_t(M) * N * 2));
  CK(cudaMemset(Cf, 0xff, size_t(M) * N * 4));
  CK(cudaMemset(Sh, 0xff, size_t(M) * N * 2));

  lrp_epilogue_t e;
  memset(&e, 0, sizeof(e));
  e.alpha = 1.f;
  e.ldc = N;
  if (mode == 0) {
    e.out = Cb; e.out_is_f32 = 0;
  } else {
    e.out = Cf; e.out_is_f32 = 1; e.shadow_bf16 = Sh; e.resid_f32 = Res; e.rowscale = rs; e.colscale = cs;
    e.bias = bias; e.alpha = 0.5f;
  }
  int rc = lrp_gemm_bf16(A, K, B, b_layout == 0 ? K : N, b_layout, M, N, K, &e, tile_n, 0);
  if (rc != 0) {
    printf("FAIL M=%d N=%d K=%d layout=%d bn=%d mode=%d: rc=%d %s\n", M, N, K, b_layout, tile_n, mode, rc,
           lrp_last_error());
    g_fail++;
    return;
  }
  dim3 g((N + 127) / 128, M);
  ref_gemm<<<g, 128>>>(A, B, b_layout, M, N, K, mode ? Res : nullptr, mode ? rs : nullptr, mode ? cs : nullptr,
                       mode ? bias : nullptr, e.alpha, Ref);
  cudaError_t ce = cudaDeviceSynchronize();
  if (ce != cudaSuccess) {
    printf("FAIL M=%d N=%d K=%d layout=%d bn=%d mode=%d: %s\n", M, N, K, b_layout, tile_n, mode,
           cudaGetErrorString(ce));
    exit(3);
  }
  std::vector<float> ref(size_t(M) * N), got(size_t(M) * N);
  CK(cudaMemcpy(ref.data(), Ref, ref.size() * 4, cudaMemcpyDeviceToHost));
  std::vector<__nv_bfloat16> gb(size_t(M) * N);
  if (mode == 0) {
    CK(cudaMemcpy(gb.data(), Cb, gb.size() * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < got.size(); ++i) got[i] = __bfloat162float(gb[i]);
  } else {
    CK(cudaMemcpy(got.data(), Cf, got.size() * 4, cudaMemcpyDeviceToHost));
  }
  double num = 0, den = 0;
  for (size_t i = 0; i < got.size(); ++i) {
    double d = double(got[i]) - ref[i];
    if (!(got[i] == got[i])) d = 1e30;
    num += d * d;
    den += double(ref[i]) * ref[i];
  }
  double rel = sqrt(num / (den + 1e-30));
  double tol = mode == 0 ? 3e-3 : 1e-5;
  double rel_sh = 0;
  if (mode == 1) {
    CK(cudaMemcpy(gb.data(), Sh, gb.size() * 2, cudaMemcpyDeviceToHost));
    double n2 = 0;
    for (size_t i = 0; i < got.size(); ++i) {
      double d = double(__bfloat162float(gb[i])) - ref[i];
      if (!(d == d)) d = 1e30;
      n2 += d * d;
    }
    rel_sh = sqrt(n2 / (den + 1e-30));
  }
  bool ok = rel < tol && rel_sh < 3e-3;
  printf("%s M=%d N=%d K=%d layout=%s bn=%d mode=%d rel_l2=%.3e shadow_rel=%.3e\n", ok ? "ok  " : "FAIL", M, N, K,
         b_layout ? "NN" : "NT", tile_n, mode, rel, rel_sh);
  if (!ok) g_fail++;
  cudaFree(A); cudaFree(B); cudaFree(Cb); cudaFree(Sh); cudaFree(Cf); cudaFree(Ref); cudaFree(Res);
  cudaFree(rs); cudaFree(cs); cudaFree(bias);
}

static void perf_case(int M, int N, int K, int b_layout, int tile_n, int mode) {
  __nv_bfloat16 *A, *B, *Cb;
  float* Cf;
  CK(cudaMalloc(&A, size_t(M) * K * 2));
  CK(cudaMalloc(&B, size_t(N) * K * 2));
  CK(cudaMalloc(&Cb, size_t(M) * N * 2));
  CK(cudaMalloc(&Cf, size_t(M) * N * 4));
  fill_bf16<<<(size_t(M) * K + 255) / 256, 256>>>(A, size_t(M) * K, 1u, 1.f);
  fill_bf16<<<(size_t(N) * K + 255) / 256, 256>>>(B, size_t(N) * K, 77u, 1.f);
  CK(cudaMemset(Cf, 0, size_t(M) * N * 4));
  lrp_epilogue_t e;
  memset(&e, 0, sizeof(e));
  e.alpha = 1.f; e.ldc = N;
  if (mode == 0) { e.out = Cb; } else { e.out = Cf; e.out_is_f32 = 1; e.resid_f32 = Cf; e.shadow_bf16 = Cb; }
  for (int i = 0; i < 3; ++i) lrp_gemm_bf16(A, K, B, b_layout == 0 ? K : N, b_layout, M, N, K, &e, tile_n, 0);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 10;
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) lrp_gemm_bf16(A, K, B, b_layout == 0 ? K : N, b_layout, M, N, K, &e, tile_n, 0);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= iters;
  printf("perf M=%d N=%d K=%d layout=%s bn=%d mode=%d: %.3f ms  %.1f TFLOP/s\n", M, N, K, b_layout ? "NN" : "NT",
         tile_n, mode, ms, 2.0 * M * N * K / ms * 1e-9);
  cudaFree(A); cudaFree(B); cudaFree(Cb); cudaFree(Cf);
}


// ---- fused eps-LRP Linear rule -------------------------------------------------------------------
__global__ void ref_z(const __nv_bfloat16* x, const __nv_bfloat16* W, const float* bias, int T, int N, int K, float* z) {
  int n = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y;
  if (n >= N) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc += __bfloat162float(x[size_t(t) * K + k]) * __bfloat162float(W[size_t(n) * K + k]);
  z[size_t(t) * N + n] = acc + (bias ? bias[n] : 0.f);
}
__global__ void ref_rin(const __nv_bfloat16* x, const __nv_bfloat16* W, const float* z, const float* R, int T, int N, int K,
                        float eps, float* rin) {
  int k = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y;
  if (k >= K) return;
  float acc = 0.f;
  for (int n = 0; n < N; ++n) {
    float s = R[size_t(t) * N + n] / (z[size_t(t) * N + n] + eps);
    acc += __bfloat162float(__float2bfloat16(s)) * __bfloat162float(W[size_t(n) * K + k]);
  }
  rin[size_t(t) * K + k] = acc * __bfloat162float(x[size_t(t) * K + k]);
}

static void eps_case(int T, int N, int K, bool with_bias, bool perf) {
  __nv_bfloat16 *x, *W, *s_ws;
  float *R, *Rin, *z, *ref, *bias;
  int32_t* flags;
  CK(cudaMalloc(&x, size_t(T) * K * 2)); CK(cudaMalloc(&W, size_t(N) * K * 2)); CK(cudaMalloc(&s_ws, size_t(T) * N * 2));
  CK(cudaMalloc(&R, size_t(T) * N * 4)); CK(cudaMalloc(&Rin, size_t(T) * K * 4)); CK(cudaMalloc(&z, size_t(T) * N * 4));
  CK(cudaMalloc(&ref, size_t(T) * K * 4)); CK(cudaMalloc(&bias, size_t(N) * 4));
  const int64_t nf = lrp_linear_eps_flags_count(T);
  CK(cudaMalloc(&flags, nf * 4));
  // positive operands keep z away from 0 so that the comparison is well conditioned
  fill_bf16<<<(size_t(T) * K + 255) / 256, 256>>>(x, size_t(T) * K, 21u, 1.f, 1.f);
  fill_bf16<<<(size_t(N) * K + 255) / 256, 256>>>(W, size_t(N) * K, 23u, 1.f, 1.f);
  fill_f32<<<(size_t(T) * N + 255) / 256, 256>>>(R, size_t(T) * N, 25u, 2.f, 0.f);
  fill_f32<<<(N + 255) / 256, 256>>>(bias, N, 27u, 1.f, 1.f);
  CK(cudaMemset(Rin, 0xff, size_t(T) * K * 4));
  CK(cudaMemset(flags, 0, nf * 4));
  int rc = lrp_linear_eps_bwd(x, W, with_bias ? bias : nullptr, R, 1, Rin, s_ws, flags, T, N, K, 1e-6f, 0);
  if (rc) { printf("FAIL eps rc=%d %s\n", rc, lrp_last_error()); g_fail++; return; }
  cudaError_t ce = cudaDeviceSynchronize();
  if (ce != cudaSuccess) { printf("FAIL eps exec %s\n", cudaGetErrorString(ce)); exit(3); }
  if (!perf) {
    ref_z<<<dim3((N + 127) / 128, T), 128>>>(x, W, with_bias ? bias : nullptr, T, N, K, z);
    ref_rin<<<dim3((K + 127) / 128, T), 128>>>(x, W, z, R, T, N, K, 1e-6f, ref);
    CK(cudaDeviceSynchronize());
    std::vector<float> a(size_t(T) * K), b(size_t(T) * K);
    CK(cudaMemcpy(a.data(), Rin, a.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), ref, b.size() * 4, cudaMemcpyDeviceToHost));
    double num = 0, den = 0;
    for (size_t i = 0; i < a.size(); ++i) { double d = double(a[i]) - b[i]; if (!(a[i] == a[i])) d = 1e30; num += d * d; den += double(b[i]) * b[i]; }
    double rel = sqrt(num / (den + 1e-30));
    bool ok = rel < 2e-3;
    printf("%s eps-linear T=%d N=%d K=%d bias=%d rel_l2=%.3e\n", ok ? "ok  " : "FAIL", T, N, K, int(with_bias), rel);
    if (!ok) g_fail++;
  } else {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 5;
    float tot = 0;
    for (int i = 0; i < iters; ++i) {
      CK(cudaMemsetAsync(flags, 0, nf * 4));
      cudaEventRecord(e0);
      lrp_linear_eps_bwd(x, W, nullptr, R, 1, Rin, s_ws, flags, T, N, K, 1e-6f, 0);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1); tot += ms;
    }
    tot /= iters;
    printf("perf eps-linear T=%d N=%d K=%d: %.3f ms  %.1f TFLOP/s (4TKN)\n", T, N, K, tot, 4.0 * T * N * K / tot * 1e-9);
  }
  cudaFree(x); cudaFree(W); cudaFree(s_ws); cudaFree(R); cudaFree(Rin); cudaFree(z); cudaFree(ref); cudaFree(bias); cudaFree(flags);
}

int main(int argc, char** argv) {
  bool perf = argc > 1 && !strcmp(argv[1], "--perf");
  if (lrp_check_device() != 0) { printf("no device: %s\n", lrp_last_error()); return 1; }
  printf("lrp_version=%d\n", lrp_version());
  if (argc > 1 && !strcmp(argv[1], "--sweep")) {
    // every GEMM shape of one Llama-3-8B layer (forward NT, LRP dgrad NN) at the token counts of micro-batch 8 / 4 / 2,
    // CTA-pair kernel (bn=2) vs one-CTA 128x256 kernel (bn=256): the table behind the per-shape dispatch in gemm_sm100.cu
    const int fwd[][2] = {{6144, 4096}, {4096, 4096}, {28672, 4096}, {4096, 14336}};
    const int bwd[][2] = {{14336, 4096}, {4096, 28672}, {4096, 4096}, {4096, 6144}};
    for (int M : {16384, 8192, 4096}) {
      for (auto& s : fwd) for (int bn : {2, 256}) perf_case(M, s[0], s[1], 0, bn, 0);
      for (auto& s : bwd) for (int bn : {2, 256}) perf_case(M, s[0], s[1], 1, bn, 1);
    }
    return 0;
  }
  const int shapes[][3] = {{128, 128, 64}, {128, 256, 128}, {256, 256, 256}, {300, 520, 200}, {1000, 1024, 4096},
                           {64, 2048, 512}};
  for (auto& s : shapes)
    for (int layout = 0; layout < 2; ++layout)
      for (int bn : {128, 256, 2})   // 2 = forced CTA-pair (cta_group::2) kernel
        for (int mode = 0; mode < 2; ++mode) check_case(s[0], s[1], s[2], layout, bn, mode);
  // multi-tile-per-CTA persistent path
  check_case(128 * 40, 256 * 8, 512, 0, 256, 0);
  check_case(128 * 40, 256 * 8, 512, 1, 256, 1);
  check_case(128 * 37, 128 * 9, 192, 1, 128, 0);
  check_case(256 * 80, 256 * 4, 512, 0, 2, 0);      // CTA pairs, several tiles per pair
  check_case(256 * 80 + 77, 256 * 4 + 8, 520, 1, 2, 1);
  check_case(4096, 4096, 1024, 0, 0, 1);            // auto -> CTA pairs
  check_case(4096, 4096, 1024, 1, 0, 0);
  eps_case(128, 256, 256, false, false);
  eps_case(300, 520, 200, true, false);
  eps_case(1500, 1024, 768, true, false);
  eps_case(4096, 2048, 1024, false, false);
  printf(g_fail ? "SELFTEST FAILED (%d)\n" : "SELFTEST PASSED\n", g_fail);
  if (perf) {
    perf_case(8192, 4096, 4096, 0, 0, 0);    // auto = CTA pairs
    perf_case(8192, 4096, 4096, 1, 0, 0);
    perf_case(16384, 28672, 4096, 0, 0, 0);
    perf_case(16384, 4096, 14336, 0, 0, 1);
    perf_case(16384, 4096, 28672, 1, 0, 1);
    perf_case(16384, 14336, 4096, 1, 0, 0);
    perf_case(16384, 6144, 4096, 0, 0, 0);
    perf_case(8192, 4096, 4096, 0, 256, 0);
    perf_case(8192, 4096, 4096, 0, 128, 0);
    perf_case(8192, 4096, 4096, 1, 256, 0);
    perf_case(8192, 4096, 4096, 1, 128, 0);
    perf_case(16384, 28672, 4096, 0, 256, 0);
    perf_case(16384, 4096, 14336, 0, 256, 1);
    perf_case(16384, 4096, 28672, 1, 256, 1);
    perf_case(16384, 14336, 4096, 1, 256, 0);
    perf_case(16384, 6144, 4096, 0, 256, 0);
    eps_case(16384, 4096, 4096, false, true);
    eps_case(16384, 14336, 4096, false, true);
  }
  return g_fail ? 1 : 0;
}
