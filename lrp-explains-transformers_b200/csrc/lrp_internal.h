// Internal host-side declarations shared by the translation units of liblrp_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/lrp_b200.h"

namespace lrp {

int set_error(int code, const char* msg);  // stores a thread-local message, returns code
int sm_count();                            // SM count of the current device (cached)
void note_launch(int n = 1);               // process-wide count of kernels this library enqueued

// Encode a 2-D bf16 tiled tensor map with 128-byte swizzle.
//   dim0 = contiguous extent (elements), dim1 = rows, ld = row pitch (elements), box0 x box1 = tile.
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t ld,
                      uint32_t box0, uint32_t box1);

// 3-D variant: dims (dim0 contiguous, dim1, dim2) with element strides stride1/stride2; box = box0 x box1 x 1.
int make_tmap_3d_bf16(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t dim2,
                      uint64_t stride1, uint64_t stride2, uint32_t box0, uint32_t box1);
int make_tmap_3d_f32(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t dim2, uint64_t stride1,
                     uint64_t stride2, uint32_t box0, uint32_t box1);

int gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, int b_layout, int M, int N, int K,
              const lrp_epilogue_t* epi, int force_bn, cudaStream_t stream);

int gemm_bf16_batched(const void* A, int64_t lda, int64_t stride_a, int a_layout, const void* B, int64_t ldb, int64_t stride_b,
                      int b_layout, int batch, int M, int N, int K, const lrp_epilogue_t* epi, int64_t stride_c, cudaStream_t stream);

int attn_bwd_v2(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv, const void* d_o,
                const float* lse, const float* delta, void* dq, void* dk, void* dv, int64_t lddq, int64_t lddk, int64_t lddv,
                const int32_t* kv_range, int B, int S, int H, int Hkv, int D, float scale, int causal, int window, float q_div,
                float k_div, float v_div, cudaStream_t st);

int attn_fwd_ws(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv, void* o, float* lse,
                const int32_t* kv_range, int B, int S, int H, int Hkv, float scale, int causal, cudaStream_t st);

int linear_eps_bwd(const void* x, const void* W, const float* bias, const void* r_out, int r_is_f32, void* r_in,
                   void* s_ws, int32_t* flags_ws, int T, int N, int K, float eps, cudaStream_t stream);

#define LRP_CHECK_LAUNCH()                                                   \
  do {                                                                       \
    cudaError_t ce__ = cudaGetLastError();                                   \
    if (ce__ != cudaSuccess) return set_error(LRP_ERR_CUDA, cudaGetErrorString(ce__)); \
    note_launch();                                                           \
  } while (0)

}  // namespace lrp
