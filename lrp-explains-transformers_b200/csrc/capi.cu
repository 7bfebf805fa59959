// C-ABI plumbing of liblrp_b200.so: error state, device query, TMA descriptor encoding and the GEMM-family
// entry points declared in include/lrp_b200.h.
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include "lrp_internal.h"

namespace lrp {

static thread_local char g_err[512] = {0};

int set_error(int code, const char* msg) {
  strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
  return code;
}

static std::atomic<long long> g_launches{0};
void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  // resolved through the runtime so the library does not link against libcuda (absent on build hosts)
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t ld, uint32_t box0,
                      uint32_t box1) {
  PFN_encodeTiled enc = get_encode();
  if (enc == nullptr) return set_error(LRP_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t dims[2] = {dim0, dim1};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[160];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (%d): dims=%llu x %llu ld=%llu box=%u x %u", int(r),
             (unsigned long long)dim0, (unsigned long long)dim1, (unsigned long long)ld, box0, box1);
    return set_error(LRP_ERR_CUDA, buf);
  }
  return LRP_OK;
}

int make_tmap_3d_bf16(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t dim2,
                      uint64_t stride1, uint64_t stride2, uint32_t box0, uint32_t box1) {
  PFN_encodeTiled enc = get_encode();
  if (enc == nullptr) return set_error(LRP_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t dims[3] = {dim0, dim1, dim2};
  cuuint64_t strides[2] = {stride1 * 2, stride2 * 2};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[200];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled(3d) failed (%d): dims=%llu x %llu x %llu strides=%llu,%llu", int(r),
             (unsigned long long)dim0, (unsigned long long)dim1, (unsigned long long)dim2,
             (unsigned long long)stride1, (unsigned long long)stride2);
    return set_error(LRP_ERR_CUDA, buf);
  }
  return LRP_OK;
}

// fp32 [dim2][dim1][dim0] tensor, 128B-swizzled boxes of 32 floats x box1 rows: the destination of the attention
// backward's dQ tile reductions (cp.reduce.async.bulk.tensor ... .add)
int make_tmap_3d_f32(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t dim2, uint64_t stride1,
                     uint64_t stride2, uint32_t box0, uint32_t box1) {
  PFN_encodeTiled enc = get_encode();
  if (enc == nullptr) return set_error(LRP_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t dims[3] = {dim0, dim1, dim2};
  cuuint64_t strides[2] = {stride1 * 4, stride2 * 4};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[200];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled(3d f32) failed (%d): dims=%llu x %llu x %llu", int(r),
             (unsigned long long)dim0, (unsigned long long)dim1, (unsigned long long)dim2);
    return set_error(LRP_ERR_CUDA, buf);
  }
  return LRP_OK;
}

}  // namespace lrp

using namespace lrp;

extern "C" {

int lrp_version(void) { return 1000; }

const char* lrp_last_error(void) { return g_err; }

int64_t lrp_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int lrp_check_device(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return set_error(LRP_ERR_NO_DEVICE, "no CUDA device visible");
  }
  int dev = 0, major = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) return set_error(LRP_ERR_NO_DEVICE, "device is not sm_100 (Blackwell B200)");
  return LRP_OK;
}

int lrp_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, int b_layout, int M, int N, int K,
                  const lrp_epilogue_t* epi, int tile_n, void* stream) {
  return gemm_bf16(A, lda, B, ldb, b_layout, M, N, K, epi, tile_n, static_cast<cudaStream_t>(stream));
}

int lrp_gemm_bf16_batched(const void* A, int64_t lda, int64_t stride_a, int a_layout, const void* B, int64_t ldb, int64_t stride_b,
                          int b_layout, int batch, int M, int N, int K, const lrp_epilogue_t* epi, int64_t stride_c, void* stream) {
  return gemm_bf16_batched(A, lda, stride_a, a_layout, B, ldb, stride_b, b_layout, batch, M, N, K, epi, stride_c,
                           static_cast<cudaStream_t>(stream));
}

int lrp_linear_fwd(const void* x, int64_t ldx, const void* W, int64_t ldw, int T, int N, int K,
                   const lrp_epilogue_t* epi, void* stream) {
  return gemm_bf16(x, ldx, W, ldw, /*NT*/ 0, T, N, K, epi, 0, static_cast<cudaStream_t>(stream));
}

int lrp_linear_dgrad_fused(const void* gy, int64_t ldg, const void* W, int64_t ldw, int T, int N, int K,
                           const lrp_epilogue_t* epi, void* stream) {
  // g_x[T,K] = g_y[T,N] W[N,K]: contraction over N, W consumed as the [Kc=N, Nc=K] "NN" operand
  return gemm_bf16(gy, ldg, W, ldw, /*NN*/ 1, T, K, N, epi, 0, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
