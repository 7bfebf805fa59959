// Micro-benchmark: chip-wide throughput of fp32 bulk tensor reductions (cp.reduce.async.bulk.tensor .add) issued in the
// access pattern of the attention backward's dQ tiles: grid (S/128 key tiles, Hkv, B); every CTA walks G query heads x
// the causal query tiles of its key tile and adds one [128 x D] fp32 tile per step from a 64 KiB smem staging buffer.
// Nothing else runs, so the result is the ceiling the dQ reductions put on the kernel (bytes of atomic payload / s).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_reduce_bench tma_reduce_bench.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// layout 0: [B,S,H,D] (row stride H*D);  layout 1: [B,H,S,D] (row stride D)
template <int D, int DEPTH>
__global__ void __launch_bounds__(128, 1) reduce_kernel(const __grid_constant__ CUtensorMap tm, int S, int H, int Hkv, int layout,
                                                        int use_red, float* base, int wide) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  const int jt = blockIdx.x, hk = blockIdx.y, b = blockIdx.z, G = H / Hkv, nq = S / 128;
  for (int i = threadIdx.x; i < DEPTH * D * 128; i += blockDim.x) ((float*)smem)[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  int cnt = 0;
  for (int g = 0; g < G; ++g)
    for (int i = jt; i < nq; ++i, ++cnt) {
      const int h = hk * G + g;
      if (use_red) {   // per-thread red.global.add.v4.f32, thread = row
        float* row = layout == 0 ? base + ((int64_t(b) * S + i * 128 + threadIdx.x) * H + h) * D
                                 : base + ((int64_t(b) * H + h) * S + i * 128 + threadIdx.x) * D;
        for (int c = 0; c < D; c += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(row + c), "f"(1.0f) : "memory");
      } else if (threadIdx.x == 0) {
        uint8_t* st = smem + (cnt % DEPTH) * (D * 128 * 4);
        for (int c = 0; c < (wide ? 1 : D / 32); ++c) {
          const int c0 = layout == 0 ? h * D + c * 32 : c * 32, c1 = i * 128, c2 = layout == 0 ? b : b * H + h;
          asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                           (uint64_t)&tm), "r"(smem_u32(st + c * 16384)), "r"(c0), "r"(c1), "r"(c2) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH - 1) : "memory");   // staging slot reusable
      }
    }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char** argv) {
  const bool occ1 = argc > 1;   // any argument: pad smem so that one CTA fits per SM (the attention kernel's occupancy)
  const int B = 4, S = 2048, H = 32, Hkv = 8, D = 128;
  PFN_encodeTiled enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
  float* buf;
  const size_t n = size_t(B) * S * H * D;
  cudaMalloc(&buf, n * 4);
  const int nq = S / 128;
  double tiles = 0;
  for (int jt = 0; jt < nq; ++jt) tiles += (nq - jt);
  tiles *= double(H / Hkv) * Hkv * B;
  const double bytes = tiles * 128 * D * 4;
  for (int layout = 0; layout < 2; ++layout)
    for (int mode = 0; mode < 5; ++mode) {   // 0: TMA depth 1, 1: TMA depth 2 (two tiles in flight), 2: red.v4, 3/4: one un-swizzled [128 x 128] box per tile, depth 1/2
      CUtensorMap tm;
      cuuint64_t dims[3], strides[2];
      if (layout == 0) { dims[0] = uint64_t(H) * D; dims[1] = S; dims[2] = B; strides[0] = uint64_t(H) * D * 4; strides[1] = uint64_t(S) * H * D * 4; }
      else { dims[0] = D; dims[1] = S; dims[2] = uint64_t(B) * H; strides[0] = D * 4; strides[1] = uint64_t(S) * D * 4; }
      const bool wide = mode >= 3;
      cuuint32_t box[3] = {wide ? 128u : 32u, 128, 1}, estr[3] = {1, 1, 1};
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       wide ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", int(r)); return 1; }
      const int depth = (mode == 1 || mode == 4) ? 2 : 1;
      const int smem = occ1 ? 200 * 1024 : depth * D * 128 * 4 + 1024;
      auto k1 = reduce_kernel<D, 1>;
      auto k2 = reduce_kernel<D, 2>;
      cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      float best = 1e9f;
      for (int rep = 0; rep < 5; ++rep) {
        cudaMemset(buf, 0, n * 4);
        cudaEventRecord(e0);
        if (depth == 2) k2<<<dim3(nq, Hkv, B), 128, smem>>>(tm, S, H, Hkv, layout, 0, buf, wide);
        else k1<<<dim3(nq, Hkv, B), 128, smem>>>(tm, S, H, Hkv, layout, mode == 2, buf, wide);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
      }
      cudaError_t ce = cudaGetLastError();
      float probe = 0;
      cudaMemcpy(&probe, buf + (size_t(S) - 1) * H * D, 4, cudaMemcpyDeviceToHost);   // last query row: all 16 key tiles added
      printf("layout %s  %-22s : %.3f ms  %.2f TB/s of fp32 atomic payload  (probe %.0f, %s)\n", layout == 0 ? "[B,S,H,D]" : "[B,H,S,D]",
             mode == 0 ? "TMA reduce, 1 in flight" : mode == 1 ? "TMA reduce, 2 in flight" : mode == 2 ? "red.global.add.v4" : mode == 3 ? "TMA wide box, 1 in flight" : "TMA wide box, 2 in flight", best, bytes / best / 1e9,
             probe, cudaGetErrorString(ce));
    }
  return 0;
}
