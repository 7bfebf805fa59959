// 4-bit NormalFloat (NF4) weight storage for the AttnLRP engine: weights live in HBM as 4-bit codes + one fp32 absmax per block of
// 64 values and are expanded to bf16 into a per-layer scratch right before the layer's tcgen05 GEMMs consume them.
//
// Reference context: every example of the reference loads the model with BitsAndBytesConfig(load_in_4bit=True, bnb_4bit_quant_type
// default "fp4"/"nf4") (examples/quantized_llama.py:13-19) and the explicit maps wrap `bnb.nn.Linear4bit` (lxt/explicit/models/
// llama.py:91-92): bitsandbytes dequantises the 4-bit weight to the compute dtype inside its matmul, forward and backward, and the
// LRP rules see an ordinary Linear.  bitsandbytes is not installed here, so the format below follows its published NF4 definition
// (QLoRA, Dettmers et al. 2023: 16 quantiles of N(0,1) normalised to [-1,1], block-wise absmax scaling, two codes per byte with the
// first element in the high nibble) and parity is pinned against an in-repo fp32 de-quantisation oracle, NOT against bnb itself.
#include "ptx_sm100.cuh"
#include "lrp_internal.h"

namespace lrp {

__constant__ float NF4_CODE[16] = {-1.0f, -0.6961928009986877f, -0.5250730514526367f, -0.39491748809814453f,
                                   -0.28444138169288635f, -0.18477343022823334f, -0.09105003625154495f, 0.0f,
                                   0.07958029955625534f, 0.16093020141124725f, 0.24611230194568634f, 0.33791524171829224f,
                                   0.44070982933044434f, 0.5626170039176941f, 0.7229568362236023f, 1.0f};

// 8 weights (4 bytes of codes) per thread -> one 16-byte bf16 store; blocksize is a multiple of 8 so the 8 share one absmax
__global__ void __launch_bounds__(256) dequant_nf4_kernel(const uint8_t* __restrict__ packed, const float* __restrict__ absmax,
                                                          __nv_bfloat16* __restrict__ out, int64_t n8, int blocksize) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n8; i += int64_t(gridDim.x) * blockDim.x) {
    const uint32_t c = *reinterpret_cast<const uint32_t*>(packed + i * 4);
    const float s = absmax[(i * 8) / blocksize];
    float f[8];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const uint32_t byte = (c >> (8 * b)) & 0xffu;
      f[2 * b] = NF4_CODE[byte >> 4] * s;
      f[2 * b + 1] = NF4_CODE[byte & 15u] * s;
    }
    *reinterpret_cast<uint4*>(out + i * 8) =
        make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  }
}

// quantiser: one thread per block of `blocksize` values: absmax, then nearest code per value (ties to the lower code index)
__global__ void __launch_bounds__(128) quant_nf4_kernel(const __nv_bfloat16* __restrict__ w, uint8_t* __restrict__ packed,
                                                        float* __restrict__ absmax, int64_t nblocks, int blocksize) {
  const int64_t blk = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (blk >= nblocks) return;
  const __nv_bfloat16* src = w + blk * blocksize;
  float m = 0.f;
  for (int i = 0; i < blocksize; ++i) m = fmaxf(m, fabsf(__bfloat162float(src[i])));
  absmax[blk] = m;
  const float inv = m > 0.f ? 1.f / m : 0.f;
  for (int i = 0; i < blocksize; i += 2) {
    uint32_t code[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float x = __bfloat162float(src[i + j]) * inv;
      int best = 0;
      float bd = fabsf(x - NF4_CODE[0]);
#pragma unroll
      for (int c = 1; c < 16; ++c) {
        const float d = fabsf(x - NF4_CODE[c]);
        if (d < bd) { bd = d; best = c; }
      }
      code[j] = uint32_t(best);
    }
    packed[(blk * blocksize + i) >> 1] = uint8_t((code[0] << 4) | code[1]);
  }
}

}  // namespace lrp

using namespace lrp;

extern "C" {

int lrp_quant_nf4(const void* w_bf16, void* packed, float* absmax, int64_t n, int blocksize, void* stream) {
  if (n <= 0 || blocksize < 8 || (blocksize % 8) != 0 || (n % blocksize) != 0)
    return set_error(LRP_ERR_ARG, "quant_nf4: n must be a positive multiple of the block size (a multiple of 8)");
  const int64_t nb = n / blocksize;
  quant_nf4_kernel<<<unsigned((nb + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      (const __nv_bfloat16*)w_bf16, (uint8_t*)packed, absmax, nb, blocksize);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

int lrp_dequant_nf4(const void* packed, const float* absmax, void* out_bf16, int64_t n, int blocksize, void* stream) {
  if (n <= 0 || blocksize < 8 || (blocksize % 8) != 0 || (n % blocksize) != 0)
    return set_error(LRP_ERR_ARG, "dequant_nf4: n must be a positive multiple of the block size (a multiple of 8)");
  if ((reinterpret_cast<uintptr_t>(packed) & 3) || (reinterpret_cast<uintptr_t>(out_bf16) & 15))
    return set_error(LRP_ERR_ARG, "dequant_nf4: packed must be 4-byte and out 16-byte aligned");
  const int64_t n8 = n / 8;
  int64_t g = (n8 + 255) / 256;
  const int64_t cap = int64_t(sm_count()) * 16;
  if (g > cap) g = cap;
  dequant_nf4_kernel<<<unsigned(g), 256, 0, static_cast<cudaStream_t>(stream)>>>((const uint8_t*)packed, absmax,
                                                                                (__nv_bfloat16*)out_bf16, n8, blocksize);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

}  // extern "C"
