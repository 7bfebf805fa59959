// Pieces shared by the one-CTA and the CTA-pair GEMM kernels: parameters, tile rasterisation and the fused epilogue of
// one 32-column accumulator chunk (reference call sites in gemm_sm100.cu's header).
#pragma once
#include "ptx_sm100.cuh"
#include "lrp_internal.h"

namespace lrp {

struct GemmParams {
  int M, N, K;
  // epilogue
  void* out;
  __nv_bfloat16* shadow;
  const float* resid;
  const float* rowscale;
  const float* colscale;
  const float* bias;
  float alpha;
  int64_t ldc;
  int out_is_f32;
  const __nv_bfloat16* gated_gu;
  __nv_bfloat16* gated_out;
  int gated_act, gated_cp;
  __nv_bfloat16* act_out;   // fused gated-MLP forward: a[m, n/2] = act(gate) * up from 32-interleaved (gate | up) column blocks
  int group_m;  // rasterisation: `group_m` m-blocks share each streamed B panel through L2
  int batch;              // strided-batched form (one-CTA kernel only): `batch` independent problems in one launch
  int64_t batch_stride_c; // element distance between consecutive problems in out / resid / shadow
};

__device__ __forceinline__ float gemm_act_eval(float x, int act) {
  if (act == LRP_ACT_SILU) return x / (1.f + __expf(-x));
  if (act == LRP_ACT_GELU_TANH) {
    const float k = 0.7978845608028654f;
    return 0.5f * x * (1.f + tanhf(k * (x + 0.044715f * x * x * x)));
  }
  return 0.5f * x * (1.f + erff(x * 0.7071067811865476f));
}

__device__ __forceinline__ void tile_coords(int t, int num_m, int num_n, int group_m, int& m_blk, int& n_blk) {
  const int tiles_per_group = group_m * num_n;
  const int group = t / tiles_per_group;
  const int first_m = group * group_m;
  const int gsize = min(group_m, num_m - first_m);
  const int in_group = t - group * tiles_per_group;
  m_blk = first_m + in_group % gsize;
  n_blk = in_group / gsize;
}

// Fused gated-MLP forward (lxt/efficient/patches.py:145-157): the weight rows are interleaved in blocks of 32 (gate rows
// 32k..32k+31, then up rows 32k..32k+31), so columns [n0, n0+32) of the accumulator are a gate block and [n0+32, n0+64) the
// matching up block.  a = act(bf16(gate)) * bf16(up), i.e. exactly what lrp_gated_act_fwd computes from the stored bf16 gu.
__device__ __forceinline__ void gemm_epilogue_act_pair(const GemmParams& p, const uint32_t (&vg)[32], const uint32_t (&vu)[32], int m,
                                                       float rs, int n0) {
  __nv_bfloat16* arow = p.act_out + int64_t(m) * (p.N >> 1) + (n0 >> 1);
#pragma unroll
  for (int j8 = 0; j8 < 4; ++j8) {
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float g = __uint_as_float(vg[j8 * 8 + j]) * rs, u = __uint_as_float(vu[j8 * 8 + j]) * rs;
      if (p.bias != nullptr) { g += p.bias[n0 + j8 * 8 + j]; u += p.bias[n0 + 32 + j8 * 8 + j]; }
      g = __bfloat162float(__float2bfloat16_rn(g));
      u = __bfloat162float(__float2bfloat16_rn(u));
      a[j] = gemm_act_eval(g, p.gated_act) * u;
    }
    *reinterpret_cast<uint4*>(arow + j8 * 8) =
        make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]), pack_bf16x2(a[4], a[5]), pack_bf16x2(a[6], a[7]));
  }
}

// v[32] = fp32 accumulators of row m, columns [n0, n0+32); rs = alpha * rowscale[m]
__device__ __forceinline__ void gemm_epilogue_chunk(const GemmParams& p, const uint32_t (&v)[32], int m, float rs, int64_t row_off,
                                                    int n0) {
#pragma unroll
  for (int j8 = 0; j8 < 4; ++j8) {
    const int n = n0 + j8 * 8;
    if (n < p.N) {  // N is a multiple of 8 (checked on the host)
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[j8 * 8 + j]) * rs;
      if (p.colscale != nullptr) {
        const float4 c0 = *reinterpret_cast<const float4*>(p.colscale + n);
        const float4 c1 = *reinterpret_cast<const float4*>(p.colscale + n + 4);
        f[0] *= c0.x; f[1] *= c0.y; f[2] *= c0.z; f[3] *= c0.w;
        f[4] *= c1.x; f[5] *= c1.y; f[6] *= c1.z; f[7] *= c1.w;
      }
      if (p.bias != nullptr) {
        const float4 b0 = *reinterpret_cast<const float4*>(p.bias + n);
        const float4 b1 = *reinterpret_cast<const float4*>(p.bias + n + 4);
        f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
        f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
      }
      if (p.resid != nullptr) {
        const float4 r0 = *reinterpret_cast<const float4*>(p.resid + row_off + n);
        const float4 r1 = *reinterpret_cast<const float4*>(p.resid + row_off + n + 4);
        f[0] += r0.x; f[1] += r0.y; f[2] += r0.z; f[3] += r0.w;
        f[4] += r1.x; f[5] += r1.y; f[6] += r1.z; f[7] += r1.w;
      }
      if (p.gated_gu != nullptr) {
        // fused gated-MLP LRP backward: f[] = g_a
        const int64_t goff = int64_t(m) * (2 * int64_t(p.N));
        const uint4 ug = *reinterpret_cast<const uint4*>(p.gated_gu + goff + n);
        const uint4 uu = *reinterpret_cast<const uint4*>(p.gated_gu + goff + p.N + n);
        const float gt[8] = {bf16_lo(ug.x), bf16_hi(ug.x), bf16_lo(ug.y), bf16_hi(ug.y),
                             bf16_lo(ug.z), bf16_hi(ug.z), bf16_lo(ug.w), bf16_hi(ug.w)};
        const float up[8] = {bf16_lo(uu.x), bf16_hi(uu.x), bf16_lo(uu.y), bf16_hi(uu.y),
                             bf16_lo(uu.z), bf16_hi(uu.z), bf16_lo(uu.w), bf16_hi(uu.w)};
        float og[8], ou[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float sv = gemm_act_eval(gt[j], p.gated_act);
          if (p.gated_cp) {
            ou[j] = f[j] * sv;
            og[j] = 0.f;
          } else {
            const float gh = f[j] * 0.5f;
            ou[j] = gh * sv;
            og[j] = (sv / (gt[j] + 1e-10f)) * (gh * up[j]);
          }
        }
        *reinterpret_cast<uint4*>(p.gated_out + goff + n) =
            make_uint4(pack_bf16x2(og[0], og[1]), pack_bf16x2(og[2], og[3]), pack_bf16x2(og[4], og[5]), pack_bf16x2(og[6], og[7]));
        *reinterpret_cast<uint4*>(p.gated_out + goff + p.N + n) =
            make_uint4(pack_bf16x2(ou[0], ou[1]), pack_bf16x2(ou[2], ou[3]), pack_bf16x2(ou[4], ou[5]), pack_bf16x2(ou[6], ou[7]));
      } else if (p.out_is_f32) {
        float* o = reinterpret_cast<float*>(p.out) + row_off + n;
        *reinterpret_cast<float4*>(o) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(f[4], f[5], f[6], f[7]);
      } else {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + row_off + n;
        *reinterpret_cast<uint4*>(o) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                                  pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
      }
      if (p.shadow != nullptr) {
        *reinterpret_cast<uint4*>(p.shadow + row_off + n) =
            make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                       pack_bf16x2(f[6], f[7]));
      }
    }
  }
}

}  // namespace lrp
