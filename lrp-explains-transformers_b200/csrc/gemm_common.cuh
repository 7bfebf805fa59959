// Pieces shared by the one-CTA and the CTA-pair GEMM kernels: parameters, tile rasterisation and the fused epilogue of
// one 32-column accumulator chunk (reference call sites in gemm_sm100.cu's header).
#pragma once
#include "ptx_sm100.cuh"
#include "lrp_internal.h"

namespace lrp {

// Epilogue stores: every output of the step's GEMMs (T = 16384 rows) is larger than the 126 MB L2 and is read again only by a later
// kernel, while the x / W operand panels are what the L2 should keep (the MMA warp waits ~9 % of its time for operands, and the panels
// are re-read ~6x from DRAM).  With GemmParams::stream_stores (default on, LRP_GEMM_STREAM_STORES=0 for A/B runs) the output stores are
// st.global.cs (evict-first), so that they do not push operand lines out.
__device__ __forceinline__ void epi_store(bool stream, uint4* dst, uint4 v) {
  if (stream) __stcs(dst, v);
  else *dst = v;
}
__device__ __forceinline__ void epi_store(bool stream, float4* dst, float4 v) {
  if (stream) __stcs(dst, v);
  else *dst = v;
}

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
  int gated_act, gated_cp, gated_layout;
  __nv_bfloat16* act_out;   // fused gated-MLP forward: a[m, n/2] = act(gate) * up from 32-interleaved (gate | up) column blocks
  // fused attention-backward prologue (O-projection dgrad): delta[b,h,s] = sum_d o[b,s,h,d] * dO[b,s,h,d] with dO = this GEMM's
  // bf16 output row (m = b*S + s, columns h*D..h*D+D-1); o has the output's layout
  const __nv_bfloat16* delta_o;
  float* delta_out;
  int delta_D, delta_S;
  int group_m;  // rasterisation: `group_m` m-blocks share each streamed B panel through L2
  int batch;              // strided-batched form (one-CTA kernel only): `batch` independent problems in one launch
  int64_t batch_stride_c; // element distance between consecutive problems in out / resid / shadow
  int stream_stores;      // evict-first output stores (see epi_store)
  long long* dbg;         // LRP_GEMM_DEBUG=1: per CTA pair {total, wait tmem_empty, wait full, tiles} cycles of the MMA warp
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

// Activation of the fused gated-MLP forward epilogue.  The epilogue warps share their SM sub-partitions' XU (MUFU) and issue slots
// with nothing else, but an IEEE division + a three-way run-time switch per element made the epilogue of the 128 x 256 gate|up
// tile outlast its 4096-deep mainloop (ncu: tensor pipe 49 %, XU 59 %; profiles/r02_ncu_reading.md).  Here the activation is a
// template parameter and costs two MUFU operations (ex2 + rcp) for SiLU and for GELU-tanh alike.
template <int ACT>
__device__ __forceinline__ float gemm_act_fast(float x) {
  if constexpr (ACT == LRP_ACT_SILU) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
    return x * r;
  } else if constexpr (ACT == LRP_ACT_GELU_TANH) {
    // 0.5 x (1 + tanh u) = x * sigmoid(2u), u = sqrt(2/pi) (x + 0.044715 x^3): the same two MUFU operations as SiLU and no
    // cancellation in the negative tail (tanh.approx's 5e-4 absolute error would be a 20 % error of gelu(-3))
    const float u2 = 1.5957691216057308f * (x + 0.044715f * x * x * x);
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * u2));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
    return x * r;
  } else {
    return 0.5f * x * (1.f + erff(x * 0.7071067811865476f));
  }
}

// Fused gated-MLP forward (lxt/efficient/patches.py:145-157): the weight rows are interleaved in blocks of 32 (gate rows
// 32k..32k+31, then up rows 32k..32k+31), so columns [n0, n0+32) of the accumulator are a gate block and [n0+32, n0+64) the
// matching up block.  a = act(bf16(gate)) * bf16(up): the same values lrp_gated_act_fwd computes from the stored bf16 gu (to
// fp32 rounding of the activation: approximate ex2 / rcp / tanh here, IEEE there; both far below the bf16 rounding of a).
// One lean pass per (gate block, up block) pair: scale, round to bf16 ONCE (the packed words are both the stored gu values and,
// re-expanded by a shift, the activation's inputs), store both gu blocks and the a block.  Kept small on purpose: the first
// version went through the generic chunk epilogue twice plus a separate activation pass and stalled on instruction fetch
// (ncu: 31 % `no_instructions`, tensor pipe 50 %; profiles/r02_ncu_reading.md).
template <int ACT>
__device__ __forceinline__ void gemm_epilogue_act_pair_t(const GemmParams& p, const uint32_t (&vg)[32], const uint32_t (&vu)[32], int m,
                                                         float rs, int64_t row_off, int n0) {
  __nv_bfloat16* grow = reinterpret_cast<__nv_bfloat16*>(p.out) + row_off + n0;
  __nv_bfloat16* arow = p.act_out + int64_t(m) * (p.N >> 1) + (n0 >> 1);
#pragma unroll
  for (int j8 = 0; j8 < 4; ++j8) {
    uint32_t pg[4], pu[4], pa[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float g0 = __uint_as_float(vg[j8 * 8 + 2 * j]) * rs, g1 = __uint_as_float(vg[j8 * 8 + 2 * j + 1]) * rs;
      float u0 = __uint_as_float(vu[j8 * 8 + 2 * j]) * rs, u1 = __uint_as_float(vu[j8 * 8 + 2 * j + 1]) * rs;
      if (p.bias != nullptr) {
        g0 += p.bias[n0 + j8 * 8 + 2 * j]; g1 += p.bias[n0 + j8 * 8 + 2 * j + 1];
        u0 += p.bias[n0 + 32 + j8 * 8 + 2 * j]; u1 += p.bias[n0 + 32 + j8 * 8 + 2 * j + 1];
      }
      pg[j] = pack_bf16x2(g0, g1);
      pu[j] = pack_bf16x2(u0, u1);
      pa[j] = pack_bf16x2(gemm_act_fast<ACT>(bf16_lo(pg[j])) * bf16_lo(pu[j]), gemm_act_fast<ACT>(bf16_hi(pg[j])) * bf16_hi(pu[j]));
    }
    epi_store(p.stream_stores != 0, reinterpret_cast<uint4*>(grow + j8 * 8), make_uint4(pg[0], pg[1], pg[2], pg[3]));
    epi_store(p.stream_stores != 0, reinterpret_cast<uint4*>(grow + 32 + j8 * 8), make_uint4(pu[0], pu[1], pu[2], pu[3]));
    epi_store(p.stream_stores != 0, reinterpret_cast<uint4*>(arow + j8 * 8), make_uint4(pa[0], pa[1], pa[2], pa[3]));
  }
}

__device__ __forceinline__ void gemm_epilogue_act_pair(const GemmParams& p, const uint32_t (&vg)[32], const uint32_t (&vu)[32], int m,
                                                       float rs, int64_t row_off, int n0) {
  if (p.gated_act == LRP_ACT_SILU) gemm_epilogue_act_pair_t<LRP_ACT_SILU>(p, vg, vu, m, rs, row_off, n0);
  else if (p.gated_act == LRP_ACT_GELU_TANH) gemm_epilogue_act_pair_t<LRP_ACT_GELU_TANH>(p, vg, vu, m, rs, row_off, n0);
  else gemm_epilogue_act_pair_t<LRP_ACT_GELU_ERF>(p, vg, vu, m, rs, row_off, n0);
}

// Fused gated-MLP LRP backward in the down-projection dgrad epilogue (lxt/efficient/patches.py:145-157, rules.py:88-127): the
// accumulator chunk is g_a[m, n0..n0+31]; with gate / up read from gated_gu,
//   g_up = (g_a / 2) * act(gate),   g_gate = (g_a / 2) * up * act(gate) / (gate + 1e-10)        (CP-LRP: g_up = g_a * act(gate), g_gate = 0)
// are written to gated_out in the same (gate | up) layout (0: halves, 1: blocks of 32 interleaved).  Lean like the forward pair:
// approximate ex2 / rcp, activation as a template parameter, no optional epilogue terms.
template <int ACT, bool CP>
__device__ __forceinline__ void gemm_epilogue_gated_bwd_t(const GemmParams& p, const uint32_t (&v)[32], int m, float rs, int n0) {
  const int64_t goff = int64_t(m) * (2 * int64_t(p.N)) + (p.gated_layout ? ((n0 >> 5) << 6) : n0);
  const int uo = p.gated_layout ? 32 : p.N;
#pragma unroll
  for (int j8 = 0; j8 < 4; ++j8) {
    const uint4 ug = *reinterpret_cast<const uint4*>(p.gated_gu + goff + j8 * 8);
    const uint4 uu = *reinterpret_cast<const uint4*>(p.gated_gu + goff + uo + j8 * 8);
    const uint32_t wg[4] = {ug.x, ug.y, ug.z, ug.w}, wu[4] = {uu.x, uu.y, uu.z, uu.w};
    uint32_t og[4], ou[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float rg[2], ru[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float g = e ? bf16_hi(wg[j]) : bf16_lo(wg[j]), u = e ? bf16_hi(wu[j]) : bf16_lo(wu[j]);
        const float ga = __uint_as_float(v[j8 * 8 + 2 * j + e]) * rs;
        const float sv = gemm_act_fast<ACT>(g);
        if constexpr (CP) {
          ru[e] = ga * sv;
          rg[e] = 0.f;
        } else {
          const float gh = 0.5f * ga;
          float r;
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(g + 1e-10f));
          ru[e] = gh * sv;
          rg[e] = (sv * r) * (gh * u);
        }
      }
      og[j] = pack_bf16x2(rg[0], rg[1]);
      ou[j] = pack_bf16x2(ru[0], ru[1]);
    }
    epi_store(p.stream_stores != 0, reinterpret_cast<uint4*>(p.gated_out + goff + j8 * 8), make_uint4(og[0], og[1], og[2], og[3]));
    epi_store(p.stream_stores != 0, reinterpret_cast<uint4*>(p.gated_out + goff + uo + j8 * 8), make_uint4(ou[0], ou[1], ou[2], ou[3]));
  }
}

__device__ __forceinline__ void gemm_epilogue_gated_bwd(const GemmParams& p, const uint32_t (&v)[32], int m, float rs, int n0) {
  if (p.gated_cp) {
    if (p.gated_act == LRP_ACT_SILU) gemm_epilogue_gated_bwd_t<LRP_ACT_SILU, true>(p, v, m, rs, n0);
    else if (p.gated_act == LRP_ACT_GELU_TANH) gemm_epilogue_gated_bwd_t<LRP_ACT_GELU_TANH, true>(p, v, m, rs, n0);
    else gemm_epilogue_gated_bwd_t<LRP_ACT_GELU_ERF, true>(p, v, m, rs, n0);
  } else {
    if (p.gated_act == LRP_ACT_SILU) gemm_epilogue_gated_bwd_t<LRP_ACT_SILU, false>(p, v, m, rs, n0);
    else if (p.gated_act == LRP_ACT_GELU_TANH) gemm_epilogue_gated_bwd_t<LRP_ACT_GELU_TANH, false>(p, v, m, rs, n0);
    else gemm_epilogue_gated_bwd_t<LRP_ACT_GELU_ERF, false>(p, v, m, rs, n0);
  }
}

// returns sum_j bf16(out[m, n0 + j]) * delta_o[m, n0 + j] over the chunk when the fused delta is requested, else 0
__device__ __forceinline__ float gemm_epilogue_chunk(const GemmParams& p, const uint32_t (&v)[32], int m, float rs, int64_t row_off,
                                                     int n0) {
  float dot = 0.f;
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
      if (p.out_is_f32) {
        float* o = reinterpret_cast<float*>(p.out) + row_off + n;
        epi_store(p.stream_stores != 0, reinterpret_cast<float4*>(o), make_float4(f[0], f[1], f[2], f[3]));
        epi_store(p.stream_stores != 0, reinterpret_cast<float4*>(o + 4), make_float4(f[4], f[5], f[6], f[7]));
      } else {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + row_off + n;
        epi_store(p.stream_stores != 0, reinterpret_cast<uint4*>(o), make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                                  pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7])));
      }
      if (p.shadow != nullptr) {
        epi_store(p.stream_stores != 0, reinterpret_cast<uint4*>(p.shadow + row_off + n),
                  make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7])));
      }
      if (p.delta_o != nullptr) {
        const uint4 uo = *reinterpret_cast<const uint4*>(p.delta_o + row_off + n);
        const float o8[8] = {bf16_lo(uo.x), bf16_hi(uo.x), bf16_lo(uo.y), bf16_hi(uo.y),
                             bf16_lo(uo.z), bf16_hi(uo.z), bf16_lo(uo.w), bf16_hi(uo.w)};
#pragma unroll
        for (int j = 0; j < 8; ++j) dot = fmaf(__bfloat162float(__float2bfloat16_rn(f[j])), o8[j], dot);
      }
    }
  }
  return dot;
}

// the fused delta: accumulate the chunk's dot product; at the last chunk of a head write delta[b, h, s]
__device__ __forceinline__ void gemm_epilogue_delta(const GemmParams& p, float& acc, float dot, int m, int n0) {
  acc += dot;
  const int n_end = n0 + 32;
  if (n_end % p.delta_D == 0) {
    const int h = n_end / p.delta_D - 1, H = p.N / p.delta_D;
    const int b = m / p.delta_S, s = m - b * p.delta_S;
    p.delta_out[(int64_t(b) * H + h) * p.delta_S + s] = acc;
    acc = 0.f;
  }
}

}  // namespace lrp
