// Shared pieces of the flash AttnLRP kernels: parameters, masking helpers, swizzled row stores and the single-thread
// UMMA issue helpers over [rows][64-column block] 128B-swizzled operand tiles.
#pragma once
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include "ptx_sm100.cuh"
#include "lrp_internal.h"

namespace lrp {

constexpr int ATT_TILE = 128;       // query rows per CTA tile == keys per tile
constexpr int ATT_THREADS = 160;   // forward: 4 soft-max warps + 1 TMA/MMA warp
constexpr int BWD_THREADS = 288;   // backward: 8 soft-max warps + 1 TMA/MMA warp
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

struct AttnParams {
  int B, S, H, Hkv, D;
  float scale, scale_log2;
  int causal, window;
  // forward
  __nv_bfloat16* o;   // [B,S,H,D]
  float* lse;         // [B,H,S]
  // backward
  const float* delta; // [B,H,S]
  float* dq_acc;      // [B,S,H,D] fp32
  __nv_bfloat16* dk;  // [B,S,Hkv,D] strided by lddk
  __nv_bfloat16* dv;
  int64_t lddk, lddv;
  float inv_k_div, inv_v_div;
  long long* dbg;  // optional timeline buffer (debug builds of the pipeline analysis), NULL otherwise
  // optional per-sequence range of VALID keys [kv_range[2b], kv_range[2b+1]) (left / right padding of a batch of prompts of
  // different lengths): keys outside it are masked for every query.  NULL = all S keys valid.
  const int* kv_range;
  // CTA order of the tile-loop kernels: the grid is 1-D and walks groups of `sched_group` sequences tile-major, so the
  // heaviest causal tiles of a group are dispatched first (longest-processing-time order; 0 = one group = whole batch)
  int sched_group;
};

// linear block id -> (tile rank t, y, z) under the order above.  `t` counts tiles in dispatch order; the caller maps it to the
// heaviest-first tile index (key tile t for the backward, query tile tiles-1-t for the forward).
__device__ __forceinline__ void sched_decode(int tiles, int Y, int Z, int group, int& t, int& y, int& z) {
  if (group < 0) {   // first-generation order (tile index fastest): kept for A/B runs, LRP_ATTN_SCHED_GROUP=-1
    t = int(blockIdx.x) % tiles;
    y = (int(blockIdx.x) / tiles) % Y;
    z = int(blockIdx.x) / (tiles * Y);
    return;
  }
  if (group == 0 || group > Z) group = Z;
  const int per_group = tiles * Y * group;
  const int ngroups = (Z + group - 1) / group;
  const int g0 = min(int(blockIdx.x) / per_group, ngroups - 1);
  const int r = int(blockIdx.x) - g0 * per_group;
  const int zc = min(group, Z - g0 * group);
  const int per_tile = Y * zc;
  t = r / per_tile;
  const int rr = r - t * per_tile;
  z = g0 * group + rr / Y;
  y = rr - (rr / Y) * Y;
}

// attn_bwd_ws.cu: warp-specialised backward (head_dim 128, no sliding window)
int attn_bwd_ws_launch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
                       const CUtensorMap& tdq, const AttnParams& p, cudaStream_t st);

inline int sched_group_default() {
  static int g = -2;
  if (g == -2) {
    const char* e = getenv("LRP_ATTN_SCHED_GROUP");
    g = e != nullptr ? atoi(e) : 2;
  }
  return g;
}

__device__ __forceinline__ void kv_bounds(const AttnParams& p, int b, int& kvlo, int& kvhi) {
  kvlo = 0;
  kvhi = p.S;
  if (p.kv_range != nullptr) {
    kvlo = max(0, p.kv_range[2 * b]);
    kvhi = min(p.S, p.kv_range[2 * b + 1]);
  }
}

__device__ __forceinline__ bool is_masked(int qpos, int kpos, int S, int causal, int window) {
  if (kpos >= S) return true;
  if (causal && kpos > qpos) return true;
  if (window > 0 && qpos - kpos >= window) return true;
  return false;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Per-row column window [lo, hi] (tile-relative) of keys that are NOT masked:
//   key (kbase + c) is visible iff lo <= c <= hi.
// kvlo / kvhi: range of valid keys of this sequence (0 / S without padding)
__device__ __forceinline__ void row_window(int qpos, int kbase, int kvlo, int kvhi, int causal, int window, int& lo, int& hi) {
  hi = kvhi - 1 - kbase;
  if (causal) hi = min(hi, qpos - kbase);
  lo = kvlo - kbase;
  if (window > 0) lo = max(lo, qpos - window + 1 - kbase);
}

template <bool MASK>
__device__ __forceinline__ float chunk_max(const uint32_t (&v)[32], int c0, int lo, int hi) {
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    float t = __uint_as_float(v[i]);
    if (MASK) t = (c0 + i > hi || c0 + i < lo) ? -INFINITY : t;
    mx = fmaxf(mx, t);
  }
  return mx;
}

// f[i] = 2^(s*scale_log2 - msub) (0 where masked); returns the chunk's sum
// 2^x on the FMA pipe (no MUFU): x = n + f with n = rint(x) taken from the low mantissa bits of x + 1.5*2^23, 2^f by a degree-3
// polynomial on [-0.5, 0.5] (max relative error 1.0e-4 = 1/40 of the bf16 rounding the result goes through; coefficient
// study in profiles/microbench/exp2_poly_accuracy.txt), exponent added by integer arithmetic.  Valid for x <= ~120.
__device__ __forceinline__ float ex2_poly3(float x) {
  x = fmaxf(x, -125.f);
  const float magic = 12582912.f;
  const float t = x + magic;
  const float f = x - (t - magic);
  const float pl = fmaf(fmaf(fmaf(0.05592204f, f, 0.24264008f), f, 0.69312102f), f, 0.99992448f);
  return __int_as_float(__float_as_int(pl) + (__float_as_int(t) << 23));
}

// POLY > 0: every POLY-th exponential of the chunk is evaluated by ex2_poly3 instead of MUFU.EX2 (16 lanes/clk/SM), which is
// what bounds soft-max pass A of the backward (16384 exponentials per 128x128 tile = 1024 MUFU cycles)
template <bool MASK, int POLY = 0>
__device__ __forceinline__ float chunk_exp(const uint32_t (&v)[32], float (&f)[32], float scale_log2, float msub, int c0,
                                           int lo, int hi) {
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float arg = fmaf(__uint_as_float(v[i]), scale_log2, -msub);
    float pe = (POLY > 0 && (i % (POLY > 0 ? POLY : 1)) == POLY - 1) ? ex2_poly3(arg) : ex2_approx(arg);
    if (MASK) pe = (c0 + i > hi || c0 + i < lo) ? 0.f : pe;
    f[i] = pe;
    sum += pe;
  }
  return sum;
}

// backward: P = 2^(s*scale_log2 - lse2), dS = P * (dP*scale - delta*scale)
template <bool MASK>
__device__ __forceinline__ void chunk_p_ds(const uint32_t (&vs)[32], const uint32_t (&vd)[32], float (&fp)[32], float (&fd)[32],
                                           float scale_log2, float lse2, float scale, float delta_s, int c0, int lo, int hi) {
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    float pe = ex2_approx(fmaf(__uint_as_float(vs[i]), scale_log2, -lse2));
    if (MASK) pe = (c0 + i > hi || c0 + i < lo) ? 0.f : pe;
    fp[i] = pe;
    fd[i] = pe * fmaf(__uint_as_float(vd[i]), scale, -delta_s);
  }
}

// write 32 consecutive bf16 columns [c*32, c*32+32) of row r into a [128][128] tile stored as two
// [128 rows][64 cols] 128B-swizzled blocks
__device__ __forceinline__ void store_row_chunk_sw128(uint8_t* tile, int r, int c, const float (&f)[32]) {
  uint8_t* rowp = tile + (c >> 1) * 16384 + r * 128;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int chunk = ((c & 1) * 4 + q) ^ (r & 7);
    *reinterpret_cast<uint4*>(rowp + chunk * 16) =
        make_uint4(pack_bf16x2(f[q * 8 + 0], f[q * 8 + 1]), pack_bf16x2(f[q * 8 + 2], f[q * 8 + 3]),
                   pack_bf16x2(f[q * 8 + 4], f[q * 8 + 5]), pack_bf16x2(f[q * 8 + 6], f[q * 8 + 7]));
  }
}

// same, from 16 already packed bf16x2 words
__device__ __forceinline__ void store_row_chunk_sw128_pk(uint8_t* tile, int r, int c, const uint32_t (&pk)[16]) {
  uint8_t* rowp = tile + (c >> 1) * 16384 + r * 128;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int chunk = ((c & 1) * 4 + q) ^ (r & 7);
    *reinterpret_cast<uint4*>(rowp + chunk * 16) = make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
  }
}

// ---- UMMA issue helpers (single thread) -------------------------------------------------------
// All loops are fully unrolled with compile-time offsets and (lo, hi) descriptor halves: the one issuing thread spends
// a few instructions per tcgen05.mma (the first version rebuilt both 64-bit descriptors per instruction, ~60 cycles
// each, which made the N=64 kernels issue-bound — profiles/r01_attn_bwd_v2_timeline.txt).
//
// C[128 x N] (+)= A_kmajor[128 x KTOT] * B_kmajor[N x KTOT]^T; tiles as [rows][64-column blocks]:
//   A_BLK / B_BLK = byte distance between consecutive 64-column blocks of the A / B tile (= rows * 128 B)
template <int N, int KTOT, int A_BLK = 16384, int B_BLK = 16384>
__device__ __forceinline__ void mma_kk(uint32_t tmem_d, uint32_t a_base, uint32_t b_base, bool acc_first) {
  constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
  constexpr uint32_t hi = sdesc_hi(1024);
  const uint32_t a_lo = sdesc_lo(a_base, 16), b_lo = sdesc_lo(b_base, 16);
#pragma unroll
  for (int kk = 0; kk < KTOT / 16; ++kk) {
    const uint32_t aoff = ((kk >> 2) * A_BLK + (kk & 3) * 32) >> 4;
    const uint32_t boff = ((kk >> 2) * B_BLK + (kk & 3) * 32) >> 4;
    tc_mma_ss_lohi(tmem_d, a_lo + aoff, hi, b_lo + boff, hi, idesc, (kk > 0 || acc_first) ? 1u : 0u);
  }
}
// C[128 x N] (+)= A_kmajor[128 x KTOT] * B_mnmajor[KTOT(k) x N];  B_LBO = byte distance between 64-column blocks of B
template <int N, int KTOT = 128, int B_LBO = 16384>
__device__ __forceinline__ void mma_kmn(uint32_t tmem_d, uint32_t a_base, uint32_t b_base, bool acc_first) {
  constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 1);
  constexpr uint32_t hi = sdesc_hi(1024);
  const uint32_t a_lo = sdesc_lo(a_base, 16), b_lo = sdesc_lo(b_base, B_LBO);
#pragma unroll
  for (int kk = 0; kk < KTOT / 16; ++kk) {
    const uint32_t aoff = ((kk >> 2) * 16384 + (kk & 3) * 32) >> 4;
    tc_mma_ss_lohi(tmem_d, a_lo + aoff, hi, b_lo + ((kk * 2048) >> 4), hi, idesc, (kk > 0 || acc_first) ? 1u : 0u);
  }
}
// C[128 x N] (+)= A_mnmajor[128(k) x 128(m)]^T * B_mnmajor[128(k) x N]
template <int N>
__device__ __forceinline__ void mma_mnmn(uint32_t tmem_d, uint32_t a_base, uint32_t b_base, bool acc_first) {
  constexpr uint32_t idesc = make_idesc_bf16(128, N, 1, 1);
  constexpr uint32_t hi = sdesc_hi(1024);
  const uint32_t a_lo = sdesc_lo(a_base, 16384), b_lo = sdesc_lo(b_base, 16384);
#pragma unroll
  for (int kk = 0; kk < 8; ++kk)
    tc_mma_ss_lohi(tmem_d, a_lo + ((kk * 2048) >> 4), hi, b_lo + ((kk * 2048) >> 4), hi, idesc, (kk > 0 || acc_first) ? 1u : 0u);
}

template <int D, int ROWS = 128>
__device__ __forceinline__ void load_tile(uint8_t* dst, const CUtensorMap* tm, uint64_t* bar, int col0, int row0, int b) {
#pragma unroll
  for (int kb = 0; kb < D / 64; ++kb) tma_load_3d(dst + kb * (ROWS * 128), tm, bar, col0 + kb * 64, row0, b);
}

}  // namespace lrp
