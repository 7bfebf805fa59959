// Flash AttnLRP backward, second generation: two software-pipelined tcgen05 kernels, no atomics.
//
// The first-generation kernel (attn_lrp.cu:attn_bwd_kernel) keeps S, dP, dV, dK in TMEM (all 512 columns) and therefore
// cannot double-buffer anything: MMA, soft-max and the dQ drain serialise (5.5 us per 128x128 tile pair, tensor pipe
// 13-25 % active).  Here the work is split so that every accumulator set leaves room for two S/dP buffers:
//
//   kernel A  dK,dV  (one CTA per 128-key tile, loops over the query tiles of 64 rows of every head of its GQA group)
//       S^T = K Q^T and dP^T = V dO^T  [128 keys x 64 queries]   -> TMEM, 2 buffers x (64 + 64) columns
//       P^T, dS^T (bf16)               -> smem, K-major A operands (keys on rows), 2 buffers
//       dV += P^T dO,  dK += dS^T Q    -> TMEM, 128 + 128 columns                       total 512 columns
//     The transposed formulation keeps M = 128 (keys) for every MMA while the S/dP footprint shrinks with the query
//     tile, and P^T / dS^T land in exactly the layout the dV / dK contractions want.
//   kernel B  dQ     (one CTA per 128-query tile of one head, loops over key tiles of 64)
//       S = Q K^T, dP = dO V^T [128 x 64] -> TMEM 2 x (64 + 64);  dS (bf16) -> smem;  dQ += dS K -> TMEM 128 columns
//     dQ is produced by exactly one CTA: no fp32 atomics, no dq workspace, no finishing kernel.
//   S and dP are recomputed by kernel B (7 instead of 5 MMA units per tile pair), which is cheaper than the
//   serialisation it removes.
//
// Roles (288 threads): warps 0-7 soft-max (warp w: TMEM lane quarter w&3, column half w>>2: each thread owns one
// 32-column chunk of one row per iteration), warp 8 lane 0 issues every TMA load and every tcgen05.mma.
// MMA issue order per iteration it:  [S,dP](it)  then  [accumulate](it-1)  — the tensor pipe always has the next
// S/dP queued while the soft-max warps work on the previous one.
#include <stdio.h>
#include <stdlib.h>
#include "attn_common.cuh"

namespace lrp {

constexpr int V2_THREADS = 288;
constexpr int V2_SM_WARPS = 8;
constexpr int QT = 64;   // kernel A: query rows per iteration;  kernel B: keys per iteration
// TMA stages of the streamed operand pair: 3 (software-pipelined issue order) for D <= 128; head_dim 256 only leaves
// room for one stage, which forces the serial order  accumulate(it-1) -> load(it) -> S/dP(it).
template <int D> struct V2Cfg { static constexpr int NSTG = D > 128 ? 1 : 3; static constexpr bool PIPE = NSTG > 1; };

__device__ __forceinline__ void bar_sync_softmax() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// write 32 bf16 columns [c*32, c*32+32) of row r of a [128 rows][64 cols] K-major 128B-swizzled tile
__device__ __forceinline__ void store_row_half_sw128(uint8_t* tile, int r, int c, const float (&f)[32]) {
  uint8_t* rowp = tile + r * 128;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int chunk = (c * 4 + q) ^ (r & 7);
    *reinterpret_cast<uint4*>(rowp + chunk * 16) =
        make_uint4(pack_bf16x2(f[q * 8 + 0], f[q * 8 + 1]), pack_bf16x2(f[q * 8 + 2], f[q * 8 + 3]),
                   pack_bf16x2(f[q * 8 + 4], f[q * 8 + 5]), pack_bf16x2(f[q * 8 + 6], f[q * 8 + 7]));
  }
}

// =================================================================================================
// kernel A: dK, dV
// =================================================================================================
// MODE 0: dV and dK together (D <= 128);  MODE 1: dV only;  MODE 2: dK only  (D = 256: 256 TMEM columns per accumulator)
template <int D, int MODE>
__global__ void __launch_bounds__(V2_THREADS, 1)
attn_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                     const __grid_constant__ CUtensorMap tmv, const __grid_constant__ CUtensorMap tmdo,
                     const AttnParams p) {
  constexpr int KV_BYTES = ATT_TILE * D * 2;  // [128 keys][D]
  constexpr int QD_BYTES = QT * D * 2;        // [64 queries][D]
  constexpr int PT_BYTES = ATT_TILE * QT * 2; // [128 keys][64 queries]
  constexpr int NSTG = V2Cfg<D>::NSTG;
  constexpr bool PIPE = V2Cfg<D>::PIPE;
  constexpr bool DO_V = MODE != 2, DO_K = MODE != 1;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem;
  uint8_t* sV = sK + KV_BYTES;
  uint8_t* sQ = sV + KV_BYTES;                 // NSTG stages
  uint8_t* sdO = sQ + NSTG * QD_BYTES;         // NSTG stages
  uint8_t* sPt = sdO + NSTG * QD_BYTES;        // 2 buffers
  uint8_t* sdSt = sPt + (DO_V ? 2 * PT_BYTES : 0);   // 2 buffers (each present only if its accumulator is)
  float2* sLD = reinterpret_cast<float2*>(sdSt + (DO_K ? 2 * PT_BYTES : 0));  // [2][64] (lse*log2e, delta*scale) per query
  uint64_t* bars = reinterpret_cast<uint64_t*>(sLD + 2 * QT);
  uint64_t* kv_full = bars;
  uint64_t* qdo_full = bars + 1;             // [NSTG]
  uint64_t* qdo_empty = qdo_full + NSTG;     // [NSTG]
  uint64_t* st_full = qdo_empty + NSTG;      // [2]
  uint64_t* st_empty = st_full + 2;          // [2]  (256 arrivals)
  uint64_t* p_full = st_empty + 2;           // [2]  (256 arrivals)
  uint64_t* p_empty = p_full + 2;            // [2]
  uint64_t* done = p_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int jt, hk, b;
  sched_decode((p.S + ATT_TILE - 1) / ATT_TILE, p.Hkv, p.B, p.sched_group, jt, hk, b);
  const int G = p.H / p.Hkv;
  const int k0 = jt * ATT_TILE;
  const int nq = (p.S + QT - 1) / QT;
  const int i_lo = p.causal ? k0 / QT : 0;
  const int i_hi = p.window > 0 ? min(nq - 1, (k0 + ATT_TILE - 1 + p.window - 1) / QT) : nq - 1;
  const int ni = max(i_hi - i_lo + 1, 0);
  const int n_it = ni * G;
  int kvlo, kvhi;
  kv_bounds(p, b, kvlo, kvhi);

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmq); tma_prefetch_desc(&tmk); tma_prefetch_desc(&tmv); tma_prefetch_desc(&tmdo);
    mbar_init(kv_full, 1);
    for (int i = 0; i < NSTG; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&st_full[i], 1); mbar_init(&st_empty[i], 256);
      mbar_init(&p_full[i], 256); mbar_init(&p_empty[i], 1);
    }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 8) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_dV = tmem_base + 256, tmem_dK = tmem_base + 256 + (MODE == 0 ? D : 0);
  // S^T buffer b at columns [64 b, 64 b + 64), dP^T buffer b at [128 + 64 b, ...)

  if (warp == 8) {
    if (lane == 0 && n_it > 0) {
      auto issue_load = [&](int it) {
        const int g = it / ni, i = i_lo + (it - g * ni), h = hk * G + g, s = it % NSTG;
        if (it >= NSTG) mbar_wait(&qdo_empty[s], ((it / NSTG) - 1) & 1);
        mbar_expect_tx(&qdo_full[s], 2 * QD_BYTES);
        load_tile<D, QT>(sQ + s * QD_BYTES, &tmq, &qdo_full[s], h * D, i * QT, b);
        load_tile<D, QT>(sdO + s * QD_BYTES, &tmdo, &qdo_full[s], h * D, i * QT, b);
      };
      mbar_expect_tx(kv_full, 2 * KV_BYTES);
      load_tile<D>(sK, &tmk, kv_full, hk * D, k0, b);
      load_tile<D>(sV, &tmv, kv_full, hk * D, k0, b);
      auto accumulate = [&](int pit, bool last) {
        const int ps = pit % NSTG, pb = pit & 1;
        mbar_wait(&p_full[pb], (pit >> 1) & 1);
        tc_fence_after();
        // dV += P^T dO ; dK += dS^T Q     (A: [128 keys][64 q] K-major, B: [64 q][D] MN-major)
        if (DO_V) mma_kmn<D, QT, QT * 128>(tmem_dV, smem_u32(sPt + pb * PT_BYTES), smem_u32(sdO + ps * QD_BYTES), pit > 0);
        if (DO_K) mma_kmn<D, QT, QT * 128>(tmem_dK, smem_u32(sdSt + pb * PT_BYTES), smem_u32(sQ + ps * QD_BYTES), pit > 0);
        if (last) {
          tc_commit(done);
        } else {
          tc_commit(&p_empty[pb]);
          tc_commit(&qdo_empty[ps]);
        }
      };
      for (int it = 0; it < min(NSTG - 1, n_it); ++it) issue_load(it);
      mbar_wait(kv_full, 0);
      for (int it = 0; it < n_it; ++it) {
        const int s = it % NSTG, tb = it & 1;
        if (!PIPE) {
          if (it >= 1) accumulate(it - 1, false);
          issue_load(it);
        }
        mbar_wait(&qdo_full[s], (it / NSTG) & 1);
        mbar_wait(&st_empty[tb], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        // S^T = K Q^T ; dP^T = V dO^T      (A: [128 keys][D] K-major, B: [64 queries][D] K-major)
        mma_kk<QT, D, 16384, QT * 128>(tmem_base + tb * QT, smem_u32(sK), smem_u32(sQ + s * QD_BYTES), false);
        mma_kk<QT, D, 16384, QT * 128>(tmem_base + 128 + tb * QT, smem_u32(sV), smem_u32(sdO + s * QD_BYTES), false);
        tc_commit(&st_full[tb]);
        if (PIPE) {
          if (it >= 1) accumulate(it - 1, false);
          // prefetch ahead; its stage is released by the accumulate-MMAs just queued above
          if (it + NSTG - 1 < n_it) issue_load(it + NSTG - 1);
        }
      }
      accumulate(n_it - 1, true);
    }
    __syncwarp();
  } else {
    const int qd = warp & 3, ch = warp >> 2;          // TMEM lane quarter, 32-column half
    const int r = qd * 32 + lane;                      // key row inside the tile
    const int key = k0 + r;
    const uint32_t lane_base = uint32_t(qd * 32) << 16;
    const int sid = threadIdx.x;                       // 0..255
    // (lse*log2e, delta*scale) of the 64 queries of an iteration: fetched from global one iteration ahead (the load
    // latency hides behind the previous iteration), staged through smem because every key row needs all 64 of them
    auto fetch_ld = [&](int it) -> float2 {
      float2 v = make_float2(0.f, 0.f);
      if (sid < QT && it < n_it) {
        const int g = it / ni, i = i_lo + (it - g * ni), h = hk * G + g;
        const int qpos = i * QT + sid;
        if (qpos < p.S) {
          const int64_t idx = (int64_t(b) * p.H + h) * p.S + qpos;
          v = make_float2(p.lse[idx] * LOG2E, p.delta[idx] * p.scale);
        }
      }
      return v;
    };
    float2 ld_next = fetch_ld(0);
    for (int it = 0; it < n_it; ++it) {
      const int g = it / ni, i = i_lo + (it - g * ni);
      const int q0 = i * QT, tb = it & 1;
      if (sid < QT) sLD[tb * QT + sid] = ld_next;
      ld_next = fetch_ld(it + 1);
      bar_sync_softmax();
      // visible query columns for this key: lo <= c <= hi (tile-relative)
      int lo = p.causal ? key - q0 : 0;
      int hi = min(p.S - 1, p.window > 0 ? key + p.window - 1 : p.S - 1) - q0;
      if (key >= kvhi || key < kvlo) { lo = 1; hi = 0; }   // padded (or out-of-range) key: attended by nobody
      const bool need_mask = (p.causal && q0 < k0 + ATT_TILE - 1) || (q0 + QT > p.S) || (k0 + ATT_TILE > kvhi) || (k0 < kvlo) ||
                             (p.window > 0 && q0 + QT - 1 - k0 >= p.window);
      const bool dbgt = p.dbg != nullptr && blockIdx.x == 0 && it < 64 && sid == 0;
      if (dbgt) p.dbg[it * 16 + 8] = clock64();
      mbar_wait(&st_full[tb], (it >> 1) & 1);
      if (dbgt) p.dbg[it * 16 + 9] = clock64();
      tc_fence_after();
      uint32_t vs[32], vd[32];
      tmem_ld32(tmem_base + lane_base + tb * QT + ch * 32, vs);
      tmem_ld32(tmem_base + lane_base + 128 + tb * QT + ch * 32, vd);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&st_empty[tb]);
      if (dbgt) p.dbg[it * 16 + 10] = clock64();
      float fp[32], fd[32];
      const float2* ld = sLD + tb * QT + ch * 32;
#pragma unroll
      for (int x = 0; x < 32; ++x) {
        const float2 l2 = ld[x];
        float pe = ex2_approx(fmaf(__uint_as_float(vs[x]), p.scale_log2, -l2.x));
        if (need_mask) pe = (ch * 32 + x > hi || ch * 32 + x < lo) ? 0.f : pe;
        fp[x] = pe;
        fd[x] = pe * fmaf(__uint_as_float(vd[x]), p.scale, -l2.y);
      }
      if (dbgt) p.dbg[it * 16 + 11] = clock64();
      mbar_wait(&p_empty[tb], ((it >> 1) & 1) ^ 1);
      if (dbgt) p.dbg[it * 16 + 12] = clock64();
      if (DO_V) store_row_half_sw128(sPt + tb * PT_BYTES, r, ch, fp);
      if (DO_K) store_row_half_sw128(sdSt + tb * PT_BYTES, r, ch, fd);
      fence_proxy_async_smem();
      mbar_arrive(&p_full[tb]);
      if (dbgt) p.dbg[it * 16 + 13] = clock64();
    }
    // epilogue: dV, dK of this key tile (column chunks split between the two warps of a lane quarter)
    if (n_it > 0) {
      mbar_wait(done, 0);
      tc_fence_after();
#pragma unroll 1
      for (int which = (DO_V ? 0 : 1); which < (DO_K ? 2 : 1); ++which) {
        const uint32_t src = which == 0 ? tmem_dV : tmem_dK;
        const float sc = which == 0 ? p.inv_v_div : p.inv_k_div;
        __nv_bfloat16* dst = which == 0 ? p.dv + (int64_t(b) * p.S + key) * p.lddv + hk * D
                                        : p.dk + (int64_t(b) * p.S + key) * p.lddk + hk * D;
#pragma unroll 1
        for (int c = ch; c < D / 32; c += 2) {
          uint32_t v[32];
          tmem_ld32(src + lane_base + c * 32, v);
          tmem_ld_wait();
          if (key < p.S) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<uint4*>(dst + c * 32 + q * 8) = make_uint4(
                  pack_bf16x2(__uint_as_float(v[q * 8]) * sc, __uint_as_float(v[q * 8 + 1]) * sc),
                  pack_bf16x2(__uint_as_float(v[q * 8 + 2]) * sc, __uint_as_float(v[q * 8 + 3]) * sc),
                  pack_bf16x2(__uint_as_float(v[q * 8 + 4]) * sc, __uint_as_float(v[q * 8 + 5]) * sc),
                  pack_bf16x2(__uint_as_float(v[q * 8 + 6]) * sc, __uint_as_float(v[q * 8 + 7]) * sc));
          }
        }
      }
    } else if (key < p.S) {
      for (int c = ch * 8; c < D; c += 16) {
        if (DO_V) *reinterpret_cast<uint4*>(p.dv + (int64_t(b) * p.S + key) * p.lddv + hk * D + c) = make_uint4(0, 0, 0, 0);
        if (DO_K) *reinterpret_cast<uint4*>(p.dk + (int64_t(b) * p.S + key) * p.lddk + hk * D + c) = make_uint4(0, 0, 0, 0);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

// =================================================================================================
// kernel B: dQ
// =================================================================================================
struct DqParams {
  __nv_bfloat16* dq;
  int64_t lddq;
  float inv_q_div;
};

template <int D>
__global__ void __launch_bounds__(V2_THREADS, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                   const __grid_constant__ CUtensorMap tmv, const __grid_constant__ CUtensorMap tmdo,
                   const AttnParams p, const DqParams dqp) {
  constexpr int QD_BYTES = ATT_TILE * D * 2;  // [128 queries][D]
  constexpr int KV_BYTES = QT * D * 2;        // [64 keys][D]
  constexpr int DS_BYTES = ATT_TILE * QT * 2; // [128 queries][64 keys]
  constexpr int NSTG = V2Cfg<D>::NSTG;
  constexpr bool PIPE = V2Cfg<D>::PIPE;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + QD_BYTES;
  uint8_t* sK = sdO + QD_BYTES;               // NSTG stages
  uint8_t* sV = sK + NSTG * KV_BYTES;         // NSTG stages
  uint8_t* sdS = sV + NSTG * KV_BYTES;        // 2 buffers
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + 2 * DS_BYTES);
  uint64_t* qdo_full = bars;
  uint64_t* kv_full = bars + 1;               // [NSTG]
  uint64_t* kv_empty = kv_full + NSTG;        // [NSTG]
  uint64_t* st_full = kv_empty + NSTG;        // [2]
  uint64_t* st_empty = st_full + 2;           // [2] (256)
  uint64_t* p_full = st_empty + 2;            // [2] (256)
  uint64_t* p_empty = p_full + 2;             // [2]
  uint64_t* done = p_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqt = (p.S + ATT_TILE - 1) / ATT_TILE;
  int trank, h, b;
  sched_decode(nqt, p.H, p.B, p.sched_group, trank, h, b);
  const int qt = nqt - 1 - trank;
  const int hk = h / (p.H / p.Hkv);
  const int q0 = qt * ATT_TILE;
  const int nkv = (p.S + QT - 1) / QT;
  const int j_hi = p.causal ? min((q0 + ATT_TILE - 1) / QT, nkv - 1) : nkv - 1;
  const int j_lo = p.window > 0 ? max(0, q0 - p.window + 1) / QT : 0;
  const int n = j_hi - j_lo + 1;
  int kvlo, kvhi;
  kv_bounds(p, b, kvlo, kvhi);

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmq); tma_prefetch_desc(&tmk); tma_prefetch_desc(&tmv); tma_prefetch_desc(&tmdo);
    mbar_init(qdo_full, 1);
    for (int i = 0; i < NSTG; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&st_full[i], 1); mbar_init(&st_empty[i], 256);
      mbar_init(&p_full[i], 256); mbar_init(&p_empty[i], 1);
    }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 8) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_dQ = tmem_base + 256;   // S buffers at [64 b], dP buffers at [128 + 64 b]

  if (warp == 8) {
    if (lane == 0) {
      auto issue_load = [&](int jj) {
        const int s = jj % NSTG;
        if (jj >= NSTG) mbar_wait(&kv_empty[s], ((jj / NSTG) - 1) & 1);
        mbar_expect_tx(&kv_full[s], 2 * KV_BYTES);
        load_tile<D, QT>(sK + s * KV_BYTES, &tmk, &kv_full[s], hk * D, (j_lo + jj) * QT, b);
        load_tile<D, QT>(sV + s * KV_BYTES, &tmv, &kv_full[s], hk * D, (j_lo + jj) * QT, b);
      };
      mbar_expect_tx(qdo_full, 2 * QD_BYTES);
      load_tile<D>(sQ, &tmq, qdo_full, h * D, q0, b);
      load_tile<D>(sdO, &tmdo, qdo_full, h * D, q0, b);
      auto accumulate = [&](int pj, bool last) {
        const int ps = pj % NSTG, pb = pj & 1;
        mbar_wait(&p_full[pb], (pj >> 1) & 1);
        tc_fence_after();
        // dQ += dS K      (A: [128 q][64 keys] K-major, B: [64 keys][D] MN-major)
        mma_kmn<D, QT, QT * 128>(tmem_dQ, smem_u32(sdS + pb * DS_BYTES), smem_u32(sK + ps * KV_BYTES), pj > 0);
        if (last) {
          tc_commit(done);
        } else {
          tc_commit(&p_empty[pb]);
          tc_commit(&kv_empty[ps]);
        }
      };
      for (int jj = 0; jj < min(NSTG - 1, n); ++jj) issue_load(jj);
      mbar_wait(qdo_full, 0);
      for (int jj = 0; jj < n; ++jj) {
        const int s = jj % NSTG, tb = jj & 1;
        if (!PIPE) {
          if (jj >= 1) accumulate(jj - 1, false);
          issue_load(jj);
        }
        mbar_wait(&kv_full[s], (jj / NSTG) & 1);
        mbar_wait(&st_empty[tb], ((jj >> 1) & 1) ^ 1);
        tc_fence_after();
        // S = Q K^T ; dP = dO V^T      (A: [128 q][D] K-major, B: [64 keys][D] K-major)
        mma_kk<QT, D, 16384, QT * 128>(tmem_base + tb * QT, smem_u32(sQ), smem_u32(sK + s * KV_BYTES), false);
        mma_kk<QT, D, 16384, QT * 128>(tmem_base + 128 + tb * QT, smem_u32(sdO), smem_u32(sV + s * KV_BYTES), false);
        tc_commit(&st_full[tb]);
        if (PIPE) {
          if (jj >= 1) accumulate(jj - 1, false);
          if (jj + NSTG - 1 < n) issue_load(jj + NSTG - 1);
        }
      }
      accumulate(n - 1, true);
    }
    __syncwarp();
  } else {
    const int qd = warp & 3, ch = warp >> 2;
    const int r = qd * 32 + lane;
    const int qpos = q0 + r;
    const uint32_t lane_base = uint32_t(qd * 32) << 16;
    const bool valid = qpos < p.S;
    float lse2 = 0.f, delta_s = 0.f;
    if (valid) {
      const int64_t idx = (int64_t(b) * p.H + h) * p.S + qpos;
      lse2 = p.lse[idx] * LOG2E;
      delta_s = p.delta[idx] * p.scale;
    }
    const bool row_ok = valid && lse2 != -INFINITY;
    if (!row_ok) lse2 = 0.f;
    for (int jj = 0; jj < n; ++jj) {
      const int kbase = (j_lo + jj) * QT, tb = jj & 1;
      int lo, hi;
      row_window(qpos, kbase, kvlo, kvhi, p.causal, p.window, lo, hi);
      if (!row_ok) { lo = 1; hi = 0; }
      const bool need_mask = (p.causal && kbase + QT - 1 > q0) || (kbase + QT > kvhi) || (kbase < kvlo) || (q0 + ATT_TILE > p.S) ||
                             (p.window > 0 && q0 + ATT_TILE - 1 - kbase >= p.window);
      mbar_wait(&st_full[tb], (jj >> 1) & 1);
      tc_fence_after();
      uint32_t vs[32], vd[32];
      tmem_ld32(tmem_base + lane_base + tb * QT + ch * 32, vs);
      tmem_ld32(tmem_base + lane_base + 128 + tb * QT + ch * 32, vd);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&st_empty[tb]);
      float fd[32];
#pragma unroll
      for (int x = 0; x < 32; ++x) {
        float pe = ex2_approx(fmaf(__uint_as_float(vs[x]), p.scale_log2, -lse2));
        if (need_mask) pe = (ch * 32 + x > hi || ch * 32 + x < lo) ? 0.f : pe;
        fd[x] = pe * fmaf(__uint_as_float(vd[x]), p.scale, -delta_s);
      }
      mbar_wait(&p_empty[tb], ((jj >> 1) & 1) ^ 1);
      store_row_half_sw128(sdS + tb * DS_BYTES, r, ch, fd);
      fence_proxy_async_smem();
      mbar_arrive(&p_full[tb]);
    }
    mbar_wait(done, 0);
    tc_fence_after();
    __nv_bfloat16* dst = dqp.dq + (int64_t(b) * p.S + qpos) * dqp.lddq + h * D;
    const float sc = dqp.inv_q_div;
#pragma unroll 1
    for (int c = ch; c < D / 32; c += 2) {
      uint32_t v[32];
      tmem_ld32(tmem_dQ + lane_base + c * 32, v);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(dst + c * 32 + q * 8) = make_uint4(
              pack_bf16x2(__uint_as_float(v[q * 8]) * sc, __uint_as_float(v[q * 8 + 1]) * sc),
              pack_bf16x2(__uint_as_float(v[q * 8 + 2]) * sc, __uint_as_float(v[q * 8 + 3]) * sc),
              pack_bf16x2(__uint_as_float(v[q * 8 + 4]) * sc, __uint_as_float(v[q * 8 + 5]) * sc),
              pack_bf16x2(__uint_as_float(v[q * 8 + 6]) * sc, __uint_as_float(v[q * 8 + 7]) * sc));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

template <int D, int MODE>
static int dkdv_smem() {
  return 2 * ATT_TILE * D * 2 + 2 * V2Cfg<D>::NSTG * QT * D * 2 + (MODE == 0 ? 4 : 2) * ATT_TILE * QT * 2 + 2 * QT * 8 + 256;
}
template <int D>
static int dq_smem() { return 2 * ATT_TILE * D * 2 + 2 * V2Cfg<D>::NSTG * QT * D * 2 + 2 * ATT_TILE * QT * 2 + 256; }

template <int D, int MODE>
static int launch_dkdv(const CUtensorMap& tq64, const CUtensorMap& tk128, const CUtensorMap& tv128, const CUtensorMap& tdo64,
                       const AttnParams& p, cudaStream_t st) {
  auto ka = attn_bwd_dkdv_kernel<D, MODE>;
  static bool done = false;
  if (!done) {
    cudaError_t ce = cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, dkdv_smem<D, MODE>());
    if (ce != cudaSuccess) return set_error(LRP_ERR_CUDA, cudaGetErrorString(ce));
    done = true;
  }
  const int tiles = (p.S + ATT_TILE - 1) / ATT_TILE;
  ka<<<dim3(tiles * p.Hkv * p.B), V2_THREADS, dkdv_smem<D, MODE>(), st>>>(tq64, tk128, tv128, tdo64, p);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

template <int D>
static int launch_v2(const CUtensorMap& tq128, const CUtensorMap& tk128, const CUtensorMap& tv128, const CUtensorMap& tdo128,
                     const CUtensorMap& tq64, const CUtensorMap& tk64, const CUtensorMap& tv64, const CUtensorMap& tdo64,
                     const AttnParams& p, const DqParams& dqp, cudaStream_t st) {
  // kernel A streams 64-row query/dO tiles against a resident 128-key tile; kernel B the other way round
  if (D <= 128) {
    if (int e = launch_dkdv<D, 0>(tq64, tk128, tv128, tdo64, p, st)) return e;
  } else {
    if (int e = launch_dkdv<D, 1>(tq64, tk128, tv128, tdo64, p, st)) return e;   // dV pass
    if (int e = launch_dkdv<D, 2>(tq64, tk128, tv128, tdo64, p, st)) return e;   // dK pass
  }
  auto kb = attn_bwd_dq_kernel<D>;
  static bool done = false;
  if (!done) {
    cudaError_t ce = cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, dq_smem<D>());
    if (ce != cudaSuccess) return set_error(LRP_ERR_CUDA, cudaGetErrorString(ce));
    done = true;
  }
  const int tiles = (p.S + ATT_TILE - 1) / ATT_TILE;
  if (dqp.inv_q_div != 0.f) {
    kb<<<dim3(tiles * p.H * p.B), V2_THREADS, dq_smem<D>(), st>>>(tq128, tk64, tv64, tdo128, p, dqp);
    LRP_CHECK_LAUNCH();
  } else {
    // CP-LRP: q is detached, dQ = 0 (strided rows of the caller's buffer)
    cudaError_t ce = cudaMemset2DAsync(dqp.dq, size_t(dqp.lddq) * 2, 0, size_t(p.H) * D * 2, size_t(p.B) * p.S, st);
    if (ce != cudaSuccess) return set_error(LRP_ERR_CUDA, cudaGetErrorString(ce));
    note_launch();
  }
  return LRP_OK;
}

// host entry used by lrp_attn_bwd (attn_lrp.cu)
int attn_bwd_v2(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv, const void* d_o,
                const float* lse, const float* delta, void* dq, void* dk, void* dv, int64_t lddq, int64_t lddk, int64_t lddv,
                const int32_t* kv_range, int B, int S, int H, int Hkv, int D, float scale, int causal, int window, float q_div,
                float k_div, float v_div, cudaStream_t st) {
  CUtensorMap tq128, tk128, tv128, tdo128, tq64, tk64, tv64, tdo64;
  const int64_t HD = int64_t(H) * D;
  struct { CUtensorMap* m; const void* ptr; uint64_t width; int64_t ld; uint32_t rows; } maps[] = {
      {&tq128, q, uint64_t(H) * D, ldq, ATT_TILE}, {&tk128, k, uint64_t(Hkv) * D, ldk, ATT_TILE},
      {&tv128, v, uint64_t(Hkv) * D, ldv, ATT_TILE}, {&tdo128, d_o, uint64_t(HD), HD, ATT_TILE},
      {&tq64, q, uint64_t(H) * D, ldq, QT},          {&tk64, k, uint64_t(Hkv) * D, ldk, QT},
      {&tv64, v, uint64_t(Hkv) * D, ldv, QT},        {&tdo64, d_o, uint64_t(HD), HD, QT}};
  for (auto& m : maps)
    if (int e = make_tmap_3d_bf16(m.m, m.ptr, m.width, S, B, m.ld, uint64_t(S) * m.ld, 64, m.rows)) return e;
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.S = S; p.H = H; p.Hkv = Hkv; p.D = D;
  p.scale = scale; p.scale_log2 = scale * LOG2E;
  p.causal = causal; p.window = window;
  p.sched_group = sched_group_default();
  p.lse = const_cast<float*>(lse);
  p.delta = delta;
  p.dk = reinterpret_cast<__nv_bfloat16*>(dk);
  p.dv = reinterpret_cast<__nv_bfloat16*>(dv);
  p.lddk = lddk; p.lddv = lddv;
  p.inv_k_div = k_div > 0.f ? 1.f / k_div : 0.f;
  p.inv_v_div = v_div > 0.f ? 1.f / v_div : 0.f;
  p.kv_range = kv_range;
  static const bool want_dbg = getenv("LRP_ATTN_DEBUG") != nullptr;
  long long* dbg_dev = nullptr;
  if (want_dbg && B * S >= 4096) {
    cudaMalloc(&dbg_dev, 64 * 16 * sizeof(long long));
    cudaMemset(dbg_dev, 0, 64 * 16 * sizeof(long long));
    p.dbg = dbg_dev;
  }
  DqParams dqp;
  dqp.dq = reinterpret_cast<__nv_bfloat16*>(dq);
  dqp.lddq = lddq;
  dqp.inv_q_div = q_div > 0.f ? 1.f / q_div : 0.f;
  const int rc = D == 256   ? launch_v2<256>(tq128, tk128, tv128, tdo128, tq64, tk64, tv64, tdo64, p, dqp, st)
                 : D == 128 ? launch_v2<128>(tq128, tk128, tv128, tdo128, tq64, tk64, tv64, tdo64, p, dqp, st)
                            : launch_v2<64>(tq128, tk128, tv128, tdo128, tq64, tk64, tv64, tdo64, p, dqp, st);
  if (dbg_dev != nullptr) {
    static bool printed = false;
    long long h[64 * 16];
    cudaDeviceSynchronize();
    cudaMemcpy(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(dbg_dev);
    if (!printed) {
      printed = true;
      printf("it | ctl: qdo_full st_empty mma1 p_full mma2 prefetch (delta cycles) | thr: wait_st ld compute wait_pe store | iter\n");
      for (int it = 1; it < 40; ++it) {
        const long long* r = h + it * 16;
        printf("%2d | %5lld %5lld %5lld %5lld %5lld %5lld | %5lld %5lld %5lld %5lld %5lld | %6lld\n", it, r[1] - r[0], r[2] - r[1],
               r[3] - r[2], r[4] - r[3], r[5] - r[4], r[6] - r[5], r[9] - r[8], r[10] - r[9], r[11] - r[10], r[12] - r[11],
               r[13] - r[12], r[0] - (h + (it - 1) * 16)[0]);
      }
    }
  }
  return rc;
}

}  // namespace lrp
