// Flash AttnLRP for sm_100a: attention forward and the LRP (Gradient x Input) backward on tcgen05 tensor
// cores with TMEM accumulators and TMA-staged, 128B-swizzled operand tiles.  No [B,H,S,S] tensor exists in HBM.
//
// Rule being implemented (reference: lxt/efficient/patches.py:171-203 `wrap_attention_forward`): the two
// attention matmuls get the uniform rule, which in GxI space is dQ/4, dK/4, dV/2 around an ordinary
// softmax-attention backward (the softmax Deep-Taylor rule of lxt/explicit/functional.py:276-322 is the
// plain softmax backward in GxI space).  CP-LRP (patches.py:249-258) is q_div = k_div = 0.
//
// Every operand tile lives in shared memory as 128 rows x (D/64) blocks of 64 bf16 (= 128 B, one swizzle
// row).  The same bytes serve as a K-major operand (contraction along the row) or as an MN-major operand
// (contraction across rows) purely by the choice of UMMA descriptor, so P and dS are written once by the
// softmax threads and consumed three times (dV = P^T dO, dK = dS^T Q, dQ = dS K).
//
// Thread roles (160 threads): warps 0-3 own one query row each (row r <-> TMEM lane r, so soft-max needs no
// shuffles); warp 4 lane 0 issues all TMA loads and all tcgen05.mma.
#include <stdio.h>
#include <stdlib.h>
#include "attn_common.cuh"

namespace lrp {

// =================================================================================================
// forward
// =================================================================================================
template <int D, int BN>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                const __grid_constant__ CUtensorMap tmv, const AttnParams p) {
  // Two CTAs are resident per SM (<= 113 KiB smem, 256 TMEM columns each): while one CTA's soft-max warps work,
  // the other CTA's S / P.V MMAs occupy the tensor pipe.
  constexpr int Q_BYTES = ATT_TILE * D * 2;
  constexpr int KV_BYTES = BN * D * 2;
  constexpr int P_BYTES = ATT_TILE * BN * 2;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;        // 2 stages
  uint8_t* sV = sK + 2 * KV_BYTES;   // 2 stages
  uint8_t* sP = sV + 2 * KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* v_full = bars + 3;    // [2]
  uint64_t* kv_empty = bars + 5;  // [2]
  uint64_t* s_full = bars + 7;
  uint64_t* p_full = bars + 8;
  uint64_t* o_full = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqt = (p.S + ATT_TILE - 1) / ATT_TILE;
  int trank, h, b;
  sched_decode(nqt, p.H, p.B, p.sched_group, trank, h, b);
  const int qt = nqt - 1 - trank;  // heaviest causal tiles first
  const int hk = h / (p.H / p.Hkv);
  const int q0 = qt * ATT_TILE;
  const int nkv = (p.S + BN - 1) / BN;
  const int j_hi = p.causal ? min((q0 + ATT_TILE - 1) / BN, nkv - 1) : nkv - 1;
  const int j_lo = p.window > 0 ? max(0, q0 - p.window + 1) / BN : 0;
  const int n = j_hi - j_lo + 1;
  int kvlo, kvhi;
  kv_bounds(p, b, kvlo, kvhi);

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmq); tma_prefetch_desc(&tmk); tma_prefetch_desc(&tmv);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&k_full[i], 1); mbar_init(&v_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 4) { tmem_alloc(tmem_slot, D > 128 ? 512 : 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(q_full, Q_BYTES);
      load_tile<D>(sQ, &tmq, q_full, h * D, q0, b);
      mbar_expect_tx(&k_full[0], KV_BYTES);
      load_tile<D, BN>(sK, &tmk, &k_full[0], hk * D, j_lo * BN, b);
      mbar_expect_tx(&v_full[0], KV_BYTES);
      load_tile<D, BN>(sV, &tmv, &v_full[0], hk * D, j_lo * BN, b);
      for (int jj = 0; jj < n; ++jj) {
        const int st = jj & 1;
        if (jj == 0) mbar_wait(q_full, 0);
        mbar_wait(&k_full[st], (jj >> 1) & 1);
        tc_fence_after();
        mma_kk<BN, D, 16384, BN * 128>(tmem_S, smem_u32(sQ), smem_u32(sK + st * KV_BYTES), false);   // S = Q K^T
        tc_commit(s_full);
        if (jj + 1 < n) {   // prefetch the next K/V tile (after S has been issued: the wait below is for P.V of tile jj-1)
          const int ns = st ^ 1;
          if (jj >= 1) mbar_wait(&kv_empty[ns], ((jj - 1) >> 1) & 1);
          mbar_expect_tx(&k_full[ns], KV_BYTES);
          load_tile<D, BN>(sK + ns * KV_BYTES, &tmk, &k_full[ns], hk * D, (j_lo + jj + 1) * BN, b);
          mbar_expect_tx(&v_full[ns], KV_BYTES);
          load_tile<D, BN>(sV + ns * KV_BYTES, &tmv, &v_full[ns], hk * D, (j_lo + jj + 1) * BN, b);
        }
        mbar_wait(p_full, jj & 1);
        mbar_wait(&v_full[st], (jj >> 1) & 1);
        tc_fence_after();
        mma_kmn<D, BN, BN * 128>(tmem_O, smem_u32(sP), smem_u32(sV + st * KV_BYTES), jj > 0);        // O += P V
        tc_commit(&kv_empty[st]);
        tc_commit(o_full);
      }
    }
    __syncwarp();
  } else {
    const int r = warp * 32 + lane;
    const int qpos = q0 + r;
    const uint32_t lane_base = uint32_t(warp * 32) << 16;
    float m_used = -INFINITY, l = 0.f;
    for (int jj = 0; jj < n; ++jj) {
      const int j = j_lo + jj;
      const int kbase = j * BN;
      const bool need_mask = (p.causal && kbase + BN - 1 > q0) || (kbase + BN > kvhi) || (kbase < kvlo) ||
                             (p.window > 0 && q0 + ATT_TILE - 1 - kbase >= p.window);
      mbar_wait(s_full, jj & 1);
      tc_fence_after();
      int lo, hi;
      row_window(qpos, kbase, kvlo, kvhi, p.causal, p.window, lo, hi);
      // pass 1: row maximum of the raw scores (scale > 0, so the order is preserved).  With 64-key tiles the row's 64
      // scores stay in registers for pass 2 (one TMEM read instead of two).
      float mx = -INFINITY;
      uint32_t sv[BN == 64 ? 2 : 1][32];
      if constexpr (BN == 64) {
        tmem_ld32(tmem_S + lane_base, sv[0]);
        tmem_ld32(tmem_S + lane_base + 32, sv[1]);
        tmem_ld_wait();
        mx = need_mask ? fmaxf(chunk_max<true>(sv[0], 0, lo, hi), chunk_max<true>(sv[1], 32, lo, hi))
                       : fmaxf(chunk_max<false>(sv[0], 0, lo, hi), chunk_max<false>(sv[1], 32, lo, hi));
      } else {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(tmem_S + lane_base + c * 32, v);
          tmem_ld_wait();
          mx = fmaxf(mx, need_mask ? chunk_max<true>(v, c * 32, lo, hi) : chunk_max<false>(v, c * 32, lo, hi));
        }
      }
      const float m_new = fmaxf(m_used, mx * p.scale_log2);
      float alpha = 1.f;
      if (jj == 0) {
        m_used = m_new;
      } else {
        mbar_wait(o_full, (jj - 1) & 1);  // P.V of the previous tile retired: sP and O may be touched
        tc_fence_after();
        const bool want = m_new > m_used + 8.f;   // lazy rescale: only when the max moved by > 2^8
        if (__any_sync(0xffffffffu, want)) {
          if (want) {
            alpha = (m_used == -INFINITY) ? 0.f : ex2_approx(m_used - m_new);
            m_used = m_new;
          }
#pragma unroll 1
          for (int c = 0; c < D / 32; ++c) {
            uint32_t v[32];
            tmem_ld32(tmem_O + lane_base + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st32(tmem_O + lane_base + c * 32, v);
          }
          tmem_st_wait();
        }
      }
      // pass 2: probabilities -> smem (bf16), running sum
      float lsum = 0.f;
      const float msub = (m_used == -INFINITY) ? 0.f : m_used;
      if constexpr (BN == 64) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float f[32];
          lsum += need_mask ? chunk_exp<true>(sv[c], f, p.scale_log2, msub, c * 32, lo, hi)
                            : chunk_exp<false>(sv[c], f, p.scale_log2, msub, c * 32, lo, hi);
          store_row_chunk_sw128(sP, r, c, f);
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t v[32];
          float f[32];
          tmem_ld32(tmem_S + lane_base + c * 32, v);
          tmem_ld_wait();
          lsum += need_mask ? chunk_exp<true>(v, f, p.scale_log2, msub, c * 32, lo, hi)
                            : chunk_exp<false>(v, f, p.scale_log2, msub, c * 32, lo, hi);
          store_row_chunk_sw128(sP, r, c, f);
        }
      }
      l = l * alpha + lsum;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    mbar_wait(o_full, (n - 1) & 1);
    tc_fence_after();
    const float inv_l = l > 0.f ? 1.f / l : 0.f;
    if (qpos < p.S) p.lse[(int64_t(b) * p.H + h) * p.S + qpos] = l > 0.f ? (m_used + log2f(l)) * LN2 : -INFINITY;
    __nv_bfloat16* orow = p.o + ((int64_t(b) * p.S + qpos) * p.H + h) * D;
#pragma unroll 1
    for (int c = 0; c < D / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(tmem_O + lane_base + c * 32, v);
      tmem_ld_wait();
      if (qpos < p.S) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[q * 8 + i]) * inv_l;
          *reinterpret_cast<uint4*>(orow + c * 32 + q * 8) =
              make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                         pack_bf16x2(f[6], f[7]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, D > 128 ? 512 : 256);
}

// =================================================================================================
// backward
// =================================================================================================
__device__ __forceinline__ void bar_sync_softmax8() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// First-generation single kernel (LRP_ATTN_BWD=v1): one tile at a time, dQ leaves as 16 red.global.add.v4.f32 per thread.
// Kept as the A/B baseline of profiles/r01_attn_bwd_v1_timeline.txt; the default is attn_bwd_pipe_kernel below.
template <int D>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                const __grid_constant__ CUtensorMap tmv, const __grid_constant__ CUtensorMap tmdo,
                const AttnParams p) {
  constexpr int TILE_BYTES = ATT_TILE * D * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + TILE_BYTES;
  uint8_t* sQ = sV + TILE_BYTES;
  uint8_t* sdO = sQ + TILE_BYTES;
  uint8_t* sP = sdO + TILE_BYTES;
  uint8_t* sdS = sP + 32768;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + 32768);
  uint64_t* kv_full = bars;
  uint64_t* qdo_full = bars + 1;
  uint64_t* s_full = bars + 2;
  uint64_t* p_full = bars + 3;
  uint64_t* dq_full = bars + 4;
  uint64_t* dq_empty = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int jt, hk, b;
  sched_decode((p.S + ATT_TILE - 1) / ATT_TILE, p.Hkv, p.B, p.sched_group, jt, hk, b);
  const int G = p.H / p.Hkv;
  const int k0 = jt * ATT_TILE;
  const int nq = (p.S + ATT_TILE - 1) / ATT_TILE;
  const int i_lo = p.causal ? jt : 0;
  const int i_hi = p.window > 0 ? min(nq - 1, (k0 + ATT_TILE - 1 + p.window - 1) / ATT_TILE) : nq - 1;
  const int ni = i_hi - i_lo + 1;
  const int n_it = ni > 0 ? ni * G : 0;
  int kvlo, kvhi;
  kv_bounds(p, b, kvlo, kvhi);

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmq); tma_prefetch_desc(&tmk); tma_prefetch_desc(&tmv); tma_prefetch_desc(&tmdo);
    mbar_init(kv_full, 1);
    mbar_init(qdo_full, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 256);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 256);
    fence_barrier_init();
  }
  if (warp == 8) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_dP = tmem_base + 128, tmem_dV = tmem_base + 256, tmem_dK = tmem_base + 256 + D;

  if (warp == 8) {
    if (lane == 0 && n_it > 0) {
      mbar_expect_tx(kv_full, 2 * TILE_BYTES);
      load_tile<D>(sK, &tmk, kv_full, hk * D, k0, b);
      load_tile<D>(sV, &tmv, kv_full, hk * D, k0, b);
      for (int it = 0; it < n_it; ++it) {
        const int g = it / ni, i = i_lo + (it - g * ni);
        const int h = hk * G + g;
        const bool dbg = p.dbg != nullptr && blockIdx.x == 0 && it < 64;
        if (dbg) p.dbg[it * 16 + 0] = clock64();
        mbar_expect_tx(qdo_full, 2 * TILE_BYTES);
        load_tile<D>(sQ, &tmq, qdo_full, h * D, i * ATT_TILE, b);
        load_tile<D>(sdO, &tmdo, qdo_full, h * D, i * ATT_TILE, b);
        if (it == 0) mbar_wait(kv_full, 0);
        mbar_wait(qdo_full, it & 1);
        if (dbg) p.dbg[it * 16 + 1] = clock64();
        if (it > 0) mbar_wait(dq_empty, (it - 1) & 1);
        if (dbg) p.dbg[it * 16 + 2] = clock64();
        tc_fence_after();
        mma_kk<128, D>(tmem_S, smem_u32(sQ), smem_u32(sK), false);    // S  = Q K^T
        mma_kk<128, D>(tmem_dP, smem_u32(sdO), smem_u32(sV), false);  // dP = dO V^T
        tc_commit(s_full);
        if (dbg) p.dbg[it * 16 + 3] = clock64();
        mbar_wait(p_full, it & 1);
        if (dbg) p.dbg[it * 16 + 4] = clock64();
        tc_fence_after();
        mma_mnmn<D>(tmem_dV, smem_u32(sP), smem_u32(sdO), it > 0);    // dV += P^T dO
        mma_mnmn<D>(tmem_dK, smem_u32(sdS), smem_u32(sQ), it > 0);    // dK += dS^T Q
        mma_kmn<D>(tmem_S, smem_u32(sdS), smem_u32(sK), false);       // dQ  = dS K   (reuses the S columns)
        tc_commit(dq_full);
        if (dbg) p.dbg[it * 16 + 5] = clock64();
        mbar_wait(dq_full, it & 1);  // Q/dO/P/dS smem reusable
        if (dbg) p.dbg[it * 16 + 6] = clock64();
      }
    }
    __syncwarp();
  } else {
    // 8 soft-max warps: warp w owns TMEM lane quarter w&3 and the 64-column half w>>2 of its rows (two warps per
    // SM sub-partition hide each other's MUFU / TMEM latencies)
    const int qd = warp & 3, ch = warp >> 2;
    const int r = qd * 32 + lane;
    const uint32_t lane_base = uint32_t(qd * 32) << 16;
    for (int it = 0; it < n_it; ++it) {
      const int g = it / ni, i = i_lo + (it - g * ni);
      const int h = hk * G + g;
      const int qpos = i * ATT_TILE + r;
      const bool valid = qpos < p.S;
      float lse2 = 0.f, delta = 0.f;
      if (valid) {
        lse2 = p.lse[(int64_t(b) * p.H + h) * p.S + qpos] * LOG2E;
        delta = p.delta[(int64_t(b) * p.H + h) * p.S + qpos];
      }
      const bool row_ok = valid && lse2 != -INFINITY;
      if (!row_ok) lse2 = 0.f;
      const bool need_mask = (p.causal && k0 + ATT_TILE - 1 > i * ATT_TILE) || (k0 + ATT_TILE > kvhi) || (k0 < kvlo) ||
                             (p.window > 0 && i * ATT_TILE + ATT_TILE - 1 - k0 >= p.window);
      const bool dbgt = p.dbg != nullptr && blockIdx.x == 0 && it < 64 && threadIdx.x == 0;
      if (dbgt) p.dbg[it * 16 + 8] = clock64();
      mbar_wait(s_full, it & 1);
      if (dbgt) p.dbg[it * 16 + 9] = clock64();
      tc_fence_after();
      int lo, hi;
      row_window(qpos, k0, kvlo, kvhi, p.causal, p.window, lo, hi);
      if (!row_ok) { lo = 1; hi = 0; }  // padded / fully masked query row: everything is masked
      const bool mask_tile = need_mask || !__all_sync(0xffffffffu, row_ok);
      const float delta_s = delta * p.scale;
#pragma unroll 1
      for (int c = ch * 2; c < ch * 2 + 2; ++c) {
        uint32_t vs[32], vd[32];
        float fp[32], fd[32];
        tmem_ld32(tmem_S + lane_base + c * 32, vs);
        tmem_ld32(tmem_dP + lane_base + c * 32, vd);
        tmem_ld_wait();
        if (mask_tile)
          chunk_p_ds<true>(vs, vd, fp, fd, p.scale_log2, lse2, p.scale, delta_s, c * 32, lo, hi);
        else
          chunk_p_ds<false>(vs, vd, fp, fd, p.scale_log2, lse2, p.scale, delta_s, c * 32, lo, hi);
        store_row_chunk_sw128(sP, r, c, fp);
        store_row_chunk_sw128(sdS, r, c, fd);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
      if (dbgt) p.dbg[it * 16 + 10] = clock64();
      mbar_wait(dq_full, it & 1);
      if (dbgt) p.dbg[it * 16 + 11] = clock64();
      tc_fence_after();
      // dQ tile -> fp32 reductions into the dq workspace.  (Releasing the TMEM columns before issuing the reductions was
      // tried and does not help: the red.global traffic saturates the SM's memory-instruction queue either way, see
      // profiles/r01_attn_bwd_v1_timeline.txt.)
        float* dqrow = p.dq_acc + ((int64_t(b) * p.S + qpos) * p.H + h) * D;
#pragma unroll 1
        for (int c = ch; c < D / 32; c += 2) {
          uint32_t v[32];
          tmem_ld32(tmem_S + lane_base + c * 32, v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              red_add_v4(dqrow + c * 32 + q * 4, __uint_as_float(v[q * 4]), __uint_as_float(v[q * 4 + 1]),
                         __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
          }
        }
        tc_fence_before();
        mbar_arrive(dq_empty);
      if (dbgt) p.dbg[it * 16 + 12] = clock64();
    }
    // dK, dV of this key tile
    const int kpos = k0 + r;
    if (n_it > 0) {
#pragma unroll 1
      for (int which = 0; which < 2; ++which) {
        const uint32_t src = which == 0 ? tmem_dV : tmem_dK;
        const float sc = which == 0 ? p.inv_v_div : p.inv_k_div;
        __nv_bfloat16* dst = which == 0 ? p.dv + (int64_t(b) * p.S + kpos) * p.lddv + hk * D
                                        : p.dk + (int64_t(b) * p.S + kpos) * p.lddk + hk * D;
#pragma unroll 1
        for (int c = ch; c < D / 32; c += 2) {
          uint32_t v[32];
          tmem_ld32(src + lane_base + c * 32, v);
          tmem_ld_wait();
          if (kpos < p.S) {
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
    } else if (kpos < p.S) {
      // key tile attended by nobody (cannot happen with causal / full attention, kept for windowed edge cases)
      for (int c = ch * 8; c < D; c += 16) {
        *reinterpret_cast<uint4*>(p.dv + (int64_t(b) * p.S + kpos) * p.lddv + hk * D + c) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(p.dk + (int64_t(b) * p.S + kpos) * p.lddk + hk * D + c) = make_uint4(0, 0, 0, 0);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

// -------------------------------------------------------------------------------------------------
// Software-pipelined form of the kernel above (default).  Same tiles and the same arithmetic, re-ordered so that the
// tensor pipe, the TMA unit and the soft-max warps overlap across consecutive (head, query-tile) iterations:
//   * Q is double-buffered (227 KiB budget: K,V,Q0,Q1,dO tiles + P/dS = 224 KiB at D=128); S_{i+1} = Q_{i+1} K^T is
//     issued right behind MMA2_i, so it runs while the soft-max warps still drain dQ_i;
//   * dQ_i accumulates in the dP columns (not the S columns), which is what lets S_{i+1} start early;
//   * the soft-max work is split in two passes: pass A (P = 2^(S*scale_log2 - lse2), the MUFU-heavy half) needs only S
//     and overlaps the dO_{i+1} load and the dP_{i+1} MMA; pass B (dS = P * (dP*scale - delta*scale)) follows dp_full;
//   * dQ_i leaves as asynchronous bulk tensor reductions from the P/dS smem (free between MMA2_i and pass A's P store).
template <int D, int POLY>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_pipe_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                     const __grid_constant__ CUtensorMap tmv, const __grid_constant__ CUtensorMap tmdo,
                     const __grid_constant__ CUtensorMap tmdq, const AttnParams p) {
  constexpr int TILE_BYTES = ATT_TILE * D * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + TILE_BYTES;
  uint8_t* sQ0 = sV + TILE_BYTES;          // Q buffer b at sQ0 + b * TILE_BYTES
  uint8_t* sdO = sQ0 + 2 * TILE_BYTES;
  uint8_t* sP = sdO + TILE_BYTES;
  uint8_t* sdS = sP + 32768;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + 32768);
  uint64_t* kv_full = bars;
  uint64_t* q_full = bars + 1;   // [2]
  uint64_t* do_full = bars + 3;
  uint64_t* s_full = bars + 4;
  uint64_t* dp_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* dq_full = bars + 7;
  uint64_t* dq_empty = bars + 8;
  uint64_t* p_ready = bars + 9;
  uint64_t* dq_ready = bars + 10;
  uint64_t* dv_done = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int jt, hk, b;
  sched_decode((p.S + ATT_TILE - 1) / ATT_TILE, p.Hkv, p.B, p.sched_group, jt, hk, b);
  const int G = p.H / p.Hkv;
  const int k0 = jt * ATT_TILE;
  const int nq = (p.S + ATT_TILE - 1) / ATT_TILE;
  const int i_lo = p.causal ? jt : 0;
  const int i_hi = p.window > 0 ? min(nq - 1, (k0 + ATT_TILE - 1 + p.window - 1) / ATT_TILE) : nq - 1;
  const int ni = i_hi - i_lo + 1;
  const int n_it = ni > 0 ? ni * G : 0;
  const bool dbg_cta = p.dbg != nullptr && blockIdx.x == 0;
  int kvlo, kvhi;
  kv_bounds(p, b, kvlo, kvhi);

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmq); tma_prefetch_desc(&tmk); tma_prefetch_desc(&tmv); tma_prefetch_desc(&tmdo);
    tma_prefetch_desc(&tmdq);
    mbar_init(kv_full, 1);
    mbar_init(q_full, 1);
    mbar_init(q_full + 1, 1);
    mbar_init(do_full, 1);
    mbar_init(s_full, 1);
    mbar_init(dp_full, 1);
    mbar_init(p_full, 256);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 256);
    mbar_init(p_ready, 256);
    mbar_init(dq_ready, 1);
    mbar_init(dv_done, 1);
    fence_barrier_init();
  }
  if (warp == 8) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_dP = tmem_base + 128, tmem_dV = tmem_base + 256, tmem_dK = tmem_base + 256 + D;

  if (warp == 8) {
    if (lane == 0 && n_it > 0) {
      auto q_coords = [&](int it, int& h, int& i) { const int g = it / ni; i = i_lo + (it - g * ni); h = hk * G + g; };
      int h, i;
      mbar_expect_tx(kv_full, 2 * TILE_BYTES);
      load_tile<D>(sK, &tmk, kv_full, hk * D, k0, b);
      load_tile<D>(sV, &tmv, kv_full, hk * D, k0, b);
      q_coords(0, h, i);
      mbar_expect_tx(q_full, TILE_BYTES);
      load_tile<D>(sQ0, &tmq, q_full, h * D, i * ATT_TILE, b);
      mbar_expect_tx(do_full, TILE_BYTES);
      load_tile<D>(sdO, &tmdo, do_full, h * D, i * ATT_TILE, b);
      if (n_it > 1) {
        q_coords(1, h, i);
        mbar_expect_tx(q_full + 1, TILE_BYTES);
        load_tile<D>(sQ0 + TILE_BYTES, &tmq, q_full + 1, h * D, i * ATT_TILE, b);
      }
      mbar_wait(kv_full, 0);
      mbar_wait(q_full, 0);
      tc_fence_after();
      mma_kk<128, D>(tmem_S, smem_u32(sQ0), smem_u32(sK), false);      // S_0 = Q_0 K^T
      tc_commit(s_full);
      for (int it = 0; it < n_it; ++it) {
        const bool dbg = dbg_cta && it < 64;
        const uint32_t sQ = smem_u32(sQ0 + (it & 1) * TILE_BYTES);
        if (dbg) p.dbg[it * 24 + 0] = clock64();
        mbar_wait(do_full, it & 1);
        if (dbg) p.dbg[it * 24 + 1] = clock64();
        if (it > 0) mbar_wait(dq_empty, (it - 1) & 1);                  // dQ_{it-1} read out of the dP columns
        if (dbg) p.dbg[it * 24 + 2] = clock64();
        tc_fence_after();
        mma_kk<128, D>(tmem_dP, smem_u32(sdO), smem_u32(sV), false);    // dP = dO V^T
        tc_commit(dp_full);
        mbar_wait(p_ready, it & 1);
        tc_fence_after();
        mma_mnmn<D>(tmem_dV, smem_u32(sP), smem_u32(sdO), it > 0);      // dV += P^T dO   (runs under pass B)
        tc_commit(dv_done);                                             // last reader of the dO tile
        mbar_wait(p_full, it & 1);
        if (dbg) p.dbg[it * 24 + 3] = clock64();
        if (it + 1 < n_it) {
          // dO_{it+1}: dV_it (the last reader of the dO tile) retired during pass B.  Issued HERE, ahead of this tile's dQ
          // reductions: the TMA unit serves its queue in order, and a load queued behind a 32 KiB reduction arrives
          // ~3k cycles later (profiles/r01_attn_bwd_pipe_timeline.txt).
          mbar_wait(dv_done, it & 1);
          q_coords(it + 1, h, i);
          mbar_expect_tx(do_full, TILE_BYTES);
          load_tile<D>(sdO, &tmdo, do_full, h * D, i * ATT_TILE, b);
        }
        tc_fence_after();
        mma_kmn<D>(tmem_dP, smem_u32(sdS), smem_u32(sK), false);        // dQ  = dS K   (into the dP columns)
        tc_commit(dq_ready);                                            // dV, dQ retired: P smem and dQ readable
        mma_mnmn<D>(tmem_dK, smem_u32(sdS), sQ, it > 0);                // dK += dS^T Q   (runs under the first dQ half-drain)
        tc_commit(dq_full);
        if (it + 1 < n_it) {
          mbar_wait(q_full + ((it + 1) & 1), ((it + 1) >> 1) & 1);
          tc_fence_after();
          mma_kk<128, D>(tmem_S, smem_u32(sQ0 + ((it + 1) & 1) * TILE_BYTES), smem_u32(sK), false);   // S_{it+1}
          tc_commit(s_full);
        }
        if (dbg) p.dbg[it * 24 + 4] = clock64();
        mbar_wait(dq_full, it & 1);                                     // MMA2_it retired: Q[it&1], P, dS reusable
        if (dbg) p.dbg[it * 24 + 5] = clock64();
        if (it + 2 < n_it) {
          q_coords(it + 2, h, i);
          mbar_expect_tx(q_full + (it & 1), TILE_BYTES);
          load_tile<D>(sQ0 + (it & 1) * TILE_BYTES, &tmq, q_full + (it & 1), h * D, i * ATT_TILE, b);
        }
      }
    }
    __syncwarp();
  } else {
    const int qd = warp & 3, ch = warp >> 2;
    const int r = qd * 32 + lane;
    const uint32_t lane_base = uint32_t(qd * 32) << 16;
    // (lse, delta) of this thread's query row are fetched one tile ahead: the global-load latency hides behind the dQ drain
    float lse_n = 0.f, delta_n = 0.f;
    auto fetch_row = [&](int it) {
      lse_n = 0.f; delta_n = 0.f;
      if (it < n_it) {
        const int g = it / ni, i = i_lo + (it - g * ni);
        const int qpos = i * ATT_TILE + r;
        if (qpos < p.S) {
          const int64_t idx = (int64_t(b) * p.H + hk * G + g) * p.S + qpos;
          lse_n = p.lse[idx];
          delta_n = p.delta[idx];
        }
      }
    };
    fetch_row(0);
    for (int it = 0; it < n_it; ++it) {
      const int g = it / ni, i = i_lo + (it - g * ni);
      const int h = hk * G + g;
      const int qpos = i * ATT_TILE + r;
      const bool valid = qpos < p.S;
      float lse2 = lse_n * LOG2E;
      const float delta = delta_n;
      const bool row_ok = valid && lse2 != -INFINITY;
      if (!row_ok) lse2 = 0.f;
      const bool need_mask = (p.causal && k0 + ATT_TILE - 1 > i * ATT_TILE) || (k0 + ATT_TILE > kvhi) || (k0 < kvlo) ||
                             (p.window > 0 && i * ATT_TILE + ATT_TILE - 1 - k0 >= p.window);
      const bool dbgt = dbg_cta && it < 64 && threadIdx.x == 0;
      if (dbgt) p.dbg[it * 24 + 8] = clock64();
      mbar_wait(s_full, it & 1);
      if (dbgt) p.dbg[it * 24 + 9] = clock64();
      tc_fence_after();
      int lo, hi;
      row_window(qpos, k0, kvlo, kvhi, p.causal, p.window, lo, hi);
      if (!row_ok) { lo = 1; hi = 0; }
      const bool mask_tile = need_mask || !__all_sync(0xffffffffu, row_ok);
      const float delta_s = delta * p.scale;
      // pass A: P for this thread's 64 columns (chunks 2ch, 2ch+1), kept in registers for pass B
      float pf[2][32];
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t vs[32];
        tmem_ld32(tmem_S + lane_base + (ch * 2 + cc) * 32, vs);
        tmem_ld_wait();
        if (mask_tile) chunk_exp<true, POLY>(vs, pf[cc], p.scale_log2, lse2, (ch * 2 + cc) * 32, lo, hi);
        else chunk_exp<false, POLY>(vs, pf[cc], p.scale_log2, lse2, (ch * 2 + cc) * 32, lo, hi);
      }
      if (dbgt) p.dbg[it * 24 + 10] = clock64();
      if (it > 0) {  // the previous tile's first dQ reduction group must have read the P half of the staging smem out
        if (threadIdx.x == 0) tma_store_wait_read<1>();
        bar_sync_softmax8();
      }
      if (dbgt) p.dbg[it * 24 + 11] = clock64();
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) store_row_chunk_sw128(sP, r, ch * 2 + cc, pf[cc]);
      fence_proxy_async_smem();
      mbar_arrive(p_ready);
      mbar_wait(dp_full, it & 1);
      if (dbgt) p.dbg[it * 24 + 12] = clock64();
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t vd[32];
        float fd[32];
        tmem_ld32(tmem_dP + lane_base + (ch * 2 + cc) * 32, vd);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) fd[e] = pf[cc][e] * fmaf(__uint_as_float(vd[e]), p.scale, -delta_s);
        if (cc == 0 && it > 0) {  // ... and the second group the dS half
          if (threadIdx.x == 0) tma_store_wait_read<0>();
          bar_sync_softmax8();
        }
        store_row_chunk_sw128(sdS, r, ch * 2 + cc, fd);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
      if (dbgt) p.dbg[it * 24 + 13] = clock64();
      fetch_row(it + 1);
      mbar_wait(dq_ready, it & 1);
      if (dbgt) p.dbg[it * 24 + 14] = clock64();
      tc_fence_after();
      // dV and dQ have retired: P is dead, so the first two [128 x 32] fp32 boxes of the dQ tile are staged in sP (128B-
      // swizzled) and leave as the first bulk-reduction group while dK += dS^T Q still reads dS
      {
        uint32_t v[32];
        tmem_ld32(tmem_dP + lane_base + ch * 32, v);
        tmem_ld_wait();
        uint8_t* rowp = sP + ch * 16384 + r * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(rowp + ((q ^ (r & 7)) * 16)) = make_uint4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
      }
      fence_proxy_async_smem();
      bar_sync_softmax8();
      if (threadIdx.x == 0) {
#pragma unroll
        for (int c = 0; c < 2; ++c) tma_reduce_add_3d(&tmdq, sP + c * 16384, h * D + c * 32, i * ATT_TILE, b);
        tma_store_commit();
      }
      if (dbgt) p.dbg[it * 24 + 16] = clock64();
      mbar_wait(dq_full, it & 1);   // dK retired: dS is dead too
      if (dbgt) p.dbg[it * 24 + 17] = clock64();
      if (D > 64) {
        uint32_t v[32];
        tmem_ld32(tmem_dP + lane_base + (ch + 2) * 32, v);
        tmem_ld_wait();
        uint8_t* rowp = sP + (ch + 2) * 16384 + r * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(rowp + ((q ^ (r & 7)) * 16)) = make_uint4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
        fence_proxy_async_smem();
      }
      tc_fence_before();
      mbar_arrive(dq_empty);
      bar_sync_softmax8();
      if (threadIdx.x == 0) {   // second bulk group (empty for D = 64): the boxes staged in sdS
#pragma unroll
        for (int c = 2; c < D / 32; ++c) tma_reduce_add_3d(&tmdq, sP + c * 16384, h * D + c * 32, i * ATT_TILE, b);
        tma_store_commit();
      }
      if (dbgt) p.dbg[it * 24 + 15] = clock64();
    }
    if (threadIdx.x == 0) tma_store_wait<0>();
    // dK, dV of this key tile
    const int kpos = k0 + r;
    if (n_it > 0) {
#pragma unroll 1
      for (int which = 0; which < 2; ++which) {
        const uint32_t src = which == 0 ? tmem_dV : tmem_dK;
        const float sc = which == 0 ? p.inv_v_div : p.inv_k_div;
        __nv_bfloat16* dst = which == 0 ? p.dv + (int64_t(b) * p.S + kpos) * p.lddv + hk * D
                                        : p.dk + (int64_t(b) * p.S + kpos) * p.lddk + hk * D;
#pragma unroll 1
        for (int c = ch; c < D / 32; c += 2) {
          uint32_t v[32];
          tmem_ld32(src + lane_base + c * 32, v);
          tmem_ld_wait();
          if (kpos < p.S) {
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
    } else if (kpos < p.S) {
      for (int c = ch * 8; c < D; c += 16) {
        *reinterpret_cast<uint4*>(p.dv + (int64_t(b) * p.S + kpos) * p.lddv + hk * D + c) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(p.dk + (int64_t(b) * p.S + kpos) * p.lddk + hk * D + c) = make_uint4(0, 0, 0, 0);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

// delta[b,h,s] = sum_d o[b,s,h,d] * dO[b,s,h,d]   (one warp per row); the same warp clears its row of the fp32 dQ
// accumulator (zero_acc != NULL), which saves the separate memset node of the atomic / bulk-reduction kernels
__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ o,
                                                         const __nv_bfloat16* __restrict__ d_o, float* __restrict__ delta,
                                                         float* __restrict__ zero_acc, int B, int S, int H, int D) {
  const int64_t row = blockIdx.x * int64_t(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t rows = int64_t(B) * S * H;
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  if (zero_acc != nullptr)
    for (int i = lane * 4; i < D; i += 128) *reinterpret_cast<float4*>(zero_acc + row * D + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  float s = 0.f;
  for (int i = lane * 2; i < D; i += 64) {
    const uint32_t a = *reinterpret_cast<const uint32_t*>(o + row * D + i);
    const uint32_t c = *reinterpret_cast<const uint32_t*>(d_o + row * D + i);
    s += bf16_lo(a) * bf16_lo(c) + bf16_hi(a) * bf16_hi(c);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) {
    const int h = int(row % H);
    const int64_t bs = row / H;
    const int sidx = int(bs % S);
    const int64_t b = bs / S;
    delta[(b * H + h) * S + sidx] = s;
  }
}

// dq[b,s,h,:] = bf16(dq_acc * inv_q_div), dq strided by lddq
// rezero = 1: the accumulator is cleared behind the read, so that the next call may skip its zero-fill (LRP_ATTN_ACC_ZERO)
__global__ void __launch_bounds__(256) attn_dq_finish_kernel(float* __restrict__ acc, __nv_bfloat16* __restrict__ dq,
                                                             int64_t lddq, int64_t rows, int HD, float inv_q_div, int rezero) {
  const int chunks = HD >> 3;
  const int64_t total = rows * chunks;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total; idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t t = idx / chunks;
    const int c = int(idx - t * chunks) * 8;
    const float4 a = *reinterpret_cast<const float4*>(acc + t * HD + c);
    const float4 b4 = *reinterpret_cast<const float4*>(acc + t * HD + c + 4);
    *reinterpret_cast<uint4*>(dq + t * lddq + c) =
        make_uint4(pack_bf16x2(a.x * inv_q_div, a.y * inv_q_div), pack_bf16x2(a.z * inv_q_div, a.w * inv_q_div),
                   pack_bf16x2(b4.x * inv_q_div, b4.y * inv_q_div), pack_bf16x2(b4.z * inv_q_div, b4.w * inv_q_div));
    if (rezero) {
      *reinterpret_cast<float4*>(acc + t * HD + c) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(acc + t * HD + c + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

template <int D, int BN>
static int fwd_smem_bytes() { return ATT_TILE * D * 2 + 4 * BN * D * 2 + ATT_TILE * BN * 2 + 128; }
template <int D>
static int bwd_smem_bytes() { return 4 * ATT_TILE * D * 2 + 65536 + 1024 + 256; }
template <int D>
static int bwd_pipe_smem_bytes() { return 5 * ATT_TILE * D * 2 + 65536 + 1024 + 256; }

static int check_common(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv, int B, int S,
                        int H, int Hkv, int D) {
  if (B <= 0 || S <= 0 || H <= 0 || Hkv <= 0) return set_error(LRP_ERR_ARG, "attn: empty problem");
  if (H % Hkv != 0) return set_error(LRP_ERR_ARG, "attn: H must be a multiple of Hkv");
  if (D != 64 && D != 128 && D != 256) return set_error(LRP_ERR_ARG, "attn: head_dim must be 64, 128 or 256");
  if ((ldq % 8) || (ldk % 8) || (ldv % 8)) return set_error(LRP_ERR_ARG, "attn: row strides must be multiples of 8");
  if ((uintptr_t(q) & 15) || (uintptr_t(k) & 15) || (uintptr_t(v) & 15))
    return set_error(LRP_ERR_ARG, "attn: q/k/v must be 16-byte aligned");
  return LRP_OK;
}

template <int D, int BN>
static int launch_fwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p,
                      cudaStream_t st) {
  auto kern = attn_fwd_kernel<D, BN>;
  static bool done = false;
  if (!done) {
    cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd_smem_bytes<D, BN>());
    if (ce != cudaSuccess) return set_error(LRP_ERR_CUDA, cudaGetErrorString(ce));
    done = true;
  }
  dim3 grid(((p.S + ATT_TILE - 1) / ATT_TILE) * p.H * p.B);
  kern<<<grid, ATT_THREADS, fwd_smem_bytes<D, BN>(), st>>>(tq, tk, tv, p);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

template <int D>
static int launch_bwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
                      const AttnParams& p, cudaStream_t st) {
  auto kern = attn_bwd_kernel<D>;
  static bool done = false;
  if (!done) {
    cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem_bytes<D>());
    if (ce != cudaSuccess) return set_error(LRP_ERR_CUDA, cudaGetErrorString(ce));
    done = true;
  }
  dim3 grid(((p.S + ATT_TILE - 1) / ATT_TILE) * p.Hkv * p.B);
  kern<<<grid, BWD_THREADS, bwd_smem_bytes<D>(), st>>>(tq, tk, tv, tdo, p);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

template <int D, int POLY>
static int launch_bwd_pipe(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
                           const CUtensorMap& tdq, const AttnParams& p, cudaStream_t st) {
  auto kern = attn_bwd_pipe_kernel<D, POLY>;
  static bool done = false;
  if (!done) {
    cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_pipe_smem_bytes<D>());
    if (ce != cudaSuccess) return set_error(LRP_ERR_CUDA, cudaGetErrorString(ce));
    done = true;
  }
  dim3 grid(((p.S + ATT_TILE - 1) / ATT_TILE) * p.Hkv * p.B);
  kern<<<grid, BWD_THREADS, bwd_pipe_smem_bytes<D>(), st>>>(tq, tk, tv, tdo, tdq, p);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

}  // namespace lrp

using namespace lrp;

extern "C" {

int lrp_attn_bwd_workspace_bytes(int B, int S, int H, int D, int64_t* dq_acc_bytes, int64_t* delta_bytes) {
  if (B <= 0 || S <= 0 || H <= 0 || D <= 0) return set_error(LRP_ERR_ARG, "attn_bwd_workspace_bytes: bad shape");
  if (dq_acc_bytes != nullptr) *dq_acc_bytes = int64_t(B) * S * H * D * 4;   // fp32 [B,S,H,D] dQ accumulator
  if (delta_bytes != nullptr) *delta_bytes = int64_t(B) * H * S * 4;         // fp32 [B,H,S] row sums of o * dO
  return LRP_OK;
}

int lrp_attn_fwd(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv, void* o, float* lse,
                 int B, int S, int H, int Hkv, int D, float scale, int causal, int window, void* stream) {
  return lrp_attn_fwd_varlen(q, k, v, ldq, ldk, ldv, o, lse, nullptr, B, S, H, Hkv, D, scale, causal, window, stream);
}

int lrp_attn_fwd_varlen(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv, void* o, float* lse,
                        const int32_t* kv_range, int B, int S, int H, int Hkv, int D, float scale, int causal, int window,
                        void* stream) {
  if (int e = check_common(q, k, v, ldq, ldk, ldv, B, S, H, Hkv, D)) return e;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    // head_dim 128 without a sliding window: the persistent two-tile kernel of attn_fwd_ws.cu (LRP_ATTN_FWD=v1 keeps the
    // first-generation kernel below for A/B measurements)
    static const bool v1 = getenv("LRP_ATTN_FWD") != nullptr && !strcmp(getenv("LRP_ATTN_FWD"), "v1");
    if (D == 128 && window <= 0 && !v1) return attn_fwd_ws(q, k, v, ldq, ldk, ldv, o, lse, kv_range, B, S, H, Hkv, scale, causal, st);
  }
  // key tile: 64 keys for D=128, 128 keys for D=64 -> 112 KiB of smem per CTA either way (2 CTAs per SM)
  const int BN = D == 64 ? 128 : 64;   // head_dim 256: 208 KiB, one CTA per SM, 320 TMEM columns
  CUtensorMap tq, tk, tv;
  if (int e = make_tmap_3d_bf16(&tq, q, uint64_t(H) * D, S, B, ldq, uint64_t(S) * ldq, 64, ATT_TILE)) return e;
  if (int e = make_tmap_3d_bf16(&tk, k, uint64_t(Hkv) * D, S, B, ldk, uint64_t(S) * ldk, 64, BN)) return e;
  if (int e = make_tmap_3d_bf16(&tv, v, uint64_t(Hkv) * D, S, B, ldv, uint64_t(S) * ldv, 64, BN)) return e;
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.S = S; p.H = H; p.Hkv = Hkv; p.D = D;
  p.scale = scale; p.scale_log2 = scale * LOG2E;
  p.causal = causal; p.window = window;
  p.sched_group = sched_group_default();
  p.o = reinterpret_cast<__nv_bfloat16*>(o);
  p.lse = lse;
  p.kv_range = kv_range;
  return D == 256 ? launch_fwd<256, 64>(tq, tk, tv, p, st) : D == 128 ? launch_fwd<128, 64>(tq, tk, tv, p, st) : launch_fwd<64, 128>(tq, tk, tv, p, st);
}

int lrp_attn_bwd(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv, const void* o,
                 const void* d_o, const float* lse, void* dq, void* dk, void* dv, int64_t lddq, int64_t lddk, int64_t lddv,
                 float* dq_acc_ws, float* delta_ws, int B, int S, int H, int Hkv, int D, float scale, int causal, int window,
                 float q_div, float k_div, float v_div, void* stream) {
  return lrp_attn_bwd_varlen(q, k, v, ldq, ldk, ldv, o, d_o, lse, dq, dk, dv, lddq, lddk, lddv, dq_acc_ws, delta_ws, nullptr, 0, B, S,
                             H, Hkv, D, scale, causal, window, q_div, k_div, v_div, stream);
}

int lrp_attn_bwd_varlen(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv, const void* o,
                        const void* d_o, const float* lse, void* dq, void* dk, void* dv, int64_t lddq, int64_t lddk, int64_t lddv,
                        float* dq_acc_ws, float* delta_ws, const int32_t* kv_range, int flags, int B, int S, int H, int Hkv, int D,
                        float scale, int causal, int window, float q_div, float k_div, float v_div, void* stream) {
  if (int e = check_common(q, k, v, ldq, ldk, ldv, B, S, H, Hkv, D)) return e;
  if ((lddq % 8) || (lddk % 8) || (lddv % 8)) return set_error(LRP_ERR_ARG, "attn_bwd: gradient strides must be multiples of 8");
  if (dq_acc_ws == nullptr || delta_ws == nullptr) return set_error(LRP_ERR_ARG, "attn_bwd: missing workspace");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUtensorMap tq, tk, tv, tdo;
  if (int e = make_tmap_3d_bf16(&tq, q, uint64_t(H) * D, S, B, ldq, uint64_t(S) * ldq, 64, ATT_TILE)) return e;
  if (int e = make_tmap_3d_bf16(&tk, k, uint64_t(Hkv) * D, S, B, ldk, uint64_t(S) * ldk, 64, ATT_TILE)) return e;
  if (int e = make_tmap_3d_bf16(&tv, v, uint64_t(Hkv) * D, S, B, ldv, uint64_t(S) * ldv, 64, ATT_TILE)) return e;
  const int64_t HD = int64_t(H) * D;
  if (int e = make_tmap_3d_bf16(&tdo, d_o, uint64_t(HD), S, B, HD, uint64_t(S) * HD, 64, ATT_TILE)) return e;
  const int64_t rows = int64_t(B) * S * H;
  const char* sel = getenv("LRP_ATTN_BWD");
  const bool two_pass = D == 256 || (sel != nullptr && !strcmp(sel, "v2"));   // head_dim 256 exists only in the two-pass v2 form
  const bool delta_ready = (flags & LRP_ATTN_DELTA_READY) != 0, acc_zero = (flags & LRP_ATTN_ACC_ZERO) != 0;
  if (!delta_ready) {
    attn_delta_kernel<<<unsigned((rows + 7) / 8), 256, 0, st>>>((const __nv_bfloat16*)o, (const __nv_bfloat16*)d_o, delta_ws,
                                                               (two_pass || acc_zero) ? nullptr : dq_acc_ws, B, S, H, D);
    LRP_CHECK_LAUNCH();
  } else if (!two_pass && !acc_zero) {
    cudaError_t ce = cudaMemsetAsync(dq_acc_ws, 0, size_t(rows) * D * sizeof(float), st);
    if (ce != cudaSuccess) return set_error(LRP_ERR_CUDA, cudaGetErrorString(ce));
    note_launch();
  }
  // LRP_ATTN_BWD=v2 selects the two-kernel pipelined, atomic-free (bit-reproducible) backward of attn_bwd_v2.cu.
  // Measured on B200 it is on par with / slightly slower than this single-kernel version (its 64-wide MMAs are
  // smem-operand bound and S/dP/exp are recomputed for dQ), so the single kernel stays the default.
  if (two_pass)
    return attn_bwd_v2(q, k, v, ldq, ldk, ldv, d_o, lse, delta_ws, dq, dk, dv, lddq, lddk, lddv, kv_range, B, S, H, Hkv, D, scale,
                       causal, window, q_div, k_div, v_div, st);
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.S = S; p.H = H; p.Hkv = Hkv; p.D = D;
  p.scale = scale; p.scale_log2 = scale * LOG2E;
  p.causal = causal; p.window = window;
  p.sched_group = sched_group_default();
  p.lse = const_cast<float*>(lse);
  p.delta = delta_ws;
  p.dq_acc = dq_acc_ws;
  p.dk = reinterpret_cast<__nv_bfloat16*>(dk);
  p.dv = reinterpret_cast<__nv_bfloat16*>(dv);
  p.lddk = lddk; p.lddv = lddv;
  p.inv_k_div = k_div > 0.f ? 1.f / k_div : 0.f;
  p.inv_v_div = v_div > 0.f ? 1.f / v_div : 0.f;
  p.kv_range = kv_range;
  long long* dbg_dev = nullptr;
  if (getenv("LRP_ATTN_DEBUG") != nullptr && int64_t(B) * S >= 4096) {
    cudaMalloc(&dbg_dev, 64 * 24 * sizeof(long long));
    cudaMemset(dbg_dev, 0, 64 * 24 * sizeof(long long));
    p.dbg = dbg_dev;
  }
  // default: the software-pipelined kernel with bulk-tensor dQ reductions; LRP_ATTN_BWD=v1 selects the first-generation
  // kernel (per-thread red.global) for A/B measurements
  const bool pipe = !(sel != nullptr && !strcmp(sel, "v1"));
  int le;
  // LRP_ATTN_BWD=ws: the warp-specialised experiment of attn_bwd_ws.cu (transposed scores, P^T in TMEM, own dQ drain warpgroup).
  // It reaches the same ~6k cycles per tile as the pipelined kernel because both are held by the rate at which the L2 takes
  // the fp32 dQ reductions (profiles/r02_attn_bwd_experiments.md), so it is not the default.
  const bool ws = D == 128 && window <= 0 && sel != nullptr && !strcmp(sel, "ws");
  if (ws) {
    CUtensorMap tdq;
    if (int e = make_tmap_3d_f32(&tdq, dq_acc_ws, uint64_t(HD), S, B, HD, uint64_t(S) * HD, 32, ATT_TILE)) return e;
    le = attn_bwd_ws_launch(tq, tk, tv, tdo, tdq, p, st);
  } else if (pipe) {
    CUtensorMap tdq;
    if (int e = make_tmap_3d_f32(&tdq, dq_acc_ws, uint64_t(HD), S, B, HD, uint64_t(S) * HD, 32, ATT_TILE)) return e;
    // LRP_ATTN_POLY=3: every 3rd exponential of soft-max pass A on the FMA pipe instead of MUFU (+2 %, off by default)
    static const bool poly = getenv("LRP_ATTN_POLY") != nullptr && atoi(getenv("LRP_ATTN_POLY")) == 3;
    if (D == 128) le = poly ? launch_bwd_pipe<128, 3>(tq, tk, tv, tdo, tdq, p, st) : launch_bwd_pipe<128, 0>(tq, tk, tv, tdo, tdq, p, st);
    else le = poly ? launch_bwd_pipe<64, 3>(tq, tk, tv, tdo, tdq, p, st) : launch_bwd_pipe<64, 0>(tq, tk, tv, tdo, tdq, p, st);
  } else {
    le = D == 128 ? launch_bwd<128>(tq, tk, tv, tdo, p, st) : launch_bwd<64>(tq, tk, tv, tdo, p, st);
  }
  if (le) return le;
  if (dbg_dev != nullptr) {
    static bool printed = false;
    long long h[64 * 24];
    cudaDeviceSynchronize();
    cudaMemcpy(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(dbg_dev);
    if (!printed && ws) {
      printed = true;
      printf("ws: clock64 stamps of CTA 0 relative to the MMA warp's iteration start\n"
             "it | mma: p_ready ds_wait ds_ready do_full dq_empty | drain: start dq_full box01 slot0_free slot1_free box3 | smx: start s_full passA dp_full ds_store dk_done end | iter\n");
      for (int it = 2; it < 14; ++it) {
        const long long* r = h + it * 24;
        const long long t0 = r[0];
        printf("%2d | %5lld %5lld %5lld %5lld %5lld | %5lld %5lld %5lld %5lld %5lld %5lld | %5lld %5lld %5lld %5lld %5lld %5lld %5lld | %6lld\n", it,
               r[1] - t0, r[2] - t0, r[3] - t0, r[4] - t0, r[5] - t0, r[6] - t0, r[7] - t0, r[8] - t0, r[9] - t0, r[10] - t0, r[11] - t0,
               r[12] - t0, r[13] - t0, r[14] - t0, r[15] - t0, r[16] - t0, r[17] - t0, r[18] - t0, r[0] - (h + (it - 1) * 24)[0]);
      }
    }
    if (!printed && pipe) {
      printed = true;
      printf("pipe: clock64 stamps of one CTA relative to the control thread's iteration start\n"
             "it | ctl: dO_in dq_empty p_full mma2_issued dq_full | thr: start s_full passA stage P_stored+dp_full passB(p_full) dq_ready drain1 dk_done end | iter\n");
      for (int it = 2; it < 12; ++it) {
        const long long* r = h + it * 24;
        const long long t0 = r[0];
        printf("%2d | %5lld %5lld %5lld %5lld %5lld | %5lld %5lld %5lld %5lld %5lld %5lld %5lld %5lld %5lld %5lld | %6lld\n", it, r[1] - t0, r[2] - t0,
               r[3] - t0, r[4] - t0, r[5] - t0, r[8] - t0, r[9] - t0, r[10] - t0, r[11] - t0, r[12] - t0, r[13] - t0, r[14] - t0,
               r[16] - t0, r[17] - t0, r[15] - t0, r[0] - (h + (it - 1) * 24)[0]);
      }
    }
    if (!printed) {
      printed = true;
      printf("v1 it | ctl: load_wait dq_empty mma1 p_full mma2 dq_full | thr: wait_s compute+store wait_dq drain | iter\n");
      for (int it = 1; it < 24; ++it) {
        const long long* r = h + it * 16;
        printf("%2d | %5lld %5lld %5lld %5lld %5lld %5lld | %5lld %5lld %5lld %5lld | %6lld\n", it, r[1] - r[0], r[2] - r[1], r[3] - r[2],
               r[4] - r[3], r[5] - r[4], r[6] - r[5], r[9] - r[8], r[10] - r[9], r[11] - r[10], r[12] - r[11],
               r[0] - (h + (it - 1) * 16)[0]);
      }
    }
  }
  const int64_t tok = int64_t(B) * S;
  const int64_t total = tok * (HD / 8);
  int64_t g = (total + 255) / 256;
  if (g > int64_t(sm_count()) * 16) g = int64_t(sm_count()) * 16;
  attn_dq_finish_kernel<<<unsigned(g), 256, 0, st>>>(dq_acc_ws, (__nv_bfloat16*)dq, lddq, tok, int(HD),
                                                    q_div > 0.f ? 1.f / q_div : 0.f, acc_zero ? 1 : 0);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

}  // extern "C"
