// Flash AttnLRP backward, third generation (head_dim 128): warp-specialised, transposed scores, P in TMEM.
//
// Rule (reference lxt/efficient/patches.py:193-203 and 249-258): the backward of the patched attention is the soft-max attention
// backward with dQ, dK, dV divided by (4, 4, 2) (AttnLRP uniform rule) or (0, 0, 1) (CP-LRP); the division is applied where the
// fp32 accumulators are converted (dK, dV here, dQ in attn_dq_finish_kernel).
//
// One CTA per 128-key tile (heavy-first 1-D grid, sched_decode) loops over the (head of the GQA group, 128-query tile) pairs that
// attend to it, like attn_bwd_pipe_kernel.  What changed is the orientation and who does what:
//   * scores are computed TRANSPOSED, S^T = K Q^T and dP^T = V dO^T (keys on the TMEM lanes, queries on the columns).  P^T (bf16)
//     then overwrites S^T in TMEM and is the A operand of dV += P^T dO straight from TMEM — no shared-memory round trip, and the
//     32 KiB P buffer of the previous kernel becomes a dedicated staging buffer for dQ;
//   * dS^T goes to shared memory once and serves both dK += dS^T Q (K-major A) and dQ = dS K (the same bytes read as an MN-major A);
//   * the fp32 dQ tile is drained by its OWN warpgroup (TMEM -> smem -> cp.reduce.async.bulk.tensor add), so the soft-max warps go
//     from pass B of tile i straight to pass A of tile i+1 (S^T(i+1) is issued right behind dV(i)).
//   warps 0-7  : soft-max warpgroups (thread <-> key row; warpgroup w handles query columns [64w, 64w+64))
//   warps 8-11 : dQ drain warpgroup (thread <-> query row), also stages (lse, delta) of the next tile in smem
//   warp  12   : tcgen05.mma issuer, owns the TMEM allocation        warp 13 : TMA producer        (warps 14, 15 idle)
// TMEM (512 columns): S^T / P^T | dP^T / dQ | dV | dK.
// smem: K, V, Q x2, dO, dS^T, stage (32 KiB each) + (lse, delta) x2 + barriers = 226.3 KiB.
// Registers are re-balanced with setmaxnreg (160 soft-max / 104 drain / 88 control = 64 Ki registers).
#include <stdlib.h>
#include "attn_common.cuh"

namespace lrp {

namespace {
constexpr int BW_D = 128;
constexpr int BW_THREADS = 512;
constexpr int BW_TILE_BYTES = ATT_TILE * BW_D * 2;   // 32 KiB
constexpr int BW_USED_BYTES = 7 * BW_TILE_BYTES + 2048 + 256;
constexpr int BW_SMEM_BYTES = 232448;                // everything the SM offers; the 1 KiB alignment slack is checked at run time

// dV[128 x N] (+)= P^T[128 keys x 128 queries] (bf16 in TMEM, two queries per column, queries [64h, 64h+64) at columns 64h ..
// 64h+31 because each soft-max warpgroup overwrites its own half of S^T) * dO_mnmajor[128 queries x N]
template <int N>
__device__ __forceinline__ void mma_ts_split(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_base, bool acc_first) {
  constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 1);
  constexpr uint32_t hi = sdesc_hi(1024);
  const uint32_t b_lo = sdesc_lo(b_base, 16384);
#pragma unroll
  for (int kk = 0; kk < 8; ++kk)
    tc_mma_ts_lohi(tmem_d, tmem_a + (kk >> 2) * 64 + (kk & 3) * 8, b_lo + ((kk * 2048) >> 4), hi, idesc, (kk > 0 || acc_first) ? 1u : 0u);
}

__device__ __forceinline__ void bar_drain() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
}  // namespace

__global__ void __launch_bounds__(BW_THREADS, 1)
attn_bwd_ws_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                   const __grid_constant__ CUtensorMap tmv, const __grid_constant__ CUtensorMap tmdo,
                   const __grid_constant__ CUtensorMap tmdq, const AttnParams p) {
  constexpr int D = BW_D;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (smem + BW_USED_BYTES > smem_raw + BW_SMEM_BYTES) __trap();     // (dynamic smem starts 1 KiB-aligned in practice)
  uint8_t* sK = smem;
  uint8_t* sV = sK + BW_TILE_BYTES;
  uint8_t* sQ = sV + BW_TILE_BYTES;            // [2]
  uint8_t* sdO = sQ + 2 * BW_TILE_BYTES;
  uint8_t* sdS = sdO + BW_TILE_BYTES;          // dS^T [128 keys][128 queries] bf16
  uint8_t* sStage = sdS + BW_TILE_BYTES;       // two 16 KiB slots: ring for the four [128 x 32] fp32 boxes of a dQ tile
  float* sLse = reinterpret_cast<float*>(sStage + BW_TILE_BYTES);   // [2][128]  lse * log2(e)   (+inf for rows without keys / beyond S)
  float* sDel = sLse + 256;                                         // [2][128]  delta * scale
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDel + 256);
  uint64_t* kv_full = bars;
  uint64_t* q_full = bars + 1;     // [2]
  uint64_t* q_empty = bars + 3;    // [2]  dK(it) retired: Q buffer reusable
  uint64_t* do_full = bars + 5;
  uint64_t* do_empty = bars + 6;   // dV(it) retired: dO buffer reusable
  uint64_t* ld_full = bars + 7;    // [2]  (lse, delta) of a tile staged by the 128 drain threads
  uint64_t* ld_empty = bars + 9;   // [2]  ... and read by the 256 soft-max threads
  uint64_t* s_full = bars + 11;
  uint64_t* dp_full = bars + 12;
  uint64_t* p_ready = bars + 13;   // P^T in TMEM (256 arrivals)
  uint64_t* ds_ready = bars + 14;  // dS^T in smem, dP^T read out of TMEM (256 arrivals)
  uint64_t* dq_full = bars + 15;   // dQ(it) retired
  uint64_t* dk_done = bars + 16;   // dK(it) retired: dS^T smem no longer read by the tensor pipe
  uint64_t* dq_empty = bars + 17;  // dQ(it) read out of TMEM (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int jt, hk, b;
  sched_decode((p.S + ATT_TILE - 1) / ATT_TILE, p.Hkv, p.B, p.sched_group, jt, hk, b);
  const int G = p.H / p.Hkv;
  const int k0 = jt * ATT_TILE;
  const int nq = (p.S + ATT_TILE - 1) / ATT_TILE;
  const int i_lo = p.causal ? jt : 0;
  const int ni = nq - i_lo;
  const int n_it = ni > 0 ? ni * G : 0;
  int kvlo, kvhi;
  kv_bounds(p, b, kvlo, kvhi);
  const bool dbg_cta = p.dbg != nullptr && blockIdx.x == 0;      // LRP_ATTN_DEBUG=1: clock64 stamps of the first (heaviest) CTA

  if (warp == 13 && lane == 0) {
    tma_prefetch_desc(&tmq); tma_prefetch_desc(&tmk); tma_prefetch_desc(&tmv); tma_prefetch_desc(&tmdo);
    tma_prefetch_desc(&tmdq);
  }
  if (warp == 12 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(q_full + i, 1);
      mbar_init(q_empty + i, 1);
      mbar_init(ld_full + i, 128);
      mbar_init(ld_empty + i, 256);
    }
    mbar_init(do_full, 1);
    mbar_init(do_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(dp_full, 1);
    mbar_init(p_ready, 256);
    mbar_init(ds_ready, 256);
    mbar_init(dq_full, 1);
    mbar_init(dk_done, 1);
    mbar_init(dq_empty, 128);
    fence_barrier_init();
  }
  if (warp == 12) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_dP = tmem_base + 128, tmem_dV = tmem_base + 256, tmem_dK = tmem_base + 384;
  auto q_coords = [&](int it, int& h, int& i) { const int g = it / ni; i = i_lo + (it - g * ni); h = hk * G + g; };

  if (warp >= 12) {
    setmaxnreg_dec<88>();
    if (warp == 13) {
      // ======================================= TMA producer =======================================
      if (lane == 0 && n_it > 0) {
        int h, i;
        mbar_expect_tx(kv_full, 2 * BW_TILE_BYTES);
        load_tile<D>(sK, &tmk, kv_full, hk * D, k0, b);
        load_tile<D>(sV, &tmv, kv_full, hk * D, k0, b);
        q_coords(0, h, i);
        mbar_expect_tx(q_full, BW_TILE_BYTES);
        load_tile<D>(sQ, &tmq, q_full, h * D, i * ATT_TILE, b);
        mbar_expect_tx(do_full, BW_TILE_BYTES);
        load_tile<D>(sdO, &tmdo, do_full, h * D, i * ATT_TILE, b);
        if (n_it > 1) {
          q_coords(1, h, i);
          mbar_expect_tx(q_full + 1, BW_TILE_BYTES);
          load_tile<D>(sQ + BW_TILE_BYTES, &tmq, q_full + 1, h * D, i * ATT_TILE, b);
        }
        for (int t = 1; t < n_it; ++t) {
          mbar_wait(do_empty, (t - 1) & 1);                      // dV(t-1) retired
          q_coords(t, h, i);
          mbar_expect_tx(do_full, BW_TILE_BYTES);
          load_tile<D>(sdO, &tmdo, do_full, h * D, i * ATT_TILE, b);
          if (t + 1 < n_it) {
            mbar_wait(q_empty + ((t + 1) & 1), ((t - 1) >> 1) & 1);   // dK(t-1) retired: its Q buffer takes Q(t+1)
            q_coords(t + 1, h, i);
            mbar_expect_tx(q_full + ((t + 1) & 1), BW_TILE_BYTES);
            load_tile<D>(sQ + ((t + 1) & 1) * BW_TILE_BYTES, &tmq, q_full + ((t + 1) & 1), h * D, i * ATT_TILE, b);
          }
        }
      }
      __syncwarp();
    } else if (warp == 12) {
      // ======================================= MMA issuer =======================================
      if (lane == 0 && n_it > 0) {
        const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), adO = smem_u32(sdO), adS = smem_u32(sdS);
        mbar_wait(kv_full, 0);
        mbar_wait(q_full, 0);
        tc_fence_after();
        mma_kk<128, D>(tmem_S, aK, smem_u32(sQ), false);                    // S^T(0) = K Q(0)^T
        tc_commit(s_full);
        mbar_wait(do_full, 0);
        tc_fence_after();
        mma_kk<128, D>(tmem_dP, aV, adO, false);                            // dP^T(0) = V dO(0)^T
        tc_commit(dp_full);
        for (int it = 0; it < n_it; ++it) {
          const uint32_t aQ = smem_u32(sQ + (it & 1) * BW_TILE_BYTES);
          const bool dbg = dbg_cta && it < 64;
          if (dbg) p.dbg[it * 24 + 0] = clock64();
          mbar_wait(p_ready, it & 1);
          if (dbg) p.dbg[it * 24 + 1] = clock64();
          tc_fence_after();
          mma_ts_split<D>(tmem_dV, tmem_S, adO, it > 0);                    // dV += P^T dO
          tc_commit(do_empty);
          if (it + 1 < n_it) {                                              // S^T(it+1): executes behind dV(it), which read P^T
            mbar_wait(q_full + ((it + 1) & 1), ((it + 1) >> 1) & 1);
            tc_fence_after();
            mma_kk<128, D>(tmem_S, aK, smem_u32(sQ + ((it + 1) & 1) * BW_TILE_BYTES), false);
            tc_commit(s_full);
          }
          if (dbg) p.dbg[it * 24 + 2] = clock64();
          mbar_wait(ds_ready, it & 1);
          if (dbg) p.dbg[it * 24 + 3] = clock64();
          tc_fence_after();
          mma_mnmn<D>(tmem_dP, adS, aK, false);                             // dQ = dS K   (into the dP^T columns)
          tc_commit(dq_full);
          mma_kmn<D>(tmem_dK, adS, aQ, it > 0);                             // dK += dS^T Q
          tc_commit(dk_done);
          tc_commit(q_empty + (it & 1));
          if (it + 1 < n_it) {
            mbar_wait(do_full, (it + 1) & 1);
            if (dbg) p.dbg[it * 24 + 4] = clock64();
            mbar_wait(dq_empty, it & 1);                                    // dQ(it) read out of the dP^T columns
            if (dbg) p.dbg[it * 24 + 5] = clock64();
            tc_fence_after();
            mma_kk<128, D>(tmem_dP, aV, adO, false);                        // dP^T(it+1)
            tc_commit(dp_full);
          }
        }
      }
      __syncwarp();
    }
  } else if (warp >= 8) {
    // ======================================= dQ drain warpgroup =======================================
    setmaxnreg_dec<104>();
    const int rq = (warp - 8) * 32 + lane;                     // query row of the tile == TMEM lane
    const uint32_t lane_base = uint32_t((warp & 3) * 32) << 16;
    auto post_ld = [&](int t) {
      int h, i;
      q_coords(t, h, i);
      const int qpos = i * ATT_TILE + rq;
      float l2 = INFINITY, ds = 0.f;
      if (qpos < p.S) {
        const int64_t idx = (int64_t(b) * p.H + h) * p.S + qpos;
        const float l = p.lse[idx];
        if (l != -INFINITY) { l2 = l * LOG2E; ds = p.delta[idx] * p.scale; }
      }
      sLse[(t & 1) * 128 + rq] = l2;
      sDel[(t & 1) * 128 + rq] = ds;
      mbar_arrive(ld_full + (t & 1));
    };
    if (n_it > 0) post_ld(0);
    if (n_it > 1) post_ld(1);
    // The fp32 dQ tile [128 x 128] leaves as four [128 x 32] boxes through a ring of two 16 KiB staging slots (box c -> slot c & 1), one
    // bulk-reduction group per box.  Boxes 2 and 3 wait in registers for their slot, so the dP^T / dQ columns of TMEM are released
    // ~0.6k cycles after dQ retired, long before the reduction engine (~0.8k cycles per box) has taken the tile.
    auto stage_box = [&](uint8_t* slot, const uint32_t (&v)[32]) {
      uint8_t* r0 = slot + rq * 128;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<uint4*>(r0 + ((q ^ (rq & 7)) * 16)) = make_uint4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
    };
    for (int it = 0; it < n_it; ++it) {
      int h, i;
      q_coords(it, h, i);
      const bool dbgd = dbg_cta && it < 64 && rq == 0;
      if (dbgd) p.dbg[it * 24 + 6] = clock64();
      mbar_wait(dq_full, it & 1);
      if (dbgd) p.dbg[it * 24 + 7] = clock64();
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld32(tmem_dP + lane_base, v0);
      tmem_ld32(tmem_dP + lane_base + 32, v1);
      tmem_ld_wait();
      if (rq == 0) tma_store_wait_read<0>();   // the previous tile's boxes 2, 3 have left the slots
      bar_drain();
      stage_box(sStage, v0);
      stage_box(sStage + 16384, v1);
      tmem_ld32(tmem_dP + lane_base + 64, v0);
      tmem_ld32(tmem_dP + lane_base + 96, v1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(dq_empty);                   // the dP^T / dQ columns are free for dP^T(it+1)
      fence_proxy_async_smem();
      bar_drain();
      if (rq == 0) {
        tma_reduce_add_3d(&tmdq, sStage, h * D, i * ATT_TILE, b);
        tma_store_commit();
        tma_reduce_add_3d(&tmdq, sStage + 16384, h * D + 32, i * ATT_TILE, b);
        tma_store_commit();
        if (dbgd) p.dbg[it * 24 + 8] = clock64();
        tma_store_wait_read<1>();              // box 0 read: slot 0 takes box 2
        if (dbgd) p.dbg[it * 24 + 9] = clock64();
      }
      bar_drain();
      stage_box(sStage, v0);
      fence_proxy_async_smem();
      bar_drain();
      if (rq == 0) {
        tma_reduce_add_3d(&tmdq, sStage, h * D + 64, i * ATT_TILE, b);
        tma_store_commit();
        tma_store_wait_read<1>();              // box 1 read: slot 1 takes box 3
        if (dbgd) p.dbg[it * 24 + 10] = clock64();
      }
      bar_drain();
      stage_box(sStage + 16384, v1);
      fence_proxy_async_smem();
      bar_drain();
      if (rq == 0) {
        tma_reduce_add_3d(&tmdq, sStage + 16384, h * D + 96, i * ATT_TILE, b);
        tma_store_commit();
        if (dbgd) p.dbg[it * 24 + 11] = clock64();
      }
      if (it + 2 < n_it) {
        mbar_wait(ld_empty + (it & 1), (it >> 1) & 1);          // pass B of tile it has read this (lse, delta) buffer
        post_ld(it + 2);
      }
    }
    if (rq == 0) tma_store_wait<0>();
  } else {
    // ======================================= soft-max warpgroups =======================================
    setmaxnreg_inc<160>();
    const int qd = warp & 3, ch = warp >> 2;
    const int rk = qd * 32 + lane;                             // key row of the tile == TMEM lane
    const uint32_t lane_base = uint32_t(qd * 32) << 16;
    const int kpos = k0 + rk;
    const bool row_valid = kpos >= kvlo && kpos < kvhi;        // (kvhi <= S)
    for (int it = 0; it < n_it; ++it) {
      int h, i;
      q_coords(it, h, i);
      const int buf = it & 1;
      const bool need_mask = (p.causal && k0 + ATT_TILE - 1 > i * ATT_TILE) || (k0 + ATT_TILE > kvhi) || (k0 < kvlo);
      const int clo = p.causal ? kpos - i * ATT_TILE : 0;      // first visible query column of this key row
      const bool dbgc = dbg_cta && it < 64 && threadIdx.x == 0;
      if (dbgc) p.dbg[it * 24 + 12] = clock64();
      mbar_wait(ld_full + buf, (it >> 1) & 1);
      mbar_wait(s_full, it & 1);
      if (dbgc) p.dbg[it * 24 + 13] = clock64();
      tc_fence_after();
      // pass A: P^T for this thread's 64 query columns, fp32 copy kept for pass B
      float pf[2][32];
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c0 = (ch * 2 + cc) * 32;
        uint32_t vs[32], pk[16];
        tmem_ld32(tmem_S + lane_base + c0, vs);
        tmem_ld_wait();
        const float4* l4 = reinterpret_cast<const float4*>(sLse + buf * 128 + c0);
#pragma unroll
        for (int e4 = 0; e4 < 8; ++e4) {
          const float4 L = l4[e4];
          const float ls[4] = {L.x, L.y, L.z, L.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int e = e4 * 4 + j;
            float pe = ex2_approx(fmaf(__uint_as_float(vs[e]), p.scale_log2, -ls[j]));
            if (need_mask) pe = (row_valid && c0 + e >= clo) ? pe : 0.f;
            pf[cc][e] = pe;
          }
        }
#pragma unroll
        for (int e = 0; e < 32; e += 2) pk[e >> 1] = pack_bf16x2(pf[cc][e], pf[cc][e + 1]);
        tmem_st16(tmem_S + lane_base + ch * 64 + cc * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_ready);
      if (dbgc) p.dbg[it * 24 + 14] = clock64();
      // pass B: dS^T = P^T * (dP^T * scale - delta * scale)
      mbar_wait(dp_full, it & 1);
      if (dbgc) p.dbg[it * 24 + 15] = clock64();
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c0 = (ch * 2 + cc) * 32;
        uint32_t vd[32];
        float fd[32];
        tmem_ld32(tmem_dP + lane_base + c0, vd);
        tmem_ld_wait();
        const float4* d4 = reinterpret_cast<const float4*>(sDel + buf * 128 + c0);
#pragma unroll
        for (int e4 = 0; e4 < 8; ++e4) {
          const float4 Dl = d4[e4];
          const float dl[4] = {Dl.x, Dl.y, Dl.z, Dl.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int e = e4 * 4 + j;
            fd[e] = pf[cc][e] * fmaf(__uint_as_float(vd[e]), p.scale, -dl[j]);
          }
        }
        if (cc == 0 && dbgc) p.dbg[it * 24 + 16] = clock64();
        if (cc == 0 && it > 0) mbar_wait(dk_done, (it - 1) & 1);    // dQ(it-1), dK(it-1) retired: the tensor pipe no longer reads dS^T
        if (cc == 0 && dbgc) p.dbg[it * 24 + 17] = clock64();
        store_row_chunk_sw128(sdS, rk, ch * 2 + cc, fd);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(ds_ready);
      mbar_arrive(ld_empty + buf);
      if (dbgc) p.dbg[it * 24 + 18] = clock64();
    }
    // dK, dV of this key tile
    if (n_it > 0) {
      mbar_wait(dk_done, (n_it - 1) & 1);
      tc_fence_after();
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
  if (warp == 12) tmem_dealloc(tmem_base, 512);
}

// launch of the kernel above; maps built by lrp_attn_bwd_varlen (128-row boxes for q / k / v / dO, fp32 32-column boxes for dQ)
int attn_bwd_ws_launch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
                       const CUtensorMap& tdq, const AttnParams& p, cudaStream_t st) {
  static bool done = false;
  if (!done) {
    cudaError_t ce = cudaFuncSetAttribute(attn_bwd_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BW_SMEM_BYTES);
    if (ce != cudaSuccess) return set_error(LRP_ERR_CUDA, cudaGetErrorString(ce));
    done = true;
  }
  dim3 grid(((p.S + ATT_TILE - 1) / ATT_TILE) * p.Hkv * p.B);
  attn_bwd_ws_kernel<<<grid, BW_THREADS, BW_SMEM_BYTES, st>>>(tq, tk, tv, tdo, tdq, p);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

}  // namespace lrp
