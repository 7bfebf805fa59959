// Flash AttnLRP forward, second generation (head_dim 128): persistent, warp-specialised, two query tiles in flight.
//
// Rule (reference lxt/efficient/patches.py:193-203): the forward of the patched attention is the ordinary soft-max
// attention; the LRP rule only rescales dQ, dK, dV in the backward, so this kernel computes O and the log-sum-exp.
//
// One CTA per SM walks a static list of work items (batch, head, PAIR of 128-row query tiles), heaviest causal pairs first.
//   warps 0-3 : soft-max warpgroup of query tile 0   (thread r <-> query row r <-> TMEM lane r: no shuffles)
//   warps 4-7 : soft-max warpgroup of query tile 1
//   warp  8   : tcgen05.mma issuer (one lane), owns the TMEM allocation
//   warp  9   : TMA producer (one lane)
// TMEM (512 columns): S0 | S1 | O0 | O1, 128 fp32 columns each.  P_i (bf16) overwrites the first 64 columns of S_i and
// is consumed from TMEM as the A operand of O_i += P_i V (no shared-memory round trip).  While warpgroup i turns
// S_i(j) into P_i(j), the tensor pipe runs O_{1-i} += P_{1-i} V(j) and S_{1-i}(j+1) = Q_{1-i} K(j+1)^T.
// tcgen05.mma executes in issue order, so S_i(j+1) may be issued right behind O_i += P_i(j) V(j); its commit also
// covers that P.V, which is what allows the soft-max warps to rescale O_i themselves (lazily, only when the running
// maximum moved by more than 2^8) without a separate correction warpgroup.
// smem: Q0, Q1 (32 KiB each; reused to stage the bf16 O tile for the TMA store), K x2 stages, V x2 stages = 192 KiB.
#include <stdlib.h>
#include "attn_common.cuh"

namespace lrp {

namespace {
constexpr int WS_D = 128;
constexpr int WS_THREADS = 320;
constexpr int WS_TILE_BYTES = ATT_TILE * WS_D * 2;   // 32 KiB
constexpr int WS_SMEM_BYTES = 6 * WS_TILE_BYTES + 1024 + 256;

struct WorkItem {
  int b, h, hk, q0;
  int n[2];   // number of 128-key tiles each of the two query tiles attends to
  int nmax;
};

__device__ __forceinline__ WorkItem decode_item(const AttnParams& p, int w, int pairs) {
  WorkItem it;
  const int BH = p.B * p.H;
  const int pr = w / BH, r = w - pr * BH;
  const int pp = p.causal ? pairs - 1 - pr : pr;   // heaviest causal pairs first
  it.h = r % p.H;
  it.b = r / p.H;
  it.hk = it.h / (p.H / p.Hkv);
  it.q0 = pp * 2 * ATT_TILE;
  const int nkv = (p.S + ATT_TILE - 1) / ATT_TILE;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int q0i = it.q0 + i * ATT_TILE;
    it.n[i] = q0i >= p.S ? 0 : (p.causal ? min(q0i / ATT_TILE + 1, nkv) : nkv);
  }
  it.nmax = max(it.n[0], it.n[1]);
  return it;
}

// O[128 x N] (+)= P[128 x KTOT] (bf16 in TMEM, two k per column) * V_mnmajor[KTOT(k) x N]
template <int N, int KTOT, int B_LBO>
__device__ __forceinline__ void mma_ts_kmn(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_base, bool acc_first) {
  constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 1);
  constexpr uint32_t hi = sdesc_hi(1024);
  const uint32_t b_lo = sdesc_lo(b_base, B_LBO);
#pragma unroll
  for (int kk = 0; kk < KTOT / 16; ++kk)
    tc_mma_ts_lohi(tmem_d, tmem_a + kk * 8, b_lo + ((kk * 2048) >> 4), hi, idesc, (kk > 0 || acc_first) ? 1u : 0u);
}

__device__ __forceinline__ void bar_sync_wg(int wg) {   // named barriers 1 / 2: the 128 threads of one soft-max warpgroup
  if (wg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
  else asm volatile("bar.sync 2, 128;" ::: "memory");
}
}  // namespace

__global__ void __launch_bounds__(WS_THREADS, 1)
attn_fwd_ws_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                   const __grid_constant__ CUtensorMap tmv, const __grid_constant__ CUtensorMap tmo, const AttnParams p,
                   const int n_items, const int pairs) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                          // [2] query tiles / O staging
  uint8_t* sK = sQ + 2 * WS_TILE_BYTES;        // [2] stages
  uint8_t* sV = sK + 2 * WS_TILE_BYTES;        // [2] stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * WS_TILE_BYTES);
  uint64_t* q_full = bars;         // [2]
  uint64_t* q_empty = bars + 2;    // [2]
  uint64_t* k_full = bars + 4;     // [2]
  uint64_t* k_empty = bars + 6;    // [2]
  uint64_t* v_full = bars + 8;     // [2]
  uint64_t* v_empty = bars + 10;   // [2]
  uint64_t* s_full = bars + 12;    // [2]  S_i(j) complete (also: every earlier MMA retired)
  uint64_t* p_full = bars + 14;    // [2]  P_i(j) written by the 128 soft-max threads
  uint64_t* o_full = bars + 16;    // [2]  last P.V of the item retired
  uint64_t* o_empty = bars + 18;   // [2]  O_i read out of TMEM by the epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 9 && lane == 0) {
    tma_prefetch_desc(&tmq); tma_prefetch_desc(&tmk); tma_prefetch_desc(&tmv); tma_prefetch_desc(&tmo);
  }
  if (warp == 8 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);  mbar_init(&q_empty[i], 1);
      mbar_init(&k_full[i], 1);  mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);  mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);  mbar_init(&p_full[i], 128);
      mbar_init(&o_full[i], 1);  mbar_init(&o_empty[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 8) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 9) {
    // ======================================= TMA producer =======================================
    if (lane == 0) {
      uint32_t kc = 0, vc = 0, qc0 = 0, qc1 = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
        const WorkItem it = decode_item(p, w, pairs);
        auto load_k = [&](int j) {
          const int st = kc & 1;
          mbar_wait(&k_empty[st], ((kc >> 1) & 1) ^ 1);
          mbar_expect_tx(&k_full[st], WS_TILE_BYTES);
          load_tile<WS_D>(sK + st * WS_TILE_BYTES, &tmk, &k_full[st], it.hk * WS_D, j * ATT_TILE, it.b);
          ++kc;
        };
        auto load_v = [&](int j) {
          const int st = vc & 1;
          mbar_wait(&v_empty[st], ((vc >> 1) & 1) ^ 1);
          mbar_expect_tx(&v_full[st], WS_TILE_BYTES);
          load_tile<WS_D>(sV + st * WS_TILE_BYTES, &tmv, &v_full[st], it.hk * WS_D, j * ATT_TILE, it.b);
          ++vc;
        };
        // K(0), V(0) first: their stages free up long before the previous item's epilogue releases the Q tiles
        load_k(0);
        load_v(0);
        if (it.n[0] > 0) {
          mbar_wait(&q_empty[0], (qc0 & 1) ^ 1);
          mbar_expect_tx(&q_full[0], WS_TILE_BYTES);
          load_tile<WS_D>(sQ, &tmq, &q_full[0], it.h * WS_D, it.q0, it.b);
          ++qc0;
        }
        if (it.n[1] > 0) {
          mbar_wait(&q_empty[1], (qc1 & 1) ^ 1);
          mbar_expect_tx(&q_full[1], WS_TILE_BYTES);
          load_tile<WS_D>(sQ + WS_TILE_BYTES, &tmq, &q_full[1], it.h * WS_D, it.q0 + ATT_TILE, it.b);
          ++qc1;
        }
        for (int j = 1; j < it.nmax; ++j) {
          load_k(j);
          load_v(j);
        }
      }
    }
    __syncwarp();
  } else if (warp == 8) {
    // ======================================= MMA issuer =======================================
    if (lane == 0) {
      uint32_t kc = 0, vc = 0, qc[2] = {0, 0}, pc[2] = {0, 0}, oc[2] = {0, 0};
      const uint32_t aQ[2] = {smem_u32(sQ), smem_u32(sQ + WS_TILE_BYTES)};
      const uint32_t tS[2] = {tmem_base, tmem_base + 128};
      const uint32_t tO[2] = {tmem_base + 256, tmem_base + 384};
      for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
        const WorkItem it = decode_item(p, w, pairs);
        // ---- S_i(0) = Q_i K(0)^T
        mbar_wait(&k_full[kc & 1], (kc >> 1) & 1);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (it.n[i] > 0) {
            mbar_wait(&q_full[i], qc[i] & 1);
            ++qc[i];
            tc_fence_after();
            mma_kk<128, WS_D>(tS[i], aQ[i], smem_u32(sK + (kc & 1) * WS_TILE_BYTES), false);
            tc_commit(&s_full[i]);
          }
        }
        tc_commit(&k_empty[kc & 1]);
        ++kc;
        for (int j = 0; j < it.nmax; ++j) {
          const bool have_next = j + 1 < it.nmax;
          mbar_wait(&v_full[vc & 1], (vc >> 1) & 1);
          if (have_next) mbar_wait(&k_full[kc & 1], (kc >> 1) & 1);
          const uint32_t aV = smem_u32(sV + (vc & 1) * WS_TILE_BYTES);
          const uint32_t aK = smem_u32(sK + (kc & 1) * WS_TILE_BYTES);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            if (j < it.n[i]) {
              mbar_wait(&p_full[i], pc[i] & 1);
              ++pc[i];
              if (j == 0) mbar_wait(&o_empty[i], (oc[i] & 1) ^ 1);   // the previous item's epilogue has read O_i
              tc_fence_after();
              mma_ts_kmn<WS_D, ATT_TILE, ATT_TILE * 128>(tO[i], tS[i], aV, j > 0);     // O_i += P_i(j) V(j)
              if (j + 1 < it.n[i]) {
                mma_kk<128, WS_D>(tS[i], aQ[i], aK, false);                            // S_i(j+1) = Q_i K(j+1)^T
                tc_commit(&s_full[i]);
              } else {
                tc_commit(&o_full[i]);
                ++oc[i];
              }
            }
          }
          tc_commit(&v_empty[vc & 1]);
          ++vc;
          if (have_next) {
            tc_commit(&k_empty[kc & 1]);
            ++kc;
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ======================================= soft-max warpgroups =======================================
    const int i = warp >> 2;
    const int wq = warp & 3;
    const int r = wq * 32 + lane;
    const uint32_t lane_base = uint32_t(wq * 32) << 16;
    const uint32_t tS = tmem_base + i * 128 + lane_base;
    const uint32_t tO = tmem_base + 256 + i * 128 + lane_base;
    uint8_t* sO = sQ + i * WS_TILE_BYTES;
    uint32_t s_cnt = 0, o_cnt = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const WorkItem it = decode_item(p, w, pairs);
      const int n = it.n[i];
      if (n == 0) continue;
      const int q0i = it.q0 + i * ATT_TILE;
      const int qpos = q0i + r;
      int kvlo, kvhi;
      kv_bounds(p, it.b, kvlo, kvhi);
      float m_used = -INFINITY, l = 0.f;
      for (int j = 0; j < n; ++j) {
        const int k0 = j * ATT_TILE;
        const bool need_mask = (p.causal && k0 + ATT_TILE - 1 > q0i) || (k0 + ATT_TILE > kvhi) || (k0 < kvlo);
        mbar_wait(&s_full[i], s_cnt & 1);
        ++s_cnt;
        tc_fence_after();
        uint32_t s[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(tS + c * 32, s[c]);
        tmem_ld_wait();
        if (need_mask) {
          int hi = kvhi - 1 - k0;
          if (p.causal) hi = min(hi, qpos - k0);
          const int lo = kvlo - k0;
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (c * 32 + e > hi || c * 32 + e < lo) s[c][e] = 0xff800000u;   // -inf
        }
        float mx0 = __uint_as_float(s[0][0]), mx1 = __uint_as_float(s[0][1]);
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int e = (c == 0 ? 2 : 0); e + 1 < 32; e += 4) {
            mx0 = fmax3(mx0, __uint_as_float(s[c][e]), __uint_as_float(s[c][e + 1]));
            if (e + 3 < 32) mx1 = fmax3(mx1, __uint_as_float(s[c][e + 2]), __uint_as_float(s[c][e + 3]));
          }
        const float m_new = fmaxf(m_used, fmaxf(mx0, mx1) * p.scale_log2);
        float alpha = 1.f;
        if (j == 0) {
          m_used = m_new;
        } else {
          // s_full(j) also certifies that O_i += P_i(j-1) V(j-1) retired, and the next P.V waits for this thread's
          // p_full arrival: O_i is stable here.  Lazy rescale: only when the maximum moved by more than 2^8.
          const bool want = m_new > m_used + 8.f;
          if (__any_sync(0xffffffffu, want)) {
            if (want) {
              alpha = (m_used == -INFINITY) ? 0.f : ex2_approx(m_used - m_new);
              m_used = m_new;
            }
#pragma unroll 1
            for (int c = 0; c < WS_D / 32; ++c) {
              uint32_t v[32];
              tmem_ld32(tO + c * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * alpha);
              tmem_st32(tO + c * 32, v);
            }
          }
        }
        const float msub = (m_used == -INFINITY) ? 0.f : m_used;
        float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(s[c][e]), p.scale_log2, -msub));
            const float p1 = ex2_approx(fmaf(__uint_as_float(s[c][e + 1]), p.scale_log2, -msub));
            sum0 += p0;
            sum1 += p1;
            pk[e >> 1] = pack_bf16x2(p0, p1);
          }
          tmem_st16(tS + c * 16, pk);   // keys [32c, 32c+32) of this row as 16 bf16x2 columns (all of S is in registers)
        }
        l = l * alpha + (sum0 + sum1);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[i]);
      }
      // ---- epilogue: O_i / l -> bf16 -> smem (the Q_i tile is dead) -> TMA store; log-sum-exp
      mbar_wait(&o_full[i], o_cnt & 1);
      ++o_cnt;
      tc_fence_after();
      const float inv_l = l > 0.f ? 1.f / l : 0.f;
#pragma unroll 1
      for (int c = 0; c < WS_D / 32; ++c) {
        uint32_t v[32];
        float f[32];
        tmem_ld32(tO + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(v[e]) * inv_l;
        store_row_chunk_sw128(sO, r, c, f);
      }
      tc_fence_before();
      mbar_arrive(&o_empty[i]);
      fence_proxy_async_smem();
      bar_sync_wg(i);
      if (r == 0) {
        tma_store_3d(&tmo, sO, it.h * WS_D, q0i, it.b);
        tma_store_3d(&tmo, sO + 16384, it.h * WS_D + 64, q0i, it.b);
        tma_store_commit();
        tma_store_wait_read<0>();
        mbar_arrive(&q_empty[i]);
      }
      if (qpos < p.S) p.lse[(int64_t(it.b) * p.H + it.h) * p.S + qpos] = l > 0.f ? (m_used + log2f(l)) * LN2 : -INFINITY;
    }
    if (r == 0) tma_store_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

// host entry: head_dim 128, causal or full attention, no sliding window (the first-generation kernel handles the rest)
int attn_fwd_ws(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv, void* o, float* lse,
                const int32_t* kv_range, int B, int S, int H, int Hkv, float scale, int causal, cudaStream_t st) {
  CUtensorMap tq, tk, tv, to;
  const int64_t HD = int64_t(H) * WS_D;
  if (int e = make_tmap_3d_bf16(&tq, q, uint64_t(HD), S, B, ldq, uint64_t(S) * ldq, 64, ATT_TILE)) return e;
  if (int e = make_tmap_3d_bf16(&tk, k, uint64_t(Hkv) * WS_D, S, B, ldk, uint64_t(S) * ldk, 64, ATT_TILE)) return e;
  if (int e = make_tmap_3d_bf16(&tv, v, uint64_t(Hkv) * WS_D, S, B, ldv, uint64_t(S) * ldv, 64, ATT_TILE)) return e;
  if (int e = make_tmap_3d_bf16(&to, o, uint64_t(HD), S, B, HD, uint64_t(S) * HD, 64, ATT_TILE)) return e;
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.S = S; p.H = H; p.Hkv = Hkv; p.D = WS_D;
  p.scale = scale; p.scale_log2 = scale * LOG2E;
  p.causal = causal; p.window = 0;
  p.o = reinterpret_cast<__nv_bfloat16*>(o);
  p.lse = lse;
  p.kv_range = kv_range;
  static bool done = false;
  if (!done) {
    cudaError_t ce = cudaFuncSetAttribute(attn_fwd_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM_BYTES);
    if (ce != cudaSuccess) return set_error(LRP_ERR_CUDA, cudaGetErrorString(ce));
    done = true;
  }
  const int pairs = (S + 2 * ATT_TILE - 1) / (2 * ATT_TILE);
  const int64_t n_items = int64_t(pairs) * B * H;
  if (n_items > 0x7fffffff) return set_error(LRP_ERR_ARG, "attn_fwd: too many tiles");
  const int grid = n_items < sm_count() ? int(n_items) : sm_count();
  attn_fwd_ws_kernel<<<grid, WS_THREADS, WS_SMEM_BYTES, st>>>(tq, tk, tv, to, p, int(n_items), pairs);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

}  // namespace lrp
