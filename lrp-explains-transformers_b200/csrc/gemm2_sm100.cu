// CTA-pair (cta_group::2) form of the persistent tcgen05 GEMM of gemm_sm100.cu — same operands, epilogue and call sites.
//
// Two CTAs of a 2-cluster (the two SMs of one TPC) own one 256 x 256 output tile.  Each CTA stages its own 128 rows of A and
// its own HALF (128 of the 256 n) of B per k-block, the even ("leader") CTA issues one 256x256x16 tcgen05.mma.cta_group::2
// per 16 k, and the accumulator rows [0,128) / [128,256) land in the TMEM of CTA 0 / CTA 1.  Compared with the one-CTA
// 128x256 tile this moves 32 KiB instead of 48 KiB from L2 into each SM per 128x256x64 of MMA work (and reads half the B
// bytes from smem), which is what the one-CTA kernel's remaining ~18 % of idle tensor time was waiting on
// (profiles/r01_ncu_summary.md).  Ring: 6 stages x (16 KiB A + 16 KiB B) per CTA; TMEM: 2 accumulator stages x 256 columns.
//
//   warp 0 (both CTAs) : TMA producer; every load is accounted on the LEADER's full barrier (cp.async.bulk.tensor ... .cta_group::2)
//   warp 1 (leader)    : MMA issuer; tcgen05.commit ... .multicast::cluster releases the smem stage / publishes the
//                        accumulator in both CTAs
//   warps 2..5 (both)  : epilogue of the CTA's own 128 rows; arrive on the leader's tmem_empty barrier
#include <stdlib.h>
#include <stdio.h>
#include <stdlib.h>
#include "gemm_common.cuh"

namespace lrp {

namespace {
constexpr int P_BM = 128;        // rows of A per CTA (256 per pair)
constexpr int P_BN = 256;        // tile width (128 rows of B staged per CTA)
constexpr int P_BK = 64;
constexpr int P_THREADS = 192;
constexpr int P_A_BYTES = P_BM * P_BK * 2;
constexpr int P_B_BYTES = (P_BN / 2) * P_BK * 2;
constexpr int P_STAGE_BYTES = P_A_BYTES + P_B_BYTES;
constexpr int P_STAGES = 6;
constexpr int P_SMEM_BYTES = P_STAGES * P_STAGE_BYTES + 1024 + 256;
}  // namespace

template <bool B_MN>
__global__ void __launch_bounds__(P_THREADS, 1)
gemm_bf16_pair_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P_STAGES * P_STAGE_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]  used in the leader: its producer's arrival + both CTAs' bytes
  uint64_t* empty_bar = bars + P_STAGES;        // [STAGES]  per CTA, multicast commit
  uint64_t* tmem_full = bars + 2 * P_STAGES;    // [2]       per CTA, multicast commit
  uint64_t* tmem_empty = tmem_full + 2;         // [2]       used in the leader: 8 epilogue warps of the pair
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  const int num_m = (p.M + 2 * P_BM - 1) / (2 * P_BM);
  const int num_n = (p.N + P_BN - 1) / P_BN;
  const int num_tiles = num_m * num_n;
  const int num_k = (p.K + P_BK - 1) / P_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < P_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();   // barrier inits and the TMEM allocation of both CTAs are visible before any cross-CTA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair; t < num_tiles; t += num_pairs) {
        int m_blk, n_blk;
        tile_coords(t, num_m, num_n, p.group_m, m_blk, n_blk);
        const int m0 = m_blk * (2 * P_BM) + int(rank) * P_BM;
        const int n0 = n_blk * P_BN + int(rank) * (P_BN / 2);
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * P_STAGE_BYTES;
          uint8_t* sb = sa + P_A_BYTES;
          const uint32_t lead_full = mapa_u32(smem_u32(&full_bar[stage]), 0);
          // Only the leader arrives (expecting the bytes of BOTH CTAs).  The peer's bytes may land first: the phase cannot
          // complete before the leader's arrival, and the peer cannot run a whole phase ahead because its empty barrier is
          // released by the leader's commit.  (A release.cluster arrive from the peer per k-block cost ~700 cycles and
          // serialised the ring: 765 instead of 1450+ TFLOP/s.)
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * P_STAGE_BYTES);
          tma_load_2d_pair(sa, &tma_a, lead_full, kb * P_BK, m0);
          if constexpr (!B_MN) {
            tma_load_2d_pair(sb, &tma_b, lead_full, kb * P_BK, n0);
          } else {
#pragma unroll
            for (int j = 0; j < P_BN / 2 / 64; ++j)
              tma_load_2d_pair(sb + j * (64 * P_BK * 2), &tma_b, lead_full, n0 + j * 64, kb * P_BK);
          }
          if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * P_BM, P_BN, 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const bool dbg = p.dbg != nullptr;
      long long t_start = 0, w_empty = 0, w_full = 0, n_tiles = 0;
      if (dbg) t_start = clock64();
      for (int t = pair; t < num_tiles; t += num_pairs) {
        long long c0 = 0;
        if (dbg) c0 = clock64();
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        if (dbg) { w_empty += clock64() - c0; ++n_tiles; }
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * P_BN;
        for (int kb = 0; kb < num_k; ++kb) {
          if (dbg) c0 = clock64();
          mbar_wait(&full_bar[stage], phase);
          if (dbg) w_full += clock64() - c0;
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = smem_u32(smem + stage * P_STAGE_BYTES);
            const uint32_t sb = sa + P_A_BYTES;
            const uint64_t adesc = make_sdesc_sw128(sa, 16, 1024);
            const uint64_t bdesc = B_MN ? make_sdesc_sw128(sb, 64 * P_BK * 2, 1024) : make_sdesc_sw128(sb, 16, 1024);
#pragma unroll
            for (int k = 0; k < P_BK / 16; ++k) {
              const uint64_t a_adv = uint64_t((k * 16 * 2) >> 4);
              const uint64_t b_adv = B_MN ? uint64_t((k * 16 * 128) >> 4) : uint64_t((k * 16 * 2) >> 4);
              tc_mma_ss_pair(tmem_d, adesc + a_adv, bdesc + b_adv, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            tc_commit_pair(&empty_bar[stage]);                     // both CTAs' smem stage reusable
            if (kb == num_k - 1) tc_commit_pair(&tmem_full[acc]);  // accumulator complete in both CTAs
          }
          __syncwarp();
          if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (dbg && lane == 0) {
        p.dbg[pair * 4 + 0] = clock64() - t_start;
        p.dbg[pair * 4 + 1] = w_empty;
        p.dbg[pair * 4 + 2] = w_full;
        p.dbg[pair * 4 + 3] = n_tiles;
      }
    }
  } else {
    // ================= epilogue warps (both CTAs, own 128 rows) =================
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = pair; t < num_tiles; t += num_pairs) {
      int m_blk, n_blk;
      tile_coords(t, num_m, num_n, p.group_m, m_blk, n_blk);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int m = m_blk * (2 * P_BM) + int(rank) * P_BM + q * 32 + lane;
      const bool row_ok = m < p.M;
      const float rs = (p.rowscale != nullptr && row_ok) ? p.rowscale[m] * p.alpha : p.alpha;
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * P_BN;
      const int64_t row_off = int64_t(m) * p.ldc;
      if (p.act_out != nullptr) {
        // gate|up forward with act(gate) * up fused: 64-column (gate block, up block) pairs
#pragma unroll 1
        for (int c = 0; c < P_BN / 64; ++c) {
          uint32_t vg[32], vu[32];
          tmem_ld32(taddr + c * 64, vg);
          tmem_ld32(taddr + c * 64 + 32, vu);
          tmem_ld_wait();
          const int n0 = n_blk * P_BN + c * 64;
          if (row_ok && n0 < p.N) gemm_epilogue_act_pair(p, vg, vu, m, rs, row_off, n0);
        }
      } else if (p.gated_gu != nullptr) {
        // down-projection dgrad with the gated-MLP backward rules fused: g_a never reaches HBM
#pragma unroll 1
        for (int c = 0; c < P_BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(taddr + c * 32, v);
          tmem_ld_wait();
          const int n0 = n_blk * P_BN + c * 32;
          if (row_ok && n0 < p.N) gemm_epilogue_gated_bwd(p, v, m, rs, n0);
        }
      } else {
        float dacc = 0.f;
#pragma unroll 1
        for (int c = 0; c < P_BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(taddr + c * 32, v);
          tmem_ld_wait();
          const int n0 = n_blk * P_BN + c * 32;
          if (row_ok && n0 < p.N) {
            const float dot = gemm_epilogue_chunk(p, v, m, rs, row_off, n0);
            if (p.delta_o != nullptr) gemm_epilogue_delta(p, dacc, dot, m, n0);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();   // neither CTA may exit (or free TMEM) while the other still depends on its smem / TMEM / barriers
  if (warp == 1) tmem_dealloc_pair(tmem_base, 512);
}

template <bool B_MN>
static int launch_pair(const void* A, int64_t lda, const void* B, int64_t ldb, const GemmParams& p, cudaStream_t stream) {
  CUtensorMap ta, tb;
  if (int e = make_tmap_2d_bf16(&ta, A, uint64_t(p.K), uint64_t(p.M), uint64_t(lda), 64, P_BM)) return e;
  if (!B_MN) {
    if (int e = make_tmap_2d_bf16(&tb, B, uint64_t(p.K), uint64_t(p.N), uint64_t(ldb), 64, P_BN / 2)) return e;
  } else {
    if (int e = make_tmap_2d_bf16(&tb, B, uint64_t(p.N), uint64_t(p.K), uint64_t(ldb), 64, P_BK)) return e;
  }
  auto kern = gemm_bf16_pair_kernel<B_MN>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES);
    if (ce != cudaSuccess) return set_error(LRP_ERR_CUDA, cudaGetErrorString(ce));
    attr_done = true;
  }
  const int num_tiles = ((p.M + 2 * P_BM - 1) / (2 * P_BM)) * ((p.N + P_BN - 1) / P_BN);
  const int max_pairs = sm_count() / 2;
  const int pairs = num_tiles < max_pairs ? num_tiles : max_pairs;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(unsigned(2 * pairs));
  cfg.blockDim = dim3(P_THREADS);
  cfg.dynamicSmemBytes = P_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static const bool want_dbg = getenv("LRP_GEMM_DEBUG") != nullptr;
  GemmParams pd = p;
  long long* dbg_dev = nullptr;
  if (want_dbg) {
    cudaMalloc(&dbg_dev, pairs * 4 * sizeof(long long));
    cudaMemset(dbg_dev, 0, pairs * 4 * sizeof(long long));
    pd.dbg = dbg_dev;
  }
  cudaError_t ce = cudaLaunchKernelEx(&cfg, kern, ta, tb, pd);
  if (ce != cudaSuccess) return set_error(LRP_ERR_CUDA, cudaGetErrorString(ce));
  note_launch();
  if (dbg_dev != nullptr) {   // debug only: synchronous read-back, averages over the CTA pairs
    cudaDeviceSynchronize();
    long long* h = new long long[pairs * 4];
    cudaMemcpy(h, dbg_dev, pairs * 4 * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(dbg_dev);
    double tot = 0, we = 0, wf = 0, nt = 0;
    for (int i = 0; i < pairs; ++i) { tot += h[i * 4]; we += h[i * 4 + 1]; wf += h[i * 4 + 2]; nt += h[i * 4 + 3]; }
    delete[] h;
    const double ideal = nt * (double(p.K) / 16.0) * 128.0;       // 128 cycles per 256x256x16 pair MMA
    printf("gemm pair dbg M=%d N=%d K=%d %s: MMA warp cycles/pair %.0f, waiting for a free accumulator %.1f %%, waiting for operands %.1f %%, "
           "ideal MMA time %.1f %% (tiles/pair %.2f)\n", p.M, p.N, p.K, B_MN ? "NN" : "NT", tot / pairs, 100 * we / tot, 100 * wf / tot,
           100 * ideal / tot, nt / pairs);
  }
  return LRP_OK;
}

int gemm_bf16_pair(const void* A, int64_t lda, const void* B, int64_t ldb, int b_layout, const GemmParams& p, cudaStream_t stream) {
  GemmParams q = p;
  q.group_m = p.group_m > 1 ? p.group_m / 2 : 1;   // group_m counts 128-row blocks; the pair's blocks are 256 rows
  return b_layout == 0 ? launch_pair<false>(A, lda, B, ldb, q, stream) : launch_pair<true>(A, lda, B, ldb, q, stream);
}

}  // namespace lrp
