// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[M,N] = epilogue( A[M,K] (bf16, row-major)  x  B )           fp32 accumulation in TMEM
//     B layout 0 ("NT"): B is [N,K] row-major  (y = x W^T, the nn.Linear forward)
//     B layout 1 ("NN"): B is [K,N] row-major  (g_x = g_y W,  the LRP / GxI backward of nn.Linear:
//                        the weight is consumed in its stored layout through an MN-major smem descriptor)
//   epilogue:  out = resid + alpha * acc * rowscale[m] * colscale[n] + bias[n]     (every term optional)
//              written as bf16 or fp32, plus an optional bf16 shadow copy.
//
// This is the single kernel every Linear on the AttnLRP path runs through (reference call sites:
// transformers modeling_llama.py:183,262-264,288 forward; autograd dgrad of the same in backward; the
// fused epilogue terms replace lxt/efficient/patches.py:111-123 (RMSNorm backward g*w*rstd), the residual
// adds, and lxt/efficient/rules.py:125-127 (divide_gradient) which otherwise are separate HBM round trips).
//
// Structure (one CTA per SM, 192 threads):
//   warp 0      : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1      : MMA issuer     (one elected lane issues tcgen05.mma 128xBNx16, accumulators in TMEM,
//                                 tcgen05.commit releases smem stages / publishes accumulators)
//   warps 2..5  : epilogue       (tcgen05.ld TMEM -> registers -> fused epilogue -> global)
//   TMEM holds two accumulator stages so the epilogue of tile i overlaps the mainloop of tile i+1.
#include <stdlib.h>
#include "gemm_common.cuh"

namespace lrp {

static int stream_stores_default() {
  static const int v = getenv("LRP_GEMM_STREAM_STORES") != nullptr ? atoi(getenv("LRP_GEMM_STREAM_STORES")) : 1;
  return v;
}


constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 192;

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
};

// A_MN: the A operand is stored contraction-major, i.e. the caller holds A^T as a [K, M] row-major tensor (the `a^T s`
//       contraction of the matmul rule, lxt/explicit/functional.py:393-408) and it is consumed through an MN-major descriptor.
// BATCHED: 3-D tensor maps, tile index runs over batch x tiles (strided-batched problems in ONE launch).
template <int BN, bool B_MN, bool A_MN = false, bool BATCHED = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES]
  uint64_t* empty_bar = bars + Cfg::STAGES;        // [STAGES]
  uint64_t* tmem_full = bars + 2 * Cfg::STAGES;    // [2]
  uint64_t* tmem_empty = tmem_full + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (p.M + BM - 1) / BM;
  const int num_n = (p.N + BN - 1) / BN;
  const int tiles_per_problem = num_m * num_n;
  const int num_tiles = tiles_per_problem * (BATCHED ? p.batch : 1);
  const int num_k = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int m_blk, n_blk;
        const int bz = BATCHED ? t / tiles_per_problem : 0;
        tile_coords(BATCHED ? t - bz * tiles_per_problem : t, num_m, num_n, p.group_m, m_blk, n_blk);
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          if constexpr (BATCHED) {
            if constexpr (!A_MN) {
              tma_load_3d(sa, &tma_a, &full_bar[stage], kb * BK, m_blk * BM, bz);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)
                tma_load_3d(sa + j * (64 * BK * 2), &tma_a, &full_bar[stage], m_blk * BM + j * 64, kb * BK, bz);
            }
            if constexpr (!B_MN) {
              tma_load_3d(sb, &tma_b, &full_bar[stage], kb * BK, n_blk * BN, bz);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_3d(sb + j * (64 * BK * 2), &tma_b, &full_bar[stage], n_blk * BN + j * 64, kb * BK, bz);
            }
          } else {
            tma_load_2d(sa, &tma_a, &full_bar[stage], kb * BK, m_blk * BM);
            if constexpr (!B_MN) {
              tma_load_2d(sb, &tma_b, &full_bar[stage], kb * BK, n_blk * BN);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_2d(sb + j * (64 * BK * 2), &tma_b, &full_bar[stage], n_blk * BN + j * 64, kb * BK);
            }
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint64_t adesc = A_MN ? make_sdesc_sw128(sa, 64 * BK * 2, 1024) : make_sdesc_sw128(sa, 16, 1024);
          const uint64_t bdesc = B_MN ? make_sdesc_sw128(sb, 64 * BK * 2, 1024) : make_sdesc_sw128(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t a_adv = A_MN ? uint64_t((k * UMMA_K * 128) >> 4) : uint64_t((k * UMMA_K * 2) >> 4);
            const uint64_t b_adv = B_MN ? uint64_t((k * UMMA_K * 128) >> 4) : uint64_t((k * UMMA_K * 2) >> 4);
            tc_mma_ss(tmem_d, adesc + a_adv, bdesc + b_adv, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);                     // smem stage reusable once these MMAs retire
          if (kb == num_k - 1) tc_commit(&tmem_full[acc]);  // accumulator complete
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ================= epilogue warps =================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int m_blk, n_blk;
      const int bz = BATCHED ? t / tiles_per_problem : 0;
      tile_coords(BATCHED ? t - bz * tiles_per_problem : t, num_m, num_n, p.group_m, m_blk, n_blk);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int m = m_blk * BM + q * 32 + lane;
      const bool row_ok = m < p.M;
      const float rs = (p.rowscale != nullptr && row_ok) ? p.rowscale[m] * p.alpha : p.alpha;
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * BN;
      const int64_t row_off = int64_t(m) * p.ldc + (BATCHED ? int64_t(bz) * p.batch_stride_c : 0);
      if (p.act_out != nullptr) {
        // gate|up forward with act(gate) * up fused: 64-column (gate block, up block) pairs
#pragma unroll 1
        for (int c = 0; c < BN / 64; ++c) {
          uint32_t vg[32], vu[32];
          tmem_ld32(taddr + c * 64, vg);
          tmem_ld32(taddr + c * 64 + 32, vu);
          tmem_ld_wait();
          const int n0 = n_blk * BN + c * 64;
          if (row_ok && n0 < p.N) gemm_epilogue_act_pair(p, vg, vu, m, rs, row_off, n0);
        }
      } else if (p.gated_gu != nullptr) {
        // down-projection dgrad with the gated-MLP backward rules fused: g_a never reaches HBM
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(taddr + c * 32, v);
          tmem_ld_wait();
          const int n0 = n_blk * BN + c * 32;
          if (row_ok && n0 < p.N) gemm_epilogue_gated_bwd(p, v, m, rs, n0);
        }
      } else {
        float dacc = 0.f;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(taddr + c * 32, v);
          tmem_ld_wait();
          const int n0 = n_blk * BN + c * 32;
          if (row_ok && n0 < p.N) {
            const float dot = gemm_epilogue_chunk(p, v, m, rs, row_off, n0);
            if (p.delta_o != nullptr) gemm_epilogue_delta(p, dacc, dot, m, n0);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <int BN, bool B_MN>
static int launch_gemm(const void* A, int64_t lda, const void* B, int64_t ldb, const GemmParams& p,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap ta, tb;
  // A: [M, K] row-major, box = 64 (k) x 128 (m)
  if (int e = make_tmap_2d_bf16(&ta, A, uint64_t(p.K), uint64_t(p.M), uint64_t(lda), 64, BM)) return e;
  if (!B_MN) {
    // B: [N, K] row-major, box = 64 (k) x BN (n)
    if (int e = make_tmap_2d_bf16(&tb, B, uint64_t(p.K), uint64_t(p.N), uint64_t(ldb), 64, BN)) return e;
  } else {
    // B: [K, N] row-major, box = 64 (n) x 64 (k)
    if (int e = make_tmap_2d_bf16(&tb, B, uint64_t(p.N), uint64_t(p.K), uint64_t(ldb), 64, BK)) return e;
  }
  auto kern = gemm_bf16_kernel<BN, B_MN>;
  static bool attr_done = false;  // idempotent; benign race
  if (!attr_done) {
    cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (ce != cudaSuccess) return set_error(LRP_ERR_CUDA, cudaGetErrorString(ce));
    attr_done = true;
  }
  const int num_tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  const int grid = num_tiles < sm_count() ? num_tiles : sm_count();
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(ta, tb, p);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

int gemm_bf16_pair(const void* A, int64_t lda, const void* B, int64_t ldb, int b_layout, const GemmParams& p, cudaStream_t stream);

// strided-batched launch of the one-CTA 128 x 128 kernel: 3-D tensor maps (dim2 = problem index)
template <bool B_MN, bool A_MN>
static int launch_gemm_batched(const void* A, int64_t lda, int64_t sa, const void* B, int64_t ldb, int64_t sb, const GemmParams& p,
                               cudaStream_t stream) {
  constexpr int BN = 128;
  using Cfg = GemmCfg<BN>;
  CUtensorMap ta, tb;
  const uint64_t nb = uint64_t(p.batch);
  if (!A_MN) {   // A: [batch][M][K], box 64 (k) x 128 (m)
    if (int e = make_tmap_3d_bf16(&ta, A, uint64_t(p.K), uint64_t(p.M), nb, uint64_t(lda), uint64_t(sa), 64, BM)) return e;
  } else {       // A^T stored: [batch][K][M], box 64 (m) x 64 (k)
    if (int e = make_tmap_3d_bf16(&ta, A, uint64_t(p.M), uint64_t(p.K), nb, uint64_t(lda), uint64_t(sa), 64, BK)) return e;
  }
  if (!B_MN) {
    if (int e = make_tmap_3d_bf16(&tb, B, uint64_t(p.K), uint64_t(p.N), nb, uint64_t(ldb), uint64_t(sb), 64, BN)) return e;
  } else {
    if (int e = make_tmap_3d_bf16(&tb, B, uint64_t(p.N), uint64_t(p.K), nb, uint64_t(ldb), uint64_t(sb), 64, BK)) return e;
  }
  auto kern = gemm_bf16_kernel<BN, B_MN, A_MN, true>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (ce != cudaSuccess) return set_error(LRP_ERR_CUDA, cudaGetErrorString(ce));
    attr_done = true;
  }
  const int64_t num_tiles = int64_t((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN) * p.batch;
  const int grid = num_tiles < sm_count() ? int(num_tiles) : sm_count();
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(ta, tb, p);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

int gemm_bf16_batched(const void* A, int64_t lda, int64_t stride_a, int a_layout, const void* B, int64_t ldb, int64_t stride_b,
                      int b_layout, int batch, int M, int N, int K, const lrp_epilogue_t* epi, int64_t stride_c, cudaStream_t stream) {
  if (batch <= 0 || M <= 0 || N <= 0 || K <= 0) return set_error(LRP_ERR_ARG, "gemm_batched: empty problem");
  if ((N % 8) != 0 || (K % 8) != 0 || (lda % 8) != 0 || (ldb % 8) != 0 || (stride_a % 8) != 0 || (stride_b % 8) != 0 ||
      (a_layout != 0 && (M % 8) != 0))
    return set_error(LRP_ERR_ARG, "gemm_batched: N, K, leading dimensions and batch strides must be multiples of 8");
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15))
    return set_error(LRP_ERR_ARG, "gemm_batched: A/B must be 16-byte aligned");
  if (epi == nullptr || epi->out == nullptr || epi->gated_gu != nullptr || epi->act_out != nullptr || (epi->ldc % 8) != 0)
    return set_error(LRP_ERR_ARG, "gemm_batched: needs a plain epilogue with an output");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.stream_stores = stream_stores_default();
  p.M = M; p.N = N; p.K = K;
  p.out = epi->out;
  p.shadow = reinterpret_cast<__nv_bfloat16*>(epi->shadow_bf16);
  p.resid = epi->resid_f32;
  p.rowscale = epi->rowscale;
  p.colscale = epi->colscale;
  p.bias = epi->bias;
  p.alpha = epi->alpha;
  p.ldc = epi->ldc;
  p.out_is_f32 = epi->out_is_f32;
  p.group_m = 16;
  p.batch = batch;
  p.batch_stride_c = stride_c;
  if (a_layout == 0)
    return b_layout == 0 ? launch_gemm_batched<false, false>(A, lda, stride_a, B, ldb, stride_b, p, stream)
                         : launch_gemm_batched<true, false>(A, lda, stride_a, B, ldb, stride_b, p, stream);
  return b_layout == 0 ? launch_gemm_batched<false, true>(A, lda, stride_a, B, ldb, stride_b, p, stream)
                       : launch_gemm_batched<true, true>(A, lda, stride_a, B, ldb, stride_b, p, stream);
}

int gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, int b_layout, int M, int N, int K,
              const lrp_epilogue_t* epi, int force_bn, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return set_error(LRP_ERR_ARG, "gemm: empty problem");
  if ((N % 8) != 0 || (K % 8) != 0) return set_error(LRP_ERR_ARG, "gemm: N and K must be multiples of 8");
  if ((lda % 8) != 0 || (ldb % 8) != 0) return set_error(LRP_ERR_ARG, "gemm: lda/ldb must be multiples of 8");
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15))
    return set_error(LRP_ERR_ARG, "gemm: A/B must be 16-byte aligned");
  if (epi == nullptr || (epi->out == nullptr && epi->gated_gu == nullptr)) return set_error(LRP_ERR_ARG, "gemm: missing output");
  if (epi->gated_gu != nullptr && (epi->gated_out == nullptr || epi->gated_act < 0 || epi->gated_act > 2))
    return set_error(LRP_ERR_ARG, "gemm: fused gated backward needs gated_out and a valid activation");
  if ((epi->ldc % 8) != 0) return set_error(LRP_ERR_ARG, "gemm: ldc must be a multiple of 8");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.stream_stores = stream_stores_default();
  p.batch = 1;
  p.M = M; p.N = N; p.K = K;
  p.out = epi->out;
  p.shadow = reinterpret_cast<__nv_bfloat16*>(epi->shadow_bf16);
  p.resid = epi->resid_f32;
  p.rowscale = epi->rowscale;
  p.colscale = epi->colscale;
  p.bias = epi->bias;
  p.alpha = epi->alpha;
  p.ldc = epi->ldc;
  p.out_is_f32 = epi->out_is_f32;
  p.gated_gu = reinterpret_cast<const __nv_bfloat16*>(epi->gated_gu);
  p.gated_out = reinterpret_cast<__nv_bfloat16*>(epi->gated_out);
  p.gated_act = epi->gated_act;
  p.gated_cp = epi->gated_cp;
  p.gated_layout = epi->gated_layout;
  if (epi->gated_gu != nullptr && ((N % 32) != 0 || epi->colscale != nullptr || epi->bias != nullptr || epi->resid_f32 != nullptr ||
                                   epi->act_out != nullptr || (epi->gated_layout != 0 && epi->gated_layout != 1)))
    return set_error(LRP_ERR_ARG, "gemm: the fused gated backward needs N % 32 == 0 and no colscale / bias / resid term");
  p.act_out = reinterpret_cast<__nv_bfloat16*>(epi->act_out);
  p.delta_o = reinterpret_cast<const __nv_bfloat16*>(epi->delta_o);
  p.delta_out = epi->delta_out;
  p.delta_D = epi->delta_head_dim;
  p.delta_S = epi->delta_seq;
  if (p.delta_o != nullptr) {
    if (p.delta_out == nullptr || p.delta_D <= 0 || (p.delta_D % 32) != 0 || p.delta_D > 256 || (256 % p.delta_D) != 0 || (N % p.delta_D) != 0 ||
        p.delta_S <= 0 || (M % p.delta_S) != 0 || epi->out == nullptr || epi->out_is_f32 || epi->act_out != nullptr || epi->gated_gu != nullptr)
      return set_error(LRP_ERR_ARG, "gemm: the fused attention delta needs a bf16 [B*S, H*D] output with D in {32,64,128,256}");
  }
  if (p.act_out != nullptr) {
    if ((N % 64) != 0 || epi->out == nullptr || epi->out_is_f32 || epi->gated_gu != nullptr || epi->gated_act < 0 || epi->gated_act > 2 ||
        epi->colscale != nullptr || epi->resid_f32 != nullptr)
      return set_error(LRP_ERR_ARG, "gemm: the fused gated forward needs a bf16 output, N % 64 == 0 and no colscale / resid term");
  }
  bool group_forced = false;
  {
    // 16 m-blocks per group measured best on B200 across the Llama shapes (sweep 8/16/32/64 in
    // profiles/r01_gemm_group_m_sweep.txt): A panels of a group stay L2-resident while B panels stream past them.
    static const int forced = getenv("LRP_GROUP_M") ? atoi(getenv("LRP_GROUP_M")) : 0;
    const int g = forced > 0 ? forced : 16;
    p.group_m = g;
    group_forced = forced > 0;
  }
  int bn = force_bn;
  if (p.delta_o != nullptr && bn != 2 && N >= 256) bn = force_bn = 0;
  if (bn == 0) {
    // 256-wide tiles when they still fill the machine; otherwise 128-wide for more parallelism
    const int64_t tiles256 = int64_t((M + BM - 1) / BM) * ((N + 255) / 256);
    bn = (N >= 256 && tiles256 >= sm_count()) ? 256 : 128;
    if (p.delta_o != nullptr && bn < p.delta_D) bn = 256;   // a tile must hold whole heads
  }
  if (force_bn == 2) return gemm_bf16_pair(A, lda, B, ldb, b_layout, p, stream);   // forced CTA-pair kernel (tests)
  if (bn == 256 && force_bn == 0) {
    // CTA pairs (gemm2_sm100.cu) when the 256 x 256 tiles still fill the 74 pairs; LRP_GEMM_PAIR=0 keeps the one-CTA kernel
    static const int pair_mode = getenv("LRP_GEMM_PAIR") ? atoi(getenv("LRP_GEMM_PAIR")) : 1;
    const int64_t tiles_pair = int64_t((M + 255) / 256) * ((N + 255) / 256);
    if (pair_mode != 0 && tiles_pair >= sm_count() / 2) {
      // the pair kernel measured best with 4096-row groups (sweep 8/16/32/64 in units of 128 rows: 15.7 / 16.2 / 16.4 / 15.6 attr/s)
      // (LRP_GROUP_M_DEEPK=<n> sets another group for deep contractions, K >= 8192: the down projection re-reads its A operand
      //  5-6x from DRAM — 2.8-3.8 GB per launch against 0.59 GB algorithmic — but 8 / 16 / 32 measured the same step rate,
      //  16.8 / 16.9 / 17.0 attributions/s, so the default stays 32)
      static const int deepk = getenv("LRP_GROUP_M_DEEPK") ? atoi(getenv("LRP_GROUP_M_DEEPK")) : 0;
      if (!group_forced) p.group_m = (K >= 8192 && deepk > 0) ? deepk : 32;
      return gemm_bf16_pair(A, lda, B, ldb, b_layout, p, stream);
    }
  }
  if (bn == 256) {
    return b_layout == 0 ? launch_gemm<256, false>(A, lda, B, ldb, p, stream)
                         : launch_gemm<256, true>(A, lda, B, ldb, p, stream);
  } else if (bn == 128) {
    return b_layout == 0 ? launch_gemm<128, false>(A, lda, B, ldb, p, stream)
                         : launch_gemm<128, true>(A, lda, B, ldb, p, stream);
  }
  return set_error(LRP_ERR_ARG, "gemm: unsupported tile width");
}

}  // namespace lrp
