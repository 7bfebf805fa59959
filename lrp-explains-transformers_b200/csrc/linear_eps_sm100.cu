// Fused relevance-space eps-LRP rule of nn.Linear in ONE persistent launch (tcgen05 / TMA / TMEM).
//
//   reference: lxt/explicit/functional.py:325-364 (linear_epsilon_fn.backward)
//       z = x W^T + b ;  s = R_out / (z + eps)   [_stabilize = plain "+ eps", functional.py:266-273]
//       R_in = x  (s W)
//
// The rule is two dependent contractions.  A 128-row block of the R_in accumulator spans the whole K extent
// (128 x K fp32 = 2 MiB at K = 4096, TMEM holds 256 KiB), so z cannot stay on chip for the second contraction
// without recomputing it K/256 times.  Instead one persistent kernel runs a two-phase tile program:
//     phase 1 tile (m, n):  z = x[m] W[n]^T over K  ->  epilogue  s = R/(z + b + eps)  -> s_ws (bf16)
//     phase 2 tile (m, k):  acc = s[m] W[:, k] over N (W read in place, MN-major descriptor)
//                           -> epilogue  R_in = x[m, k]  acc
// Tiles are ordered in groups of GROUP m-blocks, phase 2 of group g trailing phase 1 of group g+1, so `s` for a
// group is produced and consumed while still resident in the 126 MB L2.  The phase-1 -> phase-2 dependency is a
// per-m-block counter in global memory (release by the epilogue warps, acquire by the TMA producer followed by a
// generic->async proxy fence).  All CTAs are co-resident (grid <= #SMs) and every CTA walks the tile list in
// order, so the waits cannot deadlock.  Total work 4 T K N flops, one launch, no second GEMM call.
#include "ptx_sm100.cuh"
#include "lrp_internal.h"

namespace lrp {

namespace eps {
constexpr int BM = 128, BN = 256, BK = 64, UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
constexpr int THREADS = 192;
constexpr int GROUP = 8;  // m-blocks per dependency group
}  // namespace eps

struct EpsParams {
  int T, N, K;
  float eps;
  const __nv_bfloat16* x;
  const float* bias;
  const void* r_out;
  void* r_in;
  __nv_bfloat16* s_ws;
  int32_t* flags;
  int r_is_f32;
};

struct EpsTile {
  int phase;  // 1 or 2
  int m_blk, c_blk;
};

// Tile program.  Per group g of GROUP m-blocks there are P1 = gsize*n1 phase-1 tiles and P2 = gsize*n2 phase-2
// tiles.  Emission order:  P1(0), P1(1), P2(0), P1(2), P2(1), ..., P2(last).
__device__ __forceinline__ EpsTile eps_tile(int t, int num_m, int n1, int n2) {
  using namespace eps;
  const int ngroups = (num_m + GROUP - 1) / GROUP;
  // walk the segments; ngroups is small (T/1024), so a linear walk per tile is cheap relative to a tile
  int seg_start = 0;
  for (int step = 0; step <= ngroups; ++step) {
    // segment A: phase-1 tiles of group `step` (if it exists)
    if (step < ngroups) {
      const int first = step * GROUP, gsize = min(GROUP, num_m - first);
      const int cnt = gsize * n1;
      if (t < seg_start + cnt) {
        const int r = t - seg_start;
        return EpsTile{1, first + r % gsize, r / gsize};
      }
      seg_start += cnt;
    }
    // segment B: phase-2 tiles of group `step-1`
    if (step >= 1) {
      const int first = (step - 1) * GROUP, gsize = min(GROUP, num_m - first);
      const int cnt = gsize * n2;
      if (t < seg_start + cnt) {
        const int r = t - seg_start;
        return EpsTile{2, first + r % gsize, r / gsize};
      }
      seg_start += cnt;
    }
  }
  return EpsTile{0, 0, 0};
}

__device__ __forceinline__ int ld_acquire(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(eps::THREADS, 1)
linear_eps_kernel(const __grid_constant__ CUtensorMap tm_x,    // x  [T,K]  box 64 x 128
                  const __grid_constant__ CUtensorMap tm_w1,   // W  [N,K]  box 64(k) x 256(n)     (phase 1, K-major B)
                  const __grid_constant__ CUtensorMap tm_s,    // s  [T,N]  box 64 x 128
                  const __grid_constant__ CUtensorMap tm_w2,   // W  [N,K]  box 64(k) x 64(n)      (phase 2, MN-major B)
                  const EpsParams p) {
  using namespace eps;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (p.T + BM - 1) / BM;
  const int n1 = (p.N + BN - 1) / BN;  // phase-1 column tiles (over N)
  const int n2 = (p.K + BN - 1) / BN;  // phase-2 column tiles (over K)
  const int num_tiles = num_m * (n1 + n2);
  const int kblocks1 = (p.K + BK - 1) / BK;
  const int kblocks2 = (p.N + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x); tma_prefetch_desc(&tm_w1); tma_prefetch_desc(&tm_s); tma_prefetch_desc(&tm_w2);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const EpsTile tl = eps_tile(t, num_m, n1, n2);
        if (tl.phase == 2) {
          // all n1 column tiles (4 epilogue warps each) of this row block must have published s
          while (ld_acquire(p.flags + tl.m_blk) < 4 * n1) __nanosleep(64);
          asm volatile("fence.proxy.async;" ::: "memory");
        }
        const int nk = tl.phase == 1 ? kblocks1 : kblocks2;
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          if (tl.phase == 1) {
            tma_load_2d(sa, &tm_x, &full_bar[stage], kb * BK, tl.m_blk * BM);
            tma_load_2d(sb, &tm_w1, &full_bar[stage], kb * BK, tl.c_blk * BN);
          } else {
            tma_load_2d(sa, &tm_s, &full_bar[stage], kb * BK, tl.m_blk * BM);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sb + j * (64 * BK * 2), &tm_w2, &full_bar[stage], tl.c_blk * BN + j * 64, kb * BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc1 = make_idesc_bf16(BM, BN, 0, 0);
    constexpr uint32_t idesc2 = make_idesc_bf16(BM, BN, 0, 1);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const EpsTile tl = eps_tile(t, num_m, n1, n2);
      const int nk = tl.phase == 1 ? kblocks1 : kblocks2;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < nk; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
          const uint64_t adesc = make_sdesc_sw128(sa, 16, 1024);
          const bool mn = tl.phase == 2;
          const uint64_t bdesc = mn ? make_sdesc_sw128(sb, 64 * BK * 2, 1024) : make_sdesc_sw128(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t a_adv = uint64_t((k * UMMA_K * 2) >> 4);
            const uint64_t b_adv = mn ? uint64_t((k * UMMA_K * 128) >> 4) : uint64_t((k * UMMA_K * 2) >> 4);
            tc_mma_ss(tmem_d, adesc + a_adv, bdesc + b_adv, mn ? idesc2 : idesc1, (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);
          if (kb == nk - 1) tc_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const EpsTile tl = eps_tile(t, num_m, n1, n2);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int m = tl.m_blk * BM + q * 32 + lane;
      const bool row_ok = m < p.T;
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * BN;
      const int width = tl.phase == 1 ? p.N : p.K;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(taddr + c * 32, v);
        tmem_ld_wait();
        const int n0 = tl.c_blk * BN + c * 32;
        if (row_ok && n0 < width) {
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            const int n = n0 + j8 * 8;
            if (n >= width) continue;
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[j8 * 8 + j]);
            if (tl.phase == 1) {
              // s = R / (z + b + eps)
              float r[8];
              if (p.r_is_f32) {
                const float* rp = reinterpret_cast<const float*>(p.r_out) + int64_t(m) * p.N + n;
                const float4 a = *reinterpret_cast<const float4*>(rp);
                const float4 b4 = *reinterpret_cast<const float4*>(rp + 4);
                r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b4.x; r[5] = b4.y; r[6] = b4.z; r[7] = b4.w;
              } else {
                const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.r_out) + int64_t(m) * p.N + n);
                r[0] = bf16_lo(u.x); r[1] = bf16_hi(u.x); r[2] = bf16_lo(u.y); r[3] = bf16_hi(u.y);
                r[4] = bf16_lo(u.z); r[5] = bf16_hi(u.z); r[6] = bf16_lo(u.w); r[7] = bf16_hi(u.w);
              }
              if (p.bias != nullptr) {
                const float4 b0 = *reinterpret_cast<const float4*>(p.bias + n);
                const float4 b1 = *reinterpret_cast<const float4*>(p.bias + n + 4);
                f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
                f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = r[j] / (f[j] + p.eps);
              *reinterpret_cast<uint4*>(p.s_ws + int64_t(m) * p.N + n) =
                  make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
            } else {
              // R_in = x * acc
              const uint4 u = *reinterpret_cast<const uint4*>(p.x + int64_t(m) * p.K + n);
              f[0] *= bf16_lo(u.x); f[1] *= bf16_hi(u.x); f[2] *= bf16_lo(u.y); f[3] *= bf16_hi(u.y);
              f[4] *= bf16_lo(u.z); f[5] *= bf16_hi(u.z); f[6] *= bf16_lo(u.w); f[7] *= bf16_hi(u.w);
              if (p.r_is_f32) {
                float* o = reinterpret_cast<float*>(p.r_in) + int64_t(m) * p.K + n;
                *reinterpret_cast<float4*>(o) = make_float4(f[0], f[1], f[2], f[3]);
                *reinterpret_cast<float4*>(o + 4) = make_float4(f[4], f[5], f[6], f[7]);
              } else {
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.r_in) + int64_t(m) * p.K + n) =
                    make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
              }
            }
          }
        }
      }
      tc_fence_before();
      if (tl.phase == 1) __threadfence();  // release: this warp's s rows are visible before the counter bump
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&tmem_empty[acc]);
        if (tl.phase == 1) atomicAdd(p.flags + tl.m_blk, 1);  // 4 warp-arrivals per finished phase-1 tile
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int linear_eps_bwd(const void* x, const void* W, const float* bias, const void* r_out, int r_is_f32, void* r_in,
                   void* s_ws, int32_t* flags_ws, int T, int N, int K, float epsv, cudaStream_t stream) {
  using namespace eps;
  if (T <= 0 || N <= 0 || K <= 0) return set_error(LRP_ERR_ARG, "linear_eps_bwd: empty problem");
  if ((N % 8) || (K % 8)) return set_error(LRP_ERR_ARG, "linear_eps_bwd: N and K must be multiples of 8");
  if (!x || !W || !r_out || !r_in || !s_ws || !flags_ws) return set_error(LRP_ERR_ARG, "linear_eps_bwd: null pointer");
  CUtensorMap tx, tw1, ts, tw2;
  if (int e = make_tmap_2d_bf16(&tx, x, K, T, K, 64, BM)) return e;
  if (int e = make_tmap_2d_bf16(&tw1, W, K, N, K, 64, BN)) return e;
  if (int e = make_tmap_2d_bf16(&ts, s_ws, N, T, N, 64, BM)) return e;
  if (int e = make_tmap_2d_bf16(&tw2, W, K, N, K, 64, BK)) return e;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t ce = cudaFuncSetAttribute(linear_eps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (ce != cudaSuccess) return set_error(LRP_ERR_CUDA, cudaGetErrorString(ce));
    attr_done = true;
  }
  EpsParams p;
  p.T = T; p.N = N; p.K = K; p.eps = epsv;
  p.x = reinterpret_cast<const __nv_bfloat16*>(x);
  p.bias = bias;
  p.r_out = r_out; p.r_in = r_in;
  p.s_ws = reinterpret_cast<__nv_bfloat16*>(s_ws);
  p.flags = flags_ws;
  p.r_is_f32 = r_is_f32;
  const int num_m = (T + BM - 1) / BM;
  const int num_tiles = num_m * ((N + BN - 1) / BN + (K + BN - 1) / BN);
  const int grid = num_tiles < sm_count() ? num_tiles : sm_count();
  linear_eps_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(tx, tw1, ts, tw2, p);
  LRP_CHECK_LAUNCH();
  return LRP_OK;
}

}  // namespace lrp

extern "C" {

int64_t lrp_linear_eps_flags_count(int T) { return int64_t((T + lrp::eps::BM - 1) / lrp::eps::BM); }

int lrp_linear_eps_bwd(const void* x, const void* W, const float* bias, const void* r_out, int r_is_f32, void* r_in,
                       void* s_ws, int32_t* flags_ws, int T, int N, int K, float eps, void* stream) {
  return lrp::linear_eps_bwd(x, W, bias, r_out, r_is_f32, r_in, s_ws, flags_ws, T, N, K, eps,
                             static_cast<cudaStream_t>(stream));
}

}  // extern "C"
