// Persistent warp-specialised tcgen05 GEMM for sm_100a:
//     C[M,N] = epilogue(A[M,K] (fp16, row-major)  x  W[N,K]^T (fp16, row-major = nn.Linear layout))
// with fp32 accumulation in TMEM.  Replaces the cuBLAS SGEMM + separate bias / GELU / residual
// kernels that HF BertModel issues per layer (reference entry
// MemVul/custom_PTM_embedder.py:228).
//
// Roles (384 threads, 1 CTA / SM, grid = min(#tiles, #SMs), static round-robin tile schedule):
//   warp 0      : TMA producer   (cp.async.bulk.tensor 2D, SWIZZLE_128B, STAGES-deep smem ring)
//   warp 1      : MMA issuer     (one elected lane; tcgen05.mma cta_group::1 kind::f16, M=128 N=BN K=16)
//   warp 2      : TMEM allocator (2 accumulator stages x BN fp32 columns)
//   warps 4..11 : epilogue       (tcgen05.ld 32x32b.x32 -> bias / erf-GELU / fp32 residual -> global)
// Pipelines: smem full/empty mbarriers (TMA <-> MMA) and TMEM full/empty mbarriers (MMA <-> epilogue),
// so the epilogue of tile i overlaps the main loop of tile i+1.
#pragma once
#include "ptx.cuh"

namespace mv {

enum GemmEpilogue : int {
  EPI_BIAS_F16 = 0,        // out fp16 = acc + bias                      (K2: QKV projection)
  EPI_BIAS_GELU_F16 = 1,   // out fp16 = gelu_erf(acc + bias)            (K5: FFN up)
  EPI_BIAS_RESID_F32 = 2,  // out fp32 = acc + bias + resid (fp32)       (K4/K6 pre-LayerNorm sum)
};

template <int BN>
struct GemmCfg {
  static constexpr int BM = 128;
  static constexpr int BK = 64;                         // 64 fp16 = 128 B = one swizzle row
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;              // two accumulator stages
  static constexpr int THREADS = 384;
  static constexpr int EPI_WARPS = 8;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(BN == 128 || BN == 256, "BN must be 128 or 256");
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

template <int BN, int EPI>
__global__ void __launch_bounds__(GemmCfg<BN>::THREADS, 1)
gemm_f16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                        int M, int N, int K, const float* __restrict__ bias, const float* resid, void* out,
                        int ldo, const int* __restrict__ m_dev) {
  using Cfg = GemmCfg<BN>;
  if (m_dev) M = min(M, __ldg(m_dev));     // packed (var-len) batches: the row count lives on the device
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tfull_bar = empty_bar + Cfg::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = static_cast<int>(threadIdx.x & 31);

  if (warp_idx == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], Cfg::EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_n = N / BN;
  const int tiles_m = (M + Cfg::BM - 1) / Cfg::BM;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = K / Cfg::BK;

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / tiles_n, n_blk = tile % tiles_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * Cfg::BK, m_blk * Cfg::BM, kEvictNormal);
          tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * Cfg::BK, n_blk * BN, kEvictLast);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(Cfg::BM, BN, false, false);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);       // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);              // TMA bytes have landed
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t a_desc = umma_desc_sw128(sa);
          const uint64_t b_desc = umma_desc_sw128(sa + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < Cfg::BK / 16; ++k) {
            // +32 B per K=16 step inside the 128 B swizzle row  (encoded >>4 => +2)
            umma_f16_ss(d_tmem, a_desc + static_cast<uint64_t>(k * 2), b_desc + static_cast<uint64_t>(k * 2), idesc,
                        (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);                  // frees the smem slot when the MMAs retire
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tfull_bar[acc]);                      // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp_idx >= 4) {
    // ===================== epilogue =====================
    const int ew = warp_idx - 4;
    const int quarter = ew & 3;            // == warp_idx % 4: the TMEM lane quarter this warp may touch
    const int half_sel = ew >> 2;          // which half of the BN columns
    constexpr int COLS_PER_WARP = BN / 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / tiles_n, n_blk = tile % tiles_n;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * Cfg::BM + quarter * 32 + lane;
      const bool row_ok = row < M;
#pragma unroll 1
      for (int c = 0; c < COLS_PER_WARP / 32; ++c) {
        const int col0 = half_sel * COLS_PER_WARP + c * 32;
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                               static_cast<uint32_t>(acc * BN + col0), r);
        tmem_wait_ld();
        const int gcol = n_blk * BN + col0;
        const float4* bias4 = reinterpret_cast<const float4*>(bias + gcol);
        if constexpr (EPI == EPI_BIAS_RESID_F32) {
          float* orow = reinterpret_cast<float*>(out) + static_cast<size_t>(row) * ldo + gcol;
          const float* rrow = resid ? resid + static_cast<size_t>(row) * ldo + gcol : nullptr;   // nullptr: fp32 output without a residual (MEMVUL_EPI_BIAS_F32)
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = __ldg(bias4 + j);
              const float4 x = rrow ? *reinterpret_cast<const float4*>(rrow + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
              float4 o;
              o.x = __uint_as_float(r[4 * j + 0]) + b.x + x.x;
              o.y = __uint_as_float(r[4 * j + 1]) + b.y + x.y;
              o.z = __uint_as_float(r[4 * j + 2]) + b.z + x.z;
              o.w = __uint_as_float(r[4 * j + 3]) + b.w + x.w;
              *reinterpret_cast<float4*>(orow + 4 * j) = o;
            }
          }
        } else {
          __half* orow = reinterpret_cast<__half*>(out) + static_cast<size_t>(row) * ldo + gcol;
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 b0 = __ldg(bias4 + 2 * j), b1 = __ldg(bias4 + 2 * j + 1);
              float v[8];
              v[0] = __uint_as_float(r[8 * j + 0]) + b0.x;
              v[1] = __uint_as_float(r[8 * j + 1]) + b0.y;
              v[2] = __uint_as_float(r[8 * j + 2]) + b0.z;
              v[3] = __uint_as_float(r[8 * j + 3]) + b0.w;
              v[4] = __uint_as_float(r[8 * j + 4]) + b1.x;
              v[5] = __uint_as_float(r[8 * j + 5]) + b1.y;
              v[6] = __uint_as_float(r[8 * j + 6]) + b1.z;
              v[7] = __uint_as_float(r[8 * j + 7]) + b1.w;
              if constexpr (EPI == EPI_BIAS_GELU_F16) {
#pragma unroll
                for (int t = 0; t < 8; ++t) v[t] = gelu_erf(v[t]);
              }
              uint4 pk;
              pk.x = pack_half2(v[0], v[1]);
              pk.y = pack_half2(v[2], v[3]);
              pk.z = pack_half2(v[4], v[5]);
              pk.w = pack_half2(v[6], v[7]);
              *reinterpret_cast<uint4*>(orow + 8 * j) = pk;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  // ===================== teardown =====================
  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace mv
