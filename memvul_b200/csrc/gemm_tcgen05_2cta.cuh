// CTA-pair (cta_group::2) variant of the persistent tcgen05 GEMM: two SMs of a TPC cooperate on a 256 x 256
// output tile.  Each CTA stages its own 128 A rows and HALF of the 256 B rows per K-block (32 KB/stage instead
// of 48 KB), the leader CTA issues tcgen05.mma.cta_group::2 (M=256, N=256, K=16) which reads A/B from both
// CTAs' shared memory and accumulates each CTA's 128 rows into its own TMEM.  Per SM this cuts both the TMA
// write traffic and the MMA operand read traffic of shared memory by a third versus the single-CTA kernel,
// whose ncu capture (profiles/r01b_*) showed the tensor pipe idle ~50 % of the time with the smem ring full.
//
//   warp 0      : TMA producer (both CTAs; transaction bytes are credited to the LEADER's full barrier)
//   warp 1      : MMA issuer   (leader CTA only); tcgen05.commit multicast frees the smem slot in both CTAs
//   warp 2      : TMEM allocator (cta_group::2 alloc/dealloc, same warp in both CTAs)
//   warps 4..11 : epilogue of the CTA's own 128 x 256 half.  Every warp owns a 32-row x 128-column strip and
//                 moves it through a private 128B-swizzled shared-memory staging buffer:
//                   fp16 outputs : registers -> smem (conflict-free STS.128) -> TMA store of 32 x 64 boxes
//                   fp32 + resid : TMA load of the residual 32 x 32 box (ping-pong buffers, the next box always
//                                  in flight) -> add in place -> TMA store
//                 so global traffic is full 128 B lines issued by the TMA engine instead of 32 partial sectors
//                 per LSU instruction (the r01b mainloop-only experiment showed the old per-lane epilogue cost
//                 30-70 % of each GEMM).  The TMEM chunk for step c+1 is in flight while chunk c is processed.
// Epilogue warps of both CTAs arrive on the leader's "TMEM empty" barrier (remote mbarrier arrive).
#pragma once
#include "gemm_tcgen05.cuh"

namespace mv {

#ifndef MV_GEMM_STG2
#define MV_GEMM_STG2 0          // 1: two staging buffers per epilogue warp for the fp16-output epilogues.  Measured r02h:
                                // QKV 87.9 -> 90.2 us, FFN-up 124.6 -> 126.7 us, i.e. the store cost is NOT a warp waiting for
                                // its staging buffer to drain; kept for the record
#endif
#ifndef MV_GELU_PACKED
#define MV_GELU_PACKED 1        // erf-GELU epilogue on fp32 pairs (FFMA2 / FMUL2): fewer issue slots per element
#endif

template <int EPI>
struct Gemm2Cfg {
  static constexpr bool RESID = (EPI == EPI_BIAS_RESID_F32);
  static constexpr int BM = 256, BM_CTA = 128, BN = 256, BN_CTA = 128, BK = 64;
#ifndef MV_GEMM_RESID_STAGES
#define MV_GEMM_RESID_STAGES 4    // experiment (r02o): 5 = five A/B stages for the residual epilogue too, paid for with ONE
                                  // staging buffer per epilogue warp (serialised residual load -> add -> store)
#endif
  static constexpr int STAGES = RESID ? MV_GEMM_RESID_STAGES : 5;            // measured: 3 stages starve the K=3072 main loop
  static constexpr int A_BYTES = BM_CTA * BK * 2;          // 16 KB
  static constexpr int B_BYTES = BN_CTA * BK * 2;          // 16 KB (half of the tile's B rows)
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;    // 32 KB per CTA
  static constexpr int TMEM_COLS = 2 * BN;                 // two accumulator stages of 256 fp32 columns
  static constexpr int THREADS = 384;
  static constexpr int EPI_WARPS = 8;
  static constexpr int STG_BYTES = 4096;                   // 32 rows x 128 B, SWIZZLE_128B
  // RESID: ping-pong, next residual box in flight.  fp16 outputs: ping-pong too (MV_GEMM_STG2) -- the TMA unit serves an
  // SM's requests in order, so an output box queues behind every main-loop stage in flight (~2.5 k cycles) and a warp
  // that must see its single staging buffer drained before packing the next box spends the epilogue waiting
  static constexpr int STG_BUFS = ((RESID && MV_GEMM_RESID_STAGES < 5) || MV_GEMM_STG2) ? 2 : 1;
  static constexpr int OFF_STG = STAGES * STAGE_BYTES;
  // bias slices: a private 128-float copy per epilogue warp (the warps are not in lock step across tiles).  With two
  // fp16 staging buffers the 4 KB no longer fit next to five main-loop stages; those epilogues read the bias through
  // the read-only cache instead (same address in every lane: one broadcast transaction)
  static constexpr bool BIAS_SMEM = RESID || !MV_GEMM_STG2;
  static constexpr int OFF_BIAS = OFF_STG + EPI_WARPS * STG_BUFS * STG_BYTES;
  static constexpr int OFF_BAR = OFF_BIAS + (BIAS_SMEM ? EPI_WARPS * 128 * 4 : 0);
  static constexpr int SMEM_BYTES = OFF_BAR + 512;
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared-memory limit");
};

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
gemm_f16_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                             const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res,
                             int M, int N, int K, const float* __restrict__ bias, int store, void* out_ptr,
                             const int* __restrict__ m_dev) {
  using Cfg = Gemm2Cfg<EPI>;
  if (m_dev) M = min(M, __ldg(m_dev));     // packed (var-len) batches: the row count lives on the device
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tfull_bar = empty_bar + Cfg::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* res_bar = tempty_bar + 2;                      // [EPI_WARPS][2] residual-landed barriers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 2 * Cfg::EPI_WARPS);
  float* bias_smem = reinterpret_cast<float*>(smem + Cfg::OFF_BIAS);

  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = static_cast<int>(threadIdx.x & 31);
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int idle_tma = (store >> 8) & 3, idle_mma = (store >> 10) & 3, idle_epi = (store >> 12) & 3;     // how the single-lane warps wait (mbar_wait_idle)
  store &= 0xff;

  if (warp_idx == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_out);
    if constexpr (Cfg::RESID) prefetch_tmap(&tmap_res);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);        // leader's producer arrives once with the pair's total byte count
      mbar_init(&empty_bar[i], 1);       // one multicast tcgen05.commit per use
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * Cfg::EPI_WARPS);     // epilogue warps of BOTH CTAs
    }
    for (int i = 0; i < 2 * Cfg::EPI_WARPS; ++i) mbar_init(&res_bar[i], 1);
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();                    // barrier inits + TMEM allocation visible to the peer CTA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_n = N / Cfg::BN;
  const int tiles_m = (M + Cfg::BM - 1) / Cfg::BM;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = K / Cfg::BK;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp_idx == 0) {
    // ===================== TMA producer (both CTAs) =====================
    // warp-uniform loop, one elected lane issues (same reason as the MMA warp below: no per-instruction waterfall)
    {
      const bool issuer = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m_blk = tile / tiles_n, n_blk = tile % tiles_n;
        const int row_a = m_blk * Cfg::BM + static_cast<int>(cta_rank) * Cfg::BM_CTA;
        const int row_b = n_blk * Cfg::BN + static_cast<int>(cta_rank) * Cfg::BN_CTA;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_idle(&empty_bar[stage], phase ^ 1u, idle_tma);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          if (issuer) {
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
            tma_load_2d_pair(sa, &tmap_a, &full_bar[stage], kb * Cfg::BK, row_a, kEvictNormal);
            tma_load_2d_pair(sa + Cfg::A_BYTES, &tmap_b, &full_bar[stage], kb * Cfg::BK, row_b, kEvictLast);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    // The whole warp walks the loop (uniform control flow, so the descriptor arithmetic stays on the uniform datapath)
    // and one elected lane issues: with `if (lane == 0)` around everything ptxas wrapped EVERY tcgen05.mma in an
    // ELECT / 5x R2UR / BRA.U.ANY waterfall (~16 SASS instructions per MMA, ~95 per K-block), and under the
    // erf-GELU epilogue warps sharing this scheduler the issue thread fell behind the tensor pipe (r01n/r01o).
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_f16(Cfg::BM, Cfg::BN, false, false);
      const bool issuer = elect_one();
      const uint32_t smem_base = smem_u32(smem);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        mbar_wait_idle(&tempty_bar[acc], acc_phase ^ 1u, idle_mma);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * Cfg::BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_idle(&full_bar[stage], phase, idle_mma);
          tc_fence_after();
          const uint32_t sa = smem_base + static_cast<uint32_t>(stage * Cfg::STAGE_BYTES);
          const uint64_t a_desc = umma_desc_sw128(sa);
          const uint64_t b_desc = umma_desc_sw128(sa + Cfg::A_BYTES);
          if (issuer) {
#pragma unroll
            for (int k = 0; k < Cfg::BK / 16; ++k)
              umma_f16_ss_pair(d_tmem, a_desc + static_cast<uint64_t>(k * 2), b_desc + static_cast<uint64_t>(k * 2), idesc,
                               (kb | k) != 0 ? 1u : 0u);
            umma_commit_pair(&empty_bar[stage], 0b11);     // frees the slot in both CTAs
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
        if (issuer) umma_commit_pair(&tfull_bar[acc], 0b11);   // accumulators of both CTAs are complete
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp_idx >= 4) {
    // ===================== epilogue (own 128 rows x 256 columns) =====================
    const int ew = warp_idx - 4;
    const int quarter = ew & 3;                    // TMEM lane quarter == rows quarter*32 .. +31 of the CTA's 128
    const int half_sel = ew >> 2;                  // column half of the 256-wide tile
    constexpr int COLS_PER_WARP = Cfg::BN / 2;     // 128
    constexpr int NCHUNK = COLS_PER_WARP / 32;     // 4 TMEM chunks of 32 fp32 columns
    float* my_bias = bias_smem + (Cfg::BIAS_SMEM ? ew * 128 : 0);
    uint8_t* my_stg = smem + Cfg::OFF_STG + ew * Cfg::STG_BUFS * Cfg::STG_BYTES;
    uint64_t* my_res_bar = res_bar + 2 * ew;
    const uint32_t sw = static_cast<uint32_t>(lane & 7);          // 128B-swizzle phase of this lane's staging row
    uint8_t* my_row0 = my_stg + lane * 128;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t gc = 0;                               // running residual-chunk counter: buffer gc & 1, parity (gc >> 1) & 1
    uint32_t bx = 0;                               // running fp16 output-box counter: staging buffer bx & 1

    auto strip_row0 = [&](int tile) { return (tile / tiles_n) * Cfg::BM + static_cast<int>(cta_rank) * Cfg::BM_CTA + quarter * 32; };
    auto strip_col0 = [&](int tile) { return (tile % tiles_n) * Cfg::BN + half_sel * COLS_PER_WARP; };

    // residual chunk q of this warp's sequence: tile = cluster_id + (q / 4) * num_clusters, columns (q % 4) * 32
    auto issue_res_load = [&](uint32_t q) {
      const int t = cluster_id + static_cast<int>(q / NCHUNK) * num_clusters;
      if (t < num_tiles) {
        const uint32_t nb = Cfg::STG_BUFS == 2 ? (q & 1u) : 0u;
        mbar_arrive_expect_tx(&my_res_bar[nb], Cfg::STG_BYTES);
        tma_load_2d(my_stg + nb * Cfg::STG_BYTES, &tmap_res, &my_res_bar[nb], strip_col0(t) + static_cast<int>(q % NCHUNK) * 32,
                    strip_row0(t), kEvictFirst);
      }
    };
    // store == 5: fp32 output WITHOUT a residual (MEMVUL_EPI_BIAS_F32, the accuracy mode's Q K V / FFN-up products):
    // same staging / TMA-store path, nothing is loaded and the staging buffer is only written
    const bool no_resid = (store == 5);
    if constexpr (Cfg::RESID) {
      if (lane == 0 && store && !no_resid) {
        issue_res_load(0);
        if (Cfg::STG_BUFS == 2) issue_res_load(1);
      }
    }
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int row0 = strip_row0(tile), col0 = strip_col0(tile);
      // bias slice of this warp -> smem (independent of the accumulator: overlaps the wait below)
      if constexpr (Cfg::BIAS_SMEM) {
        *reinterpret_cast<float4*>(my_bias + lane * 4) = __ldg(reinterpret_cast<const float4*>(bias + col0) + lane);
        __syncwarp();
      }
      mbar_wait_idle(&tfull_bar[acc], acc_phase, idle_epi);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                              static_cast<uint32_t>(acc * Cfg::BN + half_sel * COLS_PER_WARP);
      uint32_t r[2][32];
      tmem_ld_32x32b_x32(t_addr, r[0]);
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c) {
        tmem_wait_ld();
        if (c + 1 < NCHUNK) tmem_ld_32x32b_x32(t_addr + (c + 1) * 32, r[(c + 1) & 1]);
        const uint32_t(&v)[32] = r[c & 1];
        const float* bsm = Cfg::BIAS_SMEM ? my_bias + c * 32 : bias + col0 + c * 32;
        if constexpr (Cfg::RESID) {
          const uint32_t b = Cfg::STG_BUFS == 2 ? (gc & 1u) : 0u;
          uint8_t* rowp = my_row0 + b * Cfg::STG_BYTES;
          if (store) {
            if (!no_resid) mbar_wait_idle(&my_res_bar[b], (Cfg::STG_BUFS == 2 ? (gc >> 1) : gc) & 1u, idle_epi);   // residual chunk has landed
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              float4* p = reinterpret_cast<float4*>(rowp + ((static_cast<uint32_t>(u) ^ sw) << 4));
              const float4 x = no_resid ? make_float4(0.f, 0.f, 0.f, 0.f) : *p;
              const float4 bb = *reinterpret_cast<const float4*>(bsm + 4 * u);
              float4 o;
              o.x = __uint_as_float(v[4 * u + 0]) + bb.x + x.x;
              o.y = __uint_as_float(v[4 * u + 1]) + bb.y + x.y;
              o.z = __uint_as_float(v[4 * u + 2]) + bb.z + x.z;
              o.w = __uint_as_float(v[4 * u + 3]) + bb.w + x.w;
              *p = o;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmap_out, my_stg + b * Cfg::STG_BYTES, col0 + c * 32, row0);
              bulk_commit_group();
              bulk_wait_read_all();                // this store has left the buffer: refill it with chunk gc+2
              if (!no_resid) issue_res_load(gc + Cfg::STG_BUFS);
            }
          }
          ++gc;
        } else {
          // fp16: two 32-column TMEM chunks fill one 32 x 64 staging box (128 B per row)
          const bool direct = (store == 2);          // experiment: 256-bit per-lane global stores, no smem staging
          uint8_t* box_row0 = my_row0 + ((MV_GEMM_STG2 ? (bx & 1u) : 0u) * Cfg::STG_BYTES);
          if ((c & 1) == 0 && store == 1) {
            if (lane == 0) {
              if (MV_GEMM_STG2) bulk_wait_read_1();   // the box before the previous one has left this buffer
              else bulk_wait_read_all();             // previous box has been read out of the staging buffer
            }
            __syncwarp();
          }
          if (store && store != 4) {
            uint32_t pkd[16];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float4 b0 = *reinterpret_cast<const float4*>(bsm + 8 * u);
              const float4 b1 = *reinterpret_cast<const float4*>(bsm + 8 * u + 4);
              float x[8];
              x[0] = __uint_as_float(v[8 * u + 0]) + b0.x;
              x[1] = __uint_as_float(v[8 * u + 1]) + b0.y;
              x[2] = __uint_as_float(v[8 * u + 2]) + b0.z;
              x[3] = __uint_as_float(v[8 * u + 3]) + b0.w;
              x[4] = __uint_as_float(v[8 * u + 4]) + b1.x;
              x[5] = __uint_as_float(v[8 * u + 5]) + b1.y;
              x[6] = __uint_as_float(v[8 * u + 6]) + b1.z;
              x[7] = __uint_as_float(v[8 * u + 7]) + b1.w;
              if constexpr (EPI == EPI_BIAS_GELU_F16) {
#if MV_GELU_PACKED
#pragma unroll
                for (int t = 0; t < 8; t += 2) gelu_erf_fast_x2(x[t], x[t + 1]);
#else
#pragma unroll
                for (int t = 0; t < 8; ++t) x[t] = gelu_erf_fast(x[t]);
#endif
              }
              pkd[4 * u + 0] = pack_half2(x[0], x[1]);
              pkd[4 * u + 1] = pack_half2(x[2], x[3]);
              pkd[4 * u + 2] = pack_half2(x[4], x[5]);
              pkd[4 * u + 3] = pack_half2(x[6], x[7]);
            }
            if (direct) {
              const int grow = row0 + lane;
              if (grow < M) {
                __half* orow = reinterpret_cast<__half*>(out_ptr) + static_cast<size_t>(grow) * N + col0 + c * 32;
                const uint32_t(&lo)[8] = *reinterpret_cast<const uint32_t(*)[8]>(&pkd[0]);
                const uint32_t(&hi)[8] = *reinterpret_cast<const uint32_t(*)[8]>(&pkd[8]);
                st_global_v8(orow, lo);
                st_global_v8(orow + 16, hi);
              }
            } else {
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const uint32_t unit = static_cast<uint32_t>((c & 1) * 4 + u);
                *reinterpret_cast<uint4*>(box_row0 + ((unit ^ sw) << 4)) =
                    make_uint4(pkd[4 * u], pkd[4 * u + 1], pkd[4 * u + 2], pkd[4 * u + 3]);
              }
              if ((c & 1) && store != 3) {
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                  tma_store_2d(&tmap_out, box_row0 - lane * 128, col0 + (c >> 1) * 64, row0);
                  bulk_commit_group();
                }
              }
              if (c & 1) ++bx;
            }
          } else if (store == 4 && (c & 1)) {          // experiment: stores without the math / staging writes
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmap_out, my_stg, col0 + (c >> 1) * 64, row0);
              bulk_commit_group();
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();                                      // all lanes done with TMEM and with my_bias
      if (lane == 0) mbar_arrive_cluster(&tempty_bar[acc], 0);     // leader's barrier (remote for CTA 1)
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (lane == 0) bulk_wait_read_all();                 // staging smem must outlive the last TMA store's read
  }

  // ===================== teardown =====================
  tc_fence_before();
  cluster_sync_all();                    // nobody may still target the peer's smem / TMEM
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace mv
