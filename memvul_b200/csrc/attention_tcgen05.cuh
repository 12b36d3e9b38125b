// Fused self-attention for one (sequence, head, 128-query tile) per CTA on tcgen05:
//     ctx = softmax(Q K^T / sqrt(64) + key_mask) V          
// replacing HF BertSelfAttention's matmul / div / add-mask / softmax / matmul chain
// (transformers 4.1.0, entered from MemVul/custom_PTM_embedder.py:228) that round-trips a
// [B,12,S,S] fp32 score tensor through HBM.
//
// Input  : qkv fp16 [B*S, 3*H] row-major (Q | K | V column blocks, head h at columns h*64)
//          read through ONE 2-D TMA map with box {64 cols, 128 rows}, SWIZZLE_128B.
// Output : ctx fp16 [B*S, H] row-major (head h at columns h*64).
// Masking: keys >= len[b] are excluded.  The reference adds -10000 to their scores, whose
//          exp underflows to exactly 0 in fp32, so exclusion is bit-equivalent; fully padded
//          key blocks and fully padded query tiles are skipped (padded query rows are never
//          consumed: BertPooler reads row 0 only, MemVul/model_memory.py:99).
//
// Warps 0-3 : softmax (one query row per thread; S read from TMEM, P written to smem as the
//             fp16 A operand of the second MMA, running max / sum in fp32, O rescaled in TMEM)
// Warp 4    : lane 0 issues the TMA loads and both tcgen05.mma streams.
// Key blocks of 64: S = Q K_j^T is DOUBLE-buffered in TMEM (2 x 64 columns) and P in smem (2 x 16 KB), so the tensor
// core computes S_{j+1} while the softmax warps work on S_j (the r01e capture showed them stalled 25 % of the time
// on the S barrier with a single buffer).  Footprint: 112 KB smem (Q 16 K, 4-stage K/V ring 64 K, P 32 K) and 256
// TMEM columns (S 2 x 64, O 64), so TWO CTAs are resident per SM.  Key length <= 512 (8 blocks of 64).
#pragma once
#include "ptx.cuh"

namespace mv {

struct AttnCfg {
  static constexpr int BQ = 128, BKV = 64, DH = 64, MAX_KB = 8, KV_STAGES = 4;
  static constexpr int Q_BYTES = 128 * 64 * 2;             // 16 KB: {64 x 128} fp16 box
  static constexpr int KV_BYTES = BKV * 64 * 2;            // 8 KB: {64 x 64} fp16 box
  static constexpr int P_BYTES = 128 * BKV * 2;            // 16 KB: 128 x 64 fp16 = one swizzled K-chunk
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + Q_BYTES;
  static constexpr int OFF_V = OFF_K + KV_STAGES * KV_BYTES;
  static constexpr int OFF_P = OFF_V + KV_STAGES * KV_BYTES;
  static constexpr int OFF_BAR = OFF_P + 2 * P_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 512;         // 115,200 B: two CTAs fit in 228 KB
  static constexpr int THREADS = 160;
  static constexpr int TMEM_COLS = 256;
  static constexpr int TM_S = 0, TM_O = 128;
};

__global__ void __launch_bounds__(AttnCfg::THREADS, 2)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_kv,
                         const int* __restrict__ lens,
                         __half* __restrict__ ctx, int S, int H) {
  using C = AttnCfg;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int len = lens[b];
  const int q0 = qt * C::BQ;
  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = static_cast<int>(threadIdx.x & 31);
  const size_t row_base = static_cast<size_t>(b) * S;

  if (q0 >= len) {
    // fully padded query tile: deterministic zeros, no tensor work
    const int rows = min(C::BQ, S - q0);
    for (int i = threadIdx.x; i < rows * 8; i += blockDim.x) {
      const int r = i >> 3, u = i & 7;
      *reinterpret_cast<uint4*>(ctx + (row_base + q0 + r) * H + h * C::DH + u * 8) = make_uint4(0, 0, 0, 0);
    }
    return;
  }

  extern __shared__ __align__(1024) uint8_t smem[];        // SWIZZLE_128B tiles need 1024-byte alignment
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* q_full = bars;                           // [1]
  uint64_t* k_full = bars + 1;                       // [8]
  uint64_t* v_full = k_full + C::MAX_KB;             // [8]
  uint64_t* s_full = v_full + C::MAX_KB;             // [8]  QK^T of block j complete
  uint64_t* p_full = s_full + C::MAX_KB;             // [8]  P_j in smem, S_j drained, O rescaled   (4 warp arrivals)
  uint64_t* pv_done = p_full + C::MAX_KB;            // [8]  P_j V_j accumulated into O
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + C::MAX_KB);

  const int nkb = (len + C::BKV - 1) / C::BKV;     // 1..8 key blocks; every barrier is used once (parity 0)

  if (warp_idx == 4) {
    if (lane == 0) {
      prefetch_tmap(&tmap_qkv);
      prefetch_tmap(&tmap_kv);
      mbar_init(q_full, 1);
      for (int j = 0; j < C::MAX_KB; ++j) {
        mbar_init(&k_full[j], 1);
        mbar_init(&v_full[j], 1);
        mbar_init(&s_full[j], 1);
        mbar_init(&p_full[j], 4);
        mbar_init(&pv_done[j], 1);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp_idx == 4) {
    if (lane == 0) {
      // ---------------- TMA: Q, then the first KV_STAGES K/V blocks ----------------
      const int row_q = static_cast<int>(row_base) + q0;
      auto load_kv = [&](int j) {
        const int st = j % C::KV_STAGES;
        const int row_k = static_cast<int>(row_base) + j * C::BKV;
        mbar_arrive_expect_tx(&k_full[j], C::KV_BYTES);
        tma_load_2d(smem + C::OFF_K + st * C::KV_BYTES, &tmap_kv, &k_full[j], H + h * C::DH, row_k, kEvictLast);
        mbar_arrive_expect_tx(&v_full[j], C::KV_BYTES);
        tma_load_2d(smem + C::OFF_V + st * C::KV_BYTES, &tmap_kv, &v_full[j], 2 * H + h * C::DH, row_k, kEvictLast);
      };
      mbar_arrive_expect_tx(q_full, C::Q_BYTES);
      tma_load_2d(smem + C::OFF_Q, &tmap_qkv, q_full, h * C::DH, row_q, kEvictFirst);
      for (int j = 0; j < nkb && j < C::KV_STAGES; ++j) load_kv(j);
      // ---------------- MMA issue ----------------
      constexpr uint32_t idesc_qk = umma_idesc_f16(128, C::BKV, false, false);   // S = Q K^T   (both K-major)
      constexpr uint32_t idesc_pv = umma_idesc_f16(128, 64, false, true);        // O += P V    (V is N-major)
      const uint64_t q_desc = umma_desc_sw128(smem_u32(smem + C::OFF_Q));
      auto issue_qk = [&](int j) {
        mbar_wait(&k_full[j], 0);
        tc_fence_after();
        const uint64_t k_desc = umma_desc_sw128(smem_u32(smem + C::OFF_K + (j % C::KV_STAGES) * C::KV_BYTES));
        const uint32_t d = tmem_base + C::TM_S + static_cast<uint32_t>((j & 1) * C::BKV);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ss(d, q_desc + static_cast<uint64_t>(k * 2), k_desc + static_cast<uint64_t>(k * 2), idesc_qk,
                      k != 0 ? 1u : 0u);
        umma_commit(&s_full[j]);
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      if (nkb > 1) issue_qk(1);
      for (int j = 0; j < nkb; ++j) {
        mbar_wait(&p_full[j], 0);          // P_j in smem, S[j&1] drained, O rescaled
        mbar_wait(&v_full[j], 0);
        tc_fence_after();
        const uint32_t p_addr = smem_u32(smem + C::OFF_P + (j & 1) * C::P_BYTES);
        const uint32_t v_addr = smem_u32(smem + C::OFF_V + (j % C::KV_STAGES) * C::KV_BYTES);
#pragma unroll
        for (int kk = 0; kk < C::BKV / 16; ++kk) {
          // A = P: K-major 64-wide chunk, 32 B per K=16 step.  B = V: N-major (one 128 B swizzle row per key),
          // 16 keys = 2048 B per step.
          const uint64_t a_desc = umma_desc_sw128(p_addr) + static_cast<uint64_t>(kk * 2);
          const uint64_t b_desc = umma_desc_sw128(v_addr + kk * 2048);
          umma_f16_ss(tmem_base + C::TM_O, a_desc, b_desc, idesc_pv, (j | kk) != 0 ? 1u : 0u);
        }
        umma_commit(&pv_done[j]);
        if (j + 2 < nkb) issue_qk(j + 2);  // S[j&1] is free (p_full[j]); runs under the softmax of block j+1
        if (j + C::KV_STAGES < nkb) {      // recycle this K/V stage once Q K_j^T and P_j V_j have retired
          mbar_wait(&pv_done[j], 0);
          load_kv(j + C::KV_STAGES);
        }
      }
    }
  } else {
    // ======================= softmax warps: thread <-> query row =======================
    const int r = warp_idx * 32 + lane;                       // row in tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(warp_idx * 32) << 16;
    const float c = 1.4426950408889634f * 0.125f;             // log2(e) / sqrt(64)
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nkb; ++j) {
      mbar_wait(&s_full[j], 0);
      tc_fence_after();
      uint32_t s[2][32];
      const uint32_t s_addr = tmem_base + lane_addr + C::TM_S + static_cast<uint32_t>((j & 1) * C::BKV);
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) tmem_ld_32x32b_x32(s_addr + cc * 32, s[cc]);
      tmem_wait_ld();
      const int valid = min(C::BKV, len - j * C::BKV);         // >= 1
      if (valid < C::BKV) {                                    // only the last key block of a sequence is ragged
#pragma unroll
        for (int cc = 0; cc < 2; ++cc)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (cc * 32 + i >= valid) s[cc][i] = 0xff800000u;   // -inf: exp2 -> 0, never the max
      }
      // row max: 4 independent chains of 3-input max (one chain of dependent FMNMX would be latency-bound)
      float mx4[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t* sp = &s[q >> 1][(q & 1) * 16];
        float m = __uint_as_float(sp[0]);
#pragma unroll
        for (int i = 1; i < 15; i += 2) m = max3(m, __uint_as_float(sp[i]), __uint_as_float(sp[i + 1]));
        mx4[q] = fmaxf(m, __uint_as_float(sp[15]));
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      // Lazy rescaling: keep exponentiating against the stale reference m_run until some row's maximum has grown by
      // more than 2^8 relative to it (P <= 256 stays far inside fp16, O / l accumulate in fp32 and the common factor
      // cancels in O / l).  With a fresh maximum in almost every block, rescaling O in TMEM every time cost as many
      // FMULs as the exponentials themselves plus a TMEM load/store round trip on the critical path.
      const bool grow = (mx - m_run) * c > 8.0f;               // true on the first block (m_run = -inf)
      const bool any_grow = __any_sync(0xffffffffu, grow);
      const float m_new = grow ? mx : m_run;
      const float mc = m_new * c;
      uint8_t* p_row = smem + C::OFF_P + (j & 1) * C::P_BYTES + r * 128;     // P[j&1]: PV_{j-2} retired long ago
      float l4[4] = {0.f, 0.f, 0.f, 0.f};    // independent partial sums (ILP)
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {              // 8 columns -> one 16 B unit of the swizzled row
          float e[8];
#pragma unroll
          for (int t = 0; t < 8; ++t)
            e[t] = ex2_approx(fmaf(__uint_as_float(s[cc][u * 8 + t]), c, -mc));    // ex2(-inf) = 0 for masked keys
          l4[0] += e[0] + e[1];
          l4[1] += e[2] + e[3];
          l4[2] += e[4] + e[5];
          l4[3] += e[6] + e[7];
          uint4 pk;
          pk.x = pack_half2(e[0], e[1]);
          pk.y = pack_half2(e[2], e[3]);
          pk.z = pack_half2(e[4], e[5]);
          pk.w = pack_half2(e[6], e[7]);
          const int unit = cc * 4 + u;             // 16 B unit inside the 64-column row
          *reinterpret_cast<uint4*>(p_row + ((unit ^ (r & 7)) << 4)) = pk;
        }
      }
      const float l_blk = (l4[0] + l4[1]) + (l4[2] + l4[3]);
      const float alpha = ex2_approx((m_run - m_new) * c);     // 0 on the first block (m_run = -inf)
      if (j > 0) {
        mbar_wait(&pv_done[j - 1], 0);                         // O holds blocks 0..j-1
        tc_fence_after();
        if (any_grow) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t o[32];
            const uint32_t o_addr = tmem_base + lane_addr + C::TM_O + half * 32;
            tmem_ld_32x32b_x32(o_addr, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32b_x32(o_addr, o);
          }
          tmem_wait_st();
        }
      }
      l_run = l_run * alpha + l_blk;
      m_run = m_new;
      fence_proxy_async_smem();        // P (generic-proxy stores) -> visible to the tensor core's async proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[j]);
    }
    // ---------------- O / l -> ctx ----------------
    mbar_wait(&pv_done[nkb - 1], 0);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    const int q = q0 + r;
    __half* orow = ctx + (row_base + q) * H + h * C::DH;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(tmem_base + lane_addr + C::TM_O + half * 32, o);
      tmem_wait_ld();
      if (q < S) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint4 pk;
          pk.x = pack_half2(__uint_as_float(o[8 * u + 0]) * inv_l, __uint_as_float(o[8 * u + 1]) * inv_l);
          pk.y = pack_half2(__uint_as_float(o[8 * u + 2]) * inv_l, __uint_as_float(o[8 * u + 3]) * inv_l);
          pk.z = pack_half2(__uint_as_float(o[8 * u + 4]) * inv_l, __uint_as_float(o[8 * u + 5]) * inv_l);
          pk.w = pack_half2(__uint_as_float(o[8 * u + 6]) * inv_l, __uint_as_float(o[8 * u + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + half * 32 + u * 8) = pk;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

}  // namespace mv
