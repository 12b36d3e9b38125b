// Fused self-attention on tcgen05, persistent over (sequence, head, 128-query tile) work items:
//     ctx = softmax(Q K^T / sqrt(64) + key_mask) V          
// replacing HF BertSelfAttention's matmul / div / add-mask / softmax / matmul chain
// (transformers 4.1.0, entered from MemVul/custom_PTM_embedder.py:228) that round-trips a
// [B,12,S,S] fp32 score tensor through HBM.
//
// Input  : qkv fp16 [B*S, 3*H] row-major (Q | K | V column blocks, head h at columns h*64), read through two 2-D
//          TMA maps over the same matrix (boxes {64 cols, 128 rows} for Q and {64, 64} for K/V), SWIZZLE_128B.
// Output : ctx fp16 [B*S, H] row-major (head h at columns h*64), written by TMA from the retired P rows (full tiles).
// Masking: keys >= len[b] are excluded.  The reference adds -10000 to their scores, whose exp underflows to exactly
//          0 in fp32, so exclusion is bit-equivalent; fully padded key blocks and fully padded query tiles are skipped
//          (padded query rows are never consumed: BertPooler reads row 0 only, MemVul/model_memory.py:99).
//
// Warps 0-3 : softmax (one query row per thread; S read from TMEM, P written to smem as the fp16 A operand of the
//             second MMA, running max / sum in fp32, O rescaled in TMEM only when a row maximum grew by > 2^8)
// Warp 4    : lane 0 issues both tcgen05.mma streams        Warp 5 : lane 0 issues every TMA load
// The CTA treats its whole life as ONE stream of 64-key blocks: S = Q K^T is double-buffered in TMEM and P in smem,
// K/V blocks flow through a 4-stage ring, and all barriers are rings whose parity follows the global block / item
// counters.  So the next item's Q, K, V loads and first two Q K^T products overlap the current item's tail, and the
// per-CTA set-up (TMEM allocation, barrier init, first-load latency: ~3.5 us of the 12 us a non-persistent CTA spent
// on a 512-token tile in r01g) is paid once per CTA instead of once per tile.
// Footprint: 112 KB smem (Q 16 K, 4-stage K/V ring 64 K, P 2 x 16 K) and 256 TMEM columns (S 2 x 64, O 64): TWO CTAs
// per SM.  (Double-buffering Q at the price of a 3-stage K/V ring measured 5 % slower.)  Key length <= 512.
#pragma once
#include "ptx.cuh"

namespace mv {

struct AttnCfg {
  static constexpr int BQ = 128, BKV = 64, DH = 64, KV_STAGES = 4;
  static constexpr int Q_BYTES = 128 * 64 * 2;             // 16 KB: {64 x 128} fp16 box
  static constexpr int KV_BYTES = BKV * 64 * 2;            // 8 KB: {64 x 64} fp16 box
  static constexpr int P_BYTES = 128 * BKV * 2;            // 16 KB: 128 x 64 fp16 = one swizzled K-chunk
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + Q_BYTES;
  static constexpr int OFF_V = OFF_K + KV_STAGES * KV_BYTES;
  static constexpr int OFF_P = OFF_V + KV_STAGES * KV_BYTES;
  static constexpr int OFF_BAR = OFF_P + 2 * P_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 512;         // 115,200 B: two CTAs fit in 228 KB
  static constexpr int THREADS = 192;
  static constexpr int TMEM_COLS = 256;
  static constexpr int TM_S = 0, TM_O = 128;
};

__global__ void __launch_bounds__(AttnCfg::THREADS, 2)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_kv,
                         const __grid_constant__ CUtensorMap tmap_ctx,
                         const int* __restrict__ lens, const int* __restrict__ row_start, __half* __restrict__ ctx,
                         int B, int S, int H, int n_qt, int wait_mode,
                         unsigned long long* __restrict__ trace) {
  using C = AttnCfg;
  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = static_cast<int>(threadIdx.x & 31);
  const int n_heads = H / C::DH;
  const int n_items = B * n_heads * n_qt;                  // n_qt = query tiles per sequence that are computed
  // debug only (MEMVUL_ATT_TRACE): CTA 0 stamps clock64() at the phase boundaries of its first 128 key blocks
  const bool tracing = trace != nullptr && blockIdx.x == 0;
  auto stamp = [&](uint32_t g, int slot, int k) {
    if (tracing && g < 128u) trace[slot * 1024 + g * 8 + k] = static_cast<unsigned long long>(clock64());
  };
  const int idle_tma = wait_mode & 3, idle_mma = (wait_mode >> 2) & 3, idle_sm = (wait_mode >> 4) & 3;   // see mbar_wait_idle
  const int stagger = wait_mode >> 10;                     // cycles the SM's second CTA delays its soft-max stream (0 = off)
  const bool early = ((wait_mode >> 8) & 1) != 0;          // soft-max warps test pv_done / s_full early (non-blocking)
  const bool spec = ((wait_mode >> 9) & 1) != 0;           // exponentials start against the stale row maximum (see below)

  extern __shared__ __align__(1024) uint8_t smem[];        // SWIZZLE_128B tiles need 1024-byte alignment
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* q_full = bars;                           // [1]  Q tile of item `it` landed            (TMA tx)
  uint64_t* q_empty = bars + 1;                      // [1]  last Q K^T of the item retired        (tcgen05.commit)
  uint64_t* k_full = bars + 2;                       // [KV_STAGES]
  uint64_t* v_full = k_full + C::KV_STAGES;          // [KV_STAGES]
  uint64_t* kv_empty = v_full + C::KV_STAGES;        // [KV_STAGES]  P V of that block retired    (tcgen05.commit)
  uint64_t* s_full = kv_empty + C::KV_STAGES;        // [2]  Q K^T of block g in S[g&1]
  uint64_t* p_full = s_full + 2;                     // [2]  P_g in smem, S[g&1] drained, O rescaled (4 warp arrivals)
  uint64_t* pv_done = p_full + 2;                    // [2]  P_g V_g accumulated into O
  uint64_t* o_free = pv_done + 2;                    // [1]  O of the previous item read out       (4 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 1);

  if (warp_idx == 4) {
    if (lane == 0) {
      prefetch_tmap(&tmap_qkv);
      prefetch_tmap(&tmap_kv);
      prefetch_tmap(&tmap_ctx);
      mbar_init(q_full, 1);
      mbar_init(q_empty, 1);
      for (int i = 0; i < C::KV_STAGES; ++i) {
        mbar_init(&k_full[i], 1);
        mbar_init(&v_full[i], 1);
        mbar_init(&kv_empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s_full[i], 1);
        mbar_init(&p_full[i], 4);
        mbar_init(&pv_done[i], 1);
      }
      mbar_init(o_free, 4);
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

  // Every role walks the same item sequence; `g` counts key blocks and `it` non-skipped items over the CTA's life.
  // Token-major row of sequence b's first token: b*S in the padded layout, row_start[b] in the packed (var-len) one.
  auto decode = [&](int item, int& b, int& h, int& q0, int& len, int& row_base) {
    const int qt = item % n_qt;
    h = (item / n_qt) % n_heads;
    b = item / (n_qt * n_heads);
    q0 = qt * C::BQ;
    len = lens[b];
    row_base = row_start ? row_start[b] : b * S;
  };

  if (warp_idx == 5) {
    // ============================== TMA producer ==============================
    // The whole warp walks the schedule (uniform control flow keeps the address arithmetic on the uniform datapath);
    // one elected lane issues.  Under `if (lane == 0)` ptxas wrapped every TMA / MMA instruction in an
    // ELECT / R2UR / BRA.U.ANY waterfall (~16 SASS instructions per tcgen05.mma).
    {
      const bool issuer = elect_one();
      uint32_t g = 0, it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int b, h, q0, len, row_base;
        decode(item, b, h, q0, len, row_base);
        if (q0 >= len) continue;
        const int nkb = (len + C::BKV - 1) / C::BKV;
        mbar_wait_idle(q_empty, (it & 1u) ^ 1u, idle_tma);     // previous item's last Q K^T has retired
        if (issuer) {
          mbar_arrive_expect_tx(q_full, C::Q_BYTES);
          tma_load_2d(smem + C::OFF_Q, &tmap_qkv, q_full, h * C::DH, row_base + q0, kEvictFirst);
        }
        for (int j = 0; j < nkb; ++j, ++g) {
          const uint32_t st = g % C::KV_STAGES;
          mbar_wait_idle(&kv_empty[st], ((g / C::KV_STAGES) & 1u) ^ 1u, idle_tma);
          const int row_k = row_base + j * C::BKV;
          if (issuer) {
            mbar_arrive_expect_tx(&k_full[st], C::KV_BYTES);
            tma_load_2d(smem + C::OFF_K + st * C::KV_BYTES, &tmap_kv, &k_full[st], H + h * C::DH, row_k, kEvictLast);
            mbar_arrive_expect_tx(&v_full[st], C::KV_BYTES);
            tma_load_2d(smem + C::OFF_V + st * C::KV_BYTES, &tmap_kv, &v_full[st], 2 * H + h * C::DH, row_k, kEvictLast);
          }
        }
        ++it;
      }
    }
  } else if (warp_idx == 4) {
    // ============================== MMA issuer ==============================
    {
      const bool issuer = elect_one();
      const uint32_t smem_base = smem_u32(smem);
      constexpr uint32_t idesc_qk = umma_idesc_f16(128, C::BKV, false, false);   // S = Q K^T   (both K-major)
      constexpr uint32_t idesc_pv = umma_idesc_f16(128, 64, false, true);        // O += P V    (V is N-major)
      uint32_t g0 = 0, it = 0;                                  // g0 = global index of the item's first block
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int b, h, q0, len, row_base;
        decode(item, b, h, q0, len, row_base);
        if (q0 >= len) continue;
        const int nkb = (len + C::BKV - 1) / C::BKV;
        const uint64_t q_desc = umma_desc_sw128(smem_base + C::OFF_Q);
        auto issue_qk = [&](int j) {
          const uint32_t g = g0 + static_cast<uint32_t>(j);
          const uint32_t st = g % C::KV_STAGES;
          mbar_wait_idle(&k_full[st], (g / C::KV_STAGES) & 1u, idle_mma);
          tc_fence_after();
          const uint64_t k_desc = umma_desc_sw128(smem_base + C::OFF_K + st * C::KV_BYTES);
          const uint32_t d = tmem_base + C::TM_S + (g & 1u) * C::BKV;
          if (issuer) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ss(d, q_desc + static_cast<uint64_t>(k * 2), k_desc + static_cast<uint64_t>(k * 2), idesc_qk,
                          k != 0 ? 1u : 0u);
            umma_commit(&s_full[g & 1u]);
            if (j == nkb - 1) umma_commit(q_empty);            // Q may be overwritten once this product retires
          }
        };
        mbar_wait_idle(q_full, it & 1u, idle_mma);
        // S[g&1] of the first two blocks is free: the previous item's last two p_full phases were waited on below.
        issue_qk(0);
        if (nkb > 1) issue_qk(1);
        for (int j = 0; j < nkb; ++j) {
          const uint32_t g = g0 + static_cast<uint32_t>(j);
          const uint32_t st = g % C::KV_STAGES;
          mbar_wait_idle(&p_full[g & 1u], (g >> 1) & 1u, idle_mma);   // P_g in smem, S[g&1] drained, O rescaled
          if (issuer) stamp(g, 1, 0);
          if (j == 0) mbar_wait_idle(o_free, (it & 1u) ^ 1u, idle_mma);   // previous item's O has been read out
          mbar_wait_idle(&v_full[st], (g / C::KV_STAGES) & 1u, idle_mma);
          tc_fence_after();
          if (issuer) stamp(g, 1, 1);
          const uint32_t p_addr = smem_base + C::OFF_P + (g & 1u) * C::P_BYTES;
          const uint32_t v_addr = smem_base + C::OFF_V + st * C::KV_BYTES;
          if (issuer) {
#pragma unroll
            for (int kk = 0; kk < C::BKV / 16; ++kk) {
              // A = P: K-major 64-wide chunk, 32 B per K=16 step.  B = V: N-major (one 128 B swizzle row per key),
              // 16 keys = 2048 B per step.
              const uint64_t a_desc = umma_desc_sw128(p_addr) + static_cast<uint64_t>(kk * 2);
              const uint64_t b_desc = umma_desc_sw128(v_addr + kk * 2048);
              umma_f16_ss(tmem_base + C::TM_O, a_desc, b_desc, idesc_pv, (j | kk) != 0 ? 1u : 0u);
            }
            umma_commit(&pv_done[g & 1u]);
            umma_commit(&kv_empty[st]);                        // K/V stage reusable once Q K_g^T and P_g V_g retired
          }
          if (issuer) stamp(g, 1, 2);
          if (j + 2 < nkb) issue_qk(j + 2);                    // S[g&1] is free; runs under the softmax of block g+1
          if (issuer) stamp(g, 1, 3);
        }
        g0 += static_cast<uint32_t>(nkb);
        ++it;
      }
    }
  } else {
    // ======================= softmax warps: thread <-> query row =======================
    const int r = warp_idx * 32 + lane;                       // row in tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(warp_idx * 32) << 16;
    const float c = 1.4426950408889634f * 0.125f;             // log2(e) / sqrt(64)
    uint32_t g = 0;
    bool store_pending = false;                                // this warp has a ctx TMA store reading its P rows
    // lens[] of the NEXT item is loaded one item ahead: the dependent global load (~650 cycles in the r01p trace) sat
    // on the serial path between two items
    if (stagger > 0) {
      // Which of the SM's two CTAs am I?  The hardware warp slot: the first CTA's six warps occupy slots 0-5.  (A wrong
      // guess only loses the interleaving, never correctness.)
      uint32_t wslot;
      asm volatile("mov.u32 %0, %%warpid;" : "=r"(wslot));
      if (wslot >= static_cast<uint32_t>(C::THREADS / 32)) {
        const long long t0 = clock64();
        while (clock64() - t0 < stagger) {}
      }
    }
    int len_next = 0, rb_next = 0;
    if (static_cast<int>(blockIdx.x) < n_items) {
      const int b0 = blockIdx.x / (n_qt * n_heads);
      len_next = lens[b0];
      rb_next = row_start ? row_start[b0] : b0 * S;
    }
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int qt = item % n_qt;
      const int h = (item / n_qt) % n_heads;
      const int q0 = qt * C::BQ;
      const int len = len_next;
      const size_t row_base = static_cast<size_t>(rb_next);
      if (item + static_cast<int>(gridDim.x) < n_items) {
        const int bn = (item + gridDim.x) / (n_qt * n_heads);
        len_next = lens[bn];
        rb_next = row_start ? row_start[bn] : bn * S;
      }
      // rows of this sequence that exist in the token-major matrix: S (padded layout, padded rows are written too so
      // that they stay finite) or len (packed layout: the next row already belongs to the next sequence)
      const int row_limit = row_start ? len : S;
      if (warp_idx == 0 && lane == 0 && len >= 0) stamp(g, 1, 7);
      if (q0 >= len) {
        // fully padded query tile: deterministic zeros, no tensor work (the packed layout has no such rows)
        const int rows = row_start ? 0 : min(C::BQ, S - q0);
        for (int i = threadIdx.x; i < rows * 8; i += 128) {
          const int rr = i >> 3, u = i & 7;
          *reinterpret_cast<uint4*>(ctx + (row_base + q0 + rr) * H + h * C::DH + u * 8) = make_uint4(0, 0, 0, 0);
        }
        continue;
      }
      const int nkb = (len + C::BKV - 1) / C::BKV;
      float m_run = -INFINITY, l_run = 0.f;
      if (store_pending) {                                     // issued >= one item epilogue + prologue ago: never stalls
        if (lane == 0) bulk_wait_read_all();
        __syncwarp();
        store_pending = false;
      }
      uint32_t s_ok = 0;                                       // early (non-blocking) test of this block's s_full, see below
      for (int j = 0; j < nkb; ++j, ++g) {
        if (warp_idx == 0 && lane == 0) stamp(g, 0, 0);
        if (!__all_sync(0xffffffffu, s_ok != 0u)) mbar_wait_idle(&s_full[g & 1u], (g >> 1) & 1u, idle_sm);
        tc_fence_after();
        if (warp_idx == 0 && lane == 0) stamp(g, 0, 1);
        uint32_t s[2][32];
        const uint32_t s_addr = tmem_base + lane_addr + C::TM_S + (g & 1u) * C::BKV;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) tmem_ld_32x32b_x32(s_addr + cc * 32, s[cc]);
        tmem_wait_ld();
        if (warp_idx == 0 && lane == 0) stamp(g, 0, 2);
        const int valid = min(C::BKV, len - j * C::BKV);       // >= 1
        if (valid < C::BKV) {                                  // only the last key block of a sequence is ragged
#pragma unroll
          for (int cc = 0; cc < 2; ++cc)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (cc * 32 + i >= valid) s[cc][i] = 0xff800000u;   // -inf: exp2 -> 0, never the max
        }
        // Early phase tests: P V of block g-1 was issued a whole block ago and Q K^T of block g+1 right after it, so both
        // barriers have almost always completed when they are needed -- but a blocking try_wait still costs its ~90-180
        // cycle round trip on this warp's serial chain.  test_wait is scoreboarded: issued before the last exponentials
        // its latency hides under them, and the blocking wait is only the fallback.
        uint32_t pv_ok = 0;
        uint8_t* p_row = smem + C::OFF_P + (g & 1u) * C::P_BYTES + r * 128;   // P[g&1]: P V of block g-2 retired long ago
        // P = 2^(s c - mc) -> swizzled fp16 row of the P buffer; returns the row's partial sum
        auto exp_pass = [&](float mc) -> float {
          float l4[4] = {0.f, 0.f, 0.f, 0.f};                 // independent partial sums (ILP)
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {                      // 8 columns -> one 16 B unit of the swizzled row
              if (cc == 1 && u == 3 && early && j > 0) pv_ok = mbar_test_wait(&pv_done[(g - 1) & 1u], ((g - 1) >> 1) & 1u);
              float e[8];
#pragma unroll
              for (int t = 0; t < 8; ++t)
                e[t] = ex2_approx(fmaf(__uint_as_float(s[cc][u * 8 + t]), c, -mc));   // ex2(-inf) = 0 for masked keys
              l4[0] += e[0] + e[1];
              l4[1] += e[2] + e[3];
              l4[2] += e[4] + e[5];
              l4[3] += e[6] + e[7];
              uint4 pk;
              pk.x = pack_half2(e[0], e[1]);
              pk.y = pack_half2(e[2], e[3]);
              pk.z = pack_half2(e[4], e[5]);
              pk.w = pack_half2(e[6], e[7]);
              const int unit = cc * 4 + u;                     // 16 B unit inside the 64-column row
              *reinterpret_cast<uint4*>(p_row + ((unit ^ (r & 7)) << 4)) = pk;
            }
          }
          return (l4[0] + l4[1]) + (l4[2] + l4[3]);
        };
        // row max: 4 independent chains of 3-input max (one chain of dependent FMNMX would be latency-bound)
        auto row_max = [&]() -> float {
          float mx4[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t* sp = &s[q >> 1][(q & 1) * 16];
            float m = __uint_as_float(sp[0]);
#pragma unroll
            for (int i = 1; i < 15; i += 2) m = max3(m, __uint_as_float(sp[i]), __uint_as_float(sp[i + 1]));
            mx4[q] = fmaxf(m, __uint_as_float(sp[15]));
          }
          return fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        };
        // Lazy rescaling: keep exponentiating against the stale reference m_run until some row's maximum has grown by
        // more than 2^8 relative to it (P <= 256 stays far inside fp16, O / l accumulate in fp32 and the common factor
        // cancels in O / l).  With a fresh maximum in almost every block, rescaling O in TMEM every time cost as many
        // FMULs as the exponentials themselves plus a TMEM load/store round trip on the critical path.
        // Speculation (blocks after an item's first): since the stale reference is almost always kept, the exponentials
        // start against it IMMEDIATELY and the row maximum (170 cycles of dependent FMNMX on this warp's serial chain) is
        // computed in their shadow; only a row whose maximum did jump by > 2^8 repeats its 64 exponentials.
        float mx, l_blk;
        bool grow;
        if (spec && j > 0) {
          l_blk = exp_pass(m_run * c);
          mx = row_max();
          grow = (mx - m_run) * c > 8.0f;
          if (grow) l_blk = exp_pass(mx * c);
        } else {
          mx = row_max();
          grow = (mx - m_run) * c > 8.0f;                      // true on the first block (m_run = -inf)
          l_blk = exp_pass((grow ? mx : m_run) * c);
        }
        const bool any_grow = __any_sync(0xffffffffu, grow);
        const float m_new = grow ? mx : m_run;
        if (warp_idx == 0 && lane == 0) stamp(g, 0, 3);
        if (warp_idx == 0 && lane == 0) stamp(g, 0, 4);
        const float alpha = ex2_approx((m_run - m_new) * c);   // 0 on the first block (m_run = -inf), else 1 unless grown
        s_ok = (early && j + 1 < nkb) ? mbar_test_wait(&s_full[(g + 1) & 1u], ((g + 1) >> 1) & 1u) : 0u;   // consumed at the next loop top
        if (j > 0) {
          if (!__all_sync(0xffffffffu, pv_ok != 0u))
            mbar_wait_idle(&pv_done[(g - 1) & 1u], ((g - 1) >> 1) & 1u, idle_sm);   // O holds blocks 0..j-1 of this item
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
        if (warp_idx == 0 && lane == 0) stamp(g, 0, 5);
        l_run = l_run * alpha + l_blk;
        m_run = m_new;
        fence_proxy_async_smem();        // P (generic-proxy stores) -> visible to the tensor core's async proxy
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g & 1u]);
        if (warp_idx == 0 && lane == 0) stamp(g, 0, 6);
      }
      // ---------------- O / l -> ctx ----------------
      mbar_wait_idle(&pv_done[(g - 1) & 1u], ((g - 1) >> 1) & 1u, idle_sm);
      tc_fence_after();
      if (warp_idx == 0 && lane == 0) stamp(g - 1, 0, 7);
      const float inv_l = 1.0f / l_run;
      const int q = q0 + r;
      __half* orow = ctx + (row_base + q) * H + h * C::DH;
      uint32_t o[2][32];
      tmem_ld_32x32b_x32(tmem_base + lane_addr + C::TM_O, o[0]);
      tmem_ld_32x32b_x32(tmem_base + lane_addr + C::TM_O + 32, o[1]);
      tmem_wait_ld();
      if (warp_idx == 0 && lane == 0) stamp(g - 1, 1, 4);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);                      // the next item's first P V may overwrite O now
      if (warp_idx == 0 && lane == 0) stamp(g - 1, 1, 5);
      if (q0 + C::BQ <= row_limit) {
        // Full tile: stage the warp's 32 rows in its quarter of the P buffer the item's last block used (its P V has
        // retired) and let the TMA engine write them.  The per-lane version -- every lane storing 8 x 16 B into its own
        // row, 1,536 B apart -- kept the warp ~1,900 cycles in the LSU at every item boundary (r01p trace).
        uint8_t* stg = smem + C::OFF_P + ((g - 1) & 1u) * C::P_BYTES + warp_idx * 4096;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 pk;
            pk.x = pack_half2(__uint_as_float(o[half][8 * u + 0]) * inv_l, __uint_as_float(o[half][8 * u + 1]) * inv_l);
            pk.y = pack_half2(__uint_as_float(o[half][8 * u + 2]) * inv_l, __uint_as_float(o[half][8 * u + 3]) * inv_l);
            pk.z = pack_half2(__uint_as_float(o[half][8 * u + 4]) * inv_l, __uint_as_float(o[half][8 * u + 5]) * inv_l);
            pk.w = pack_half2(__uint_as_float(o[half][8 * u + 6]) * inv_l, __uint_as_float(o[half][8 * u + 7]) * inv_l);
            const int unit = half * 4 + u;
            *reinterpret_cast<uint4*>(stg + lane * 128 + ((unit ^ (lane & 7)) << 4)) = pk;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmap_ctx, stg, h * C::DH, static_cast<int>(row_base) + q0 + warp_idx * 32);
          bulk_commit_group();
        }
        store_pending = true;
      } else if (q < row_limit) {                              // ragged last tile: later rows belong to the next sequence
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 pk;
            pk.x = pack_half2(__uint_as_float(o[half][8 * u + 0]) * inv_l, __uint_as_float(o[half][8 * u + 1]) * inv_l);
            pk.y = pack_half2(__uint_as_float(o[half][8 * u + 2]) * inv_l, __uint_as_float(o[half][8 * u + 3]) * inv_l);
            pk.z = pack_half2(__uint_as_float(o[half][8 * u + 4]) * inv_l, __uint_as_float(o[half][8 * u + 5]) * inv_l);
            pk.w = pack_half2(__uint_as_float(o[half][8 * u + 6]) * inv_l, __uint_as_float(o[half][8 * u + 7]) * inv_l);
            *reinterpret_cast<uint4*>(orow + half * 32 + u * 8) = pk;
          }
        }
      }
      if (warp_idx == 0 && lane == 0) stamp(g - 1, 1, 6);
    }
    if (store_pending && lane == 0) bulk_wait_read_all();      // the staging rows must outlive the last store's read
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

}  // namespace mv
