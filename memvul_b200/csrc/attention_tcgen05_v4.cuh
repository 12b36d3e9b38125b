// Fused self-attention on tcgen05, TWO QUERY TILES PER CTA IN EXPLICIT PING-PONG:
//     ctx = softmax(Q K^T / sqrt(64) + key_mask) V          (HF BertSelfAttention as entered from
//     MemVul/custom_PTM_embedder.py:224-228).  Same data contract and numerics as attention_tcgen05.cuh / _v3.cuh.
//
// Why: r02q (profiles/r02q_attention_v3.md).  Independent co-resident CTAs fall into lock-step: they slow each other down
// exactly while they are all exponentiating (one MUFU pipe per scheduler), finish that phase together and then wait
// together -- the per-CTA block period of the three-stream kernel is ~3 x (exp phase) + (rest), the MUFU pipe idles ~45 %.
// Streams of DIFFERENT CTAs cannot be kept out of phase; two query tiles inside ONE CTA can:
//   * a work item is (sequence, head, PAIR of 128-query tiles); contexts A (warps 0-3) and B (warps 4-7) run the soft-max of
//     tile 2t and 2t+1 against the SAME K / V ring (half the K / V shared-memory and L2 traffic per query);
//   * soft-max warp i of A and of B sit on the same scheduler and pass a TOKEN (two mbarriers per pair of warps): a warp
//     exponentiates only while it holds the token, so one of the two is always in its MUFU phase while the other does its
//     waits, TMEM load, row maximum, fences and hand-over.  Strict alternation A, B, A, B (both tiles see the same keys,
//     hence the same number of blocks).  A tile pair with only one live tile (odd tile count, short sequence, the
//     CLS-only last layer) runs context A without the token.
//   * per context: S single-buffered in TMEM and released as soon as it is in registers (`s_free`), P single-buffered,
//     O; K and V rings (3 stages) are freed by the LAST product that reads them (the commit after context B's).
// Warps: 0-3 soft-max A, 4-7 soft-max B, 8 MMA issuer (both contexts, fixed order per key block:
// QK_A(g+1), PV_A(g), QK_B(g+1), PV_B(g)), 9 TMA producer, 10-11 idle (complete the third warpgroup for setmaxnreg).
// Two CTAs per SM: 24 warps = 6 per scheduler, launched at 80 registers; soft-max warpgroups rise to 104, the third drops
// to 32 (per scheduler and CTA 104 + 104 + 32 = 3 x 80).  112.5 KB of shared memory, 256 TMEM columns.  Key length <= 512.
#pragma once
#include "ptx.cuh"

namespace mv {

struct Attn4Cfg {
  static constexpr int BQ = 128, BKV = 64, DH = 64, KV_STAGES = 3;
  static constexpr int Q_BYTES = 128 * 64 * 2;             // 16 KB: {64 x 128} fp16 box
  static constexpr int KV_BYTES = BKV * 64 * 2;            // 8 KB: {64 x 64} fp16 box
  static constexpr int P_BYTES = 128 * BKV * 2;            // 16 KB: 128 x 64 fp16 = one swizzled K-chunk
  static constexpr int OFF_Q = 0;                          // [2 contexts]
  static constexpr int OFF_K = OFF_Q + 2 * Q_BYTES;
  static constexpr int OFF_V = OFF_K + KV_STAGES * KV_BYTES;
  static constexpr int OFF_P = OFF_V + KV_STAGES * KV_BYTES;   // [2 contexts]
  static constexpr int OFF_BAR = OFF_P + 2 * P_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 512;         // 115,200 B
  static_assert(2 * (SMEM_BYTES + 1024) <= 233472, "attention v4 must fit two CTAs per SM");
  static constexpr int THREADS = 384;
  static constexpr int CTAS_PER_SM = 2;
  static constexpr int REGS_LAUNCH = 80, REGS_SOFTMAX = 104, REGS_AUX = 32;
  static_assert(2 * REGS_SOFTMAX + REGS_AUX == 3 * REGS_LAUNCH, "the register pool of a CTA must balance");
  static constexpr int TMEM_COLS = 256;                    // context X: S at 128 X, O at 128 X + 64
};

__global__ void __launch_bounds__(Attn4Cfg::THREADS, Attn4Cfg::CTAS_PER_SM)
attention_tcgen05_v4_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_kv,
                            const __grid_constant__ CUtensorMap tmap_ctx,
                            const int* __restrict__ lens, const int* __restrict__ row_start, __half* __restrict__ ctx,
                            int B, int S, int H, int n_qt, int wait_mode, int use_token) {
  using C = Attn4Cfg;
  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = static_cast<int>(threadIdx.x & 31);
  const int n_heads = H / C::DH;
  const int n_qp = (n_qt + 1) >> 1;                        // tile pairs per (sequence, head)
  const int n_items = B * n_heads * n_qp;
  const int idle_tma = wait_mode & 3, idle_mma = (wait_mode >> 2) & 3, idle_sm = (wait_mode >> 4) & 3;   // see mbar_wait_idle

  extern __shared__ __align__(1024) uint8_t smem[];        // SWIZZLE_128B tiles need 1024-byte alignment
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* q_full = bars;                           // [1]  the item's Q tile(s) landed                (TMA tx)
  uint64_t* q_empty = bars + 1;                      // [1]  the item's last Q K^T retired              (tcgen05.commit)
  uint64_t* k_full = bars + 2;                       // [KV_STAGES]
  uint64_t* v_full = k_full + C::KV_STAGES;          // [KV_STAGES]
  uint64_t* k_empty = v_full + C::KV_STAGES;         // [KV_STAGES]  last Q K^T on that stage retired  (tcgen05.commit)
  uint64_t* v_empty = k_empty + C::KV_STAGES;        // [KV_STAGES]  last P V on that stage retired    (tcgen05.commit)
  uint64_t* s_full = v_empty + C::KV_STAGES;         // [2 ctx]  Q K^T of the context's block in S
  uint64_t* s_free = s_full + 2;                     // [2 ctx]  S is in the soft-max registers        (4 warp arrivals)
  uint64_t* p_full = s_free + 2;                     // [2 ctx]  P in smem, O rescaled                 (4 warp arrivals)
  uint64_t* pv_done = p_full + 2;                    // [2 ctx]  P V accumulated into O
  uint64_t* o_free = pv_done + 2;                    // [2 ctx]  O of the context's previous item read out (4 warp arrivals)
  uint64_t* tok_a = o_free + 2;                      // [4 warps]  B's warp i finished its exponentials -> A's warp i may start
  uint64_t* tok_b = tok_a + 4;                       // [4 warps]  A's warp i finished                   -> B's warp i may start
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tok_b + 4);

  if (warp_idx == 8) {
    if (lane == 0) {
      prefetch_tmap(&tmap_qkv);
      prefetch_tmap(&tmap_kv);
      prefetch_tmap(&tmap_ctx);
      mbar_init(q_full, 1);
      mbar_init(q_empty, 1);
      for (int i = 0; i < C::KV_STAGES; ++i) {
        mbar_init(&k_full[i], 1);
        mbar_init(&v_full[i], 1);
        mbar_init(&k_empty[i], 1);
        mbar_init(&v_empty[i], 1);
      }
      for (int x = 0; x < 2; ++x) {
        mbar_init(&s_full[x], 1);
        mbar_init(&s_free[x], 4);
        mbar_init(&p_full[x], 4);
        mbar_init(&pv_done[x], 1);
        mbar_init(&o_free[x], 4);
      }
      for (int i = 0; i < 4; ++i) {
        mbar_init(&tok_a[i], 1);
        mbar_init(&tok_b[i], 1);
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

  // Every role walks the same item sequence.  An item is live when its first tile has rows (q0a < len); context B is live
  // when the pair has a second tile with rows.  `g` counts key blocks of live items (ring parity), `it` live items; context
  // B keeps its own block / item counters for its single-buffered barriers.
  auto decode = [&](int item, int& b, int& h, int& q0a, int& len, int& row_base, bool& b_act) {
    const int tp = item % n_qp;
    h = (item / n_qp) % n_heads;
    b = item / (n_qp * n_heads);
    q0a = tp * 2 * C::BQ;
    len = lens[b];
    row_base = row_start ? row_start[b] : b * S;
    b_act = (2 * tp + 1 < n_qt) && (q0a + C::BQ < len);
  };

  if (warp_idx >= 10) {
    setmaxnreg_dec<C::REGS_AUX>();                             // idle: present only so that warpgroup 2 is complete
  } else if (warp_idx == 9) {
    // ============================== TMA producer (warp-uniform walk, one elected issuing lane) ==============================
    setmaxnreg_dec<C::REGS_AUX>();
    const bool issuer = elect_one();
    uint32_t g = 0, it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int b, h, q0a, len, row_base;
      bool b_act;
      decode(item, b, h, q0a, len, row_base, b_act);
      if (q0a >= len) continue;
      const int nkb = (len + C::BKV - 1) / C::BKV;
      mbar_wait_idle(q_empty, (it & 1u) ^ 1u, idle_tma);       // previous item's last Q K^T has retired
      if (issuer) {
        mbar_arrive_expect_tx(q_full, b_act ? 2 * C::Q_BYTES : C::Q_BYTES);
        tma_load_2d(smem + C::OFF_Q, &tmap_qkv, q_full, h * C::DH, row_base + q0a, kEvictFirst);
        if (b_act) tma_load_2d(smem + C::OFF_Q + C::Q_BYTES, &tmap_qkv, q_full, h * C::DH, row_base + q0a + C::BQ, kEvictFirst);
      }
      for (int j = 0; j < nkb; ++j, ++g) {
        const uint32_t st = g % C::KV_STAGES;
        const uint32_t par = ((g / C::KV_STAGES) & 1u) ^ 1u;
        const int row_k = row_base + j * C::BKV;
        mbar_wait_idle(&k_empty[st], par, idle_tma);
        if (issuer) {
          mbar_arrive_expect_tx(&k_full[st], C::KV_BYTES);
          tma_load_2d(smem + C::OFF_K + st * C::KV_BYTES, &tmap_kv, &k_full[st], H + h * C::DH, row_k, kEvictLast);
        }
        mbar_wait_idle(&v_empty[st], par, idle_tma);
        if (issuer) {
          mbar_arrive_expect_tx(&v_full[st], C::KV_BYTES);
          tma_load_2d(smem + C::OFF_V + st * C::KV_BYTES, &tmap_kv, &v_full[st], 2 * H + h * C::DH, row_k, kEvictLast);
        }
      }
      ++it;
    }
  } else if (warp_idx == 8) {
    // ============================== MMA issuer (both contexts) ==============================
    setmaxnreg_dec<C::REGS_AUX>();
    const bool issuer = elect_one();
    const uint32_t smem_base = smem_u32(smem);
    constexpr uint32_t idesc_qk = umma_idesc_f16(128, C::BKV, false, false);   // S = Q K^T   (both K-major)
    constexpr uint32_t idesc_pv = umma_idesc_f16(128, 64, false, true);        // O += P V    (V is N-major)
    uint32_t g0 = 0, it = 0;                                    // ring block counter / live items
    uint32_t cb0 = 0, itb = 0;                                  // context B's block / item counters
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int b, h, q0a, len, row_base;
      bool b_act;
      decode(item, b, h, q0a, len, row_base, b_act);
      if (q0a >= len) continue;
      const int nkb = (len + C::BKV - 1) / C::BKV;
      // S_x = Q_x K_j^T.  `last` = this is the last product that reads K_j (then the stage, and after the item's last block
      // the Q tiles, are released by the commit, which covers every MMA issued so far).
      auto issue_qk = [&](int x, int j, bool last) {
        const uint32_t g = g0 + static_cast<uint32_t>(j);
        const uint32_t st = g % C::KV_STAGES;
        const uint32_t c = (x == 0 ? g0 : cb0) + static_cast<uint32_t>(j);    // the context's own block counter
        if (c > 0) mbar_wait_idle(&s_free[x], (c - 1) & 1u, idle_mma);        // its previous S is in registers
        mbar_wait_idle(&k_full[st], (g / C::KV_STAGES) & 1u, idle_mma);
        tc_fence_after();
        const uint64_t q_desc = umma_desc_sw128(smem_base + C::OFF_Q + x * C::Q_BYTES);
        const uint64_t k_desc = umma_desc_sw128(smem_base + C::OFF_K + st * C::KV_BYTES);
        const uint32_t d = tmem_base + static_cast<uint32_t>(x * 128);
        if (issuer) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(d, q_desc + static_cast<uint64_t>(k * 2), k_desc + static_cast<uint64_t>(k * 2), idesc_qk,
                        k != 0 ? 1u : 0u);
          umma_commit(&s_full[x]);
          if (last) {
            umma_commit(&k_empty[st]);
            if (j == nkb - 1) umma_commit(q_empty);
          }
        }
      };
      // O_x += P_x V_j.  `last` as above for the V stage.
      auto issue_pv = [&](int x, int j, bool last) {
        const uint32_t g = g0 + static_cast<uint32_t>(j);
        const uint32_t st = g % C::KV_STAGES;
        const uint32_t c = (x == 0 ? g0 : cb0) + static_cast<uint32_t>(j);
        mbar_wait_idle(&p_full[x], c & 1u, idle_mma);           // P in smem, O rescaled
        if (j == 0) mbar_wait_idle(&o_free[x], ((x == 0 ? it : itb) & 1u) ^ 1u, idle_mma);   // previous O read out
        mbar_wait_idle(&v_full[st], (g / C::KV_STAGES) & 1u, idle_mma);
        tc_fence_after();
        const uint64_t p_desc = umma_desc_sw128(smem_base + C::OFF_P + x * C::P_BYTES);
        const uint32_t v_addr = smem_base + C::OFF_V + st * C::KV_BYTES;
        const uint32_t d = tmem_base + static_cast<uint32_t>(x * 128 + 64);
        if (issuer) {
#pragma unroll
          for (int kk = 0; kk < C::BKV / 16; ++kk) {
            // A = P: K-major 64-wide chunk, 32 B per K=16 step.  B = V: N-major (one 128 B swizzle row per key),
            // 16 keys = 2048 B per step.
            const uint64_t b_desc = umma_desc_sw128(v_addr + kk * 2048);
            umma_f16_ss(d, p_desc + static_cast<uint64_t>(kk * 2), b_desc, idesc_pv, (j | kk) != 0 ? 1u : 0u);
          }
          umma_commit(&pv_done[x]);
          if (last) umma_commit(&v_empty[st]);
        }
      };
      mbar_wait_idle(q_full, it & 1u, idle_mma);
      issue_qk(0, 0, !b_act);
      if (b_act) issue_qk(1, 0, true);
      for (int j = 0; j < nkb; ++j) {
        if (j + 1 < nkb) issue_qk(0, j + 1, !b_act);            // as soon as S_A is in registers: runs under A's soft-max
        issue_pv(0, j, !b_act);
        if (b_act) {
          if (j + 1 < nkb) issue_qk(1, j + 1, true);
          issue_pv(1, j, true);
        }
      }
      g0 += static_cast<uint32_t>(nkb);
      ++it;
      if (b_act) {
        cb0 += static_cast<uint32_t>(nkb);
        ++itb;
      }
    }
  } else {
    // ======================= soft-max warps: context x = warp / 4, thread <-> query row =======================
    setmaxnreg_inc<C::REGS_SOFTMAX>();
    const int x = warp_idx >> 2, wi = warp_idx & 3;
    const int r = wi * 32 + lane;                             // row in tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(wi * 32) << 16;
    const uint32_t tm_s = tmem_base + lane_addr + static_cast<uint32_t>(x * 128);
    const uint32_t tm_o = tm_s + 64;
    const float c = 1.4426950408889634f * 0.125f;             // log2(e) / sqrt(64)
    uint8_t* const p_row = smem + C::OFF_P + x * C::P_BYTES + r * 128;   // this thread's swizzled P row
    uint64_t* const my_tok = x == 0 ? &tok_a[wi] : &tok_b[wi];          // completes when the partner warp passes the token
    uint64_t* const other_tok = x == 0 ? &tok_b[wi] : &tok_a[wi];
    uint32_t g = 0;                                            // this context's block counter
    uint32_t tk = 0;                                           // token rounds (blocks of paired items)
    bool store_pending = false;                                // this warp has a ctx TMA store reading its P rows
    int len_next = 0, rb_next = 0;                             // lens[] / row_start[] are loaded one item ahead
    if (static_cast<int>(blockIdx.x) < n_items) {
      const int b0 = blockIdx.x / (n_qp * n_heads);
      len_next = lens[b0];
      rb_next = row_start ? row_start[b0] : b0 * S;
    }
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int tp = item % n_qp;
      const int h = (item / n_qp) % n_heads;
      const int len = len_next;
      const size_t row_base = static_cast<size_t>(rb_next);
      if (item + static_cast<int>(gridDim.x) < n_items) {
        const int bn = (item + gridDim.x) / (n_qp * n_heads);
        len_next = lens[bn];
        rb_next = row_start ? row_start[bn] : bn * S;
      }
      const int tile = 2 * tp + x;
      if (tile >= n_qt) continue;                              // the pair has no second tile
      const int q0 = tile * C::BQ;
      const bool paired = (2 * tp + 1 < n_qt) && (tp * 2 * C::BQ + C::BQ < len);   // both contexts live: token protocol
      const int row_limit = row_start ? len : S;               // rows of this sequence that exist in the token-major matrix
      if (q0 >= len) {
        // fully padded query tile: deterministic zeros, no tensor work (the packed layout has no such rows)
        const int rows = row_start ? 0 : min(C::BQ, S - q0);
        for (int i = threadIdx.x & 127; i < rows * 8; i += 128) {
          const int rr = i >> 3, u = i & 7;
          *reinterpret_cast<uint4*>(ctx + (row_base + q0 + rr) * H + h * C::DH + u * 8) = make_uint4(0, 0, 0, 0);
        }
        continue;
      }
      const int nkb = (len + C::BKV - 1) / C::BKV;
      float m_run = -INFINITY, l_run = 0.f;
      uint32_t s_ok = 0;                                       // early (non-blocking) test of the next block's s_full
      for (int j = 0; j < nkb; ++j, ++g) {
        if (!__all_sync(0xffffffffu, s_ok != 0u)) mbar_wait_idle(&s_full[x], g & 1u, idle_sm);
        tc_fence_after();
        uint32_t s[2][32];
        tmem_ld_32x32b_x32(tm_s, s[0]);
        tmem_ld_32x32b_x32(tm_s + 32, s[1]);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[x]);                // Q K^T of this context's next block may overwrite S now
        uint32_t pv_ok = j > 0 ? mbar_test_wait(&pv_done[x], (g - 1) & 1u) : 1u;   // scoreboarded: hides under the row maximum
        const int valid = min(C::BKV, len - j * C::BKV);       // >= 1
        if (valid < C::BKV) {                                  // only the last key block of a sequence is ragged
#pragma unroll
          for (int cc = 0; cc < 2; ++cc)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (cc * 32 + i >= valid) s[cc][i] = 0xff800000u;   // -inf: exp2 -> 0, never the max
        }
        // row max: 4 independent chains of 3-input max
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
        // Lazy rescaling (as in the first kernel): keep the stale reference m_run until a row maximum has grown by > 2^8.
        const bool grow = (mx - m_run) * c > 8.0f;             // true on the first block (m_run = -inf)
        const float m_new = grow ? mx : m_run;
        const float mc = m_new * c;
        const bool any_grow = __any_sync(0xffffffffu, grow);
        const float alpha = ex2_approx((m_run - m_new) * c);   // 0 on the first block, else 1 unless grown
        // P's buffer must be free: the previous block's P V retired (long ago), or, on an item's first block, the previous
        // item's ctx store has read this warp's staging rows.
        if (j == 0) {
          if (store_pending) {
            if (lane == 0) bulk_wait_read_all();
            __syncwarp();
            store_pending = false;
          }
        } else {
          if (!__all_sync(0xffffffffu, pv_ok != 0u)) mbar_wait_idle(&pv_done[x], (g - 1) & 1u, idle_sm);
          tc_fence_after();
        }
        // ---- the MUFU phase: only while holding the pair's token ----
        if (paired && use_token) mbar_wait_idle(my_tok, x == 0 ? ((tk & 1u) ^ 1u) : (tk & 1u), idle_sm);
        float l4[4] = {0.f, 0.f, 0.f, 0.f};                   // independent partial sums (ILP)
#pragma unroll
        for (int unit = 0; unit < 8; ++unit) {                 // 8 columns -> one 16 B unit of the swizzled row
          float e[8];
#pragma unroll
          for (int t = 0; t < 8; ++t)
            e[t] = ex2_approx(fmaf(__uint_as_float(s[unit >> 2][(unit & 3) * 8 + t]), c, -mc));   // ex2(-inf) = 0 for masked keys
          l4[0] += e[0] + e[1];
          l4[1] += e[2] + e[3];
          l4[2] += e[4] + e[5];
          l4[3] += e[6] + e[7];
          uint4 pk;
          pk.x = pack_half2(e[0], e[1]);
          pk.y = pack_half2(e[2], e[3]);
          pk.z = pack_half2(e[4], e[5]);
          pk.w = pack_half2(e[6], e[7]);
          *reinterpret_cast<uint4*>(p_row + ((unit ^ (r & 7)) << 4)) = pk;
        }
        if (paired) {
          if (use_token) {
            __syncwarp();
            if (lane == 0) mbar_arrive(other_tok);             // the partner warp may exponentiate now
          }
          ++tk;
        }
        const float l_blk = (l4[0] + l4[1]) + (l4[2] + l4[3]);
        s_ok = (j + 1 < nkb) ? mbar_test_wait(&s_full[x], (g + 1) & 1u) : 0u;   // consumed at the next loop top
        if (j > 0 && any_grow) {                               // O holds blocks 0..j-1 of this item (pv_done waited above)
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(tm_o + half * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32b_x32(tm_o + half * 32, o);
          }
          tmem_wait_st();
        }
        l_run = l_run * alpha + l_blk;
        m_run = m_new;
        fence_proxy_async_smem();        // P (generic-proxy stores) -> visible to the tensor core's async proxy
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[x]);
      }
      // ---------------- O / l -> ctx ----------------
      mbar_wait_idle(&pv_done[x], (g - 1) & 1u, idle_sm);
      tc_fence_after();
      const float inv_l = 1.0f / l_run;
      const int q = q0 + r;
      __half* orow = ctx + (row_base + q) * H + h * C::DH;
      uint32_t o[2][32];
      tmem_ld_32x32b_x32(tm_o, o[0]);
      tmem_ld_32x32b_x32(tm_o + 32, o[1]);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[x]);                  // the context's next first P V may overwrite O now
      if (q0 + C::BQ <= row_limit) {
        // Full tile: stage the warp's 32 rows in its own quarter of the P buffer (the item's last P V has retired) and let
        // the TMA engine write them; the next item's first P store waits for the read (store_pending).
        uint8_t* stg = smem + C::OFF_P + x * C::P_BYTES + wi * 4096;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 w;
            w.x = pack_half2(__uint_as_float(o[half][8 * u + 0]) * inv_l, __uint_as_float(o[half][8 * u + 1]) * inv_l);
            w.y = pack_half2(__uint_as_float(o[half][8 * u + 2]) * inv_l, __uint_as_float(o[half][8 * u + 3]) * inv_l);
            w.z = pack_half2(__uint_as_float(o[half][8 * u + 4]) * inv_l, __uint_as_float(o[half][8 * u + 5]) * inv_l);
            w.w = pack_half2(__uint_as_float(o[half][8 * u + 6]) * inv_l, __uint_as_float(o[half][8 * u + 7]) * inv_l);
            const int unit = half * 4 + u;
            *reinterpret_cast<uint4*>(stg + lane * 128 + ((unit ^ (lane & 7)) << 4)) = w;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmap_ctx, stg, h * C::DH, static_cast<int>(row_base) + q0 + wi * 32);
          bulk_commit_group();
        }
        store_pending = true;
      } else if (q < row_limit) {                              // ragged last tile: later rows belong to the next sequence
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 w;
            w.x = pack_half2(__uint_as_float(o[half][8 * u + 0]) * inv_l, __uint_as_float(o[half][8 * u + 1]) * inv_l);
            w.y = pack_half2(__uint_as_float(o[half][8 * u + 2]) * inv_l, __uint_as_float(o[half][8 * u + 3]) * inv_l);
            w.z = pack_half2(__uint_as_float(o[half][8 * u + 4]) * inv_l, __uint_as_float(o[half][8 * u + 5]) * inv_l);
            w.w = pack_half2(__uint_as_float(o[half][8 * u + 6]) * inv_l, __uint_as_float(o[half][8 * u + 7]) * inv_l);
            *reinterpret_cast<uint4*>(orow + half * 32 + u * 8) = w;
          }
        }
      }
    }
    if (store_pending && lane == 0) bulk_wait_read_all();      // the staging rows must outlive the last store's read
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

}  // namespace mv
