// Fused self-attention on tcgen05, THREE resident soft-max streams per SM:
//     ctx = softmax(Q K^T / sqrt(64) + key_mask) V          (HF BertSelfAttention as entered from
//     MemVul/custom_PTM_embedder.py:224-228).  Same data contract, item order and numerics as attention_tcgen05.cuh
// (one qkv matrix read through two TMA maps, 64-key blocks, lazy rescaling by 2^8, persistent CTAs); what changed
// follows the r02o probes of that kernel: alone on an SM one CTA needs 1,589 cycles per key block of which only 657
// are exponentials, two co-resident CTAs interleave to 1,850 cycles per two blocks -- the MUFU pipe is busy 55 % of the
// time and the kernel is bound by the SERIAL per-block chain of a CTA (wait S, TMEM load, exponentials, hand-over,
// wait), not by a pipe.  More independent chains per SM is what helps, and the limits are per-SM resources:
//   registers   3 CTAs x 192 threads -> 112 per thread (first kernel: 167).  The soft-max therefore keeps only the score
//               row (64) and, as the scores are consumed, the packed fp16 probabilities (32) in registers.
//   TMEM        3 x 128 columns: S (64, SINGLE-buffered) + O (64).  S is released to the MMA warp as soon as the four
//               soft-max warps hold it in registers (`s_free`), so Q K^T of block g+1 still runs under the soft-max of g.
//   shared mem  3 x 66 KB: Q 16 K, K ring 2 x 8 K, V ring 2 x 8 K, P 16 K (SINGLE-buffered).  K and V stages are freed
//               separately (K(g) by the commit of Q K_g^T, V(g) by P_g V_g), so the K of block g+2 is requested a whole
//               block period before its product is issued.  P is single-buffered: the probabilities of block g are
//               computed into registers first and stored once P_(g-1) V_(g-1) has retired (tested non-blocking under the
//               exponentials; it has long happened).
// Optionally (template POLY) POLY of every 8 exponentials are evaluated on the FMA pipe (Cody-Waite split + degree-4
// minimax polynomial, relative error 2.7e-6 -- below the fp16 rounding of P) to relieve the MUFU pipe when three streams
// saturate it.
// Warps 0-3: soft-max (one query row per thread);  warp 4: MMA issuer;  warp 5: TMA producer.  Key length <= 512.
#pragma once
#include "ptx.cuh"

namespace mv {

struct Attn3Cfg {
  static constexpr int BQ = 128, BKV = 64, DH = 64, KV_STAGES = 2;
  static constexpr int Q_BYTES = 128 * 64 * 2;             // 16 KB: {64 x 128} fp16 box
  static constexpr int KV_BYTES = BKV * 64 * 2;            // 8 KB: {64 x 64} fp16 box
  static constexpr int P_BYTES = 128 * BKV * 2;            // 16 KB: 128 x 64 fp16 = one swizzled K-chunk
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + Q_BYTES;
  static constexpr int OFF_V = OFF_K + KV_STAGES * KV_BYTES;
  static constexpr int OFF_P = OFF_V + KV_STAGES * KV_BYTES;
  static constexpr int OFF_BAR = OFF_P + P_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 256;         // 65,792 B
  static_assert(3 * (SMEM_BYTES + 1024) <= 233472, "attention v3 must fit three CTAs per SM");
  // Eight warps: 0-3 soft-max, 4 MMA issuer, 5 TMA producer, 6-7 idle -- they only complete the second warpgroup, which
  // setmaxnreg needs.  Launched at 80 registers per thread (3 x 8 warps = 6 per scheduler x 80 x 32 = 15,360 of its 16,384);
  // warpgroup 1 then drops to REGS_AUX and the soft-max warpgroup rises to REGS_SOFTMAX (REGS_AUX + REGS_SOFTMAX = 160, so
  // every scheduler still holds 3 x (120 + 40) x 32 = 15,360).  Six 112-register warps per CTA do NOT fit three times:
  // 18 warps put five on some scheduler (5 x 112 x 32 = 17,920 > 16,384) and the SM silently runs two CTAs (r02q: 160 us).
  static constexpr int THREADS = 256;
  static constexpr int CTAS_PER_SM = 3;
  static constexpr int REGS_LAUNCH = 80, REGS_SOFTMAX = 120, REGS_AUX = 40;
  static_assert(REGS_SOFTMAX + REGS_AUX == 2 * REGS_LAUNCH, "the register pool of a CTA must balance");
  static constexpr int TMEM_COLS = 128, TMEM_COLS_P = 32;
  static constexpr int TM_S = 0, TM_O = 64;
};

// 2^x on the FMA pipe, x <= ~8 (lazy rescaling keeps the argument below 8); -inf (masked key) -> 2^-126 * p ~ 1e-38, which
// rounds to 0 in the fp16 P and is invisible in the fp32 row sum (>= 2^-8).
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;                          // 1.5 * 2^23: round-to-nearest integer in the low mantissa bits
  const float n = t - 12582912.0f;
  const float f = x - n;                                    // [-0.5, 0.5]
  float p = 0.009570102207362652f;
  p = fmaf(p, f, 0.05591785907745361f);
  p = fmaf(p, f, 0.240247443318367f);
  p = fmaf(p, f, 0.6931217908859253f);
  p = fmaf(p, f, 0.9999992847442627f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));   // p * 2^n (n's integer bits sit at the bottom of t)
}

template <int POLY, bool PTMEM>
__global__ void __launch_bounds__(Attn3Cfg::THREADS, Attn3Cfg::CTAS_PER_SM)
attention_tcgen05_v3_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_kv,
                            const __grid_constant__ CUtensorMap tmap_ctx,
                            const int* __restrict__ lens, const int* __restrict__ row_start, __half* __restrict__ ctx,
                            int B, int S, int H, int n_qt, int wait_mode) {
  using C = Attn3Cfg;
  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = static_cast<int>(threadIdx.x & 31);
  const int n_heads = H / C::DH;
  const int n_items = B * n_heads * n_qt;                  // n_qt = query tiles per sequence that are computed
  const int idle_tma = wait_mode & 3, idle_mma = (wait_mode >> 2) & 3, idle_sm = (wait_mode >> 4) & 3;   // see mbar_wait_idle

  extern __shared__ __align__(1024) uint8_t smem[];        // SWIZZLE_128B tiles need 1024-byte alignment
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* q_full = bars;                           // [1]  Q tile of item `it` landed            (TMA tx)
  uint64_t* q_empty = bars + 1;                      // [1]  last Q K^T of the item retired        (tcgen05.commit)
  uint64_t* k_full = bars + 2;                       // [KV_STAGES]
  uint64_t* v_full = k_full + C::KV_STAGES;          // [KV_STAGES]
  uint64_t* k_empty = v_full + C::KV_STAGES;         // [KV_STAGES]  Q K^T of that block retired  (tcgen05.commit)
  uint64_t* v_empty = k_empty + C::KV_STAGES;        // [KV_STAGES]  P V of that block retired    (tcgen05.commit)
  uint64_t* s_full = v_empty + C::KV_STAGES;         // [1]  Q K^T of block g in S
  uint64_t* s_free = s_full + 1;                     // [1]  S of block g is in the soft-max registers (4 warp arrivals)
  uint64_t* p_full = s_free + 1;                     // [1]  P_g in smem, O rescaled                   (4 warp arrivals)
  uint64_t* pv_done = p_full + 1;                    // [1]  P_g V_g accumulated into O
  uint64_t* o_free = pv_done + 1;                    // [1]  O of the previous item read out           (4 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 1);   // [0] S | O (128 columns), [1] P (32 columns, PTMEM only)

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
        mbar_init(&k_empty[i], 1);
        mbar_init(&v_empty[i], 1);
      }
      mbar_init(s_full, 1);
      mbar_init(s_free, 4);
      mbar_init(p_full, 4);
      mbar_init(pv_done, 1);
      mbar_init(o_free, 4);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    if (PTMEM) tmem_alloc(tmem_slot + 1, C::TMEM_COLS_P);    // 3 x (128 + 32) = 480 of the SM's 512 columns
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_p = PTMEM ? tmem_slot[1] : 0u;        // P as the TMEM A operand of O += P V (fp16 pairs, 32 columns)

  // Every role walks the same item sequence; `g` counts key blocks and `it` non-skipped items over the CTA's life.  All
  // single-buffered barriers complete one phase per key block (parity g & 1), the two-stage rings one per two blocks.
  auto decode = [&](int item, int& b, int& h, int& q0, int& len, int& row_base) {
    const int qt = item % n_qt;
    h = (item / n_qt) % n_heads;
    b = item / (n_qt * n_heads);
    q0 = qt * C::BQ;
    len = lens[b];
    row_base = row_start ? row_start[b] : b * S;
  };

  if (warp_idx >= 6) {
    setmaxnreg_dec<C::REGS_AUX>();                             // idle: present only so that warpgroup 1 is complete
  } else if (warp_idx == 5) {
    // ============================== TMA producer (warp-uniform walk, one elected issuing lane) ==============================
    setmaxnreg_dec<C::REGS_AUX>();
    const bool issuer = elect_one();
    uint32_t g = 0, it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int b, h, q0, len, row_base;
      decode(item, b, h, q0, len, row_base);
      if (q0 >= len) continue;
      const int nkb = (len + C::BKV - 1) / C::BKV;
      mbar_wait_idle(q_empty, (it & 1u) ^ 1u, idle_tma);       // previous item's last Q K^T has retired
      if (issuer) {
        mbar_arrive_expect_tx(q_full, C::Q_BYTES);
        tma_load_2d(smem + C::OFF_Q, &tmap_qkv, q_full, h * C::DH, row_base + q0, kEvictFirst);
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
  } else if (warp_idx == 4) {
    // ============================== MMA issuer ==============================
    setmaxnreg_dec<C::REGS_AUX>();
    const bool issuer = elect_one();
    const uint32_t smem_base = smem_u32(smem);
    constexpr uint32_t idesc_qk = umma_idesc_f16(128, C::BKV, false, false);   // S = Q K^T   (both K-major)
    constexpr uint32_t idesc_pv = umma_idesc_f16(128, 64, false, true);        // O += P V    (V is N-major)
    const uint64_t q_desc = umma_desc_sw128(smem_base + C::OFF_Q);
    const uint64_t p_desc = umma_desc_sw128(smem_base + C::OFF_P);
    uint32_t g0 = 0, it = 0;                                    // g0 = global index of the item's first block
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int b, h, q0, len, row_base;
      decode(item, b, h, q0, len, row_base);
      if (q0 >= len) continue;
      const int nkb = (len + C::BKV - 1) / C::BKV;
      auto issue_qk = [&](int j) {
        const uint32_t g = g0 + static_cast<uint32_t>(j);
        const uint32_t st = g % C::KV_STAGES;
        if (g > 0) mbar_wait_idle(s_free, (g - 1) & 1u, idle_mma);            // S of block g-1 is in registers
        mbar_wait_idle(&k_full[st], (g / C::KV_STAGES) & 1u, idle_mma);
        tc_fence_after();
        const uint64_t k_desc = umma_desc_sw128(smem_base + C::OFF_K + st * C::KV_BYTES);
        if (issuer) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(tmem_base + C::TM_S, q_desc + static_cast<uint64_t>(k * 2), k_desc + static_cast<uint64_t>(k * 2),
                        idesc_qk, k != 0 ? 1u : 0u);
          umma_commit(s_full);
          umma_commit(&k_empty[st]);                            // the K stage is reusable once this product retires
          if (j == nkb - 1) umma_commit(q_empty);               // and so is Q after the item's last one
        }
      };
      mbar_wait_idle(q_full, it & 1u, idle_mma);
      issue_qk(0);
      for (int j = 0; j < nkb; ++j) {
        const uint32_t g = g0 + static_cast<uint32_t>(j);
        const uint32_t st = g % C::KV_STAGES;
        if (j + 1 < nkb) issue_qk(j + 1);                       // as soon as S_g is in registers: runs under the soft-max of g
        mbar_wait_idle(p_full, g & 1u, idle_mma);               // P_g in smem, O rescaled
        if (j == 0) mbar_wait_idle(o_free, (it & 1u) ^ 1u, idle_mma);   // previous item's O has been read out
        mbar_wait_idle(&v_full[st], (g / C::KV_STAGES) & 1u, idle_mma);
        tc_fence_after();
        const uint32_t v_addr = smem_base + C::OFF_V + st * C::KV_BYTES;
        if (issuer) {
#pragma unroll
          for (int kk = 0; kk < C::BKV / 16; ++kk) {
            // A = P: K-major 64-wide chunk, 32 B per K=16 step.  B = V: N-major (one 128 B swizzle row per key),
            // 16 keys = 2048 B per step.
            const uint64_t b_desc = umma_desc_sw128(v_addr + kk * 2048);
            if (PTMEM)
              umma_f16_ts(tmem_base + C::TM_O, tmem_p + static_cast<uint32_t>(kk * 8), b_desc, idesc_pv, (j | kk) != 0 ? 1u : 0u);
            else
              umma_f16_ss(tmem_base + C::TM_O, p_desc + static_cast<uint64_t>(kk * 2), b_desc, idesc_pv,
                          (j | kk) != 0 ? 1u : 0u);
          }
          umma_commit(pv_done);
          umma_commit(&v_empty[st]);
        }
      }
      g0 += static_cast<uint32_t>(nkb);
      ++it;
    }
  } else {
    // ======================= soft-max warps: thread <-> query row =======================
    setmaxnreg_inc<C::REGS_SOFTMAX>();
    const int r = warp_idx * 32 + lane;                       // row in tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(warp_idx * 32) << 16;
    const float c = 1.4426950408889634f * 0.125f;             // log2(e) / sqrt(64)
    uint8_t* const p_row = smem + C::OFF_P + r * 128;         // this thread's swizzled P row
    uint32_t g = 0;
    bool store_pending = false;                                // this warp has a ctx TMA store reading its P rows
    int len_next = 0, rb_next = 0;                             // lens[] / row_start[] are loaded one item ahead
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
      const int row_limit = row_start ? len : S;               // rows of this sequence that exist in the token-major matrix
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
      uint32_t s_ok = 0;                                       // early (non-blocking) test of the next block's s_full
      for (int j = 0; j < nkb; ++j, ++g) {
        if (!__all_sync(0xffffffffu, s_ok != 0u)) mbar_wait_idle(s_full, g & 1u, idle_sm);
        tc_fence_after();
        uint32_t s[2][32];
        const uint32_t s_addr = tmem_base + lane_addr + C::TM_S;
        tmem_ld_32x32b_x32(s_addr, s[0]);
        tmem_ld_32x32b_x32(s_addr + 32, s[1]);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free);                    // Q K^T of block g+1 may overwrite S now
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
        // P = 2^(s c - mc) -> packed fp16 in registers (the score registers die as they are consumed)
        uint4 pk[8];
        float l4[4] = {0.f, 0.f, 0.f, 0.f};                   // independent partial sums (ILP)
        uint32_t pv_ok = 0;
#pragma unroll
        for (int unit = 0; unit < 8; ++unit) {                 // 8 columns -> one 16 B unit of the swizzled row
          if (unit == 5 && j > 0) pv_ok = mbar_test_wait(pv_done, (g - 1) & 1u);   // scoreboarded: hides under the exponentials
          float e[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const float x = fmaf(__uint_as_float(s[unit >> 2][(unit & 3) * 8 + t]), c, -mc);
            e[t] = (t >= 8 - POLY) ? ex2_poly(x) : ex2_approx(x);   // ex2(-inf) = 0 for masked keys
          }
          l4[0] += e[0] + e[1];
          l4[1] += e[2] + e[3];
          l4[2] += e[4] + e[5];
          l4[3] += e[6] + e[7];
          pk[unit].x = pack_half2(e[0], e[1]);
          pk[unit].y = pack_half2(e[2], e[3]);
          pk[unit].z = pack_half2(e[4], e[5]);
          pk[unit].w = pack_half2(e[6], e[7]);
        }
        const float l_blk = (l4[0] + l4[1]) + (l4[2] + l4[3]);
        s_ok = (j + 1 < nkb) ? mbar_test_wait(s_full, (g + 1) & 1u) : 0u;   // consumed at the next loop top
        if (j == 0) {
          if (store_pending) {                                 // the previous item's ctx store still reads this warp's P rows
            if (lane == 0) bulk_wait_read_all();
            __syncwarp();
            store_pending = false;
          }
        } else {
          if (!__all_sync(0xffffffffu, pv_ok != 0u)) mbar_wait_idle(pv_done, (g - 1) & 1u, idle_sm);   // P free, O holds blocks 0..j-1
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
#pragma unroll
        for (int unit = 0; unit < 8; ++unit)
          if (!PTMEM) *reinterpret_cast<uint4*>(p_row + ((unit ^ (r & 7)) << 4)) = pk[unit];
        if (PTMEM) {
          // P never touches shared memory: 32 packed fp16 pairs per row -> 32 TMEM columns (word w = keys 2w, 2w+1), read by
          // the tensor core as the A operand.  Saves the 16 KB of STS, the proxy fence and the MMA's 16 KB re-read per block
          // (r02q ncu: the TC pipe is busy 48 % of the cycles against 28 % of math -- its shared-memory operand fetches).
          uint32_t pw[32];
#pragma unroll
          for (int unit = 0; unit < 8; ++unit) {
            pw[unit * 4 + 0] = pk[unit].x;
            pw[unit * 4 + 1] = pk[unit].y;
            pw[unit * 4 + 2] = pk[unit].z;
            pw[unit * 4 + 3] = pk[unit].w;
          }
          tmem_st_32x32b_x32(tmem_p + lane_addr, pw);
          tmem_wait_st();
        }
        l_run = l_run * alpha + l_blk;
        m_run = m_new;
        if (!PTMEM) fence_proxy_async_smem();   // P (generic-proxy stores) -> visible to the tensor core's async proxy
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
      }
      // ---------------- O / l -> ctx ----------------
      mbar_wait_idle(pv_done, (g - 1) & 1u, idle_sm);
      tc_fence_after();
      const float inv_l = 1.0f / l_run;
      const int q = q0 + r;
      __half* orow = ctx + (row_base + q) * H + h * C::DH;
      uint32_t o[2][32];
      tmem_ld_32x32b_x32(tmem_base + lane_addr + C::TM_O, o[0]);
      tmem_ld_32x32b_x32(tmem_base + lane_addr + C::TM_O + 32, o[1]);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);                      // the next item's first P V may overwrite O now
      if (q0 + C::BQ <= row_limit) {
        // Full tile: stage the warp's 32 rows in its own quarter of the P buffer (the item's last P V has retired) and let
        // the TMA engine write them; the next item's first P store waits for the read (store_pending).
        uint8_t* stg = smem + C::OFF_P + warp_idx * 4096;
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
          tma_store_2d(&tmap_ctx, stg, h * C::DH, static_cast<int>(row_base) + q0 + warp_idx * 32);
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
  if (warp_idx == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
    if (PTMEM) tmem_dealloc(tmem_p, C::TMEM_COLS_P);
  }
}

}  // namespace mv
