// Fused self-attention on tcgen05, second generation (r02):  ctx = softmax(Q K^T / 8 + key_mask) V   (SURVEY.md 2.2 K3,
// HF BertSelfAttention as entered from MemVul/custom_PTM_embedder.py:224-228).  Same data contract as
// attention_tcgen05.cuh (one qkv matrix read through two TMA maps, 64-key blocks, lazy rescaling, persistent CTAs, two
// CTAs per SM); what changed follows the r01p / r02b phase traces of that kernel:
//
//   * the block period (1,600 cycles) was the serial chain of ONE soft-max warp per 32 rows -- wait S, row max (190),
//     64 exponentials + pack + STS per thread (800), hand-over -- with only two such warps per scheduler.  Now EIGHT
//     soft-max warps: warps w and w+4 share the 32 rows of TMEM lane quarter w&3 and each exponentiates HALF of the 64
//     keys.  Both read the whole score row for the maximum (the row maximum is exact, so the two threads of a row agree
//     bit for bit on max / grow / alpha without any exchange); the partial row sums meet once per item through 1 KB of
//     shared memory.  The per-block chain per warp halves and every scheduler has four warps to interleave.
//   * the single MMA warp needed ~450 cycles to issue P.V and ~590 for Q.K^T(g+2), back to back on the critical path.
//     Now two issuers: warp 8 loads Q and issues S = Q K^T (running up to two blocks ahead of the soft-max), warp 9
//     issues O += P V and, right after, the K/V loads of block g+3 (it owns the ring slot P.V(g-1) just freed).
//   * the item boundary (2,400 cycles per 512-key item) made the next item's first P.V wait for the read-out of O.
//     O is double-buffered in the 64 spare TMEM columns (S 2x64 + O 2x64 = 256), and the read-out is split over the
//     eight warps (32 columns each).
//   * P[g&1] may be rewritten as soon as P.V(g-2) has retired -- that is what the soft-max now waits for at the top of
//     a block (it has long happened); P.V(g-1) is awaited only on the rare blocks that rescale O.
#pragma once
#include "ptx.cuh"

namespace mv {

struct Attn2Cfg {
  static constexpr int BQ = 128, BKV = 64, DH = 64, KV_STAGES = 4;
  static constexpr int Q_BYTES = 128 * 64 * 2;             // 16 KB: {64 x 128} fp16 box
  static constexpr int KV_BYTES = BKV * 64 * 2;            // 8 KB: {64 x 64} fp16 box
  static constexpr int P_BYTES = 128 * BKV * 2;            // 16 KB: 128 x 64 fp16 = one swizzled K-chunk
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + Q_BYTES;
  static constexpr int OFF_V = OFF_K + KV_STAGES * KV_BYTES;
  static constexpr int OFF_P = OFF_V + KV_STAGES * KV_BYTES;
  static constexpr int OFF_BAR = OFF_P + 2 * P_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 256;         // 114,944 B
  // two CTAs per SM need 2 x (SMEM_BYTES + 1 KB reserved) <= 228 KB: r02d shipped 115,968 B, got ONE CTA per SM and ran
  // at half speed -- keep the assert
  static_assert(2 * (SMEM_BYTES + 1024) <= 233472, "attention v2 must fit two CTAs per SM");
  static constexpr int SM_WARPS = 8;
  static constexpr int THREADS = (SM_WARPS + 2) * 32;      // 320
  static constexpr int TMEM_COLS = 256;
  static constexpr int TM_S = 0, TM_O = 128;
};

// the (item, key block) sequence every role walks; items with q0 >= len are skipped by everybody
struct AttnCursor {
  int item, step, n_items, n_qt, n_heads, S;
  const int* lens; const int* row_start;
  int b, h, q0, len, row_base, nkb, j;      // current item / block
  uint32_t g, it;                           // global block / item counters (over the CTA's life)
  bool valid;
  __device__ __forceinline__ void load_item() {
    while (item < n_items) {
      const int qt = item % n_qt;
      h = (item / n_qt) % n_heads;
      b = item / (n_qt * n_heads);
      q0 = qt * Attn2Cfg::BQ;
      len = lens[b];
      if (q0 < len) {
        row_base = row_start ? row_start[b] : b * S;
        nkb = (len + Attn2Cfg::BKV - 1) / Attn2Cfg::BKV;
        j = 0;
        valid = true;
        return;
      }
      item += step;
    }
    valid = false;
  }
  __device__ __forceinline__ void init(int first, int stride, int n_items_, int n_qt_, int n_heads_, int S_, const int* lens_,
                                       const int* row_start_) {
    item = first; step = stride; n_items = n_items_; n_qt = n_qt_; n_heads = n_heads_; S = S_;
    lens = lens_; row_start = row_start_;
    g = 0; it = 0;
    load_item();
  }
  __device__ __forceinline__ void next_block() {        // advance one key block (and to the next item after the last)
    ++g;
    if (++j == nkb) { ++it; item += step; load_item(); }
  }
};

__global__ void __launch_bounds__(Attn2Cfg::THREADS, 2)
attention_tcgen05_v2_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_kv,
                            const __grid_constant__ CUtensorMap tmap_ctx,
                            const int* __restrict__ lens, const int* __restrict__ row_start, __half* __restrict__ ctx,
                            int B, int S, int H, int n_qt, int wait_mode,
                            unsigned long long* __restrict__ trace) {
  using C = Attn2Cfg;
  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = static_cast<int>(threadIdx.x & 31);
  const int n_heads = H / C::DH;
  const int n_items = B * n_heads * n_qt;
  // debug only (MEMVUL_ATT_TRACE): CTA 0 stamps clock64() at the phase boundaries of its first 128 key blocks
  const bool tracing = trace != nullptr && blockIdx.x == 0;
  auto stamp = [&](uint32_t g, int slot, int k) {
    if (tracing && g < 128u) trace[slot * 1024 + g * 8 + k] = static_cast<unsigned long long>(clock64());
  };
  const int idle_qk = wait_mode & 3, idle_pv = (wait_mode >> 2) & 3, idle_sm = (wait_mode >> 4) & 3;   // see mbar_wait_idle

  extern __shared__ __align__(1024) uint8_t smem[];        // SWIZZLE_128B tiles need 1024-byte alignment
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* q_full = bars;                           // [1]  Q tile of item `it` landed            (TMA tx)
  uint64_t* q_empty = bars + 1;                      // [1]  last Q K^T of the item retired        (tcgen05.commit)
  uint64_t* k_full = bars + 2;                       // [KV_STAGES]
  uint64_t* v_full = k_full + C::KV_STAGES;          // [KV_STAGES]
  uint64_t* kv_empty = v_full + C::KV_STAGES;        // [KV_STAGES]  P V of that block retired    (tcgen05.commit)
  uint64_t* s_full = kv_empty + C::KV_STAGES;        // [2]  Q K^T of block g in S[g&1]
  uint64_t* p_full = s_full + 2;                     // [2]  P_g in smem, S[g&1] drained (8 warp arrivals)
  uint64_t* pv_done = p_full + 2;                    // [2]  P_g V_g accumulated into O
  uint64_t* o_free = pv_done + 2;                    // [2]  O[it&1] of item it read out (8 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

  if (warp_idx == C::SM_WARPS) {
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
        mbar_init(&p_full[i], C::SM_WARPS);
        mbar_init(&pv_done[i], 1);
        mbar_init(&o_free[i], C::SM_WARPS);
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
  const uint32_t smem_base = smem_u32(smem);

  if (warp_idx == C::SM_WARPS) {
    // ============================== warp 8: Q loads + S = Q K^T ==============================
    // warp-uniform loops, one elected lane issues (keeps the descriptor arithmetic on the uniform datapath)
    const bool issuer = elect_one();
    constexpr uint32_t idesc_qk = umma_idesc_f16(128, C::BKV, false, false);   // both operands K-major
    AttnCursor c;
    c.init(blockIdx.x, gridDim.x, n_items, n_qt, n_heads, S, lens, row_start);
    const uint64_t q_desc = umma_desc_sw128(smem_base + C::OFF_Q);
    while (c.valid) {
      if (c.j == 0) {
        mbar_wait_idle(q_empty, (c.it & 1u) ^ 1u, idle_qk);     // previous item's last Q K^T has retired
        if (issuer) {
          mbar_arrive_expect_tx(q_full, C::Q_BYTES);
          tma_load_2d(smem + C::OFF_Q, &tmap_qkv, q_full, c.h * C::DH, c.row_base + c.q0, kEvictFirst);
        }
        mbar_wait_idle(q_full, c.it & 1u, idle_qk);
      }
      const uint32_t g = c.g, st = g % C::KV_STAGES;
      mbar_wait_idle(&k_full[st], (g / C::KV_STAGES) & 1u, idle_qk);
      if (g >= 2) mbar_wait_idle(&p_full[g & 1u], ((g - 2) >> 1) & 1u, idle_qk);   // S[g&1] drained by the soft-max of block g-2
      tc_fence_after();
      const uint64_t k_desc = umma_desc_sw128(smem_base + C::OFF_K + st * C::KV_BYTES);
      const uint32_t d = tmem_base + C::TM_S + (g & 1u) * C::BKV;
      if (issuer) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ss(d, q_desc + static_cast<uint64_t>(k * 2), k_desc + static_cast<uint64_t>(k * 2), idesc_qk, k != 0 ? 1u : 0u);
        umma_commit(&s_full[g & 1u]);
        if (c.j == c.nkb - 1) umma_commit(q_empty);            // Q may be overwritten once this product retires
      }
      c.next_block();
    }
  } else if (warp_idx == C::SM_WARPS + 1) {
    // ============================== warp 9: O += P V, K/V loads ==============================
    const bool issuer = elect_one();
    constexpr uint32_t idesc_pv = umma_idesc_f16(128, 64, false, true);        // V is N-major
    AttnCursor c, ld;                                           // compute cursor and load cursor (runs 3 blocks ahead)
    c.init(blockIdx.x, gridDim.x, n_items, n_qt, n_heads, S, lens, row_start);
    ld = c;
    auto load_block = [&]() {                                   // K and V of the load cursor's block into its ring slot
      const uint32_t st = ld.g % C::KV_STAGES;
      mbar_wait_idle(&kv_empty[st], ((ld.g / C::KV_STAGES) & 1u) ^ 1u, idle_pv);
      const int row_k = ld.row_base + ld.j * C::BKV;
      if (issuer) {
        mbar_arrive_expect_tx(&k_full[st], C::KV_BYTES);
        tma_load_2d(smem + C::OFF_K + st * C::KV_BYTES, &tmap_kv, &k_full[st], H + ld.h * C::DH, row_k, kEvictLast);
        mbar_arrive_expect_tx(&v_full[st], C::KV_BYTES);
        tma_load_2d(smem + C::OFF_V + st * C::KV_BYTES, &tmap_kv, &v_full[st], 2 * H + ld.h * C::DH, row_k, kEvictLast);
      }
      ld.next_block();
    };
#pragma unroll 1
    for (int i = 0; i < C::KV_STAGES - 1 && ld.valid; ++i) load_block();       // prologue: blocks 0, 1, 2
    while (c.valid) {
      const uint32_t g = c.g, st = g % C::KV_STAGES;
      const uint32_t ob = c.it & 1u;
      mbar_wait_idle(&p_full[g & 1u], (g >> 1) & 1u, idle_pv);                 // P_g in smem
      if (issuer) stamp(g, 1, 0);
      if (c.j == 0 && c.it >= 2) mbar_wait_idle(&o_free[ob], ((c.it - 2) >> 1) & 1u, idle_pv);   // O[ob] of item it-2 read out
      mbar_wait_idle(&v_full[st], (g / C::KV_STAGES) & 1u, idle_pv);
      tc_fence_after();
      const uint32_t p_addr = smem_base + C::OFF_P + (g & 1u) * C::P_BYTES;
      const uint32_t v_addr = smem_base + C::OFF_V + st * C::KV_BYTES;
      if (issuer) {
#pragma unroll
        for (int kk = 0; kk < C::BKV / 16; ++kk) {
          // A = P: K-major 64-wide chunk, 32 B per K=16 step.  B = V: N-major (one 128 B swizzle row per key), 2048 B per step.
          const uint64_t a_desc = umma_desc_sw128(p_addr) + static_cast<uint64_t>(kk * 2);
          const uint64_t b_desc = umma_desc_sw128(v_addr + kk * 2048);
          umma_f16_ss(tmem_base + C::TM_O + ob * 64, a_desc, b_desc, idesc_pv, (c.j | kk) != 0 ? 1u : 0u);
        }
        umma_commit(&pv_done[g & 1u]);
        umma_commit(&kv_empty[st]);      // Q K_g^T retired before the soft-max read S_g, i.e. before P_g existed: K and V are free
        stamp(g, 1, 2);
      }
      c.next_block();
      if (ld.valid) load_block();        // block g+3: its ring slot is the one P V(g-1) has freed
    }
  } else {
    // ======================= soft-max warps 0-7: thread <-> (query row, half of the keys) =======================
    const int quarter = warp_idx & 3;                           // TMEM lane quarter == rows quarter*32 .. +31 of the tile
    const int hs = warp_idx >> 2;                               // which 32 of the block's 64 keys this thread exponentiates
    const int r = quarter * 32 + lane;                          // row in tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const float cs = 1.4426950408889634f * 0.125f;              // log2(e) / sqrt(64)
    const uint32_t pair_bar = 1u + static_cast<uint32_t>(quarter);   // named barrier of the two warps sharing these rows
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory"); };
    uint32_t g = 0, it = 0;
    bool store_pending = false;                                 // (hs == 0) this warp has a ctx TMA store reading staged rows
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
      if (item + static_cast<int>(gridDim.x) < n_items) {      // one item ahead: the dependent load is off the serial path
        const int bn = (item + gridDim.x) / (n_qt * n_heads);
        len_next = lens[bn];
        rb_next = row_start ? row_start[bn] : bn * S;
      }
      const int row_limit = row_start ? len : S;               // rows of this sequence that exist in the token-major matrix
      if (q0 >= len) {
        // fully padded query tile (padded layout only): deterministic zeros, no tensor work
        const int rows = row_start ? 0 : min(C::BQ, S - q0);
        for (int i = threadIdx.x; i < rows * 8; i += C::SM_WARPS * 32) {
          const int rr = i >> 3, u = i & 7;
          *reinterpret_cast<uint4*>(ctx + (row_base + q0 + rr) * H + h * C::DH + u * 8) = make_uint4(0, 0, 0, 0);
        }
        continue;
      }
      const int nkb = (len + C::BKV - 1) / C::BKV;
      const uint32_t ob = it & 1u;
      float m_run = -INFINITY, l_run = 0.f;
      // the staged ctx rows of the previous item must have left shared memory before P is written again (the store was
      // issued a whole item boundary ago: this never stalls); the partner warp learns it through the pair barrier
      if (hs == 0 && store_pending) {
        if (lane == 0) bulk_wait_read_all();
        store_pending = false;
      }
      pair_sync();
      for (int j = 0; j < nkb; ++j, ++g) {
        const bool tr = warp_idx == 0 && lane == 0;
        if (tr) stamp(g, 0, 0);
        if (g >= 2) mbar_wait_idle(&pv_done[g & 1u], ((g - 2) >> 1) & 1u, idle_sm);   // P[g&1] is free: P V(g-2) retired
        mbar_wait_idle(&s_full[g & 1u], (g >> 1) & 1u, idle_sm);
        tc_fence_after();
        if (tr) stamp(g, 0, 1);
        const uint32_t s_addr = tmem_base + lane_addr + C::TM_S + (g & 1u) * C::BKV;
        const int valid = min(C::BKV, len - j * C::BKV);        // >= 1; only the last block of a sequence is ragged
        uint32_t s[32];
        float mx;
        {
          // the OTHER half of the row first: only its maximum is needed
          tmem_ld_32x32b_x32(s_addr + (1 - hs) * 32, s);
          tmem_wait_ld();
          if (valid < C::BKV) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if ((1 - hs) * 32 + i >= valid) s[i] = 0xff800000u;
          }
          float m4[2];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const uint32_t* sp = &s[q * 16];
            float m = __uint_as_float(sp[0]);
#pragma unroll
            for (int i = 1; i < 15; i += 2) m = max3(m, __uint_as_float(sp[i]), __uint_as_float(sp[i + 1]));
            m4[q] = fmaxf(m, __uint_as_float(sp[15]));
          }
          mx = fmaxf(m4[0], m4[1]);
        }
        tmem_ld_32x32b_x32(s_addr + hs * 32, s);
        tmem_wait_ld();
        if (valid < C::BKV) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (hs * 32 + i >= valid) s[i] = 0xff800000u;        // -inf: exp2 -> 0, never the max
        }
        {
          float m4[2];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const uint32_t* sp = &s[q * 16];
            float m = __uint_as_float(sp[0]);
#pragma unroll
            for (int i = 1; i < 15; i += 2) m = max3(m, __uint_as_float(sp[i]), __uint_as_float(sp[i + 1]));
            m4[q] = fmaxf(m, __uint_as_float(sp[15]));
          }
          mx = max3(mx, m4[0], m4[1]);
        }
        // Lazy rescaling (see attention_tcgen05.cuh): exponentiate against the stale reference m_run until some row's
        // maximum has grown by more than 2^8.  Both threads of a row hold the same m_run / mx, hence the same decision.
        const bool grow = (mx - m_run) * cs > 8.0f;             // true on the first block (m_run = -inf)
        const bool any_grow = __any_sync(0xffffffffu, grow);
        const float m_new = grow ? mx : m_run;
        if (tr) stamp(g, 0, 2);
        const float mc = m_new * cs;
        uint8_t* p_row = smem + C::OFF_P + (g & 1u) * C::P_BYTES + r * 128;
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int u = 0; u < 4; ++u) {                           // 8 columns -> one 16 B unit of the swizzled row
          float e[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) e[t] = ex2_approx(fmaf(__uint_as_float(s[u * 8 + t]), cs, -mc));
          l4[0] += e[0] + e[1];
          l4[1] += e[2] + e[3];
          l4[2] += e[4] + e[5];
          l4[3] += e[6] + e[7];
          uint4 pk;
          pk.x = pack_half2(e[0], e[1]);
          pk.y = pack_half2(e[2], e[3]);
          pk.z = pack_half2(e[4], e[5]);
          pk.w = pack_half2(e[6], e[7]);
          const int unit = hs * 4 + u;                          // 16 B unit inside the 64-column row
          *reinterpret_cast<uint4*>(p_row + ((unit ^ (r & 7)) << 4)) = pk;
        }
        const float l_blk = (l4[0] + l4[1]) + (l4[2] + l4[3]);
        if (tr) stamp(g, 0, 3);
        const float alpha = ex2_approx((m_run - m_new) * cs);   // 0 on the first block, else 1 unless grown
        if (j > 0 && any_grow) {
          mbar_wait_idle(&pv_done[(g - 1) & 1u], ((g - 1) >> 1) & 1u, idle_sm);   // O holds blocks 0..j-1 of this item
          tc_fence_after();
          uint32_t o[32];
          const uint32_t o_addr = tmem_base + lane_addr + C::TM_O + ob * 64 + hs * 32;   // this thread's half of the row
          tmem_ld_32x32b_x32(o_addr, o);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_32x32b_x32(o_addr, o);
          tmem_wait_st();
        }
        l_run = l_run * alpha + l_blk;
        m_run = m_new;
        fence_proxy_async_smem();        // P (generic-proxy stores) -> visible to the tensor core's async proxy
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g & 1u]);
        if (tr) stamp(g, 0, 4);
      }
      // ---------------- O / l -> ctx ----------------
      // partial row sums meet in the OTHER P buffer (P[g&1]: its last reader P V(g-2) retired before P V(g-1), which
      // is awaited first; its next writer is the next item's first block, after two more pair barriers)
      float* lsum = reinterpret_cast<float*>(smem + C::OFF_P + (g & 1u) * C::P_BYTES + r * 128);
      mbar_wait_idle(&pv_done[(g - 1) & 1u], ((g - 1) >> 1) & 1u, idle_sm);   // in-order retirement: P V(g-2) is done too
      tc_fence_after();
      lsum[hs] = l_run;
      if (warp_idx == 0 && lane == 0) stamp(g - 1, 0, 5);
      uint32_t o[32];
      tmem_ld_32x32b_x32(tmem_base + lane_addr + C::TM_O + ob * 64 + hs * 32, o);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[ob]);                  // item it+2 may accumulate into O[ob] now
      pair_sync();                                              // partner's partial sum is in shared memory
      const float inv_l = 1.0f / (l_run + lsum[1 - hs]);
      const int q = q0 + r;
      uint4 pk[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        pk[u].x = pack_half2(__uint_as_float(o[8 * u + 0]) * inv_l, __uint_as_float(o[8 * u + 1]) * inv_l);
        pk[u].y = pack_half2(__uint_as_float(o[8 * u + 2]) * inv_l, __uint_as_float(o[8 * u + 3]) * inv_l);
        pk[u].z = pack_half2(__uint_as_float(o[8 * u + 4]) * inv_l, __uint_as_float(o[8 * u + 5]) * inv_l);
        pk[u].w = pack_half2(__uint_as_float(o[8 * u + 6]) * inv_l, __uint_as_float(o[8 * u + 7]) * inv_l);
      }
      if (q0 + C::BQ <= row_limit) {
        // full tile: stage the quarter's 32 rows in its rows of the P buffer the item's last block used (its P V has
        // retired), each warp its 64-byte half, and let the TMA engine write the 32 x 64 box
        uint8_t* stg = smem + C::OFF_P + ((g - 1) & 1u) * C::P_BYTES + quarter * 4096;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int unit = hs * 4 + u;
          *reinterpret_cast<uint4*>(stg + lane * 128 + ((unit ^ (lane & 7)) << 4)) = pk[u];
        }
        fence_proxy_async_smem();
        pair_sync();                                            // both halves are staged
        if (hs == 0) {
          if (lane == 0) {
            tma_store_2d(&tmap_ctx, stg, h * C::DH, static_cast<int>(row_base) + q0 + quarter * 32);
            bulk_commit_group();
          }
          store_pending = true;
        }
      } else {
        if (q < row_limit) {                                    // ragged last tile: later rows belong to the next sequence
          __half* orow = ctx + (row_base + q) * H + h * C::DH + hs * 32;
#pragma unroll
          for (int u = 0; u < 4; ++u) *reinterpret_cast<uint4*>(orow + u * 8) = pk[u];
        }
        pair_sync();                                            // keeps the barrier count per item uniform (lsum reuse)
      }
      if (warp_idx == 0 && lane == 0) stamp(g - 1, 0, 6);
      ++it;
    }
    if (hs == 0 && store_pending && lane == 0) bulk_wait_read_all();   // the staging rows must outlive the last store's read
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == C::SM_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

}  // namespace mv
