// Fused [CLS] pool + header + CWE-anchor match, all fp32 on CUDA cores,
// ONE cooperative launch with grid-wide syncs between phases:
//   P1  pooled = tanh(Wp . h[:,0] + bp)                       AllenNLP BertPooler    (model_memory.py:64,99)
//   P2  u      = relu(Wh . pooled + bh)                       FeedForward 768->512   (model_memory.py:70,102)
//   P3  uterm[b,c] = Wu[c] . u[b]                              first third of Linear(1536->2)
//   P4  logits[b,g,c] = uterm[b,c] + vterm[g,c] + sum_k Wd[c,k] |u[b,k] - v[g,k]|   (model_memory.py:135-141)
//       p = softmax_c(logits)                                  (:142)
//       best[b] = argmax_g p[b,g,same]  (first maximum wins)   (:144-145)
//   P5  best_idx[b], best_probs[b,:] = p[b, best_idx[b], :]     (:146-147)
// The reference materialises a [B,G,1536] concat tensor; here the separable form is used:
// the anchor bank v [G,512] is read once per (b-chunk) pass with 128-bit coalesced loads, kept in
// registers, and only the |u - v| term is O(B*G*512).  vterm = Wv . v is precomputed when the bank is
// built (bank_vterm_kernel).
//
// P4 work item = (4 anchors) x (chunk of BC queries), one warp per item, K split across the 32 lanes.
#pragma once
#include <cooperative_groups.h>
#include "ptx.cuh"

namespace mv {
namespace cg = cooperative_groups;

struct PoolMatchParams {
  // inputs
  const float* cls;        // row b at cls + b * cls_stride  ([CLS] hidden state, H floats)
  long long cls_stride;
  const float* wp; const float* bp;      // [H,H], [H]
  const float* wh; const float* bh;      // [D,H], [D]
  const float* wproj;                    // [2, 3*D] = [Wu | Wv | Wd]
  const float* bank;                     // [G,D]
  const float* vterm;                    // [G,2]
  // workspace / outputs
  float* pooled;                         // [B,H]
  float* u;                              // [B,D]
  float* uterm;                          // [B,2]
  unsigned long long* best_key;          // [B]
  float* logits;                         // [B,G,2]
  float* probs;                          // [B,G,2]
  int* best_idx;                         // [B]
  float* best_probs;                     // [B,2]
  int B, G, H, D, same_idx, b_chunk, phase_mask, tiled;
};

// out[b,n] = act(sum_k x[b,k] W[n,k] + bias[n]).  Work item = (chunk of 8 rows b, output column n), chunk-major; a block
// takes a contiguous range of items, stages the chunk's 8 input rows in shared memory ONCE and lets its warps walk the
// range's columns: the W row lives in registers, the 8 dot products are independent chains whose warp reductions
// interleave.  (r01 read the 8 rows from global memory for every column: 151 MB of L2 traffic at B = 64 and 660 MB at
// B = 256 for the pooler alone -- ~30 us and ~130 us of the launch.)
constexpr int kDenseRows = 8, kDenseMaxK = 768;
template <int ACT /*0 tanh, 1 relu*/>
__device__ __forceinline__ void dense_rows_phase(const float* x, long long x_stride, const float* __restrict__ W,
                                                 const float* __restrict__ bias, float* out, int B, int N, int K,
                                                 float* xs /* shared [kDenseRows * kDenseMaxK] */) {
  constexpr int MAXV = kDenseMaxK / 128;        // K <= 768, multiple of 128
  constexpr int BCH = kDenseRows;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int nv = K >> 7;                        // float4 per lane
  const int nbc = (B + BCH - 1) / BCH;
  const long long items = static_cast<long long>(N) * nbc;
  const long long per = (items + gridDim.x - 1) / gridDim.x;
  const long long i0 = blockIdx.x * per, i1 = min(items, i0 + per);
  for (long long seg = i0; seg < i1;) {
    const int chunk = static_cast<int>(seg / N);
    const long long seg_end = min(i1, static_cast<long long>(chunk + 1) * N);
    const int b0 = chunk * BCH;
    __syncthreads();                            // the previous segment is done with xs
    for (int i = threadIdx.x; i < BCH * (K >> 2); i += blockDim.x) {
      const int r = i / (K >> 2), k4 = i - r * (K >> 2);
      const int bb = min(b0 + r, B - 1);        // tail rows repeat the last row; never stored
      reinterpret_cast<float4*>(xs)[r * (K >> 2) + k4] =
          *reinterpret_cast<const float4*>(x + static_cast<size_t>(bb) * x_stride + k4 * 4);
    }
    __syncthreads();
    for (long long item = seg + warp; item < seg_end; item += wpb) {
      const int n = static_cast<int>(item - static_cast<long long>(chunk) * N);
      float4 w[MAXV];
#pragma unroll
      for (int i = 0; i < MAXV; ++i)
        w[i] = (i < nv) ? __ldg(reinterpret_cast<const float4*>(W + static_cast<size_t>(n) * K + i * 128 + lane * 4))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
      float acc[BCH];
#pragma unroll
      for (int r = 0; r < BCH; ++r) acc[r] = 0.f;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        if (i < nv) {
#pragma unroll
          for (int r = 0; r < BCH; ++r) {
            const float4 xv = *reinterpret_cast<const float4*>(xs + r * K + i * 128 + lane * 4);
            acc[r] = fmaf(xv.x, w[i].x, acc[r]);
            acc[r] = fmaf(xv.y, w[i].y, acc[r]);
            acc[r] = fmaf(xv.z, w[i].z, acc[r]);
            acc[r] = fmaf(xv.w, w[i].w, acc[r]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < BCH; ++r) acc[r] = warp_sum(acc[r]);
      const float bn = bias[n];
#pragma unroll
      for (int r = 0; r < BCH; ++r) {
        if (lane == r && b0 + r < B) {
          const float v = acc[r] + bn;
          out[static_cast<size_t>(b0 + r) * N + n] = ACT == 0 ? tanhf(v) : fmaxf(v, 0.f);
        }
      }
    }
    seg = seg_end;
  }
}

__device__ __forceinline__ void match_phase(const PoolMatchParams& p, int gwarp, int nwarps, int lane,
                                            unsigned long long* sbest) {
  constexpr int MAXJ = 4;                       // D <= 512
  const int D = p.D, G = p.G, B = p.B;
  const int nd4 = D >> 2;
  const float* wd0 = p.wproj + 2 * D;
  const float* wd1 = p.wproj + 3 * D + 2 * D;
  const int n_gq = (G + 3) >> 2;
  const int n_bc = (B + p.b_chunk - 1) / p.b_chunk;
  const int items = n_gq * n_bc;
  // lane's K slice: float4 index lane + 32*j
  float4 w0[MAXJ], w1[MAXJ];
#pragma unroll
  for (int j = 0; j < MAXJ; ++j) {
    const int k4 = lane + 32 * j;
    const bool ok = k4 < nd4;
    w0[j] = ok ? __ldg(reinterpret_cast<const float4*>(wd0) + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
    w1[j] = ok ? __ldg(reinterpret_cast<const float4*>(wd1) + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  auto load_quad = [&](int item, float4 (&dst)[4][MAXJ]) {
    const int g0n = (item / n_bc) * 4;
#pragma unroll
    for (int gi = 0; gi < 4; ++gi) {
      const int g = min(g0n + gi, G - 1);       // clamp: tail anchors recompute the last row, never stored
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) {
        const int k4 = lane + 32 * j;
        dst[gi][j] = (k4 < nd4) ? __ldg(reinterpret_cast<const float4*>(p.bank + static_cast<size_t>(g) * D) + k4)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  };
  float4 v[4][MAXJ];
  for (int item = gwarp; item < items; item += nwarps) {
    const int gq = item / n_bc, bc = item - gq * n_bc;
    const int g0 = gq * 4;
    load_quad(item, v);
    const int b_end = min(B, (bc + 1) * p.b_chunk);
    for (int b = bc * p.b_chunk; b < b_end; ++b) {
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) {
        const int k4 = lane + 32 * j;
        if (k4 < nd4) {
          const float4 uu = *(reinterpret_cast<const float4*>(p.u + static_cast<size_t>(b) * D) + k4);
#pragma unroll
          for (int gi = 0; gi < 4; ++gi) {
            const float dx = fabsf(uu.x - v[gi][j].x), dy = fabsf(uu.y - v[gi][j].y);
            const float dz = fabsf(uu.z - v[gi][j].z), dw = fabsf(uu.w - v[gi][j].w);
            acc[gi * 2 + 0] = fmaf(dx, w0[j].x, acc[gi * 2 + 0]);
            acc[gi * 2 + 1] = fmaf(dx, w1[j].x, acc[gi * 2 + 1]);
            acc[gi * 2 + 0] = fmaf(dy, w0[j].y, acc[gi * 2 + 0]);
            acc[gi * 2 + 1] = fmaf(dy, w1[j].y, acc[gi * 2 + 1]);
            acc[gi * 2 + 0] = fmaf(dz, w0[j].z, acc[gi * 2 + 0]);
            acc[gi * 2 + 1] = fmaf(dz, w1[j].z, acc[gi * 2 + 1]);
            acc[gi * 2 + 0] = fmaf(dw, w0[j].w, acc[gi * 2 + 0]);
            acc[gi * 2 + 1] = fmaf(dw, w1[j].w, acc[gi * 2 + 1]);
          }
        }
      }
      // 8 partial sums x 32 lanes -> halving butterfly (9 shuffles): value (lane>>2)&7 ends in every lane
      float t4[4], t2[2], t1;
      {
        const bool up = lane & 16;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float send = up ? acc[i] : acc[4 + i];
          const float keep = up ? acc[4 + i] : acc[i];
          t4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
      }
      {
        const bool up = lane & 8;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float send = up ? t4[i] : t4[2 + i];
          const float keep = up ? t4[2 + i] : t4[i];
          t2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
      }
      {
        const bool up = lane & 4;
        const float send = up ? t2[0] : t2[1];
        const float keep = up ? t2[1] : t2[0];
        t1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
      t1 += __shfl_xor_sync(0xffffffffu, t1, 2);
      t1 += __shfl_xor_sync(0xffffffffu, t1, 1);
      const float other = __shfl_xor_sync(0xffffffffu, t1, 4);     // class-1 partner of a class-0 lane
      unsigned long long key = 0ull;
      if ((lane & 7) == 0) {
        const int gi = lane >> 3;
        const int g = g0 + gi;
        if (g < G) {
          const float l0 = t1 + p.uterm[b * 2 + 0] + __ldg(p.vterm + g * 2 + 0);
          const float l1 = other + p.uterm[b * 2 + 1] + __ldg(p.vterm + g * 2 + 1);
          const float m = fmaxf(l0, l1);
          const float e0 = expf(l0 - m), e1 = expf(l1 - m);
          const float inv = 1.0f / (e0 + e1);
          const float p0 = e0 * inv, p1 = e1 * inv;
          const size_t o = (static_cast<size_t>(b) * G + g) * 2;
          *reinterpret_cast<float2*>(p.logits + o) = make_float2(l0, l1);
          *reinterpret_cast<float2*>(p.probs + o) = make_float2(p0, p1);
          const float ps = p.same_idx == 0 ? p0 : p1;
          key = (static_cast<unsigned long long>(__float_as_uint(ps)) << 32) |
                static_cast<unsigned long long>(0xFFFFFFFFu - static_cast<unsigned>(g));
        }
      }
      // max over the 4 anchors of this warp (keys are 0 in the other lanes), one atomic per (b, quad)
      key = max(key, __shfl_xor_sync(0xffffffffu, key, 8));
      key = max(key, __shfl_xor_sync(0xffffffffu, key, 16));
      // block-level arg-max first (shared-memory atomics), one global atomic per (block, query) afterwards: with one
      // global atomicMax per (query, anchor quad) the B target addresses serialised B*G/4 L2 atomics.
      if (lane == 0) {
        if (sbest) atomicMax(sbest + b, key);
        else atomicMax(p.best_key + b, key);
      }
    }
  }
}


// Generic match for feature widths the register-blocked phase does not cover (D > 512: ``use_header=False`` matches the
// 768-wide pooled vectors directly, model_memory.py:69,101 -- unused by the shipped configs, so simple beats fast):
// one warp per (query, anchor) pair, K over the lanes.
__device__ __forceinline__ void match_phase_generic(const PoolMatchParams& p, int gwarp, int nwarps, int lane) {
  const int D = p.D, G = p.G;
  const float* wd0 = p.wproj + 2 * D;
  const float* wd1 = p.wproj + 3 * D + 2 * D;
  const long long items = static_cast<long long>(p.B) * G;
  for (long long item = gwarp; item < items; item += nwarps) {
    const int b = static_cast<int>(item / G), g = static_cast<int>(item - static_cast<long long>(b) * G);
    const float* u = p.u + static_cast<size_t>(b) * D;
    const float* v = p.bank + static_cast<size_t>(g) * D;
    float a0 = 0.f, a1 = 0.f;
    for (int k = lane; k < D; k += 32) {
      const float d = fabsf(u[k] - __ldg(v + k));
      a0 = fmaf(d, __ldg(wd0 + k), a0);
      a1 = fmaf(d, __ldg(wd1 + k), a1);
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    if (lane == 0) {
      const float l0 = a0 + p.uterm[b * 2 + 0] + __ldg(p.vterm + g * 2 + 0);
      const float l1 = a1 + p.uterm[b * 2 + 1] + __ldg(p.vterm + g * 2 + 1);
      const float m = fmaxf(l0, l1);
      const float e0 = expf(l0 - m), e1 = expf(l1 - m);
      const float inv = 1.0f / (e0 + e1);
      const float p0 = e0 * inv, p1 = e1 * inv;
      const size_t o = (static_cast<size_t>(b) * G + g) * 2;
      *reinterpret_cast<float2*>(p.logits + o) = make_float2(l0, l1);
      *reinterpret_cast<float2*>(p.probs + o) = make_float2(p0, p1);
      const float ps = p.same_idx == 0 ? p0 : p1;
      atomicMax(p.best_key + b, (static_cast<unsigned long long>(__float_as_uint(ps)) << 32) |
                                    static_cast<unsigned long long>(0xFFFFFFFFu - static_cast<unsigned>(g)));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Tiled match for LARGE problems (BASELINE config 4: 256 queries x 16,384 anchors): the |u - v| term is O(B*G*512)
// FP32 work, so what matters is operand reuse, like an SGEMM.  A block owns a 64-anchor tile of the bank in shared
// memory (read from HBM exactly once per pass over the queries) and sweeps 64-query tiles against it, u arriving
// in double-buffered 64-wide K chunks (cp.async); every thread keeps a 4 x 4 (query x anchor) register tile for
// both classes.  Shared-memory rows are padded (+4 floats) so the 128-bit operand loads are conflict-free.
// Wd (the |u - v| third of Linear(1536 -> 2), both classes) for the tiled match, in CONSTANT memory: every lane of a warp
// multiplies by the same Wd[c][k], so the FFMAs take it as a constant-bank / uniform operand and the two LDS.128 per
// 4-wide K step that fetched it from shared memory (2 of 10: the phase ran at the shared-memory wavefront limit, 0.50 of
// the FP32 lane peak in r01) disappear.  Written by memvul_pool_match with a stream-ordered device-to-device copy before
// every tiled launch; one model per device at a time may use the tiled path concurrently.
__constant__ float c_match_wd[2 * 512];

struct MatchTileCfg {
  static constexpr int BT = 64, GT = 64, KC = 64, D = 512;
  static constexpr int V_LD = D + 4, U_LD = KC + 4;
  static constexpr int OFF_V = 0;                                   // [GT][V_LD] floats
  static constexpr int OFF_U = OFF_V + GT * V_LD * 4;               // [2][BT][U_LD]
  static constexpr int SMEM_BYTES = OFF_U + 2 * BT * U_LD * 4;      // 166,912 B
};

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = smem_u32(smem_dst);
  const int sz = valid ? 16 : 0;                                    // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void match_phase_tiled(const PoolMatchParams& p, uint8_t* smem, unsigned long long* sbest) {
  using T = MatchTileCfg;
  float* sv = reinterpret_cast<float*>(smem + T::OFF_V);
  float* su = reinterpret_cast<float*>(smem + T::OFF_U);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int B = p.B, G = p.G;
  const int n_bt = (B + T::BT - 1) / T::BT, n_gt = (G + T::GT - 1) / T::GT;
  const int items = n_bt * n_gt;
  // contiguous item ranges per block, anchor-tile major: consecutive items of a block share the bank tile
  const int per = (items + gridDim.x - 1) / gridDim.x;
  const int it0 = blockIdx.x * per, it1 = min(items, it0 + per);
  int cur_gt = -1;
  for (int item = it0; item < it1; ++item) {
    const int gt = item / n_bt, bt = item - gt * n_bt;
    const int g0 = gt * T::GT, b0 = bt * T::BT;
    __syncthreads();                                                // previous item is done with sv / su
    if (gt != cur_gt) {                                             // stream this anchor tile in once
      cur_gt = gt;
      for (int c = tid; c < T::GT * (T::D / 4); c += blockDim.x) {
        const int r = c / (T::D / 4), k4 = c % (T::D / 4);
        const int g = g0 + r;
        cp_async_16(sv + r * T::V_LD + k4 * 4, p.bank + static_cast<size_t>(min(g, G - 1)) * T::D + k4 * 4, g < G);
      }
    }
    auto load_u = [&](int kc, int buf) {
      for (int c = tid; c < T::BT * (T::KC / 4); c += blockDim.x) {
        const int r = c / (T::KC / 4), k4 = c % (T::KC / 4);
        const int b = b0 + r;
        cp_async_16(su + (buf * T::BT + r) * T::U_LD + k4 * 4,
                    p.u + static_cast<size_t>(min(b, B - 1)) * T::D + kc * T::KC + k4 * 4, b < B);
      }
      cp_async_commit();
    };
    load_u(0, 0);
    float acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.f;
    constexpr int NKC = T::D / T::KC;
    for (int kc = 0; kc < NKC; ++kc) {
      if (kc + 1 < NKC) { load_u(kc + 1, (kc + 1) & 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
      __syncthreads();
      const float* ub = su + ((kc & 1) * T::BT) * T::U_LD;
#pragma unroll 4
      for (int k4 = 0; k4 < T::KC / 4; ++k4) {
        const int k = kc * T::KC + k4 * 4;
        const float4 w0 = *reinterpret_cast<const float4*>(c_match_wd + k);            // constant bank: no LDS
        const float4 w1 = *reinterpret_cast<const float4*>(c_match_wd + T::D + k);
        float4 uu[4], vv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) uu[i] = *reinterpret_cast<const float4*>(ub + (ty + 16 * i) * T::U_LD + k4 * 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) vv[j] = *reinterpret_cast<const float4*>(sv + (tx + 16 * j) * T::V_LD + k);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float dx = fabsf(uu[i].x - vv[j].x), dy = fabsf(uu[i].y - vv[j].y);
            const float dz = fabsf(uu[i].z - vv[j].z), dw = fabsf(uu[i].w - vv[j].w);
            acc[i][j][0] = fmaf(dx, w0.x, fmaf(dy, w0.y, fmaf(dz, w0.z, fmaf(dw, w0.w, acc[i][j][0]))));
            acc[i][j][1] = fmaf(dx, w1.x, fmaf(dy, w1.y, fmaf(dz, w1.z, fmaf(dw, w1.w, acc[i][j][1]))));
          }
      }
      __syncthreads();                                              // everyone is done with this u buffer
    }
    // epilogue: + Wu.u + Wv.v, softmax, stores, per-query arg-max over this anchor tile
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int b = b0 + ty + 16 * i;
      unsigned long long key = 0ull;
      if (b < B) {
        const float ut0 = p.uterm[b * 2 + 0], ut1 = p.uterm[b * 2 + 1];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int g = g0 + tx + 16 * j;
          if (g < G) {
            const float l0 = acc[i][j][0] + ut0 + __ldg(p.vterm + g * 2 + 0);
            const float l1 = acc[i][j][1] + ut1 + __ldg(p.vterm + g * 2 + 1);
            const float m = fmaxf(l0, l1);
            const float e0 = expf(l0 - m), e1 = expf(l1 - m);
            const float inv = 1.0f / (e0 + e1);
            const float p0 = e0 * inv, p1 = e1 * inv;
            const size_t o = (static_cast<size_t>(b) * G + g) * 2;
            *reinterpret_cast<float2*>(p.logits + o) = make_float2(l0, l1);
            *reinterpret_cast<float2*>(p.probs + o) = make_float2(p0, p1);
            const float ps = p.same_idx == 0 ? p0 : p1;
            const unsigned long long kk = (static_cast<unsigned long long>(__float_as_uint(ps)) << 32) |
                                          static_cast<unsigned long long>(0xFFFFFFFFu - static_cast<unsigned>(g));
            key = max(key, kk);
          }
        }
      }
      // the 16 lanes sharing ty hold the other anchors of query b
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) key = max(key, __shfl_xor_sync(0xffffffffu, key, o));
      if (tx == 0 && b < B && key) {
        if (sbest) atomicMax(sbest + b, key);
        else atomicMax(p.best_key + b, key);
      }
    }
  }
}

constexpr int kMaxSmemBest = 1024;     // queries per launch whose running arg-max lives in shared memory (8 KB)
enum : int { PM_POOL = 1, PM_HEADER = 2, PM_UTERM = 4, PM_MATCH = 8, PM_FINAL = 16, PM_ALL = 31 };

__global__ void __launch_bounds__(256) pool_match_kernel(const PoolMatchParams p) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int gwarp = blockIdx.x * wpb + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * wpb;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthreads = gridDim.x * blockDim.x;
  const bool multi = (p.phase_mask & (p.phase_mask - 1)) != 0;     // >1 phase => cooperative launch
  bool need_sync = false;
  auto phase_sync = [&]() {
    if (need_sync && multi) cg::this_grid().sync();
    need_sync = true;
  };
  __shared__ __align__(16) float dense_xs[kDenseRows * kDenseMaxK];        // 24 KB: the 8 input rows of a dense-phase segment
  if (p.phase_mask & PM_POOL) {
    phase_sync();
    dense_rows_phase<0>(p.cls, p.cls_stride, p.wp, p.bp, p.pooled, p.B, p.H, p.H, dense_xs);
    if (p.best_key)
      for (int b = gtid; b < p.B; b += nthreads) p.best_key[b] = 0ull;
  }
  if (p.phase_mask & PM_HEADER) {
    phase_sync();
    dense_rows_phase<1>(p.pooled, p.H, p.wh, p.bh, p.u, p.B, p.D, p.H, dense_xs);
  }
  if (p.phase_mask & PM_UTERM) {
    phase_sync();
    for (int w = gwarp; w < p.B * 2; w += nwarps) {
      const int b = w >> 1, c = w & 1;
      const float* wu = p.wproj + c * 3 * p.D;
      float acc = 0.f;
      for (int k = lane; k < p.D; k += 32) acc = fmaf(p.u[static_cast<size_t>(b) * p.D + k], __ldg(wu + k), acc);
      acc = warp_sum(acc);
      if (lane == 0) p.uterm[w] = acc;
    }
    if (!(p.phase_mask & PM_POOL))
      for (int b = gtid; b < p.B; b += nthreads) p.best_key[b] = 0ull;
  }
  if (p.phase_mask & PM_MATCH) {
    phase_sync();
    __shared__ unsigned long long sbest_buf[kMaxSmemBest];
    unsigned long long* sbest = p.B <= kMaxSmemBest ? sbest_buf : nullptr;
    if (sbest) {
      for (int b = threadIdx.x; b < p.B; b += blockDim.x) sbest[b] = 0ull;
      __syncthreads();
    }
    if (p.tiled) {
      extern __shared__ __align__(16) uint8_t dyn_smem[];
      match_phase_tiled(p, dyn_smem, sbest);
    } else if (p.D > 512) {
      match_phase_generic(p, gwarp, nwarps, lane);
    } else {
      match_phase(p, gwarp, nwarps, lane, sbest);
    }
    if (sbest) {
      __syncthreads();
      for (int b = threadIdx.x; b < p.B; b += blockDim.x)
        if (sbest[b]) atomicMax(p.best_key + b, sbest[b]);
    }
  }
  if (p.phase_mask & PM_FINAL) {
    phase_sync();
    for (int b = gtid; b < p.B; b += nthreads) {
      const unsigned long long key = p.best_key[b];
      const int g = static_cast<int>(0xFFFFFFFFu - static_cast<unsigned>(key & 0xFFFFFFFFull));
      p.best_idx[b] = g;
      const size_t o = (static_cast<size_t>(b) * p.G + g) * 2;
      p.best_probs[b * 2 + 0] = p.probs[o];
      p.best_probs[b * 2 + 1] = p.probs[o + 1];
    }
  }
}

// vterm[g,c] = Wv[c] . bank[g]   (second third of Linear(1536->2); once per bank build)
__global__ void __launch_bounds__(256) bank_vterm_kernel(const float* __restrict__ bank,
                                                         const float* __restrict__ wproj, float* __restrict__ vterm,
                                                         int G, int D) {
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= G) return;
  const float* wv0 = wproj + D;
  const float* wv1 = wproj + 3 * D + D;
  float a0 = 0.f, a1 = 0.f;
  for (int k = lane; k < D; k += 32) {
    const float x = bank[static_cast<size_t>(g) * D + k];
    a0 = fmaf(x, __ldg(wv0 + k), a0);
    a1 = fmaf(x, __ldg(wv1 + k), a1);
  }
  a0 = warp_sum(a0);
  a1 = warp_sum(a1);
  if (lane == 0) {
    vterm[g * 2 + 0] = a0;
    vterm[g * 2 + 1] = a1;
  }
}

// MemVul-m head (model_single.py:62-65,88-90): logits[b,c] = Wc[c] . h[b]; probs = softmax.  One warp per sample.
__global__ void __launch_bounds__(256) single_head_kernel(const float* __restrict__ hfeat,
                                                          const float* __restrict__ wc, float* __restrict__ logits,
                                                          float* __restrict__ probs, int B, int D) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  float a0 = 0.f, a1 = 0.f;
  for (int k = lane; k < D; k += 32) {
    const float x = hfeat[static_cast<size_t>(b) * D + k];
    a0 = fmaf(x, __ldg(wc + k), a0);
    a1 = fmaf(x, __ldg(wc + D + k), a1);
  }
  a0 = warp_sum(a0);
  a1 = warp_sum(a1);
  if (lane == 0) {
    const float m = fmaxf(a0, a1);
    const float e0 = expf(a0 - m), e1 = expf(a1 - m);
    const float inv = 1.0f / (e0 + e1);
    logits[b * 2] = a0; logits[b * 2 + 1] = a1;
    probs[b * 2] = e0 * inv; probs[b * 2 + 1] = e1 * inv;
  }
}

}  // namespace mv
