// Opt-in accuracy mode of the encoder (MEMVUL_ENC_PRECISE): fp32-grade results from the SAME tcgen05 GEMM kernels.
//
// tcgen05 has no fp32 MMA; the default path rounds every GEMM operand to fp16 (11-bit significand), which keeps the
// match logits within 1e-3 of the fp32 reference for heads of PyTorch's default scale (2.7e-4 measured over 1,024
// S=512 rows) but not for heads 4x larger (profiles/r02h_precision.json).  Here every operand is split into two fp16
// numbers, a = a_hi + a_lo (22 significand bits), and the three significant partial products of a GEMM are obtained
// by CONCATENATING ALONG K:
//     A' = [A_hi | A_lo | A_hi]   (M x 3K)        W' = [W_hi | W_hi | W_lo]   (N x 3K)
//     A' W'^T = A_hi W_hi^T + A_lo W_hi^T + A_hi W_lo^T          (a_lo w_lo ~ 2^-22 relative is dropped)
// so the unchanged kind::f16 kernels (fp16 products are exact in the fp32 accumulator) run a K' = 3K problem.  Between
// the GEMMs the activations stay fp32: the kernels below re-split them, apply the erf-GELU in fp32, and run the
// attention (QK^T, softmax, PV: 4 % of the encoder's FLOPs) in fp32 on the CUDA cores -- a flash-style kernel that is
// several times slower than the tcgen05 one, which is the price of this mode and why it is opt-in.
// Measured (profiles/r02n_precision_split.json, r02n_precise_bench.json): max |u - u_ref| 1.8e-5 instead of 7.7e-4, max
// |logit error| 6.7e-6 / 2.7e-5 / 1.1e-4 at head scales x1 / x4 / x16 (default path: 2.7e-4 / 1.1e-3 / 4.3e-3); 36.5 ms
// instead of 6.6 ms per 64 x 512 batch.  The remaining error is the tensor core's own fp32 accumulation (partial sums
// are aligned and truncated inside every K = 16 MMA), not the operand split.
// Replaces the same reference calls as the fast path (custom_PTM_embedder.py:224-228 -> HF BertLayer).
#pragma once
#include "ptx.cuh"

namespace mv {

// out [M, 3K] fp16 = [hi | lo | hi] of act(x [M, K] fp32);  ACT 0: identity, 1: erf-GELU (HF "gelu", exact erff).
// Rows >= *m_dev (packed batches: device-side row count, rounded up to the 256-row GEMM tile) are skipped.
template <int ACT>
__global__ void __launch_bounds__(256) split3_rows_kernel(const float* __restrict__ x, __half* __restrict__ out, int M,
                                                          int K, const int* __restrict__ m_dev) {
  const int rows = m_dev ? min(M, ((__ldg(m_dev) + 255) >> 8) << 8) : M;
  const int k4n = K >> 2;
  const long long total = static_cast<long long>(rows) * k4n;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int row = static_cast<int>(i / k4n), c4 = static_cast<int>(i - static_cast<long long>(row) * k4n);
    float4 v = *reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * K + c4 * 4);
    if (ACT == 1) {
      v.x = 0.5f * v.x * (1.0f + erff(v.x * 0.70710678118654752f));
      v.y = 0.5f * v.y * (1.0f + erff(v.y * 0.70710678118654752f));
      v.z = 0.5f * v.z * (1.0f + erff(v.z * 0.70710678118654752f));
      v.w = 0.5f * v.w * (1.0f + erff(v.w * 0.70710678118654752f));
    }
    const __half hx = __float2half_rn(v.x), hy = __float2half_rn(v.y), hz = __float2half_rn(v.z), hw = __float2half_rn(v.w);
    const __half lx = __float2half_rn(v.x - __half2float(hx)), ly = __float2half_rn(v.y - __half2float(hy));
    const __half lz = __float2half_rn(v.z - __half2float(hz)), lw = __float2half_rn(v.w - __half2float(hw));
    uint2 hi, lo;
    hi.x = static_cast<uint32_t>(__half_as_ushort(hx)) | (static_cast<uint32_t>(__half_as_ushort(hy)) << 16);
    hi.y = static_cast<uint32_t>(__half_as_ushort(hz)) | (static_cast<uint32_t>(__half_as_ushort(hw)) << 16);
    lo.x = static_cast<uint32_t>(__half_as_ushort(lx)) | (static_cast<uint32_t>(__half_as_ushort(ly)) << 16);
    lo.y = static_cast<uint32_t>(__half_as_ushort(lz)) | (static_cast<uint32_t>(__half_as_ushort(lw)) << 16);
    __half* o = out + static_cast<size_t>(row) * 3 * K + c4 * 4;
    *reinterpret_cast<uint2*>(o) = hi;
    *reinterpret_cast<uint2*>(o + K) = lo;
    *reinterpret_cast<uint2*>(o + 2 * K) = hi;
  }
}

// fp32 self-attention, ctx = softmax(Q K^T / 8 + key_mask) V per head (head_dim 64), flash-style on the CUDA cores.
// qkv fp32 [rows, 3H] (Q | K | V column blocks, head h at columns h*64), ctx fp32 [rows, H]; sequence b occupies rows
// b*S.. (padded layout) or row_start[b].. (packed layout), lens[b] of them valid.  One block = (64-query tile, head,
// sequence); thread (ty, tx) of a 16 x 16 grid owns queries 4ty..4ty+3 and keys / output dims 4tx..4tx+3.
// Shared tiles are stored [d][index] (Q, K) / [key][d] (V) / [key][query] (P) so that every inner-loop operand is one
// conflict-free 128-bit load.  Online softmax with exact expf; keys >= len are excluded (the reference's -10000 mask
// underflows to exactly 0 in fp32: attention_tcgen05.cuh header).
struct AttnF32Cfg {
  static constexpr int BQ = 64, BK = 64, DH = 64, LD = 68;                // LD: padded row of the [64][64] tiles
  static constexpr int SMEM_BYTES = 4 * 64 * LD * 4;                       // Qt, Kt, Vs, Pt: 69,632 B
};

__global__ void __launch_bounds__(256) attention_f32_kernel(const float* __restrict__ qkv, const int* __restrict__ lens,
                                                            const int* __restrict__ row_start, float* __restrict__ ctx,
                                                            int B, int S, int H, int n_qt) {
  using C = AttnF32Cfg;
  extern __shared__ __align__(16) float sm_f32[];
  float* Qt = sm_f32;                       // [d][query]
  float* Kt = Qt + 64 * C::LD;              // [d][key]
  float* Vs = Kt + 64 * C::LD;              // [key][d]
  float* Pt = Vs + 64 * C::LD;              // [key][query]
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int len = lens[b];
  const size_t row_base = row_start ? static_cast<size_t>(row_start[b]) : static_cast<size_t>(b) * S;
  const int row_limit = row_start ? len : S;                               // rows of this sequence that exist
  const int q0 = qt * C::BQ;
  if (q0 >= row_limit) return;
  const size_t ld = static_cast<size_t>(3) * H;
  if (q0 >= len) {                                                         // fully padded query tile: zeros
    for (int i = tid; i < 64 * 16; i += 256) {
      const int r = i >> 4, u = i & 15;
      if (q0 + r < row_limit) *reinterpret_cast<float4*>(ctx + (row_base + q0 + r) * H + h * 64 + u * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }
  // Q tile, transposed into [d][query]; rows past the sequence repeat its last row (never stored)
  for (int i = tid; i < 64 * 16; i += 256) {
    const int r = i >> 4, u = i & 15;
    const int qr = min(q0 + r, row_limit - 1);
    const float4 v = *reinterpret_cast<const float4*>(qkv + (row_base + qr) * ld + h * 64 + u * 4);
    Qt[(u * 4 + 0) * C::LD + r] = v.x; Qt[(u * 4 + 1) * C::LD + r] = v.y;
    Qt[(u * 4 + 2) * C::LD + r] = v.z; Qt[(u * 4 + 3) * C::LD + r] = v.w;
  }
  float o[4][4], m_run[4], l_run[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -INFINITY; l_run[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  }
  const int nkb = (len + C::BK - 1) / C::BK;
  for (int kb = 0; kb < nkb; ++kb) {
    const int k0 = kb * C::BK;
    __syncthreads();                                                       // previous tile's P V is done with Kt / Vs / Pt
    for (int i = tid; i < 64 * 16; i += 256) {
      const int r = i >> 4, u = i & 15;
      const int kr = min(k0 + r, len - 1);                                 // keys >= len are masked below
      const float* src = qkv + (row_base + kr) * ld + h * 64 + u * 4;
      const float4 kv = *reinterpret_cast<const float4*>(src + H);
      const float4 vv = *reinterpret_cast<const float4*>(src + 2 * H);
      Kt[(u * 4 + 0) * C::LD + r] = kv.x; Kt[(u * 4 + 1) * C::LD + r] = kv.y;
      Kt[(u * 4 + 2) * C::LD + r] = kv.z; Kt[(u * 4 + 3) * C::LD + r] = kv.w;
      *reinterpret_cast<float4*>(Vs + r * C::LD + u * 4) = vv;
    }
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
    for (int d = 0; d < 64; ++d) {
      const float4 q = *reinterpret_cast<const float4*>(Qt + d * C::LD + ty * 4);
      const float4 k = *reinterpret_cast<const float4*>(Kt + d * C::LD + tx * 4);
      const float qa[4] = {q.x, q.y, q.z, q.w}, ka[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(qa[i], ka[j], s[i][j]);
    }
    // scores / sqrt(64), key mask, online softmax over the 16 threads that share the query rows
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[i][j] = (k0 + tx * 4 + j < len) ? s[i][j] * 0.125f : -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int off = 1; off < 16; off <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float m_new = fmaxf(m_run[i], mx);                             // finite: every tile has >= 1 valid key
      const float alpha = expf(m_run[i] - m_new);                          // 0 on the first tile (m_run = -inf)
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) { s[i][j] = expf(s[i][j] - m_new); rs += s[i][j]; }
#pragma unroll
      for (int off = 1; off < 16; off <<= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
      l_run[i] = l_run[i] * alpha + rs;
      m_run[i] = m_new;
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] *= alpha;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float4*>(Pt + (tx * 4 + j) * C::LD + ty * 4) = make_float4(s[0][j], s[1][j], s[2][j], s[3][j]);
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < 64; ++j) {
      const float4 p = *reinterpret_cast<const float4*>(Pt + j * C::LD + ty * 4);
      const float4 v = *reinterpret_cast<const float4*>(Vs + j * C::LD + tx * 4);
      const float pa[4] = {p.x, p.y, p.z, p.w}, va[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int dd = 0; dd < 4; ++dd) o[i][dd] = fmaf(pa[i], va[dd], o[i][dd]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + ty * 4 + i;
    if (q < row_limit) {
      const float inv = 1.0f / l_run[i];
      *reinterpret_cast<float4*>(ctx + (row_base + q) * H + h * 64 + tx * 4) =
          make_float4(o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv);
    }
  }
}

// hidden_out[b*S] = x[row_start[b]]: the [CLS] rows of a packed residual stream into their slots of the padded tensor
__global__ void __launch_bounds__(192) scatter_cls_rows_kernel(const float* __restrict__ x, const int* __restrict__ row_start,
                                                               float* __restrict__ out, int B, int S, int H) {
  const int b = blockIdx.x;
  const float4* src = reinterpret_cast<const float4*>(x + static_cast<size_t>(row_start[b]) * H);
  float4* dst = reinterpret_cast<float4*>(out + static_cast<size_t>(b) * S * H);
  for (int i = threadIdx.x; i < H / 4; i += blockDim.x) dst[i] = src[i];
}

}  // namespace mv
