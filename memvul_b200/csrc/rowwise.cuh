// Row-wise HBM-bound kernels: embedding gather + LayerNorm and the
// LayerNorm that follows the two residual GEMMs (second half of rows K4 / K6).
// One warp per token row; every lane owns NV float4 (NV = H / 128), loads are 128-bit and
// fully coalesced, statistics are two-pass in registers (mean, then centred variance) exactly
// like the reference's fp32 LayerNorm (eps 1e-12, HF BertEmbeddings / BertSelfOutput / BertOutput).
// Each kernel writes the fp32 residual stream AND the fp16 copy the next tcgen05 GEMM reads.
#pragma once
#include "ptx.cuh"

namespace mv {

template <int NV>
__device__ __forceinline__ void ln_store(float4 (&v)[NV], const float* __restrict__ gamma,
                                         const float* __restrict__ beta, float eps, float* __restrict__ out32,
                                         __half* __restrict__ out16, int lane) {
  constexpr int H = NV * 128;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) * (1.0f / H);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / H) + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = i * 128 + lane * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col));
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + col));
    float4 o;
    o.x = v[i].x * rstd * g.x + b.x;
    o.y = v[i].y * rstd * g.y + b.y;
    o.z = v[i].z * rstd * g.z + b.z;
    o.w = v[i].w * rstd * g.w + b.w;
    if (out32) *reinterpret_cast<float4*>(out32 + col) = o;
    if (out16) {
      uint2 pk;
      pk.x = pack_half2(o.x, o.y);
      pk.y = pack_half2(o.z, o.w);
      *reinterpret_cast<uint2*>(out16 + col) = pk;
    }
  }
}

// y [M,H] fp32 (pre-LN residual sum)  ->  x32 [M,H] fp32, x16 [M,H] fp16.   In-place (x32 == y) is fine.
// x32_stride: elements between consecutive output rows of x32 (H for a dense matrix; S*H scatters row b to the
// [CLS] slot of sequence b in a [B,S,H] tensor).
template <int NV>
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const float* y, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps, float* x32,
                                                             long long x32_stride, __half* __restrict__ x16, int M) {
  constexpr int H = NV * 128;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* src = y + static_cast<size_t>(row) * H;
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = *reinterpret_cast<const float4*>(src + i * 128 + lane * 4);
  ln_store<NV>(v, gamma, beta, eps, x32 ? x32 + static_cast<size_t>(row) * x32_stride : nullptr,
               x16 ? x16 + static_cast<size_t>(row) * H : nullptr, lane);
}

// Gather one row per sequence (row b*S of a [B*S, H] matrix) into dense [B, H] matrices: the [CLS] rows that the last
// encoder layer's output projection / FFN actually need (BertPooler reads hidden[:,0] only, model_memory.py:99).
__global__ void __launch_bounds__(256) gather_cls_rows_kernel(const float* __restrict__ x32, const __half* __restrict__ c16,
                                                              float* __restrict__ x32_cls, __half* __restrict__ c16_cls,
                                                              int B, int S, int H) {
  const int b = blockIdx.x;
  const size_t src = static_cast<size_t>(b) * S * H, dst = static_cast<size_t>(b) * H;
  for (int i = threadIdx.x * 4; i < H; i += blockDim.x * 4) {
    *reinterpret_cast<float4*>(x32_cls + dst + i) = *reinterpret_cast<const float4*>(x32 + src + i);
    *reinterpret_cast<uint2*>(c16_cls + dst + i) = *reinterpret_cast<const uint2*>(c16 + src + i);
  }
}

// K1: LN(word[ids] + pos[s] + type[tt]).  ids / type_ids are int64 [B*S] as AllenNLP's
// PretrainedTransformerIndexer produces them; type_ids == nullptr means all-zero
// (custom_PTM_embedder.py:199-202).  Out-of-range ids are clamped to [0, vocab) on the device;
// the host wrapper rejects them up front.
template <int NV>
__global__ void __launch_bounds__(256) embed_layernorm_kernel(
    const long long* __restrict__ ids, const long long* __restrict__ type_ids, const float* __restrict__ word,
    const float* __restrict__ pos, const float* __restrict__ type, const float* __restrict__ gamma,
    const float* __restrict__ beta, float eps, float* __restrict__ x32, __half* __restrict__ x16, int M, int S,
    int vocab, int type_vocab) {
  constexpr int H = NV * 128;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  long long id = ids[row];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  long long tt = type_ids ? type_ids[row] : 0;
  tt = tt < 0 ? 0 : (tt >= type_vocab ? type_vocab - 1 : tt);
  const float* w = word + static_cast<size_t>(id) * H;
  const float* p = pos + static_cast<size_t>(row % S) * H;
  const float* t = type + static_cast<size_t>(tt) * H;
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = i * 128 + lane * 4;
    const float4 a = __ldg(reinterpret_cast<const float4*>(w + col));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p + col));
    const float4 c = __ldg(reinterpret_cast<const float4*>(t + col));
    v[i] = make_float4((a.x + b.x) + c.x, (a.y + b.y) + c.y, (a.z + b.z) + c.z, (a.w + b.w) + c.w);
  }
  ln_store<NV>(v, gamma, beta, eps, x32 + static_cast<size_t>(row) * H, x16 + static_cast<size_t>(row) * H, lane);
}

}  // namespace mv
