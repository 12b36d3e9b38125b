// Micro-benchmark behind the anchor-match inner loop (pool_match.cuh, tiled phase): which instruction mix does the
// |u - v| . Wd term at the highest rate on sm_100a?  One "pair-k" = one (query, anchor, k) element for BOTH classes.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/fma_probe tools/fma_probe.cu && gpurun_out/fma_probe
// Variants (16 independent (query, anchor) pairs per thread like the 4 x 4 register tile, operands in registers):
//   0  FADD + 2 x FFMA(|d|, w, acc), w in vector registers           (the r02k kernel: 3-register FFMA)
//   1  same, w read as a kernel-uniform constant-bank operand          (2-register FFMA)
//   2  max form: FMNMX + 2 x FFMA(m, w, acc), w in vector registers
//   3  max form, w constant-bank operand
//   4  max form, FFMA2 packed over two consecutive k (acc pairs = even / odd k), w in vector register pairs
//   5  |d| form with FADD2 for the subtraction + 2 x FFMA2 on |d| pairs (abs by LOP on the ALU pipe)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__constant__ float c_w[2 * 64];

__device__ __forceinline__ unsigned long long pk(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long fsub2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("sub.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

template <int V>
__global__ void __launch_bounds__(256) probe(const float* __restrict__ in, float* __restrict__ out, int iters) {
  // 4 u values x 4 v values per k step, 8 k per outer iteration; operands come from registers refreshed by cheap
  // integer-free recurrences so that the compiler cannot hoist the arithmetic
  float u[4][8], v[4][8], w0[8], w1[8];
  const int t = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) { u[i][k] = in[(t + i * 8 + k) & 1023]; v[i][k] = in[(t * 3 + i * 8 + k + 77) & 1023]; }
#pragma unroll
  for (int k = 0; k < 8; ++k) { w0[k] = in[k + 5]; w1[k] = in[k + 300]; }
  float acc[4][4][2];
  unsigned long long acc2[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j][0] = acc[i][j][1] = 0.f; acc2[i][j][0] = acc2[i][j][1] = 0ull; }
  for (int it = 0; it < iters; ++it) {
    const int cb = (it & 7) * 8;                       // constant-bank offset: uniform, dynamic
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
      const float wa0 = (V == 1 || V == 3) ? c_w[cb + k] : w0[k], wa1 = (V == 1 || V == 3) ? c_w[64 + cb + k] : w1[k];
      const float wb0 = (V == 1 || V == 3) ? c_w[cb + k + 1] : w0[k + 1], wb1 = (V == 1 || V == 3) ? c_w[64 + cb + k + 1] : w1[k + 1];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (V == 0 || V == 1) {
            const float da = fabsf(u[i][k] - v[j][k]), db = fabsf(u[i][k + 1] - v[j][k + 1]);
            acc[i][j][0] = fmaf(da, wa0, acc[i][j][0]); acc[i][j][1] = fmaf(da, wa1, acc[i][j][1]);
            acc[i][j][0] = fmaf(db, wb0, acc[i][j][0]); acc[i][j][1] = fmaf(db, wb1, acc[i][j][1]);
          } else if (V == 2 || V == 3) {
            const float ma = fmaxf(u[i][k], v[j][k]), mb = fmaxf(u[i][k + 1], v[j][k + 1]);
            acc[i][j][0] = fmaf(ma, wa0, acc[i][j][0]); acc[i][j][1] = fmaf(ma, wa1, acc[i][j][1]);
            acc[i][j][0] = fmaf(mb, wb0, acc[i][j][0]); acc[i][j][1] = fmaf(mb, wb1, acc[i][j][1]);
          } else if (V == 4) {
            const float ma = fmaxf(u[i][k], v[j][k]), mb = fmaxf(u[i][k + 1], v[j][k + 1]);
            const unsigned long long m2 = pk(ma, mb);
            acc2[i][j][0] = ffma2(m2, pk(w0[k], w0[k + 1]), acc2[i][j][0]);
            acc2[i][j][1] = ffma2(m2, pk(w1[k], w1[k + 1]), acc2[i][j][1]);
          } else {
            unsigned long long d2 = fsub2(pk(u[i][k], u[i][k + 1]), pk(v[j][k], v[j][k + 1]));
            d2 &= 0x7fffffff7fffffffull;
            acc2[i][j][0] = ffma2(d2, pk(w0[k], w0[k + 1]), acc2[i][j][0]);
            acc2[i][j][1] = ffma2(d2, pk(w1[k], w1[k + 1]), acc2[i][j][1]);
          }
        }
    }
    // perturb one operand per iteration so iterations are not identical (cheap: 8 FADD per 384+ FMA-class instructions)
#pragma unroll
    for (int k = 0; k < 8; ++k) u[0][k] += 1e-3f;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a, b, c, d;
      upk(acc2[i][j][0], a, b); upk(acc2[i][j][1], c, d);
      s += acc[i][j][0] + acc[i][j][1] + a + b + c + d;
    }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int V>
void run(const char* name, const float* in, float* out, int sms, int bps) {
  const int iters = 4096, blocks = sms * bps;     // bps x 256 threads per SM = 2 bps warps per SMSP
  probe<V><<<blocks, 256>>>(in, out, 16);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    probe<V><<<blocks, 256>>>(in, out, iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double pair_k = double(blocks) * 256 * iters * 8 * 16;
  printf("variant %d (%d warps/SMSP) %-48s %8.3f ms  %7.2f T pair-k/s  (x3 = %6.2f T lane-instr/s of the |d| form)\n", V, 2 * bps, name, best,
         pair_k / best / 1e9, 3 * pair_k / best / 1e9);
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("%s, %d SMs, %d MHz\n", p.name, p.multiProcessorCount, p.clockRate / 1000);
  float *in, *out;
  CK(cudaMalloc(&in, 1024 * 4)); CK(cudaMalloc(&out, 148 * 4 * 256 * 4));
  float h[1024];
  for (int i = 0; i < 1024; ++i) h[i] = (float)((i * 2654435761u) >> 8 & 0xffff) / 65536.f;
  CK(cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice));
  CK(cudaMemcpyToSymbol(c_w, h, sizeof(float) * 128));
  const int sms = p.multiProcessorCount;
  for (int bps = 1; bps <= 2; ++bps) {
    run<0>("FADD + 2 FFMA(|d|), w in registers", in, out, sms, bps);
    run<1>("FADD + 2 FFMA(|d|), w uniform register", in, out, sms, bps);
    run<2>("FMNMX + 2 FFMA, w in registers", in, out, sms, bps);
    run<3>("FMNMX + 2 FFMA, w uniform register", in, out, sms, bps);
    run<4>("2 FMNMX + 2 FFMA2 per 2 k (k-packed)", in, out, sms, bps);
    run<5>("FADD2 + 2 LOP + 2 FFMA2 per 2 k", in, out, sms, bps);
  }
  return 0;
}
