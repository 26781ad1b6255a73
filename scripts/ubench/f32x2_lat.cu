// Microbenchmark: latency / per-warp throughput of packed FP32 (FFMA2, FADD2) at LOW occupancy:
// K independent dependent-chains per thread, W warps per SM (1 CTA per SM).
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 2048
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 x, u64 a, u64 b) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(x), "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fadd2(u64 x, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(x), "l"(b)); return d; }
__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
template <int K, int MODE> __global__ void kern(float *out, float a, float b, long long *cyc)
{
  u64 x[K];
  u64 aa = pk2(a, a), bb = pk2(b, b);
#pragma unroll
  for (int i = 0; i < K; i++) x[i] = pk2(threadIdx.x + i, threadIdx.x - i);
  long long t0 = clock64();
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int i = 0; i < K; i++) {
      if (MODE == 0) x[i] = ffma2(x[i], aa, bb);              // 64-bit register operands
      if (MODE == 1) x[i] = ffma2(pk2(a, a), x[i], bb);       // scalar-broadcast tap (R.F32 form)
      if (MODE == 2) x[i] = fadd2(x[i], bb);
      if (MODE == 3) { x[i] = fadd2(x[i], x[(i + 1) % K]); x[i] = ffma2(pk2(a, a), x[i], x[(i + 2) % K]); }   // add feeding fma, cross deps
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < K; i++) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int K, int MODE> void run(const char *name, int warps, float *out, long long *cyc)
{
  kern<K, MODE><<<148, warps * 32>>>(out, 1.0001f, 0.5f, cyc);
  cudaDeviceSynchronize();
  kern<K, MODE><<<148, warps * 32>>>(out, 1.0001f, 0.5f, cyc);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  double ops = (double)ITER * K * (MODE == 3 ? 2 : 1);
  printf("%-22s K=%2d warps/SM=%2d: %.2f cycles per packed op per warp; %.2f packed warp-ops/clk/SM\n", name, K, warps, c / ops, ops * warps / c);
}
int main()
{
  float *out; cudaMalloc(&out, 148 * 1024 * 4);
  long long *cyc; cudaMalloc(&cyc, 8);
#define ALLK(MODE, name, W) run<1, MODE>(name, W, out, cyc); run<2, MODE>(name, W, out, cyc); run<4, MODE>(name, W, out, cyc); run<8, MODE>(name, W, out, cyc); run<16, MODE>(name, W, out, cyc);
  for (int W : {4, 8, 12, 16}) {
    ALLK(0, "FFMA2 reg", W)
    ALLK(1, "FFMA2 scalar tap", W)
    ALLK(2, "FADD2", W)
    ALLK(3, "FADD2->FFMA2", W)
  }
  return 0;
}
