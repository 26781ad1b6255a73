// Microbenchmark: scalar FFMA vs packed fma.rn.f32x2 issue rate on sm_100a (per SM, per clock).
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
__global__ void k_ffma(float *out, float a, float b)
{
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = __fmaf_rn(x[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long x, unsigned long long a, unsigned long long b)
{
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(x), "l"(a), "l"(b));
  return d;
}
__global__ void k_ffma2(float *out, float a, float b)
{
  unsigned long long x[16], aa, bb;
  asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
  for (int i = 0; i < 16; i++) { float f = threadIdx.x + i; asm("mov.b64 %0, {%1, %1};" : "=l"(x[i]) : "f"(f)); }
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = ffma2(x[i], aa, bb);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: FFMA with 3 distinct register sources (x = y*z + x)
__global__ void k_ffma3(float *out, float a, float b)
{
  float x[16], y[16];
#pragma unroll
  for (int i = 0; i < 16; i++) { x[i] = threadIdx.x + i; y[i] = a * i + b; }
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = __fmaf_rn(y[i], y[(i + 5) & 15], x[i]);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fadd2(float *out, float a, float b)
{
  unsigned long long x[16], bb;
  asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
  for (int i = 0; i < 16; i++) { float f = threadIdx.x + i; asm("mov.b64 %0, {%1, %1};" : "=l"(x[i]) : "f"(f)); }
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) asm("add.rn.f32x2 %0, %1, %2;" : "=l"(x[i]) : "l"(x[i]), "l"(bb));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// shared-memory read bandwidth: LDS.32 / LDS.64 / LDS.128, conflict-free
template <int V> __global__ void k_lds(float *out)
{
  __shared__ __align__(16) float s[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) s[i] = i;
  __syncthreads();
  float acc = 0;
  int base = (threadIdx.x * V) & 8191;
  for (int it = 0; it < 1024; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      int idx = (base + u * 1024 + it * V * 32) & 8191;
      if (V == 1) acc += s[idx];
      if (V == 2) { float2 v = *reinterpret_cast<float2 *>(&s[idx]); acc += v.x + v.y; }
      if (V == 4) { float4 v = *reinterpret_cast<float4 *>(&s[idx]); acc += v.x + v.y + v.z + v.w; }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <typename F> float timeit(F f)
{
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  int sms = p.multiProcessorCount;
  float *out; cudaMalloc(&out, sms * 8 * 1024 * 4);
  const int blocks = sms * 2, thr = 512;   // 32 warps per SM
  double ghz = clk * 1e-6;
  printf("SMs %d clock %.3f GHz\n", sms, ghz);
  for (int rep = 0; rep < 2; rep++) {
    float t1 = timeit([&] { k_ffma<<<blocks, thr>>>(out, 1.0001f, 0.5f); });
    float t2 = timeit([&] { k_ffma2<<<blocks, thr>>>(out, 1.0001f, 0.5f); });
    float t3 = timeit([&] { k_ffma3<<<blocks, thr>>>(out, 1.0001f, 0.5f); });
    float t4 = timeit([&] { k_fadd2<<<blocks, thr>>>(out, 1.0001f, 0.5f); });
    double n = (double)blocks * thr * ITER * 16;
    printf("FFMA  (x*a+b)  : %.3f ms  %.1f fma/clk/SM (at nominal clock)\n", t1, n / (t1 * 1e-3) / (ghz * 1e9) / sms);
    printf("FFMA2 packed   : %.3f ms  %.1f fma/clk/SM\n", t2, 2 * n / (t2 * 1e-3) / (ghz * 1e9) / sms);
    printf("FFMA  3-reg    : %.3f ms  %.1f fma/clk/SM\n", t3, n / (t3 * 1e-3) / (ghz * 1e9) / sms);
    printf("FADD2 packed   : %.3f ms  %.1f add/clk/SM\n", t4, 2 * n / (t4 * 1e-3) / (ghz * 1e9) / sms);
    float l1 = timeit([&] { k_lds<1><<<blocks, thr>>>(out); });
    float l2 = timeit([&] { k_lds<2><<<blocks, thr>>>(out); });
    float l4 = timeit([&] { k_lds<4><<<blocks, thr>>>(out); });
    double m = (double)blocks * thr * 1024 * 8 * 4;
    printf("LDS.32 %.1f B/clk/SM  LDS.64 %.1f  LDS.128 %.1f\n", m / (l1 * 1e-3) / (ghz * 1e9) / sms,
           2 * m / (l2 * 1e-3) / (ghz * 1e9) / sms, 4 * m / (l4 * 1e-3) / (ghz * 1e9) / sms);
  }
  return 0;
}
