// Does alternating kernels with different shared-memory carve-outs cost time?  (B200)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_small(float *o) { if (threadIdx.x == 9999) o[0] = 1; }
__global__ void k_small2(float *o) { if (threadIdx.x == 9999) o[0] = 1; }
__global__ void k_big(float *o) { extern __shared__ float s[]; if (threadIdx.x == 9999) { s[0] = 1; o[0] = s[1]; } }
template <typename F> float timeit(F f, int n)
{
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 5; i++) f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < n; i++) f();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms * 1000.f / n;
}
int main()
{
  float *o; cudaMalloc(&o, 1024);
  const int SM = 56832;
  cudaFuncSetAttribute(k_big, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);
  cudaFuncSetAttribute(k_small2, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  for (int grid : {592, 3221}) {
    printf("grid %d\n", grid);
    printf("  big only            : %.2f us/launch\n", timeit([&] { k_big<<<grid, 160, SM>>>(o); }, 200));
    printf("  small only          : %.2f us/launch\n", timeit([&] { k_small<<<grid, 160>>>(o); }, 200));
    printf("  small + big         : %.2f us/pair\n", timeit([&] { k_small<<<grid, 160>>>(o); k_big<<<grid, 160, SM>>>(o); }, 200));
    printf("  small(maxshared)+big: %.2f us/pair\n", timeit([&] { k_small2<<<grid, 160>>>(o); k_big<<<grid, 160, SM>>>(o); }, 200));
  }
  // same through a CUDA graph (as the extractor submits)
  cudaStream_t st; cudaStreamCreate(&st);
  for (int variant = 0; variant < 2; variant++) {
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
    for (int i = 0; i < 4; i++) { if (variant) k_small2<<<592, 160, 0, st>>>(o); else k_small<<<592, 160, 0, st>>>(o); k_big<<<592, 160, SM, st>>>(o); }
    cudaStreamEndCapture(st, &g);
    cudaGraphInstantiate(&ge, g, 0);
    printf("graph of 4 x (small%s + big): %.2f us/graph\n", variant ? "(maxshared)" : "", timeit([&] { cudaGraphLaunch(ge, st); }, 200));
  }
  return 0;
}
