// homography.cu -- FindHomography: RANSAC over the matches MatchSiftData produced.
//
// Behavioural spec: reference matching.cu:1000-1087 (host), :907-948 (ComputeHomographies),
// :953-996 (TestHomographies), :821-905 (8x8 inverse by LU decomposition with implicit pivoting).
// Same contract: points with score > minScore and ambiguity < maxAmbiguity are eligible; numLoops
// (rounded up to 16) 4-point samples are drawn with the C library's rand() in the reference's call
// order (so a caller that seeds srand() gets the same samples from both libraries); every
// hypothesis is scored by its number of inliers among ALL points (reprojection error < thresh,
// products rounded toward zero as in the reference); the first hypothesis with the highest count
// wins; homography[0..7] is written (homography[8] stays 1), *numMatches = inlier count.
//
// Structure (new): one gather kernel turns the AoS records into SoA coordinates + (score,
// ambiguity) pairs (one dense D2H copy instead of two strided ones), one thread per hypothesis
// solves the 8x8 system, one warp per hypothesis counts inliers, and the arg-max is taken on the
// device with a packed atomicMax -- 3 launches and 2 small copies, no cudaMalloc per call.
// Difference kept on purpose: the reference also "tests" the up-to-15 uninitialised padding
// entries behind numPts (it loops to the next multiple of 16); here only real points count.
#include "common.cuh"

#include <vector>

namespace cs {

__global__ void hg_gather_kernel(const SiftPoint *__restrict__ s, int n, float *__restrict__ coord, int stride,
                                 float2 *__restrict__ sa)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  coord[i] = s[i].xpos;
  coord[i + stride] = s[i].ypos;
  coord[i + 2 * stride] = s[i].match_xpos;
  coord[i + 3 * stride] = s[i].match_ypos;
  sa[i] = make_float2(s[i].score, s[i].ambiguity);
}

// LU decomposition with implicit pivoting + back substitution of the 8 unit vectors
// (matching.cu:821-905; the classic ludcmp/lubksb pair).  a is destroyed, inv receives a^-1.
__device__ void invert8(float (&a)[8][8], float (&inv)[8][8])
{
  int indx[8];
  float vv[8];
  int imax = 0;
  for (int i = 0; i < 8; i++) {
    float big = 0.0f;
    for (int j = 0; j < 8; j++) big = fmaxf(big, fabsf(a[i][j]));
    vv[i] = (big > 0.0f) ? (float)(1.0 / (double)big) : 1e16f;
  }
  for (int j = 0; j < 8; j++) {
    for (int i = 0; i < j; i++) {
      float sum = a[i][j];
      for (int k = 0; k < i; k++) sum = __fmaf_rn(-a[i][k], a[k][j], sum);
      a[i][j] = sum;
    }
    float big = 0.0f;
    for (int i = j; i < 8; i++) {
      float sum = a[i][j];
      for (int k = 0; k < j; k++) sum = __fmaf_rn(-a[i][k], a[k][j], sum);
      a[i][j] = sum;
      float dum = __fmul_rn(vv[i], fabsf(sum));
      if (dum >= big) { big = dum; imax = i; }
    }
    if (j != imax) {
      for (int k = 0; k < 8; k++) { float t = a[imax][k]; a[imax][k] = a[j][k]; a[j][k] = t; }
      vv[imax] = vv[j];
    }
    indx[j] = imax;
    if (a[j][j] == 0.0f) a[j][j] = 1e-16f;
    if (j != 7) {
      float dum = (float)(1.0 / (double)a[j][j]);
      for (int i = j + 1; i < 8; i++) a[i][j] = __fmul_rn(a[i][j], dum);
    }
  }
  for (int j = 0; j < 8; j++) {
    float b[8];
    for (int k = 0; k < 8; k++) b[k] = 0.0f;
    b[j] = 1.0f;
    int ii = -1;
    for (int i = 0; i < 8; i++) {
      int ip = indx[i];
      float sum = b[ip];
      b[ip] = b[i];
      if (ii != -1) {
        for (int k = ii; k < i; k++) sum = __fmaf_rn(-a[i][k], b[k], sum);
      } else if (sum != 0.0f) {
        ii = i;
      }
      b[i] = sum;
    }
    for (int i = 7; i >= 0; i--) {
      float sum = b[i];
      for (int k = i + 1; k < 8; k++) sum = __fmaf_rn(-a[i][k], b[k], sum);
      b[i] = __fdiv_rn(sum, a[i][i]);
    }
    for (int i = 0; i < 8; i++) inv[i][j] = b[i];
  }
}

// one thread per hypothesis (matching.cu:907-948)
__global__ void __launch_bounds__(64)
hg_compute_kernel(const float *__restrict__ coord, int stride, const int *__restrict__ randPts,
                  float *__restrict__ homo, int numLoops)
{
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= numLoops) return;
  float a[8][8], ia[8][8], b[8];
  for (int i = 0; i < 4; i++) {
    int pt = randPts[i * numLoops + idx];
    float x1 = coord[pt], y1 = coord[pt + stride], x2 = coord[pt + 2 * stride], y2 = coord[pt + 3 * stride];
    float *r1 = a[2 * i], *r2 = a[2 * i + 1];
    r1[0] = x1; r1[1] = y1; r1[2] = 1.0f; r1[3] = r1[4] = r1[5] = 0.0f;
    r1[6] = __fmul_rn(-x2, x1); r1[7] = __fmul_rn(-x2, y1);
    r2[0] = r2[1] = r2[2] = 0.0f; r2[3] = x1; r2[4] = y1; r2[5] = 1.0f;
    r2[6] = __fmul_rn(-y2, x1); r2[7] = __fmul_rn(-y2, y1);
    b[2 * i] = x2; b[2 * i + 1] = y2;
  }
  invert8(a, ia);
  for (int j = 0; j < 8; j++) {
    float sum = 0.0f;
    for (int i = 0; i < 8; i++) sum = __fmaf_rn(ia[j][i], b[i], sum);
    homo[j * numLoops + idx] = sum;
  }
}

// one warp per hypothesis (matching.cu:953-996): inlier count over all points, then a packed
// atomicMax (count in the high word, ~index in the low word: the FIRST best hypothesis wins,
// as in the host loop matching.cu:1068-1073)
__global__ void __launch_bounds__(256)
hg_test_kernel(const float *__restrict__ coord, int stride, int numPts, const float *__restrict__ homo,
               int numLoops, float thresh2, unsigned long long *__restrict__ best)
{
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= numLoops) return;
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = homo[i * numLoops + warp];
  int cnt = 0;
  for (int i = lane; i < numPts; i += 32) {
    float x1 = coord[i], y1 = coord[i + stride], x2 = coord[i + 2 * stride], y2 = coord[i + 3 * stride];
    float nomx = __fadd_rn(__fadd_rn(__fmul_rz(a[0], x1), __fmul_rz(a[1], y1)), a[2]);
    float nomy = __fadd_rn(__fadd_rn(__fmul_rz(a[3], x1), __fmul_rz(a[4], y1)), a[5]);
    float deno = __fadd_rn(__fadd_rn(__fmul_rz(a[6], x1), __fmul_rz(a[7], y1)), 1.0f);
    float errx = __fsub_rn(__fmul_rz(x2, deno), nomx);
    float erry = __fsub_rn(__fmul_rz(y2, deno), nomy);
    float err2 = __fadd_rn(__fmul_rz(errx, errx), __fmul_rz(erry, erry));
    if (err2 < __fmul_rz(thresh2, __fmul_rz(deno, deno))) cnt++;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0)
    atomicMax(best, ((unsigned long long)(unsigned)cnt << 32) | (unsigned long long)(0xffffffffu - (unsigned)warp));
}

}  // namespace cs

using namespace cs;

double FindHomography(SiftData &data, float *homography, int *numMatches, int numLoops, float minScore,
                      float maxAmbiguity, float thresh)
{
  *numMatches = 0;
  homography[0] = homography[4] = homography[8] = 1.0f;
  homography[1] = homography[2] = homography[3] = 0.0f;
  homography[5] = homography[6] = homography[7] = 0.0f;
#ifdef MANAGEDMEM
  SiftPoint *d_sift = data.m_data;
#else
  if (data.d_data == NULL) return 0.0;
  SiftPoint *d_sift = data.d_data;
#endif
  numLoops = iDivUp(numLoops, 16) * 16;
  const int numPts = data.numPts;
  if (numPts < 8) return 0.0;                                   // matching.cu:1016-1017
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, 0);
  const int stride = iDivUp(numPts, 16) * 16;
  float *d_coord = nullptr, *d_homo = nullptr;
  float2 *d_sa = nullptr;
  int *d_rand = nullptr;
  unsigned long long *d_best = nullptr;
  // one allocation for all scratch
  size_t off_sa = sizeof(float) * 4 * (size_t)stride;
  size_t off_rand = off_sa + sizeof(float2) * (size_t)numPts;
  size_t off_homo = off_rand + sizeof(int) * 4 * (size_t)numLoops;
  size_t off_best = off_homo + sizeof(float) * 8 * (size_t)numLoops;
  size_t total = off_best + 16;
  char *scratch = nullptr;
  if (cudaMalloc((void **)&scratch, total) != cudaSuccess) {
    fprintf(stderr, "cudasift_b200: FindHomography: cudaMalloc failed\n");
    exit(-1);
  }
  d_coord = (float *)scratch; d_sa = (float2 *)(scratch + off_sa); d_rand = (int *)(scratch + off_rand);
  d_homo = (float *)(scratch + off_homo); d_best = (unsigned long long *)(scratch + off_best);
  hg_gather_kernel<<<idivup(numPts, 256), 256>>>(d_sift, numPts, d_coord, stride, d_sa);
  std::vector<float2> h_sa(numPts);
  cudaMemcpy(h_sa.data(), d_sa, sizeof(float2) * numPts, cudaMemcpyDeviceToHost);
  std::vector<int> valid;
  valid.reserve(numPts);
  for (int i = 0; i < numPts; i++)
    if (h_sa[i].x > minScore && h_sa[i].y < maxAmbiguity) valid.push_back(i);   // matching.cu:1034-1037
  const int numValid = (int)valid.size();
  if (numValid >= 8) {
    std::vector<int> h_rand(4 * (size_t)numLoops);
    for (int i = 0; i < numLoops; i++) {                          // matching.cu:1041-1053, same rand() order
      int p1 = rand() % numValid;
      int p2 = rand() % numValid;
      int p3 = rand() % numValid;
      int p4 = rand() % numValid;
      while (p2 == p1) p2 = rand() % numValid;
      while (p3 == p1 || p3 == p2) p3 = rand() % numValid;
      while (p4 == p1 || p4 == p2 || p4 == p3) p4 = rand() % numValid;
      h_rand[i + 0 * numLoops] = valid[p1];
      h_rand[i + 1 * numLoops] = valid[p2];
      h_rand[i + 2 * numLoops] = valid[p3];
      h_rand[i + 3 * numLoops] = valid[p4];
    }
    cudaMemcpy(d_rand, h_rand.data(), sizeof(int) * 4 * numLoops, cudaMemcpyHostToDevice);
    cudaMemset(d_best, 0, sizeof(unsigned long long));
    hg_compute_kernel<<<idivup(numLoops, 64), 64>>>(d_coord, stride, d_rand, d_homo, numLoops);
    hg_test_kernel<<<idivup(numLoops * 32, 256), 256>>>(d_coord, stride, numPts, d_homo, numLoops, thresh * thresh, d_best);
    count_launch(3);
    unsigned long long best = 0;
    cudaMemcpy(&best, d_best, sizeof(best), cudaMemcpyDeviceToHost);
    const int maxCount = (int)(best >> 32);
    const int maxIndex = (int)(0xffffffffu - (unsigned)(best & 0xffffffffu));
    *numMatches = maxCount;
    cudaMemcpy2D(homography, sizeof(float), &d_homo[maxIndex], sizeof(float) * numLoops, sizeof(float), 8,
                 cudaMemcpyDeviceToHost);                           // matching.cu:1075
  } else {
    count_launch(1);
  }
  cudaFree(scratch);
  cudaEventRecord(e1, 0);
  cudaEventSynchronize(e1);
  float ms = 0.0f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    fprintf(stderr, "cudasift_b200: FindHomography failed: %s\n", cudaGetErrorString(err));
    exit(-1);
  }
  return ms;
}

extern "C" int cs_find_homography(void *d_pts, int numPts, float *homography, int *numMatches, int numLoops,
                                  float minScore, float maxAmbiguity, float thresh, double *ms)
{
  SiftData sd;
  sd.numPts = numPts; sd.maxPts = numPts;
#ifdef MANAGEDMEM
  sd.m_data = (SiftPoint *)d_pts;
#else
  sd.h_data = NULL; sd.d_data = (SiftPoint *)d_pts;
#endif
  double t = FindHomography(sd, homography, numMatches, numLoops, minScore, maxAmbiguity, thresh);
  if (ms) *ms = t;
  return 0;
}
