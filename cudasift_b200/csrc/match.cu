// match.cu -- exact FP32 brute-force matcher (the reference-arithmetic path).
//
// Behavioural spec: reference FindMaxCorr10 (matching.cu:301-397) + MatchSiftData
// (matching.cu:1090-1206).  This kernel reproduces the reference's results bit for bit:
//   * score(p1,p2) is the sequential k = 0..127 FMA chain starting from 0 (the reference's
//     SASS is a pure FFMA chain in exactly that order);
//   * a reference thread owns (row, partition) with partition = ((p2 mod 32) div 4) and
//     keeps max / second / index with strict '>' updates in increasing p2 order;
//   * only p2 < 32*floor(n2/32) are visited (quirk Q7);
//   * the final merge over the 8 partitions ignores the per-partition second scores of
//     partitions 1..7 (quirk Q9) and breaks ties towards the lowest partition (Q10).
// It is the fallback / verification path of the tensor-core matcher (match_tc.cu), used
// for small sets and for inputs the FP16 screening cannot bound.
//
// Differences from the reference kept on purpose: rows >= n1 are never written (the
// reference stores up to 31 records past the array, Q8); with no visited candidate
// (match == -1) match_xpos/ypos are set to 0 instead of reading sift2[-1] (Q7).
#include "common.cuh"

namespace cs {

#define MX_ROWS 32          // rows of set 1 per CTA
#define MX_COLS 32          // candidates of set 2 per step (the reference's block size)
#define MX_LD 33            // row stride in float4 (+1 float4 padding: conflict-free LDS.128)
#define MX_THREADS 128

struct PartState { float mx, sec; int idx; };

__device__ __forceinline__ void part_update(PartState &s, float sc, int p2)
{ // matching.cu:354-359
  if (sc > s.mx) { s.sec = s.mx; s.mx = sc; s.idx = p2; }
  else if (sc > s.sec) s.sec = sc;
}

// rows == nullptr: the CTAs handle rows [32*b, +32) of set 1, b grid-strided.  rows != nullptr: they
// handle entries [32*b, +32) of the row list (*nrows entries) -- the fallback of the tensor-core
// path for rows it could not certify.
__global__ void __launch_bounds__(MX_THREADS)
match_exact_kernel(SiftPoint *__restrict__ sift1, const SiftPoint *__restrict__ sift2, int n1, int n2,
                   const int *__restrict__ rows, const unsigned int *__restrict__ nrows,
                   const unsigned int *__restrict__ gate)
{
  // device-side switch (tensor path): *gate != 0 = the inputs were out of range for the split-FP16 screening,
  // scan every row exactly; otherwise only the rows of the list
  if (gate) {
    if (*gate != 0) rows = nullptr;
    else if (!rows) return;
  }
  __shared__ float4 s_a[MX_ROWS * MX_LD];
  __shared__ float4 s_b[MX_COLS * MX_LD];
  const int tid = threadIdx.x;
  const int rq = tid & 15;       // rows rq and rq+16 of the CTA's 32
  const int part = tid >> 4;     // partition 0..7 -> candidates 4*part..4*part+3 of each block
  if (rows) n1 = (int)*nrows;
 for (int bp1 = blockIdx.x * MX_ROWS; bp1 < n1; bp1 += gridDim.x * MX_ROWS) {
  __syncthreads();
  for (int i = tid; i < MX_ROWS * 32; i += MX_THREADS) {
    int r = i >> 5, d = i & 31;
    int p1 = min(bp1 + r, n1 - 1);
    if (rows) p1 = rows[p1];
    s_a[r * MX_LD + d] = reinterpret_cast<const float4 *>(sift1[p1].data)[d];
  }
  PartState st[2];
  st[0].mx = st[1].mx = 0.0f; st[0].sec = st[1].sec = 0.0f; st[0].idx = st[1].idx = -1;

  const int nblk = n2 / MX_COLS;    // matching.cu:325
  for (int b = 0; b < nblk; b++) {
    __syncthreads();
    for (int i = tid; i < MX_COLS * 32; i += MX_THREADS) {
      int r = i >> 5, d = i & 31;
      s_b[r * MX_LD + d] = __ldg(reinterpret_cast<const float4 *>(sift2[b * MX_COLS + r].data) + d);
    }
    __syncthreads();
    float acc[2][4];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[i][j] = 0.0f;
#pragma unroll 4
    for (int d = 0; d < 32; d++) {
      float4 a0 = s_a[rq * MX_LD + d], a1 = s_a[(rq + 16) * MX_LD + d];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        float4 bv = s_b[(4 * part + j) * MX_LD + d];
        acc[0][j] = __fmaf_rn(a0.x, bv.x, acc[0][j]);
        acc[0][j] = __fmaf_rn(a0.y, bv.y, acc[0][j]);
        acc[0][j] = __fmaf_rn(a0.z, bv.z, acc[0][j]);
        acc[0][j] = __fmaf_rn(a0.w, bv.w, acc[0][j]);
        acc[1][j] = __fmaf_rn(a1.x, bv.x, acc[1][j]);
        acc[1][j] = __fmaf_rn(a1.y, bv.y, acc[1][j]);
        acc[1][j] = __fmaf_rn(a1.z, bv.z, acc[1][j]);
        acc[1][j] = __fmaf_rn(a1.w, bv.w, acc[1][j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int p2 = b * MX_COLS + 4 * part + j;
      part_update(st[0], acc[0][j], p2);
      part_update(st[1], acc[1][j], p2);
    }
  }
  __syncthreads();
  // publish per-(row, partition) states, then one thread per row merges (matching.cu:366-396)
  float *m_mx = reinterpret_cast<float *>(s_b);
  float *m_sec = m_mx + MX_ROWS * 8;
  int *m_idx = reinterpret_cast<int *>(m_sec + MX_ROWS * 8);
#pragma unroll
  for (int i = 0; i < 2; i++) {
    int r = rq + 16 * i;
    m_mx[part * MX_ROWS + r] = st[i].mx;
    m_sec[part * MX_ROWS + r] = st[i].sec;
    m_idx[part * MX_ROWS + r] = st[i].idx;
  }
  __syncthreads();
  if (tid < MX_ROWS && bp1 + tid < n1) {
    float mx = m_mx[tid], sec = m_sec[tid];
    int idx = m_idx[tid];
    for (int y = 0; y < 8; y++) {
      int iy = m_idx[y * MX_ROWS + tid];
      float my = m_mx[y * MX_ROWS + tid];
      if (idx != iy) {
        if (my > mx) { sec = fmaxf(mx, sec); mx = my; idx = iy; }
        else if (my > sec) sec = my;
      }
    }
    SiftPoint *o = sift1 + (rows ? rows[bp1 + tid] : bp1 + tid);
    o->score = mx;
    o->match = idx;
    o->match_xpos = idx >= 0 ? sift2[idx].xpos : 0.0f;
    o->match_ypos = idx >= 0 ? sift2[idx].ypos : 0.0f;
    o->ambiguity = __fdiv_rn(sec, __fadd_rn(mx, 1e-6f));
  }
 }
}

int match_exact(SiftPoint *s1, int n1, const SiftPoint *s2, int n2, cudaStream_t st)
{
  if (n1 <= 0) return 0;
  match_exact_kernel<<<idivup(n1, MX_ROWS), MX_THREADS, 0, st>>>(s1, s2, n1, n2, nullptr, nullptr, nullptr);
  count_launch();
  CS_CUDA(cudaGetLastError());
  return 0;
}

// Fallback of the tensor path in one launch, decided on the device (no host round trip): the listed
// rows, or every row if *gate != 0.
int match_exact_fallback(SiftPoint *s1, int n1, const SiftPoint *s2, int n2, const int *rows, const unsigned int *nrows,
                         const unsigned int *gate, cudaStream_t st)
{
  if (n1 <= 0) return 0;
  const int blocks = idivup(n1, MX_ROWS) < 296 ? idivup(n1, MX_ROWS) : 296;    // grid-stride over the rows
  match_exact_kernel<<<blocks, MX_THREADS, 0, st>>>(s1, s2, n1, n2, rows, nrows, gate);
  count_launch();
  CS_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace cs
