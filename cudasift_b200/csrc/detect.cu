// detect.cu -- fused multi-scale blur + DoG + 3x3x3 extrema + sub-pixel refinement.
//
// Behavioural spec: reference LaplaceMultiMem (cudaSiftD.cu:1753-1793, host :460-487) and
// FindPointsMultiNew (cudaSiftD.cu:1292-1431, host :489-514).  The reference writes the
// 7 DoG planes of every octave to global memory (58 MB at 1080p) and reads them back up to
// three times; here one CTA stages an input tile in shared memory, computes the 8 blurred
// scales and the 7 DoG planes of a 64x16 tile entirely on chip and tests the 62x14 interior
// for extrema (in two phases of 4 scales), so the only global traffic is one read of the octave base image and the
// few keypoints found.  One launch covers all octaves (the tile index selects the level).
//
// Per-pixel arithmetic (which products are fused, order of the sums) is pinned to the
// reference's sm_100 SASS with explicit round-to-nearest intrinsics, so DoG values and
// therefore the extrema decisions are bit-identical to the reference's.
//
// Deliberate difference (documented in DESIGN.md): the reference keeps at most 32
// candidates per 30x8x1 block (cudaSiftD.cu:1371,1379); no such cap exists here.
#include "common.cuh"

namespace cs {

#define DT_W 64               // DoG tile width   (62 interior columns)
#define DT_H 16               // DoG tile height  (14 interior rows)
#define DT_IW (DT_W + 8)      // 72 staged input columns
#define DT_IH (DT_H + 8)      // 24 staged input rows
#define DT_THREADS 288           // 9 warps: the vertical pass has 72 x 4 = 288 tasks
#define DT_PH 4                                          // scales blurred per phase (2 phases)
#define DT_SMEM_V (DT_PH * DT_H * DT_IW)                 // floats: vertical results of one phase
#define DT_SMEM_DOG ((CS_LAPLACE_S - 1) * DT_H * DT_W)   // floats: the 7 DoG planes
#define DT_SMEM_IN (DT_IH * DT_IW)                       // floats: staged input tile
#define DT_SMEM_BYTES ((DT_SMEM_V + DT_SMEM_DOG + DT_SMEM_IN) * 4)   // 54016 B -> 4 CTAs per SM

// cudaSiftD.cu:1769-1772 / 1779-1788: sum = k0*c; sum += kj*(x[-j]+x[+j]), j=1..4.
// SASS: FMUL(k1,p1); FFMA(k0,c); FFMA(k2,p2); FFMA(k3,p3); FFMA(k4,p4).
__device__ __forceinline__ float lap_sym9(const float *k, float c, float p1, float p2, float p3, float p4)
{
  float s = __fmul_rn(k[1], p1);
  s = __fmaf_rn(k[0], c, s);
  s = __fmaf_rn(k[2], p2, s);
  s = __fmaf_rn(k[3], p3, s);
  s = __fmaf_rn(k[4], p4, s);
  return s;
}

// Blur the staged tile at 8 scales and leave the 7 DoG planes in s_dog[7][DT_H][DT_W].
// (x0,y0) = image coordinates of DoG element (0,0).
// If s_list != nullptr, every interior pixel whose |DoG| exceeds `thresh` in one of the five
// testable planes is appended to s_list (r*64+d), decided on the register copies of the DoG
// values -- the extrema test then only visits those pixels.
__device__ __forceinline__ void dog_tile(const float *__restrict__ img, int w, int h, int pitch,
                                         int x0, int y0, const LaplaceTaps &taps,
                                         float *s_v, float *s_dog,
                                         float thresh = 0.0f, unsigned short *s_list = nullptr, int *s_cnt = nullptr)
{
  const int tid = threadIdx.x;
  float *s_in = s_dog + DT_SMEM_DOG;

  {
    // 288 threads = 4 rows x 72 columns: thread (r0, c0) loads rows r0, r0+4, ..., r0+20 of its column.
    // Every load is issued before the first store, so the DRAM latencies overlap.
    constexpr int N = DT_IH / 4;
    static_assert(DT_THREADS == 4 * DT_IW && DT_IH % 4 == 0, "load mapping");
    const int r0 = tid / DT_IW, c0 = tid - r0 * DT_IW;
    const float *col = img + min(max(x0 + c0 - 4, 0), w - 1);
    float v[N];
#pragma unroll
    for (int k = 0; k < N; k++) v[k] = __ldg(col + (size_t)min(max(y0 + r0 + 4 * k - 4, 0), h - 1) * pitch);
#pragma unroll
    for (int k = 0; k < N; k++) s_in[(r0 + 4 * k) * DT_IW + c0] = v[k];
  }
  __syncthreads();

  // Two phases of 4 scales each (halves the shared memory of the vertical results, so that 4 CTAs
  // fit on an SM).  Vertical pass: task = (column, group of 4 rows), pair sums shared by the
  // scales.  Horizontal pass + DoG: thread = (row, 4 consecutive columns).
  const int vg = tid / DT_IW, vc = tid - vg * DT_IW;
  const int hr = tid >> 4, hc0 = (tid & 15) * 4;
  float prev[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  float amax[4] = {0.0f, 0.0f, 0.0f, 0.0f};     // max |DoG| over planes 1..5 of this thread's 4 pixels
#pragma unroll
  for (int ph = 0; ph < CS_LAPLACE_S / DT_PH; ph++) {
    {
      float in[12];
#pragma unroll
      for (int i = 0; i < 12; i++) in[i] = s_in[(4 * vg + i) * DT_IW + vc];
#pragma unroll
      for (int rr = 0; rr < 4; rr++) {
        float cc = in[rr + 4];
        float p1 = __fadd_rn(in[rr + 3], in[rr + 5]), p2 = __fadd_rn(in[rr + 2], in[rr + 6]);
        float p3 = __fadd_rn(in[rr + 1], in[rr + 7]), p4 = __fadd_rn(in[rr], in[rr + 8]);
#pragma unroll
        for (int s = 0; s < DT_PH; s++)
          s_v[(s * DT_H + 4 * vg + rr) * DT_IW + vc] = lap_sym9(taps.k[DT_PH * ph + s], cc, p1, p2, p3, p4);
      }
    }
    __syncthreads();
    if (tid < 256) {
#pragma unroll
      for (int s = 0; s < DT_PH; s++) {
        const int sg = DT_PH * ph + s;
        const float4 *p = reinterpret_cast<const float4 *>(&s_v[(s * DT_H + hr) * DT_IW + hc0]);
        float4 a = p[0], b = p[1], c = p[2];
        float v[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
        float o[4];
#pragma unroll
        for (int d = 0; d < 4; d++)
          o[d] = lap_sym9(taps.k[sg], v[d + 4], __fadd_rn(v[d + 3], v[d + 5]), __fadd_rn(v[d + 2], v[d + 6]),
                          __fadd_rn(v[d + 1], v[d + 7]), __fadd_rn(v[d], v[d + 8]));
        if (sg > 0) {
          const float4 dg = make_float4(__fsub_rn(o[0], prev[0]), __fsub_rn(o[1], prev[1]), __fsub_rn(o[2], prev[2]),
                                        __fsub_rn(o[3], prev[3]));
          *reinterpret_cast<float4 *>(&s_dog[((sg - 1) * DT_H + hr) * DT_W + hc0]) = dg;
          if (sg >= 2 && sg <= 6) {             // DoG planes 1..5 are the ones tested for extrema
            amax[0] = fmaxf(amax[0], fabsf(dg.x)); amax[1] = fmaxf(amax[1], fabsf(dg.y));
            amax[2] = fmaxf(amax[2], fabsf(dg.z)); amax[3] = fmaxf(amax[3], fabsf(dg.w));
          }
        }
#pragma unroll
        for (int d = 0; d < 4; d++) prev[d] = o[d];
      }
    }
    if (ph == CS_LAPLACE_S / DT_PH - 1 && s_list != nullptr && tid < 256 && hr >= 1 && hr <= DT_H - 2) {
      // image-border pixels can never be strict extrema in the reference (their clamped
      // neighbour is the pixel itself, cudaSiftD.cu:1308,1331-1332) -> interior only
#pragma unroll
      for (int d = 0; d < 4; d++) {
        const int dd = hc0 + d;
        if (amax[d] > thresh && dd >= 1 && dd <= DT_W - 2 && x0 + dd <= w - 2 && y0 + hr <= h - 2)
          s_list[atomicAdd(s_cnt, 1)] = (unsigned short)(hr * DT_W + dd);
      }
    }
    __syncthreads();
  }
}

// cudaSiftD.cu:1383-1429 on the shared-memory DoG tile.  d1 points at the candidate in
// plane scale+1; planes are DT_H*DT_W apart, rows DT_W apart.
__device__ __noinline__ void refine_and_store(const float *d1, int gx, int gy, int scale,
                                              const DetectLevel &L, const DetectParams &P)
{
  const int PL = DT_H * DT_W, RW = DT_W;
  const float *d0 = d1 - PL, *d2 = d1 + PL;
  float val = d1[0];
  float two = __fadd_rn(val, val);
  float dxx = __fsub_rn(__fsub_rn(two, d1[-1]), d1[1]);
  float dyy = __fsub_rn(__fsub_rn(two, d1[-RW]), d1[RW]);
  float dxy = __fmul_rn(0.25f, __fsub_rn(__fsub_rn(__fadd_rn(d1[RW + 1], d1[-RW - 1]), d1[-RW + 1]), d1[RW - 1]));
  float tra = __fadd_rn(dxx, dyy);
  float det = __fmaf_rn(dxx, dyy, -__fmul_rn(dxy, dxy));
  float tra2 = __fmul_rn(tra, tra);
  if (!(tra2 < __fmul_rn(P.edgeLimit, det))) return;
  float edge = __fdividef(tra2, det);
  float dx = __fmul_rn(0.5f, __fsub_rn(d1[1], d1[-1]));
  float dy = __fmul_rn(0.5f, __fsub_rn(d1[RW], d1[-RW]));
  float ds = __fmul_rn(0.5f, __fsub_rn(d0[0], d2[0]));
  float dss = __fsub_rn(__fsub_rn(two, d2[0]), d0[0]);
  float dxs = __fmul_rn(0.25f, __fsub_rn(__fsub_rn(__fadd_rn(d2[1], d0[-1]), d0[1]), d2[-1]));
  float dys = __fmul_rn(0.25f, __fsub_rn(__fsub_rn(__fadd_rn(d2[RW], d0[-RW]), d2[-RW]), d0[RW]));
  float idxx = __fmaf_rn(dyy, dss, -__fmul_rn(dys, dys));
  float idxy = __fmaf_rn(dys, dxs, -__fmul_rn(dxy, dss));
  float idxs = __fmaf_rn(dxy, dys, -__fmul_rn(dyy, dxs));
  float det3 = __fmaf_rn(idxs, dxs, __fmaf_rn(idxx, dxx, __fmul_rn(idxy, dxy)));
  float idet = __fdividef(1.0f, det3);
  float idyy = __fmaf_rn(dxx, dss, -__fmul_rn(dxs, dxs));
  float idys = __fmaf_rn(dxy, dxs, -__fmul_rn(dxx, dys));
  float idss = det;
  float pdx = __fmul_rn(idet, __fmaf_rn(ds, idxs, __fmaf_rn(dx, idxx, __fmul_rn(dy, idxy))));
  float pdy = __fmul_rn(idet, __fmaf_rn(ds, idys, __fmaf_rn(dy, idyy, __fmul_rn(dx, idxy))));
  float pds = __fmul_rn(idet, __fmaf_rn(idss, ds, __fmaf_rn(dx, idxs, __fmul_rn(dy, idys))));
  if (pdx < -0.5f || pdx > 0.5f || pdy < -0.5f || pdy > 0.5f || pds < -0.5f || pds > 0.5f) {
    pdx = __fdividef(dx, dxx);
    pdy = __fdividef(dy, dyy);
    pds = __fdividef(ds, dss);
  }
  float dsum = __fmaf_rn(ds, pds, __fmaf_rn(dx, pdx, __fmul_rn(dy, pdy)));
  float sc = __fmul_rn(powf(2.0f, __fdiv_rn((float)scale, (float)CS_NUM_SCALES)), exp2f(__fmul_rn(pds, P.factor)));
  if (!(sc >= L.lowestScale)) return;
  unsigned int idx = atomicAdd(&P.counters[0], 1u);
  if (idx >= (unsigned)P.maxPts) idx = P.maxPts - 1;    // cudaSiftD.cu:1421
  SiftPoint *q = P.pts + idx;
  q->xpos = __fadd_rn((float)gx, pdx);
  q->ypos = __fadd_rn((float)gy, pdy);
  q->scale = sc;
  q->sharpness = __fmaf_rn(dsum, 0.5f, val);
  q->edgeness = edge;
  q->subsampling = L.subsampling;
}

__global__ void __launch_bounds__(DT_THREADS, 4)
detect_kernel(const __grid_constant__ DetectParams P)
{
  extern __shared__ __align__(16) float smem[];
  float *s_v = smem;
  float *s_dog = smem + DT_SMEM_V;

  // which level does this tile belong to?  (levels are listed coarsest-last)
  int level = 0;
#pragma unroll 1
  for (int l = 1; l < P.numLevels; l++)
    if ((int)blockIdx.x >= P.lev[l].tileBase) level = l;
  const DetectLevel &L = P.lev[level];
  const int tile = blockIdx.x - L.tileBase;
  const int by = tile / L.tilesX, bx = tile - by * L.tilesX;
  const int x0 = bx * (DT_W - 2), y0 = by * (DT_H - 2);
  const int w = L.w, h = L.h;

  __shared__ unsigned short s_list[(DT_H - 2) * (DT_W - 2)];
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  const float thresh = P.thresh;
  dog_tile(L.img, w, h, L.pitch, x0, y0, L.taps, s_v, s_dog, thresh, s_list, &s_cnt);

  // 3x3x3 extrema, only on the pixels the blur pass flagged (|DoG| > thresh at some scale)
  const int ncand = s_cnt;
  for (int i = threadIdx.x; i < ncand; i += DT_THREADS) {
    const int rd = s_list[i];
    const int r = rd / DT_W, d = rd - r * DT_W;
    const int gx = x0 + d, gy = y0 + r;
    const float *c = s_dog + rd;
#pragma unroll 1
    for (int sc = 0; sc < CS_NUM_SCALES; sc++) {
      const float *d1 = c + (sc + 1) * (DT_H * DT_W);
      float v = d1[0];
      if (!(fabsf(v) > thresh)) continue;
      bool ext = true;
      if (v > 0.0f) {
#pragma unroll
        for (int pl = -1; pl <= 1; pl++)
#pragma unroll
          for (int dy = -1; dy <= 1; dy++)
#pragma unroll
            for (int dx = -1; dx <= 1; dx++)
              if (pl != 0 || dy != 0 || dx != 0)
                ext = ext && (v > d1[pl * (DT_H * DT_W) + dy * DT_W + dx]);
      } else {
#pragma unroll
        for (int pl = -1; pl <= 1; pl++)
#pragma unroll
          for (int dy = -1; dy <= 1; dy++)
#pragma unroll
            for (int dx = -1; dx <= 1; dx++)
              if (pl != 0 || dy != 0 || dx != 0)
                ext = ext && (v < d1[pl * (DT_H * DT_W) + dy * DT_W + dx]);
      }
      if (ext) refine_and_store(d1, gx, gy, sc, L, P);
    }
  }
}

int launch_detect(const DetectParams &p, cudaStream_t st)
{
  static bool configured = false;
  if (!configured) {
    CS_CUDA(cudaFuncSetAttribute(detect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SMEM_BYTES));
    configured = true;
  }
  if (p.totalTiles <= 0) return 0;
  detect_kernel<<<p.totalTiles, DT_THREADS, DT_SMEM_BYTES, st>>>(p);
  count_launch();
  CS_CUDA(cudaGetLastError());
  return 0;
}

// Stage-level entry point (parity tests): materialise the 7 DoG planes exactly as the
// fused detector computes them.  Plane stride = h*pitch floats (cudaSiftH.cu:184).
__global__ void __launch_bounds__(DT_THREADS)
dog_planes_kernel(const float *__restrict__ img, float *__restrict__ dog, int w, int h, int pitch,
                  int tilesX, const __grid_constant__ LaplaceTaps taps)
{
  extern __shared__ __align__(16) float smem[];
  float *s_v = smem;
  float *s_dog = smem + DT_SMEM_V;
  const int by = blockIdx.x / tilesX, bx = blockIdx.x - by * tilesX;
  const int x0 = bx * (DT_W - 2), y0 = by * (DT_H - 2);
  dog_tile(img, w, h, pitch, x0, y0, taps, s_v, s_dog);
  const size_t plane = (size_t)h * pitch;
  for (int i = threadIdx.x; i < (CS_LAPLACE_S - 1) * DT_H * DT_W; i += DT_THREADS) {
    int s = i / (DT_H * DT_W), rem = i - s * (DT_H * DT_W);
    int r = rem / DT_W, d = rem - r * DT_W;
    int gx = x0 + d, gy = y0 + r;
    if (gx < w && gy < h) dog[s * plane + (size_t)gy * pitch + gx] = s_dog[i];
  }
}

int launch_dog_planes(const float *base, float *dog, int w, int h, int pitch, const LaplaceTaps &taps, cudaStream_t st)
{
  static bool configured = false;
  if (!configured) {
    CS_CUDA(cudaFuncSetAttribute(dog_planes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SMEM_BYTES));
    configured = true;
  }
  int tilesX = idivup(w, DT_W - 2), tilesY = idivup(h, DT_H - 2);
  dog_planes_kernel<<<tilesX * tilesY, DT_THREADS, DT_SMEM_BYTES, st>>>(base, dog, w, h, pitch, tilesX, taps);
  count_launch();
  CS_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace cs
