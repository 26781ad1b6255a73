// detect.cu -- fused multi-scale blur + DoG + 3x3x3 extrema + sub-pixel refinement.
//
// Behavioural spec: reference LaplaceMultiMem (cudaSiftD.cu:1753-1793, host :460-487) and
// FindPointsMultiNew (cudaSiftD.cu:1292-1431, host :489-514).  The reference writes the
// 7 DoG planes of every octave to global memory (58 MB at 1080p) and reads them back up to
// three times; here one CTA stages an input tile in shared memory, computes the 8 blurred
// scales and the 7 DoG planes of a 64x16 tile entirely on chip and tests the 62x14 interior
// for extrema, so the only global traffic is one read of the octave base image and the few
// keypoints found.  One launch covers all octaves (the tile index selects the level).
//
// Blackwell specifics: all blur arithmetic is issued as packed FP32 pairs (fma.rn.f32x2 /
// add.rn.f32x2 -> FFMA2 / FADD2), which halves the issue slots of the ~116 FP32 operations
// per pixel that parity with the reference fixes.  A pair is (row r, row r+8) of the same
// column: the staged input, the vertical results and the DoG planes all live in shared memory
// as float2 {row r, row r+8}, so every 64-bit shared load is already an aligned operand pair
// and no register shuffling is needed.  Taps are duplicated (k,k) in the kernel parameters and
// consumed straight from uniform registers.
//
// Per-pixel arithmetic (which products are fused, order of the sums) is pinned to the
// reference's sm_100 SASS with explicit round-to-nearest operations (the packed forms round
// each half exactly like the scalar ones), so DoG values and therefore the extrema decisions
// are bit-identical to the reference's.
//
// Deliberate difference (documented in DESIGN.md): the reference keeps at most 32
// candidates per 30x8x1 block (cudaSiftD.cu:1371,1379); no such cap exists here.
#include "common.cuh"

namespace cs {

#define DT_W 64               // DoG tile width   (62 interior columns)
#define DT_H 16               // DoG tile height  (14 interior rows)
#define DT_HP (DT_H / 2)      // row pairs (r, r+8)
#define DT_IW (DT_W + 8)      // 72 staged input columns
#define DT_IH (DT_H + 8)      // 24 staged input rows
#define DT_THREADS 160        // 5 warps: the vertical pass has 72 x 2 = 144 tasks
#define DT_PH 4               // scales blurred per phase (2 phases)
#define DT_NC 8               // DoG columns per thread in the horizontal pass
#define DT_VS 74              // float2 stride of a vertical-result row (37 16-byte chunks: odd -> the 8 row
                              // pairs of a quarter warp hit 8 different bank groups)
// shared memory, in float2 units
#define DT_SM_IN (2 * DT_HP * DT_IW)                  // staged input pairs P[j] = (in[j], in[j+8]), j = 0..15
#define DT_SM_V (DT_PH * DT_HP * DT_VS)               // vertical results of one phase
#define DT_SM_DOG ((CS_LAPLACE_S - 1) * DT_HP * DT_W) // 7 DoG planes, 16-byte chunks XOR-swizzled by the row pair
#define DT_SMEM_BYTES ((DT_SM_IN + DT_SM_V + DT_SM_DOG) * 8)   // 56832 B -> 4 CTAs per SM

typedef unsigned long long f32x2;   // two packed floats: lo = row r, hi = row r+8

__device__ __forceinline__ f32x2 pk(float2 v) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y)); return r; }
__device__ __forceinline__ float2 upk(f32x2 v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// cudaSiftD.cu:1769-1772 / 1779-1788: sum = k0*c; sum += kj*(x[-j]+x[+j]), j=1..4.
// SASS: FMUL(k1,p1); FFMA(k0,c); FFMA(k2,p2); FFMA(k3,p3); FFMA(k4,p4)  -- here on both halves at once.
__device__ __forceinline__ f32x2 lap_sym9(const float2 *k, f32x2 c, f32x2 p1, f32x2 p2, f32x2 p3, f32x2 p4)
{
  f32x2 s = mul2(pk(k[1]), p1);
  s = fma2(pk(k[0]), c, s);
  s = fma2(pk(k[2]), p2, s);
  s = fma2(pk(k[3]), p3, s);
  s = fma2(pk(k[4]), p4, s);
  return s;
}

// float index of DoG element (plane, row, col) in the swizzled pair layout
__device__ __forceinline__ int dog_index(int plane, int row, int col)
{
  const int rr = row & (DT_HP - 1);
  return (((plane * DT_HP + rr) * (DT_W / 2) + ((col >> 1) ^ rr)) << 2) + ((col & 1) << 1) + (row >> 3);
}

// Blur the tile at 8 scales and leave the 7 DoG planes in s_dog (pair layout, see dog_index).
// (x0,y0) = image coordinates of DoG element (0,0).
// If s_list != nullptr, every interior pixel whose |DoG| exceeds `thresh` in one of the five
// testable planes is appended to s_list (row*64+col), decided on the register copies of the DoG
// values -- the extrema test then only visits those pixels.  s_list may alias s_in.
__device__ __forceinline__ void dog_tile(const float *__restrict__ img, int w, int h, int pitch,
                                         int x0, int y0, const LaplaceTaps &taps,
                                         float2 *s_in, float2 *s_v, float2 *s_dog,
                                         float thresh = 0.0f, unsigned short *s_list = nullptr, int *s_cnt = nullptr)
{
  const int tid = threadIdx.x;

  {
    // Task (column c, g = 0..7) loads rows g, g+8, g+16 of its column and stores the pairs P[g] and P[g+8].
    // Every load is issued before the first store, so the DRAM latencies overlap.
    constexpr int TASKS = DT_IW * DT_HP;
    constexpr int N = (TASKS + DT_THREADS - 1) / DT_THREADS;
    float a[N], b[N], c[N];
#pragma unroll
    for (int k = 0; k < N; k++) {
      const int t = tid + k * DT_THREADS;
      if (t < TASKS) {
        const int g = t / DT_IW, cc = t - g * DT_IW;
        const float *col = img + min(max(x0 + cc - 4, 0), w - 1);
        a[k] = __ldg(col + (size_t)min(max(y0 + g - 4, 0), h - 1) * pitch);
        b[k] = __ldg(col + (size_t)min(max(y0 + g + 4, 0), h - 1) * pitch);
        c[k] = __ldg(col + (size_t)min(max(y0 + g + 12, 0), h - 1) * pitch);
      }
    }
#pragma unroll
    for (int k = 0; k < N; k++) {
      const int t = tid + k * DT_THREADS;
      if (t < TASKS) {
        s_in[t] = make_float2(a[k], b[k]);                       // t = g*72 + cc
        s_in[t + DT_HP * DT_IW] = make_float2(b[k], c[k]);
      }
    }
  }
  __syncthreads();

  // Two phases of 4 scales each.  Vertical pass: task = (column, 4 row pairs), pair sums shared by the
  // scales.  Horizontal pass + DoG: task = (row pair, DT_NC consecutive columns); the 8 lanes of a
  // quarter warp take the 8 row pairs so that 128-bit shared accesses are conflict free.
  const int vh = tid / DT_IW, vc = tid - vh * DT_IW;           // vertical task (tid < 144)
  const int lane = tid & 31, warp = tid >> 5;
  const int hr = lane & 7, hg = (lane >> 3) + 4 * warp;         // horizontal task (hg < 64 / DT_NC)
  const bool hact = hg < DT_W / DT_NC;
  f32x2 prev[DT_NC];
  float2 amax[DT_NC];                                           // max |DoG| over planes 1..5
#pragma unroll
  for (int d = 0; d < DT_NC; d++) { prev[d] = 0ull; amax[d] = make_float2(0.0f, 0.0f); }

#pragma unroll
  for (int ph = 0; ph < CS_LAPLACE_S / DT_PH; ph++) {
    if (tid < 2 * DT_IW) {
      f32x2 in[12];
#pragma unroll
      for (int i = 0; i < 12; i++) in[i] = pk(s_in[(4 * vh + i) * DT_IW + vc]);
#pragma unroll
      for (int rr = 0; rr < 4; rr++) {
        const f32x2 cc = in[rr + 4];
        const f32x2 p1 = add2(in[rr + 3], in[rr + 5]), p2 = add2(in[rr + 2], in[rr + 6]);
        const f32x2 p3 = add2(in[rr + 1], in[rr + 7]), p4 = add2(in[rr], in[rr + 8]);
#pragma unroll
        for (int s = 0; s < DT_PH; s++)
          s_v[(s * DT_HP + 4 * vh + rr) * DT_VS + vc] = upk(lap_sym9(taps.k[DT_PH * ph + s], cc, p1, p2, p3, p4));
      }
    }
    __syncthreads();
    if (hact) {
#pragma unroll
      for (int s = 0; s < DT_PH; s++) {
        const int sg = DT_PH * ph + s;
        const float4 *p = reinterpret_cast<const float4 *>(&s_v[(s * DT_HP + hr) * DT_VS + DT_NC * hg]);
        f32x2 v[DT_NC + 8];
#pragma unroll
        for (int k = 0; k < (DT_NC + 8) / 2; k++) {
          const float4 q = p[k];
          v[2 * k] = pk(make_float2(q.x, q.y));
          v[2 * k + 1] = pk(make_float2(q.z, q.w));
        }
        f32x2 o[DT_NC];
#pragma unroll
        for (int d = 0; d < DT_NC; d++)
          o[d] = lap_sym9(taps.k[sg], v[d + 4], add2(v[d + 3], v[d + 5]), add2(v[d + 2], v[d + 6]),
                          add2(v[d + 1], v[d + 7]), add2(v[d], v[d + 8]));
        if (sg > 0) {
          float4 *out = reinterpret_cast<float4 *>(s_dog) + ((sg - 1) * DT_HP + hr) * (DT_W / 2);
#pragma unroll
          for (int k = 0; k < DT_NC / 2; k++) {
            const float2 d0 = upk(sub2(o[2 * k], prev[2 * k])), d1 = upk(sub2(o[2 * k + 1], prev[2 * k + 1]));
            out[((DT_NC / 2) * hg + k) ^ hr] = make_float4(d0.x, d0.y, d1.x, d1.y);
            if (sg >= 2 && sg <= 6) {             // DoG planes 1..5 are the ones tested for extrema
              amax[2 * k].x = fmaxf(amax[2 * k].x, fabsf(d0.x)); amax[2 * k].y = fmaxf(amax[2 * k].y, fabsf(d0.y));
              amax[2 * k + 1].x = fmaxf(amax[2 * k + 1].x, fabsf(d1.x)); amax[2 * k + 1].y = fmaxf(amax[2 * k + 1].y, fabsf(d1.y));
            }
          }
        }
#pragma unroll
        for (int d = 0; d < DT_NC; d++) prev[d] = o[d];
      }
    }
    if (ph == CS_LAPLACE_S / DT_PH - 1 && s_list != nullptr && hact) {
      // image-border pixels can never be strict extrema in the reference (their clamped
      // neighbour is the pixel itself, cudaSiftD.cu:1308,1331-1332) -> interior only.
      // (s_list aliases s_in: the last vertical pass is behind the barrier above.)
      unsigned int mask = 0;                      // bit d: row hr, bit 16+d: row hr+8
#pragma unroll
      for (int d = 0; d < DT_NC; d++) {
        const int dd = DT_NC * hg + d;
        if (dd >= 1 && dd <= DT_W - 2 && x0 + dd <= w - 2) {
          if (amax[d].x > thresh) mask |= 1u << d;
          if (amax[d].y > thresh) mask |= 0x10000u << d;
        }
      }
      if (!(hr >= 1 && y0 + hr <= h - 2)) mask &= 0xffff0000u;
      if (!(hr <= DT_HP - 2 && y0 + hr + DT_HP <= h - 2)) mask &= 0x0000ffffu;
      if (mask) {
        int at = atomicAdd(s_cnt, __popc(mask));
        while (mask) {
          const int b = __ffs(mask) - 1;
          mask &= mask - 1;
          s_list[at++] = (unsigned short)((hr + (b >> 4) * DT_HP) * DT_W + DT_NC * hg + (b & 15));
        }
      }
    }
    __syncthreads();
  }
}

// cudaSiftD.cu:1383-1429 on the 3x3x3 neighbourhood v[plane][row][col] around the candidate
// (v[1][1][1]; plane index = scale, scale+1, scale+2).
__device__ __noinline__ void refine_and_store(const float (&v)[3][3][3], int gx, int gy, int scale,
                                              const DetectLevel &L, const DetectParams &P)
{
  const float val = v[1][1][1];
  float two = __fadd_rn(val, val);
  float dxx = __fsub_rn(__fsub_rn(two, v[1][1][0]), v[1][1][2]);
  float dyy = __fsub_rn(__fsub_rn(two, v[1][0][1]), v[1][2][1]);
  float dxy = __fmul_rn(0.25f, __fsub_rn(__fsub_rn(__fadd_rn(v[1][2][2], v[1][0][0]), v[1][0][2]), v[1][2][0]));
  float tra = __fadd_rn(dxx, dyy);
  float det = __fmaf_rn(dxx, dyy, -__fmul_rn(dxy, dxy));
  float tra2 = __fmul_rn(tra, tra);
  if (!(tra2 < __fmul_rn(P.edgeLimit, det))) return;
  float edge = __fdividef(tra2, det);
  float dx = __fmul_rn(0.5f, __fsub_rn(v[1][1][2], v[1][1][0]));
  float dy = __fmul_rn(0.5f, __fsub_rn(v[1][2][1], v[1][0][1]));
  float ds = __fmul_rn(0.5f, __fsub_rn(v[0][1][1], v[2][1][1]));
  float dss = __fsub_rn(__fsub_rn(two, v[2][1][1]), v[0][1][1]);
  float dxs = __fmul_rn(0.25f, __fsub_rn(__fsub_rn(__fadd_rn(v[2][1][2], v[0][1][0]), v[0][1][2]), v[2][1][0]));
  float dys = __fmul_rn(0.25f, __fsub_rn(__fsub_rn(__fadd_rn(v[2][2][1], v[0][0][1]), v[2][0][1]), v[0][2][1]));
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
  extern __shared__ __align__(16) float2 smem2[];
  float2 *s_v = smem2;
  float2 *s_dog = smem2 + DT_SM_V;
  float2 *s_in = s_dog + DT_SM_DOG;
  unsigned short *s_list = reinterpret_cast<unsigned short *>(s_in);    // 868 entries, alive after the blur
  static_assert((DT_H - 2) * (DT_W - 2) * 2 <= DT_SM_IN * 8, "candidate list fits into the staging area");

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

  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  const float thresh = P.thresh;
  dog_tile(L.img, w, h, L.pitch, x0, y0, L.taps, s_in, s_v, s_dog, thresh, s_list, &s_cnt);

  // 3x3x3 extrema, only on the pixels the blur pass flagged (|DoG| > thresh at some scale)
  const float *dogf = reinterpret_cast<const float *>(s_dog);
  const int ncand = s_cnt;
  for (int i = threadIdx.x; i < ncand; i += DT_THREADS) {
    const int rd = s_list[i];
    const int r = rd / DT_W, d = rd - r * DT_W;
    int off[3][3];                                    // float offsets of the 3x3 window inside a plane
#pragma unroll
    for (int dy = 0; dy < 3; dy++)
#pragma unroll
      for (int dx = 0; dx < 3; dx++) off[dy][dx] = dog_index(0, r + dy - 1, d + dx - 1);
#pragma unroll 1
    for (int sc = 0; sc < CS_NUM_SCALES; sc++) {
      const float *pl = dogf + sc * (DT_HP * DT_W * 2);
      const float c = pl[DT_HP * DT_W * 2 + off[1][1]];
      if (!(fabsf(c) > thresh)) continue;
      float v[3][3][3];
      bool mx = true, mn = true;
#pragma unroll
      for (int p = 0; p < 3; p++)
#pragma unroll
        for (int dy = 0; dy < 3; dy++)
#pragma unroll
          for (int dx = 0; dx < 3; dx++) {
            const float t = pl[p * (DT_HP * DT_W * 2) + off[dy][dx]];
            v[p][dy][dx] = t;
            if (p != 1 || dy != 1 || dx != 1) { mx = mx && (c > t); mn = mn && (c < t); }
          }
      if (c > 0.0f ? mx : mn) refine_and_store(v, x0 + d, y0 + r, sc, L, P);
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
  extern __shared__ __align__(16) float2 smem2[];
  float2 *s_v = smem2;
  float2 *s_dog = smem2 + DT_SM_V;
  float2 *s_in = s_dog + DT_SM_DOG;
  const int by = blockIdx.x / tilesX, bx = blockIdx.x - by * tilesX;
  const int x0 = bx * (DT_W - 2), y0 = by * (DT_H - 2);
  dog_tile(img, w, h, pitch, x0, y0, taps, s_in, s_v, s_dog);
  const float *dogf = reinterpret_cast<const float *>(s_dog);
  const size_t plane = (size_t)h * pitch;
  for (int i = threadIdx.x; i < (CS_LAPLACE_S - 1) * DT_H * DT_W; i += DT_THREADS) {
    int s = i / (DT_H * DT_W), rem = i - s * (DT_H * DT_W);
    int r = rem / DT_W, d = rem - r * DT_W;
    int gx = x0 + d, gy = y0 + r;
    if (gx < w && gy < h) dog[s * plane + (size_t)gy * pitch + gx] = dogf[dog_index(s, r, d)];
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
