// pyramid2.cu -- fused, batched Gaussian pyramid: (A) LowPass + first ScaleDown in one marching kernel fed by
// TMA, (B) the remaining ScaleDowns in one launch.  Behavioural spec: reference LowPassBlock
// (cudaSiftD.cu:1986-2037, host cudaSiftH.cu:406-435) and ScaleDown (cudaSiftD.cu:84-168, host :308-338);
// per-pixel arithmetic as in pyramid.cu (pinned to the reference's sm_100 SASS), so every level is bit-identical
// to the reference's.
//
// Kernel A: a CTA (64 threads) owns a strip of 240 columns x R rows of the full-resolution level and marches down
// the rows.  Input rows arrive by TMA (cp.async.bulk.tensor.2d, 256 x 1 boxes, 8 rows in flight per CTA, row index
// clamped by the issuing thread); each thread filters 4 columns: horizontal 9-tap from the staged row, vertical
// 9-tap over its own column history (a 16-row ring in shared memory that only the owning thread touches, so no
// barrier), one 128-bit store of the level-0 row.  The same row is handed to the 5-tap ScaleDown through shared
// memory, so level 1 is produced without re-reading level 0: HBM traffic per image = read input + write level 0
// + write level 1 (the reference: + one more read of level 0).
// Kernel B: a CTA owns an 8x8 tile of the coarsest level it produces and computes the up to three levels above it
// for that tile in shared memory (85x85 -> 41x41 -> 19x19 -> 8x8), writing the part of every level it owns: the
// four dependent launches of the reference (cudaSiftH.cu:153-157) become two.
#include "common.cuh"

namespace cs {

__device__ __forceinline__ int clampi2(int v, int lo, int hi) { return min(max(v, lo), hi); }

// k[4] is the centre tap.  SASS of LowPassBlock: FMUL(k3,p1); FFMA(k4,c); FFMA(k2,p2); FFMA(k1,p3); FFMA(k0,p4).
__device__ __forceinline__ float pa_sym9(const Taps9 &t, float c, float p1, float p2, float p3, float p4)
{
  float s = __fmul_rn(t.k[3], p1);
  s = __fmaf_rn(t.k[4], c, s);
  s = __fmaf_rn(t.k[2], p2, s);
  s = __fmaf_rn(t.k[1], p3, s);
  s = __fmaf_rn(t.k[0], p4, s);
  return s;
}
// ScaleDown horizontal (cudaSiftD.cu:121): FMUL(k1,(a1+a3)); FFMA(k0,(a0+a4)); FFMA(k2,a2)
__device__ __forceinline__ float sd_h(const Taps5 &t, float a0, float a1, float a2, float a3, float a4)
{
  float s = __fmul_rn(t.k[1], __fadd_rn(a1, a3));
  s = __fmaf_rn(t.k[0], __fadd_rn(a0, a4), s);
  s = __fmaf_rn(t.k[2], a2, s);
  return s;
}
// ScaleDown vertical (cudaSiftD.cu:123): FMUL(k0,(r0+r4)); FFMA(k2,c); FFMA(k1,(r1+r3))
__device__ __forceinline__ float sd_v(const Taps5 &t, float r0, float r1, float r2, float r3, float r4)
{
  float s = __fmul_rn(t.k[0], __fadd_rn(r0, r4));
  s = __fmaf_rn(t.k[2], r2, s);
  s = __fmaf_rn(t.k[1], __fadd_rn(r1, r3), s);
  return s;
}

#define PA_THREADS 64
#define PA_IW 256          // staged input columns: x0-8 .. x0+247
#define PA_NS 8            // input rows in flight per CTA
#define PA_HR 16           // rows of horizontally filtered history per thread

__global__ void __launch_bounds__(PA_THREADS)
pyr_lowpass_sd_kernel(const __grid_constant__ PyrAParams P)
{
  __shared__ __align__(128) float s_in[PA_NS][PA_IW];
  __shared__ __align__(16) float4 s_h[PA_HR][PA_THREADS];      // [row & 15][thread]: private columns
  __shared__ __align__(16) float s_l0[2][PA_IW];               // level-0 row handed to the ScaleDown
  __shared__ __align__(8) float2 s_h2[8][PA_THREADS];          // [row & 7][thread]: private columns
  __shared__ __align__(8) uint64_t s_full[PA_NS];

  const int t = threadIdx.x;
  const int strip = blockIdx.x % P.stripsX, rb = blockIdx.x / P.stripsX, img = blockIdx.y;
  const int w = P.w, h = P.h;
  const int x0 = strip * CS_PA_OWN;
  const int R0 = rb * P.rowsPerCta, R1 = (rb == P.rowBlocks - 1) ? h : min(R0 + P.rowsPerCta, h);
  const bool sd = P.lev1 != nullptr;
  const int Y1a = R0 >> 1, Y1b = sd ? ((rb == P.rowBlocks - 1) ? P.h1 : min(R1 >> 1, P.h1)) : 0;
  // level-0 rows this CTA computes / input rows it needs
  const int ya = sd ? max(R0 - 2, 0) : R0;
  const int yb = sd ? min(max(R1 - 1, 2 * Y1b), h - 1) : R1 - 1;
  const int ia = max(ya - 4, 0), ib = min(yb + 4, h - 1);
  const CUtensorMap *map = P.inMaps + img;
  float *lev0 = P.lev0 + (size_t)img * P.lev0Stride;
  float *lev1 = sd ? P.lev1 + (size_t)img * P.lev1Stride : nullptr;

  if (t == 0) {
    for (int i = 0; i < PA_NS; i++) mbarrier_init(&s_full[i], 1);
    mbarrier_init_fence();
  }
  __syncthreads();
  if (t == 0) {                                             // the maps are kernel parameters: no descriptor fence needed
    for (int i = 0; i < PA_NS; i++)
      if (ia + i <= ib) {
        mbarrier_expect_tx(&s_full[i], PA_IW * 4);
        tma_load_2d(&s_in[i][0], map, x0 - 8, ia + i, &s_full[i]);
      }
  }

  const bool edge = (x0 == 0) || (x0 + PA_IW - 8 > w);       // some staged column lies outside the image
  const bool act = t < 62;                                   // level-0 columns cb .. cb+3
  const int cb = x0 - 4 + 4 * t;
  const bool own = t >= 1 && t <= 60;                        // owned columns x0 .. x0+239
  const int x1 = (x0 >> 1) + 2 * (t - 1);                    // this thread's two level-1 columns
  int nextY1 = Y1a;

  for (int u = ia; u <= yb + 4; u++) {
    if (u <= ib) {
      // ---------------------------------------------------------------- horizontal 9-tap of input row u
      const int seq = u - ia, slot = seq & (PA_NS - 1);
      mbarrier_wait(&s_full[slot], (seq >> 3) & 1);
      if (act) {
        float v[12];
        if (!edge) {
          const float4 *p = reinterpret_cast<const float4 *>(&s_in[slot][4 * t]);
          const float4 a = p[0], b = p[1], c = p[2];
          v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
          v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w;
        } else {
#pragma unroll
          for (int i = 0; i < 12; i++) v[i] = s_in[slot][clampi2(cb - 4 + i, 0, w - 1) - (x0 - 8)];
        }
        float o[4];
#pragma unroll
        for (int d = 0; d < 4; d++)
          o[d] = pa_sym9(P.lp, v[d + 4], __fadd_rn(v[d + 5], v[d + 3]), __fadd_rn(v[d + 6], v[d + 2]),
                         __fadd_rn(v[d + 7], v[d + 1]), __fadd_rn(v[d + 8], v[d]));
        s_h[u & (PA_HR - 1)][t] = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
    const int y = u - 4;
    const bool row = y >= ya && y <= yb;
    if (row && act) {
      // ---------------------------------------------------------------- vertical 9-tap -> level-0 row y
      float4 r[9];
#pragma unroll
      for (int j = 0; j < 9; j++) r[j] = s_h[clampi2(y - 4 + j, 0, h - 1) & (PA_HR - 1)][t];
      float4 o;
#define PAV(f) pa_sym9(P.lp, r[4].f, __fadd_rn(r[3].f, r[5].f), __fadd_rn(r[2].f, r[6].f), __fadd_rn(r[1].f, r[7].f), __fadd_rn(r[0].f, r[8].f))
      o.x = PAV(x); o.y = PAV(y); o.z = PAV(z); o.w = PAV(w);
#undef PAV
      if (own && y >= R0 && y < R1) {
        float *out = lev0 + (size_t)y * P.p0 + cb;
        if (cb + 3 < w) *reinterpret_cast<float4 *>(out) = o;
        else {
          if (cb < w) out[0] = o.x;
          if (cb + 1 < w) out[1] = o.y;
          if (cb + 2 < w) out[2] = o.z;
        }
      }
      if (sd) *reinterpret_cast<float4 *>(&s_l0[y & 1][4 * t]) = o;
    }
    __syncthreads();
    if (t == 0 && u <= ib && u + PA_NS <= ib) {               // refill the slot of row u
      const int slot = (u - ia) & (PA_NS - 1);
      mbarrier_expect_tx(&s_full[slot], PA_IW * 4);
      tma_load_2d(&s_in[slot][0], map, x0 - 8, u + PA_NS, &s_full[slot]);
    }
    if (sd && row && own) {
      // ---------------------------------------------------------------- ScaleDown: horizontal 5-tap of row y
      float a[7];
      const float *l0 = s_l0[y & 1];
#pragma unroll
      for (int i = 0; i < 7; i++) a[i] = l0[clampi2(cb - 2 + i, 0, w - 1) - (x0 - 4)];
      s_h2[y & 7][t] = make_float2(sd_h(P.sd, a[0], a[1], a[2], a[3], a[4]), sd_h(P.sd, a[2], a[3], a[4], a[5], a[6]));
    }
    if (sd && row) {
      // ---------------------------------------------------------------- ... vertical 5-tap -> level-1 rows that are complete
      while (nextY1 < Y1b && min(2 * nextY1 + 2, h - 1) <= y) {
        if (own) {
          float2 q[5];
#pragma unroll
          for (int j = 0; j < 5; j++) q[j] = s_h2[clampi2(2 * nextY1 - 2 + j, 0, h - 1) & 7][t];
          const float ox = sd_v(P.sd, q[0].x, q[1].x, q[2].x, q[3].x, q[4].x);
          const float oy = sd_v(P.sd, q[0].y, q[1].y, q[2].y, q[3].y, q[4].y);
          float *out = lev1 + (size_t)nextY1 * P.p1 + x1;
          if (x1 + 1 < P.w1) *reinterpret_cast<float2 *>(out) = make_float2(ox, oy);
          else if (x1 < P.w1) out[0] = ox;
        }
        nextY1++;
      }
    }
  }
}

const void *pyr_a_func() { return (const void *)pyr_lowpass_sd_kernel; }

int launch_pyr_a(const PyrAParams &p, int batch, cudaStream_t st)
{
  dim3 grid(p.stripsX * p.rowBlocks, batch);
  pyr_lowpass_sd_kernel<<<grid, PA_THREADS, 0, st>>>(p);
  count_launch();
  CS_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------
// Kernel B: a chain of up to 3 ScaleDowns for one 8x8 tile of the last level.
// ---------------------------------------------------------------------------------------
#define PB_THREADS 256
#define PB_T 8
#define PB_N1 (2 * PB_T + 3)      // 19
#define PB_N2 (2 * PB_N1 + 3)     // 41
#define PB_N3 (2 * PB_N2 + 3)     // 85

// One ScaleDown inside shared memory.  src: nin x nin region of level `ls` whose element (0,0) is pixel
// (sx0, sy0); dst: nout x nout region of the next level with origin (dx0, dy0) = ((sx0+2)/2, (sy0+2)/2).
// Coordinates are clamped to the image first (cudaSiftD.cu:98-99,113), then to the region.
__device__ __forceinline__ void pb_scaledown(const float *src, int nin, int sx0, int sy0, int sw, int sh,
                                             float *tmp, float *dst, int nout, int dx0, int dy0, const Taps5 &k)
{
  for (int i = threadIdx.x; i < nin * nout; i += PB_THREADS) {      // horizontal: rows of src x columns of dst
    const int r = i / nout, c = i - r * nout;
    const float *row = src + r * nin;
    float a[5];
#pragma unroll
    for (int j = 0; j < 5; j++) a[j] = row[clampi2(clampi2(2 * (dx0 + c) - 2 + j, 0, sw - 1) - sx0, 0, nin - 1)];
    tmp[r * nout + c] = sd_h(k, a[0], a[1], a[2], a[3], a[4]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nout * nout; i += PB_THREADS) {
    const int r = i / nout, c = i - r * nout;
    float q[5];
#pragma unroll
    for (int j = 0; j < 5; j++) q[j] = tmp[clampi2(clampi2(2 * (dy0 + r) - 2 + j, 0, sh - 1) - sy0, 0, nin - 1) * nout + c];
    dst[r * nout + c] = sd_v(k, q[0], q[1], q[2], q[3], q[4]);
  }
  __syncthreads();
}

// write the part of a region that this tile owns: pixels [ox, ox+on) x [oy, oy+on) of the level
__device__ __forceinline__ void pb_store(const float *reg, int n, int rx0, int ry0, float *img, int w, int h, int pitch,
                                         int ox, int oy, int on)
{
  for (int i = threadIdx.x; i < on * on; i += PB_THREADS) {
    const int r = i / on, c = i - r * on;
    const int gx = ox + c, gy = oy + r;
    if (gx < w && gy < h) img[(size_t)gy * pitch + gx] = reg[(gy - ry0) * n + (gx - rx0)];
  }
}

__global__ void __launch_bounds__(PB_THREADS)
pyr_chain_kernel(const __grid_constant__ PyrBParams P)
{
  extern __shared__ __align__(16) float pb_smem[];
  float *bufA = pb_smem;                       // up to 85 x 85
  float *bufT = bufA + PB_N3 * PB_N3;          // horizontal results, up to 85 x 41
  float *bufB = bufT + PB_N3 * PB_N2;          // up to 41 x 41
  const int img = blockIdx.z;
  const int steps = P.steps;                   // 1..3
  // region sizes from the last level upwards: n[0] = 8 (last), n[k] = 2 n[k-1] + 3
  int n[4], ox[4], oy[4];
  n[0] = PB_T; ox[0] = blockIdx.x * PB_T; oy[0] = blockIdx.y * PB_T;
  for (int k = 1; k <= steps; k++) { n[k] = 2 * n[k - 1] + 3; ox[k] = 2 * ox[k - 1] - 2; oy[k] = 2 * oy[k - 1] - 2; }
  // stage the source region (level index 0 of P = the input of the chain), clamped
  {
    const float *src = P.img[0] + (size_t)img * P.stride[0];
    const int ns = n[steps], sx0 = ox[steps], sy0 = oy[steps];
    for (int i = threadIdx.x; i < ns * ns; i += PB_THREADS) {
      const int r = i / ns, c = i - r * ns;
      bufA[i] = __ldg(src + (size_t)clampi2(sy0 + r, 0, P.h[0] - 1) * P.pitch[0] + clampi2(sx0 + c, 0, P.w[0] - 1));
    }
  }
  __syncthreads();
  float *cur = bufA, *nxt = bufB;
  for (int s = 1; s <= steps; s++) {
    const int ki = steps - s + 1, ko = steps - s;        // region indices of source and destination
    pb_scaledown(cur, n[ki], ox[ki], oy[ki], P.w[s - 1], P.h[s - 1], bufT, nxt, n[ko], ox[ko], oy[ko], P.sd);
    // owned part of level s: the pixels under this tile, 8 * 2^(steps - s) on a side
    const int on = PB_T << (steps - s);
    pb_store(nxt, n[ko], ox[ko], oy[ko], P.img[s] + (size_t)img * P.stride[s], P.w[s], P.h[s], P.pitch[s],
             blockIdx.x * on, blockIdx.y * on, on);
    __syncthreads();
    float *sw = cur; cur = nxt; nxt = sw;
  }
}

static int g_pb_configured[64];
#define PB_SMEM_BYTES ((PB_N3 * PB_N3 + PB_N3 * PB_N2 + PB_N2 * PB_N2) * 4)

int launch_pyr_b(const PyrBParams &p, int batch, cudaStream_t st)
{
  if (p.steps < 1) return 0;
  int dev = 0;
  CS_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && !g_pb_configured[dev]) {
    CS_CUDA(cudaFuncSetAttribute(pyr_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PB_SMEM_BYTES));
    g_pb_configured[dev] = 1;
  }
  // tiles must cover every level of the chain (integer halving can leave a level wider than twice the next)
  int tx = 1, ty = 1;
  for (int s = 1; s <= p.steps; s++) {
    const int on = PB_T << (p.steps - s);
    tx = tx > idivup(p.w[s], on) ? tx : idivup(p.w[s], on);
    ty = ty > idivup(p.h[s], on) ? ty : idivup(p.h[s], on);
  }
  dim3 grid(tx, ty, batch);
  pyr_chain_kernel<<<grid, PB_THREADS, PB_SMEM_BYTES, st>>>(p);
  count_launch();
  CS_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace cs
