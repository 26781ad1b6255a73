// describe.cu -- fused orientation assignment + 128-D descriptor extraction.
//
// Behavioural spec: reference ComputeOrientationsCONST (cudaSiftD.cu:972-1057) and
// ExtractSiftDescriptorsCONSTNew (cudaSiftD.cu:308-417, FastAtan2 :295-306), host side
// cudaSiftH.cu:353-382.  The reference runs two persistent kernels per octave (10 launches
// per image) and accumulates both histograms with shared-memory float atomics, whose
// summation order -- and therefore the low-order bits of every descriptor -- change from
// run to run.  Here a single launch serves all octaves: one 128-thread CTA per keypoint
// computes the orientation histogram, its one or two peaks, and then one descriptor per
// peak.  Both histograms are accumulated by an owner-computes gather in a fixed order
// (the same order oracle/sift_oracle.c uses), so results are run-to-run deterministic.
//
// Image samples go through the texture unit exactly as in the reference (bilinear,
// clamp, unnormalised coordinates, cudaSiftH.cu:186-205), so the 1.8 fixed-point
// interpolation weights are the hardware's own.
#include "common.cuh"

namespace cs {

#define DS_THREADS 128

// cudaSiftD.cu:295-306, as scheduled in the reference's SASS.
__device__ __forceinline__ float fast_atan2(float y, float x)
{
  float absx = fabsf(x), absy = fabsf(y);
  float a = __fdiv_rn(fminf(absx, absy), fmaxf(absx, absy));
  float s = __fmul_rn(a, a);
  float r = __fmaf_rn(s, -0.0464964749f, 0.15931422f);
  r = __fmaf_rn(s, r, -0.327622764f);
  r = __fmul_rn(s, r);
  r = __fmaf_rn(r, a, a);
  r = (absy > absx ? __fsub_rn(1.57079637f, r) : r);
  r = (x < 0 ? __fsub_rn(3.14159274f, r) : r);
  r = (y < 0 ? -r : r);
  return r;
}

__global__ void __launch_bounds__(DS_THREADS)
describe_kernel(const __grid_constant__ DescribeParams P)
{
  __shared__ float s_hist[64];
  __shared__ float s_gauss11[11];
  __shared__ float s_gauss16[16];
  __shared__ float s_ow[121];          // orientation sample weights
  __shared__ int s_obin[121];          // orientation sample bins
  __shared__ __align__(16) float s_g2[256][4];   // descriptor votes per sample: ul, ll, ur, lr
  __shared__ float s_angf[256];
  __shared__ int s_angi[256];
  __shared__ float s_buf[128];
  __shared__ float s_sums[4];
  __shared__ float s_ori[2];
  __shared__ int s_slot[2];
  __shared__ int s_nori;

  const int tx = threadIdx.x;
  if (tx < 16) s_gauss16[tx] = __expf(-(tx - 7.5f) * (tx - 7.5f) / 128.0f);   // cudaSiftD.cu:318

  const unsigned int found = P.counters[0];
  const int numPrim = (int)min(found, (unsigned)P.maxPts);

  for (int pt = blockIdx.x; pt < numPrim; pt += gridDim.x) {
    SiftPoint *sp = P.pts + pt;
    const float px = sp->xpos, py = sp->ypos, pscale = sp->scale, psub = sp->subsampling;
    const int level = ((__float_as_int(psub) >> 23) & 0xff) - 127;   // subsampling = 2^level
    const cudaTextureObject_t tex = P.tex[level];

    // ------------------------------------------------------------ orientation histogram
    {
      float i2sigma2 = __fdiv_rn(-1.0f, __fmul_rn(__fmul_rn(pscale, 4.5f), pscale));   // :982
      if (tx < 11) {
        float t = (float)(tx - 5);
        s_gauss11[tx] = expf(__fmul_rn(t, __fmul_rn(t, i2sigma2)));                     // :984
      }
      if (tx < 64) s_hist[tx] = 0.0f;
    }
    __syncthreads();
    if (tx < 121) {
      int yd = tx / 11, xd = tx - yd * 11;
      float xf = __fadd_rn((float)xd, __fadd_rn(px, -4.5f));
      float yf = __fadd_rn((float)yd, __fadd_rn(py, -4.5f));
      float dx = __fsub_rn(tex2D<float>(tex, __fadd_rn(xf, 1.0f), yf), tex2D<float>(tex, __fadd_rn(xf, -1.0f), yf));
      float dy = __fsub_rn(tex2D<float>(tex, xf, __fadd_rn(yf, 1.0f)), tex2D<float>(tex, xf, __fadd_rn(yf, -1.0f)));
      int bin = __float2int_rz(__fadd_rn(__fdiv_rn(__fmul_rn(16.0f, atan2f(dy, dx)), 3.1416f), 16.5f));   // :997
      if (bin > 31) bin = 0;
      if (bin < 0) bin = 0;
      float grad = __fsqrt_rn(__fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
      s_ow[tx] = __fmul_rn(__fmul_rn(grad, s_gauss11[xd]), s_gauss11[yd]);
      s_obin[tx] = bin;
    }
    __syncthreads();
    if (tx < 32) {   // owner-computes gather in sample order (deterministic)
      float acc = 0.0f;
      for (int i = 0; i < 121; i++)
        if (s_obin[i] == tx) acc = __fadd_rn(acc, s_ow[i]);
      s_hist[tx] = acc;
    }
    __syncthreads();
    const int x1m = (tx >= 1 ? tx - 1 : tx + 31), x1p = (tx <= 30 ? tx + 1 : tx - 31);
    if (tx < 32) {   // :1004-1010
      int x2m = (tx >= 2 ? tx - 2 : tx + 30), x2p = (tx <= 29 ? tx + 2 : tx - 30);
      float v = __fmaf_rn(s_hist[tx], 6.0f, __fmul_rn(4.0f, __fadd_rn(s_hist[x1m], s_hist[x1p])));
      s_hist[tx + 32] = __fadd_rn(v, __fadd_rn(s_hist[x2m], s_hist[x2p]));
    }
    __syncthreads();
    if (tx < 32) {   // :1012-1015
      float v = s_hist[32 + tx];
      s_hist[tx] = (v > s_hist[32 + x1m] && v >= s_hist[32 + x1p] ? v : 0.0f);
    }
    __syncthreads();
    if (tx == 0) {   // :1017-1053
      float maxval1 = 0.0f, maxval2 = 0.0f;
      int i1 = -1, i2 = -1;
      for (int i = 0; i < 32; i++) {
        float v = s_hist[i];
        if (v > maxval1) { maxval2 = maxval1; maxval1 = v; i2 = i1; i1 = i; }
        else if (v > maxval2) { maxval2 = v; i2 = i; }
      }
      float val1 = s_hist[32 + ((i1 + 1) & 31)], val2 = s_hist[32 + ((i1 + 31) & 31)];
      float peak = __fadd_rn((float)i1, __fdiv_rn(__fmul_rn(0.5f, __fsub_rn(val1, val2)),
                                                  __fsub_rn(__fsub_rn(__fadd_rn(maxval1, maxval1), val1), val2)));
      s_ori[0] = __fmul_rn(11.25f, (peak < 0.0f ? __fadd_rn(peak, 32.0f) : peak));
      s_slot[0] = pt;
      int nori = 1;
      if (maxval2 > __fmul_rn(0.8f, maxval1)) {
        float v1 = s_hist[32 + ((i2 + 1) & 31)], v2 = s_hist[32 + ((i2 + 31) & 31)];
        float pk = __fadd_rn((float)i2, __fdiv_rn(__fmul_rn(0.5f, __fsub_rn(v1, v2)),
                                                  __fsub_rn(__fsub_rn(__fadd_rn(maxval2, maxval2), v1), v2)));
        // Reference quirk Q1 (cudaSiftH.cu:115): secondary orientations of the finest
        // octave land beyond numPts and are never reported -> do not produce them.
        if (psub != P.finestSubsampling) {
          atomicMax(&P.counters[1], (unsigned)numPrim);
          unsigned int idx = atomicAdd(&P.counters[1], 1u);
          if (idx < (unsigned)P.maxPts) {
            s_ori[1] = __fmul_rn(11.25f, (pk < 0.0f ? __fadd_rn(pk, 32.0f) : pk));
            s_slot[1] = (int)idx;
            nori = 2;
          }
        }
      }
      s_nori = nori;
    }
    __syncthreads();
    const int nori = s_nori;

    // ------------------------------------------------------------------ descriptors
    for (int k = 0; k < nori; k++) {
      const float orientation = s_ori[k];
      float theta = __fmul_rn(2.0f * 3.1415f / 360.0f, orientation);   // :330
      float sina = __sinf(theta), cosa = __cosf(theta);
      float scale = __fmul_rn(12.0f / 16.0f, pscale);
      float ssina = __fmul_rn(scale, sina), scosa = __fmul_rn(scale, cosa);
      int has8 = 0;
#pragma unroll
      for (int rep = 0; rep < 2; rep++) {
        const int sidx = tx + rep * DS_THREADS;     // sample index = y*16 + x
        const int y = sidx >> 4, x = sidx & 15;
        float tt = x - 7.5f, yy = y - 7.5f;
        // :338-339 as contracted in the reference's SASS
        float xpos = __fadd_rn(__fmaf_rn(-ssina, yy, __fadd_rn(__fmul_rn(tt, scosa), px)), 0.5f);
        float ypos = __fadd_rn(__fmaf_rn(scosa, yy, __fmaf_rn(tt, ssina, py)), 0.5f);
        float dx = __fsub_rn(tex2D<float>(tex, __fadd_rn(xpos, cosa), __fadd_rn(ypos, sina)),
                             tex2D<float>(tex, __fsub_rn(xpos, cosa), __fsub_rn(ypos, sina)));
        float dy = __fsub_rn(tex2D<float>(tex, __fsub_rn(xpos, sina), __fadd_rn(ypos, cosa)),
                             tex2D<float>(tex, __fadd_rn(xpos, sina), __fsub_rn(ypos, cosa)));
        float grad = __fmul_rn(__fmul_rn(s_gauss16[y], s_gauss16[x]),
                               __fsqrt_rn(__fmaf_rn(dx, dx, __fmul_rn(dy, dy))));
        float angf = __fmaf_rn(fast_atan2(dy, dx), 4.0f / 3.1415f, 4.0f);   // :345
        int hori = (x + 2) / 4 - 1;
        float horf = __fsub_rn(__fmul_rn(x - 1.5f, 0.25f), (float)hori), ihorf = __fsub_rn(1.0f, horf);
        int veri = (y + 2) / 4 - 1;
        float verf = __fsub_rn(__fmul_rn(y - 1.5f, 0.25f), (float)veri), iverf = __fsub_rn(1.0f, verf);
        int angi = __float2int_rz(angf);
        angf = __fsub_rn(angf, (float)angi);
        // Quirk Q22: for dy == +0, dx < 0 (edges of saturated areas) angf = 8.0001 and angi = 8;
        // the reference then adds its "iangf" vote at flat index 8*cell + 8, i.e. into angle
        // bin 0 of the NEXT cell (cudaSiftD.cu:353-384).  Reproduced below (has8 path).
        has8 |= (angi >= 8);
        float gl = __fmul_rn(ihorf, grad), gr = __fmul_rn(horf, grad);
        float4 g2 = make_float4(__fmul_rn(iverf, gl), __fmul_rn(verf, gl), __fmul_rn(iverf, gr), __fmul_rn(verf, gr));
        *reinterpret_cast<float4 *>(s_g2[sidx]) = g2;
        s_angf[sidx] = angf;
        s_angi[sidx] = angi;
      }
      has8 = __syncthreads_or(has8);
      {  // owner-computes gather: thread = output bin (ycell, xcell, angle)
        const int cy = tx >> 5, cx = (tx >> 3) & 3, a = tx & 7;
        const int ylo = max(0, 4 * cy - 2), yhi = min(15, 4 * cy + 5);
        const int xlo = max(0, 4 * cx - 2), xhi = min(15, 4 * cx + 5);
        float acc = 0.0f;
        for (int y = ylo; y <= yhi; y++) {
          const int lower = (((y + 2) >> 2) - 1 != cy);       // sample votes into its lower cell
          for (int x = xlo; x <= xhi; x++) {
            const int right = (((x + 2) >> 2) - 1 != cx);     // ... into its right cell
            const int sidx = y * 16 + x;
            const int angi = s_angi[sidx];
            const int angp = (angi < 7 ? angi + 1 : 0);
            if (angi == a || angp == a) {
              float g2 = s_g2[sidx][2 * right + lower];
              float af = s_angf[sidx];
              float wgt = (angi == a ? __fsub_rn(1.0f, af) : af);
              acc = __fadd_rn(acc, __fmul_rn(wgt, g2));
            }
          }
        }
        if (has8 && a == 0 && tx >= 8) {
          // Q22 votes of the previous cell (flat order) land in this cell's bin 0
          const int pc = (tx >> 3) - 1, pcy = pc >> 2, pcx = pc & 3;
          const int y0 = max(0, 4 * pcy - 2), y1 = min(15, 4 * pcy + 5);
          const int x0 = max(0, 4 * pcx - 2), x1 = min(15, 4 * pcx + 5);
          for (int y = y0; y <= y1; y++) {
            const int lower = (((y + 2) >> 2) - 1 != pcy);
            for (int x = x0; x <= x1; x++) {
              const int right = (((x + 2) >> 2) - 1 != pcx);
              const int sidx = y * 16 + x;
              if (s_angi[sidx] >= 8)
                acc = __fadd_rn(acc, __fmul_rn(__fsub_rn(1.0f, s_angf[sidx]), s_g2[sidx][2 * right + lower]));
            }
          }
        }
        s_buf[tx] = acc;
      }
      // :391-409 normalise, clamp at 0.2, renormalise
      float v = s_buf[tx];
      float sum = __fmul_rn(v, v);
#pragma unroll
      for (int i = 16; i > 0; i /= 2) sum = __fadd_rn(sum, __shfl_down_sync(0xffffffffu, sum, i));
      if ((tx & 31) == 0) s_sums[tx >> 5] = sum;
      __syncthreads();
      float tsum1 = __fadd_rn(__fadd_rn(__fadd_rn(s_sums[0], s_sums[1]), s_sums[2]), s_sums[3]);
      float t1 = fminf(__fmul_rn(v, rsqrtf(tsum1)), 0.2f);
      sum = __fmul_rn(t1, t1);
#pragma unroll
      for (int i = 16; i > 0; i /= 2) sum = __fadd_rn(sum, __shfl_down_sync(0xffffffffu, sum, i));
      __syncthreads();
      if ((tx & 31) == 0) s_sums[tx >> 5] = sum;
      __syncthreads();
      float tsum2 = __fadd_rn(__fadd_rn(__fadd_rn(s_sums[0], s_sums[1]), s_sums[2]), s_sums[3]);
      SiftPoint *out = P.pts + s_slot[k];
      out->data[tx] = __fmul_rn(t1, rsqrtf(tsum2));
      if (tx == 0) {
        out->xpos = __fmul_rn(px, psub);        // :410-414
        out->ypos = __fmul_rn(py, psub);
        out->scale = __fmul_rn(pscale, psub);
        out->orientation = orientation;
        if (k == 1) {                           // :1045-1051 (copy of the primary)
          out->sharpness = sp->sharpness;
          out->edgeness = sp->edgeness;
          out->subsampling = psub;
        }
      }
      __syncthreads();
    }
  }
}

int launch_describe(const DescribeParams &p, int gridBlocks, cudaStream_t st)
{
  describe_kernel<<<gridBlocks, DS_THREADS, 0, st>>>(p);
  count_launch();
  CS_CUDA(cudaGetLastError());
  return 0;
}

// cudaSiftD.cu:753-761 (RescalePositions) for the scaleUp path; count read on device.
__global__ void rescale_kernel(SiftPoint *pts, const unsigned int *counters, int maxPts, float f)
{
  unsigned int n = max(min(counters[0], (unsigned)maxPts), min(counters[1], (unsigned)maxPts));
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    pts[i].xpos = __fmul_rn(pts[i].xpos, f);
    pts[i].ypos = __fmul_rn(pts[i].ypos, f);
    pts[i].scale = __fmul_rn(pts[i].scale, f);
  }
}

int launch_rescale(SiftPoint *pts, const unsigned int *counters, int maxPts, float f, cudaStream_t st)
{
  rescale_kernel<<<64, 128, 0, st>>>(pts, counters, maxPts, f);
  count_launch();
  CS_CUDA(cudaGetLastError());
  return 0;
}

__global__ void tex_probe_kernel(cudaTextureObject_t tex, const float *xs, const float *ys, int n, float *out)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = tex2D<float>(tex, xs[i], ys[i]);
}

int launch_tex_probe(cudaTextureObject_t tex, const float *xs, const float *ys, int n, float *out, cudaStream_t st)
{
  tex_probe_kernel<<<idivup(n, 256), 256, 0, st>>>(tex, xs, ys, n, out);
  count_launch();
  CS_CUDA(cudaGetLastError());
  return 0;
}

int make_texture(cudaTextureObject_t *tex, const float *img, int w, int h, int pitch)
{ // cudaSiftH.cu:186-205
  cudaResourceDesc res = {};
  res.resType = cudaResourceTypePitch2D;
  res.res.pitch2D.devPtr = const_cast<float *>(img);
  res.res.pitch2D.width = w;
  res.res.pitch2D.height = h;
  res.res.pitch2D.pitchInBytes = (size_t)pitch * sizeof(float);
  res.res.pitch2D.desc = cudaCreateChannelDesc<float>();
  cudaTextureDesc td = {};
  td.addressMode[0] = cudaAddressModeClamp;
  td.addressMode[1] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModeLinear;
  td.readMode = cudaReadModeElementType;
  td.normalizedCoords = 0;
  CS_CUDA(cudaCreateTextureObject(tex, &res, &td, NULL));
  return 0;
}

}  // namespace cs
