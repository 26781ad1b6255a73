// match_tc.cu -- tcgen05 tensor-core matcher with exact FP32 re-score.
//
// Behavioural spec: the results are those of reference FindMaxCorr10 / MatchSiftData
// (matching.cu:301-397, 1090-1206), bit for bit -- see match.cu for the semantics.
// The reference evaluates all N1 x N2 x 128 products on the FP32 SIMT pipe; here the dense
// contraction runs on the 5th-generation tensor cores and only the handful of candidates
// that can decide a row are re-scored with the reference's sequential FMA chain.
//
// Pipeline:
//   1. prep      both descriptor sets -> FP16, written directly in the UMMA "interleaved"
//                (no-swizzle, K-major) core-matrix layout so that a 256x128 operand tile is
//                one contiguous 64 KB blob = one cp.async.bulk; row norms for the error bound.
//   2. gemm<1>   persistent warp-specialised kernel: cp.async.bulk producer warp, single-
//                thread tcgen05.mma issuer (M=128,N=256,K=16, FP32 accumulators in TMEM, two
//                256-column accumulators ping-pong), 8 epilogue warps read the accumulators
//                with tcgen05.ld and keep, per (row, partition), the running maximum.
//   3. gemm<2>   same GEMM; the epilogue now knows every partition's maximum and emits only
//                the 4-candidate groups that lie within the error bound of it.
//   (between 2 and 3: bound  -- thread per row: emission thresholds from the pass-1 maxima)
//   4. chain     thread per surviving candidate: exact k=0..127 FMA chain
//   5. final     thread per row: the reference's per-partition update rule and 8-way merge.
// Error bound: |fp16-tensor score - exact chain| <= eps(row) = C1*|a|*max|b| + C2*(|a|+max|b|)
// (input rounding 2^-11 per operand, FP32 accumulation); every candidate within
// delta = 2*eps of a maximum that can influence (score, match, ambiguity) is re-scored, so
// indices and scores equal the reference's exactly.  Rows whose bound cannot be certified
// (non-positive maxima, list overflow) fall back to the exact SIMT kernel.
#include "common.cuh"

#include <cuda_fp16.h>

namespace cs {

#define TC_M 128            // rows per accumulator (UMMA M)
#define TC_MT 256           // rows per CTA tile (two accumulators)
#define TC_N 256            // candidates per tile (UMMA N)
#define TC_KCH 16           // 16-byte chunks (8 halves) per descriptor
#define TC_A_BYTES (TC_MT * 256)          // 64 KB
#define TC_B_BYTES (TC_N * 256)           // 64 KB
#define TC_STAGES 2
#define TC_THREADS 384
#define TC_RT 16            // exact scores kept per row before falling back
#define TC_PM 12            // floats per (row, segment) of pass-1 output (8 maxima + P0 second)
#define TC_QS 1024          // shared-memory candidate queue entries per CTA (pass 2)
#define TC_SMEM_BYTES (TC_A_BYTES + TC_STAGES * TC_B_BYTES + TC_QS * 24 + 1024)
#define TC_C1 1.06e-3f      // > 2^-10 (inputs) + 2^-15 (tensor accumulate) + 2^-17 (reference chain)
#define TC_C2 1.0e-6f       // FP16 subnormal rounding, 2^-25 * sqrt(128)

// ------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
  uint32_t addr = smem_u32(bar), ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 consecutive 32-bit columns of this thread's TMEM lane (asynchronous: pair with tc_ld_wait)
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, uint32_t (&r)[32])
{
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
}
// Wait for the outstanding tcgen05.ld's.  The registers are in/out operands so that the
// compiler cannot schedule their consumers above the wait.
__device__ __forceinline__ void tc_ld_wait(uint32_t (&r)[32])
{
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :: "memory");
}

// UMMA shared-memory descriptor, K-major, no swizzle ("interleaved" core matrices of
// 8 rows x 16 bytes).  lbo = byte distance between the two K chunks of one K=16 step,
// sbo = byte distance between consecutive 8-row groups.  Bit layout: cute::UMMA::SmemDescriptor.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: D=F32, A=B=F16, both K-major, N=256, M=128
// (cute::UMMA::InstrDescriptor: c_format [4,6), n_dim>>3 at [17,23), m_dim>>4 at [24,29)).
#define TC_IDESC ((1u << 4) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24))

// ------------------------------------------------------------------------------ workspace
struct TcPlan {
  int n1, n2v, n_mt, n_nt, total, U, runs, grid;
  int dbg;   // experiments (env CS_TC_DBG): 1 = do not reload B tiles, 2 = split bulk copies into 8 KB pieces
};
struct TcBuffers {
  __half *a16, *b16;          // packed operand tiles
  float *normA;               // per row of set 1
  float *bmax;                // [0] max norm of set 2 (float bits), [1] bad-input flag
  float *pm;                  // pass-1 maxima  [slot][TC_PM]
  float *rowthr;              // [row][8] emission thresholds (3e38 = partition cannot matter)
  float4 *qv;                 // candidate queue: the 4 tensor scores of the group
  unsigned int *qgid;         // group id (p2 / 4)
  int *qrow;
  unsigned int qcap;
  unsigned int *rcnt;         // [row] number of exact scores in the row's table
  float2 *rtab;               // [row][TC_RT] (exact score, p2 as int bits)
  int *fbRows;                // fallback row list
  unsigned int *counters;     // [0] fallback rows, [1] queue entries, [2] chains re-scored, [3] overflow
};

// ------------------------------------------------------------------------------ prep
// block = 32 rows x 16 chunks; writes chunk-major tiles: blob(tile)[chunk][row][8 halves]
// One launch converts both sets: CTAs [0, blocksA) take set 1, the rest set 2.
__global__ void __launch_bounds__(512)
tc_prep_kernel(const SiftPoint *__restrict__ ptsA, int nA, __half *__restrict__ outA, float *__restrict__ norms,
               int blocksA, const SiftPoint *__restrict__ ptsB, int nB, __half *__restrict__ outB,
               float *__restrict__ bmax)
{
  __shared__ float s_sq[16][33];
  __shared__ int s_bad;
  const int isB = (int)blockIdx.x >= blocksA;
  const SiftPoint *__restrict__ pts = isB ? ptsB : ptsA;
  const int nvalid = isB ? nB : nA, rowsPerTile = isB ? TC_N : TC_M;
  __half *__restrict__ out = isB ? outB : outA;
  const int lane = threadIdx.x & 31, c = threadIdx.x >> 5;
  const int row = (isB ? (int)blockIdx.x - blocksA : (int)blockIdx.x) * 32 + lane;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  float v[8];
  if (row < nvalid) {
    const float4 *p = reinterpret_cast<const float4 *>(pts[row].data + 8 * c);
    float4 a = __ldg(p), b = __ldg(p + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = 0.0f;
  }
  float sq = 0.0f;
  bool bad = false;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    sq = fmaf(v[i], v[i], sq);
    bad = bad || !(fabsf(v[i]) < 32768.0f);     // also catches NaN/Inf
  }
  if (bad) s_bad = 1;
  __half2 h[4];
#pragma unroll
  for (int i = 0; i < 4; i++) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
  const int tile = row / rowsPerTile, r = row - tile * rowsPerTile;
  uint4 pk;
  pk.x = *reinterpret_cast<uint32_t *>(&h[0]); pk.y = *reinterpret_cast<uint32_t *>(&h[1]);
  pk.z = *reinterpret_cast<uint32_t *>(&h[2]); pk.w = *reinterpret_cast<uint32_t *>(&h[3]);
  *reinterpret_cast<uint4 *>(out + ((size_t)(tile * TC_KCH + c) * rowsPerTile + r) * 8) = pk;
  s_sq[c][lane] = sq;
  __syncthreads();
  if (c == 0) {
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; i++) t += s_sq[i][lane];
    float nrm = sqrtf(t) * 1.0001f;
    if (!isB && row < nvalid) norms[row] = nrm;
    if (isB) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) nrm = fmaxf(nrm, __shfl_xor_sync(0xffffffffu, nrm, o));
      if (lane == 0) atomicMax(reinterpret_cast<int *>(bmax), __float_as_int(nrm));
    }
    if (lane == 0 && s_bad) atomicMax(reinterpret_cast<int *>(bmax) + 1, 1);
  }
}

// ------------------------------------------------------------------------------ shared bits
struct RowBound { float G[8]; float T2; float delta; };

// slot of (cta, run, row-in-tile)
__device__ __forceinline__ size_t tc_slot(const TcPlan &pl, int cta, int run, int rowInTile)
{
  return ((size_t)cta * pl.runs + run) * TC_MT + rowInTile;
}

// Combine the pass-1 maxima of every segment that covers m-tile `mt` for one row.
__device__ __forceinline__ void tc_row_bound(const TcPlan &pl, const float *__restrict__ pm, int mt, int rowInTile,
                                            float normA, float bmax, RowBound &rb)
{
#pragma unroll
  for (int p = 0; p < 8; p++) rb.G[p] = 0.0f;
  float t1 = 0.0f, t2 = 0.0f;
  const int u0 = mt * pl.n_nt, u1 = u0 + pl.n_nt - 1;
  for (int c = u0 / pl.U; c <= u1 / pl.U; c++) {
    const int run = mt - (c * pl.U) / pl.n_nt;
    const float *q = pm + tc_slot(pl, c, run, rowInTile) * TC_PM;
    float4 a = *reinterpret_cast<const float4 *>(q), b = *reinterpret_cast<const float4 *>(q + 4);
    float s2 = q[8];
    rb.G[0] = fmaxf(rb.G[0], a.x); rb.G[1] = fmaxf(rb.G[1], a.y); rb.G[2] = fmaxf(rb.G[2], a.z); rb.G[3] = fmaxf(rb.G[3], a.w);
    rb.G[4] = fmaxf(rb.G[4], b.x); rb.G[5] = fmaxf(rb.G[5], b.y); rb.G[6] = fmaxf(rb.G[6], b.z); rb.G[7] = fmaxf(rb.G[7], b.w);
    // top-2 of the union of the segments' (first, second) group maxima of partition 0
    t2 = fmaxf(fmaxf(t2, s2), fminf(t1, a.x));
    t1 = fmaxf(t1, a.x);
  }
  rb.T2 = t2;
  rb.delta = 2.0f * (TC_C1 * normA * bmax + TC_C2 * (normA + bmax));
}

// second largest element of the pool {G[0..7], T2}
__device__ __forceinline__ float tc_pool_second(const RowBound &rb)
{
  float m1 = rb.T2, m2 = 0.0f;
#pragma unroll
  for (int p = 0; p < 8; p++) {
    m2 = fmaxf(m2, fminf(m1, rb.G[p]));
    m1 = fmaxf(m1, rb.G[p]);
  }
  return m2;
}

// ------------------------------------------------------------------------------ GEMM
template <int PASS>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const TcPlan pl, const TcBuffers bf)
{
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *sA = smem;
  uint8_t *sB = smem + TC_A_BYTES;
  float4 *s_qv = reinterpret_cast<float4 *>(smem + TC_A_BYTES + TC_STAGES * TC_B_BYTES);   // pass-2 candidate queue
  uint2 *s_qm = reinterpret_cast<uint2 *>(s_qv + TC_QS);
  __shared__ unsigned int s_qcount, s_qbase;
  __shared__ uint64_t bar_a_full, bar_a_empty, bar_b_full[TC_STAGES], bar_b_empty[TC_STAGES];
  __shared__ uint64_t bar_acc_full[2], bar_acc_empty[2];
  __shared__ uint32_t s_tmem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x;
  const int u_begin = min(cta * pl.U, pl.total), u_end = min(u_begin + pl.U, pl.total);

  if (threadIdx.x == 0) {
    s_qcount = 0;
    mbar_init(&bar_a_full, 1); mbar_init(&bar_a_empty, 1);
    for (int s = 0; s < TC_STAGES; s++) { mbar_init(&bar_b_full[s], 1); mbar_init(&bar_b_empty[s], 1); }
    for (int h = 0; h < 2; h++) { mbar_init(&bar_acc_full[h], 1); mbar_init(&bar_acc_empty[h], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  if (warp == 0) {
    // ============================ producer: bulk copies of whole operand tiles ============================
    if (lane == 0) {
      int runIdx = 0, prevMt = -1;
      for (int u = u_begin, i = 0; u < u_end; u++, i++) {
        const int mt = u / pl.n_nt, nt = u - mt * pl.n_nt;
        if (mt != prevMt) {
          if (runIdx > 0) mbar_wait(&bar_a_empty, (runIdx - 1) & 1);
          mbar_expect_tx(&bar_a_full, TC_A_BYTES);
          bulk_g2s(sA, reinterpret_cast<const uint8_t *>(bf.a16) + (size_t)mt * TC_A_BYTES, TC_A_BYTES, &bar_a_full);
          prevMt = mt; runIdx++;
        }
        const int s = i % TC_STAGES, it = i / TC_STAGES;
        if (it > 0) mbar_wait(&bar_b_empty[s], (it - 1) & 1);
        const uint8_t *srcB = reinterpret_cast<const uint8_t *>(bf.b16) + (size_t)nt * TC_B_BYTES;
        if ((pl.dbg & 1) && it > 0) {
          mbar_arrive(&bar_b_full[s]);                  // experiment: MMA/epilogue rate without B traffic
        } else if (pl.dbg & 2) {
          mbar_expect_tx(&bar_b_full[s], TC_B_BYTES);
          for (int piece = 0; piece < 8; piece++)
            bulk_g2s(sB + s * TC_B_BYTES + piece * 8192, srcB + piece * 8192, 8192, &bar_b_full[s]);
        } else {
          mbar_expect_tx(&bar_b_full[s], TC_B_BYTES);
          bulk_g2s(sB + s * TC_B_BYTES, srcB, TC_B_BYTES, &bar_b_full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer (one thread) ============================
    if (lane == 0) {
      int runIdx = 0, prevMt = -1;
      const uint32_t aBase = smem_u32(sA), bBase = smem_u32(sB);
      for (int u = u_begin, i = 0; u < u_end; u++, i++) {
        const int mt = u / pl.n_nt;
        if (mt != prevMt) { mbar_wait(&bar_a_full, runIdx & 1); prevMt = mt; runIdx++; }
        const int s = i % TC_STAGES, it = i / TC_STAGES;
        mbar_wait(&bar_b_full[s], it & 1);
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < 2; h++) {
          mbar_wait(&bar_acc_empty[h], (i & 1) ^ 1);      // first use passes on a fresh barrier
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 8; k++) {
            // K=16 step k = chunks 2k, 2k+1; A half h: [chunk][128 rows][16 B]; B: [chunk][256 rows][16 B]
            uint64_t ad = umma_desc(aBase + h * (TC_A_BYTES / 2) + k * 2 * (TC_M * 16), TC_M * 16, 128);
            uint64_t bd = umma_desc(bBase + s * TC_B_BYTES + k * 2 * (TC_N * 16), TC_N * 16, 128);
            tc_mma_f16(tmem + h * TC_N, ad, bd, TC_IDESC, k > 0 ? 1u : 0u);
          }
          tc_commit(&bar_acc_full[h]);
        }
        tc_commit(&bar_b_empty[s]);
        const bool lastOfRun = (u + 1 == u_end) || ((u + 1) / pl.n_nt != mt);
        if (lastOfRun) tc_commit(&bar_a_empty);
      }
    }
  } else if (warp >= 4) {
    // ============================ epilogue: 8 warps, one TMEM lane (= row) per thread ============================
    const int h = (warp - 4) >> 2, q = warp & 3;
    const int rowInTile = h * TC_M + q * 32 + lane;
    const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + h * TC_N;
    float st[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};   // PASS 1: running maxima; PASS 2: emission thresholds
    float t2 = 0.0f;          // PASS 1: second largest group maximum of partition 0
    size_t slot = 0;
    int row = 0;
    int prevMt = -1, runIdx = 0;
    // one 32-column chunk: group j (columns 4j..4j+3) belongs to partition j
    auto process = [&](const uint32_t (&r)[32], int nt, int c) {
      float m[8];
#pragma unroll
      for (int j = 0; j < 8; j++)
        m[j] = fmaxf(fmaxf(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1])),
                     fmaxf(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])));
      if (PASS == 1) {
        t2 = fmaxf(t2, fminf(st[0], m[0]));
#pragma unroll
        for (int j = 0; j < 8; j++) st[j] = fmaxf(st[j], m[j]);
      } else {
        bool any = false;
#pragma unroll
        for (int j = 0; j < 8; j++) any = any || (m[j] > st[j]);
        if (any) {              // rare (about 3 groups per row over the whole sweep): one branch per chunk
#pragma unroll
          for (int j = 0; j < 8; j++)
            if (m[j] > st[j]) {
              // shared-memory queue (a global atomic with a returned index would cost ~1 us here);
              // flushed to the global queue once, at the end of the kernel
              const float4 vv = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                            __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
              const unsigned int gid = (unsigned)(nt * (TC_N / 4) + c * 8 + j);
              const unsigned int qi = atomicAdd(&s_qcount, 1u);
              if (qi < TC_QS) { s_qv[qi] = vv; s_qm[qi] = make_uint2(gid, (unsigned)row); }
              else {
                const unsigned int idx = atomicAdd(&bf.counters[1], 1u);
                if (idx < bf.qcap) { bf.qv[idx] = vv; bf.qgid[idx] = gid; bf.qrow[idx] = row; }
              }
            }
        }
      }
    };
    for (int u = u_begin, i = 0; u < u_end; u++, i++) {
      const int mt = u / pl.n_nt, nt = u - mt * pl.n_nt;
      if (mt != prevMt) {
        slot = tc_slot(pl, cta, runIdx, rowInTile);
        row = mt * TC_MT + rowInTile;
        prevMt = mt; runIdx++;
        t2 = 0.0f;
        if (PASS == 1) {
#pragma unroll
          for (int p = 0; p < 8; p++) st[p] = 0.0f;
        } else {
          if (row < pl.n1) {
            const float4 a = *reinterpret_cast<const float4 *>(bf.rowthr + (size_t)row * 8);
            const float4 b = *reinterpret_cast<const float4 *>(bf.rowthr + (size_t)row * 8 + 4);
            st[0] = a.x; st[1] = a.y; st[2] = a.z; st[3] = a.w; st[4] = b.x; st[5] = b.y; st[6] = b.z; st[7] = b.w;
          } else {
#pragma unroll
            for (int p = 0; p < 8; p++) st[p] = 3.0e38f;
          }
        }
      }
      mbar_wait(&bar_acc_full[h], i & 1);
      tc_fence_after();
      // software-pipelined TMEM reads: the load of chunk c+1 is in flight while chunk c is reduced
      uint32_t ra[32], rb[32];
      tc_ld32_issue(tbase, ra);
      tc_ld_wait(ra);
#pragma unroll 1
      for (int c = 0; c < TC_N / 32; c += 2) {
        tc_ld32_issue(tbase + (c + 1) * 32, rb);
        process(ra, nt, c);
        tc_ld_wait(rb);
        if (c + 2 < TC_N / 32) tc_ld32_issue(tbase + (c + 2) * 32, ra);
        process(rb, nt, c + 1);
        if (c + 2 < TC_N / 32) tc_ld_wait(ra);
      }
      tc_fence_before();
      mbar_arrive(&bar_acc_empty[h]);
      const bool lastOfRun = (u + 1 == u_end) || ((u + 1) / pl.n_nt != mt);
      if (lastOfRun && PASS == 1) {
        float *q4 = bf.pm + slot * TC_PM;
        *reinterpret_cast<float4 *>(q4) = make_float4(st[0], st[1], st[2], st[3]);
        *reinterpret_cast<float4 *>(q4 + 4) = make_float4(st[4], st[5], st[6], st[7]);
        q4[8] = t2;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
  if (PASS == 2) {      // flush the CTA's candidate queue: one global reservation, coalesced copy
    const unsigned int nq = min(s_qcount, (unsigned)TC_QS);
    if (threadIdx.x == 0) s_qbase = nq ? atomicAdd(&bf.counters[1], nq) : 0u;
    __syncthreads();
    for (unsigned int i = threadIdx.x; i < nq; i += TC_THREADS) {
      const unsigned int idx = s_qbase + i;
      if (idx < bf.qcap) { bf.qv[idx] = s_qv[i]; bf.qgid[idx] = s_qm[i].x; bf.qrow[idx] = (int)s_qm[i].y; }
    }
  }
}

// ------------------------------------------------------------------------------ bound
// Thread per row: which tensor scores must be re-scored exactly?  Pool = the 8 partition maxima
// plus partition 0's second best (quirk Q9); only pool elements within delta of the pool's second
// largest can become (score, ambiguity).  rowthr[p] = smallest tensor score of partition p that
// still needs an exact chain (3e38: the partition cannot matter).
__global__ void __launch_bounds__(128)
tc_bound_kernel(const TcPlan pl, const TcBuffers bf)
{
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= pl.n1) return;
  const int mt = row / TC_MT, rowInTile = row - mt * TC_MT;
  const float bmax = __int_as_float(*reinterpret_cast<const int *>(bf.bmax));
  RowBound rb;
  tc_row_bound(pl, bf.pm, mt, rowInTile, bf.normA[row], bmax, rb);
  const float A2 = tc_pool_second(rb);
  const float need = A2 - rb.delta;
  float thr[8];
  // certification: the two pool elements that decide the row must be clearly positive
  const bool certified = (A2 > 2.0f * rb.delta) && (A2 < 3.0e38f);
#pragma unroll
  for (int p = 0; p < 8; p++)
    thr[p] = (certified && rb.G[p] >= need) ? fmaxf(rb.G[p] - rb.delta, 0.0f) : 3.0e38f;
  // partition 0 also supplies its SECOND best: T2 (second largest group maximum) is a lower
  // bound of that value, so everything above T2 - delta is needed
  if (certified && rb.G[0] >= need) thr[0] = fmaxf(rb.T2 - rb.delta, 0.0f);
  *reinterpret_cast<float4 *>(bf.rowthr + (size_t)row * 8) = make_float4(thr[0], thr[1], thr[2], thr[3]);
  *reinterpret_cast<float4 *>(bf.rowthr + (size_t)row * 8 + 4) = make_float4(thr[4], thr[5], thr[6], thr[7]);
  if (!certified) bf.fbRows[atomicAdd(&bf.counters[0], 1u)] = row;
}

// ------------------------------------------------------------------------------ chain
// Thread per (queue entry, member): the reference's score, matching.cu:338-351 -- a sequential
// k = 0..127 FMA chain starting from 0 -- for every candidate that can still decide its row.
// Results go to the row's small table (slot by atomic counter; order is irrelevant, see final).
__global__ void __launch_bounds__(256)
tc_chain_kernel(const TcBuffers bf, const SiftPoint *__restrict__ sift1, const SiftPoint *__restrict__ sift2)
{
  const unsigned int n = min(bf.counters[1], bf.qcap) * 4u;
  if (blockIdx.x == 0 && threadIdx.x == 0 && bf.counters[1] > bf.qcap) atomicMax(&bf.counters[3], 1u);
  for (unsigned int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const unsigned int e = t >> 2, j = t & 3;
    const float4 vv = bf.qv[e];
    const unsigned int gid = bf.qgid[e];
    const int row = bf.qrow[e];
    const float v = (j == 0 ? vv.x : j == 1 ? vv.y : j == 2 ? vv.z : vv.w);
    if (v > bf.rowthr[(size_t)row * 8 + (gid & 7)]) {
      const float4 *a = reinterpret_cast<const float4 *>(sift1[row].data);
      const float4 *b = reinterpret_cast<const float4 *>(sift2[gid * 4 + j].data);
      float acc = 0.0f;
#pragma unroll 16
      for (int d = 0; d < 32; d++) {
        const float4 av = __ldg(a + d), bv = __ldg(b + d);
        acc = __fmaf_rn(av.x, bv.x, acc);
        acc = __fmaf_rn(av.y, bv.y, acc);
        acc = __fmaf_rn(av.z, bv.z, acc);
        acc = __fmaf_rn(av.w, bv.w, acc);
      }
      const unsigned int slot = atomicAdd(&bf.rcnt[row], 1u);
      if (slot < TC_RT) bf.rtab[(size_t)row * TC_RT + slot] = make_float2(acc, __int_as_float((int)(gid * 4 + j)));
      atomicAdd(&bf.counters[2], 1u);
    }
  }
}

// ------------------------------------------------------------------------------ final
// Thread per row.  The reference's per-partition rule (matching.cu:354-359: strict '>' in
// increasing p2) is order independent once stated as: pmax = largest score, pidx = lowest p2
// attaining it, psec = second largest of the multiset; then the 8-way merge of :378-390.
__global__ void __launch_bounds__(64)
tc_final_kernel(const TcPlan pl, const TcBuffers bf, SiftPoint *__restrict__ sift1, const SiftPoint *__restrict__ sift2)
{
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= pl.n1) return;
  float pmx[8], psec0 = 0.0f;
  int pidx[8];
#pragma unroll
  for (int p = 0; p < 8; p++) { pmx[p] = 0.0f; pidx[p] = -1; }
  const unsigned int n = bf.rcnt[row];
  if (n > TC_RT) {      // more near-ties than the table holds: exact scan of this row instead
    bf.fbRows[atomicAdd(&bf.counters[0], 1u)] = row;
    return;
  }
  const float2 *tab = bf.rtab + (size_t)row * TC_RT;
  for (unsigned int i = 0; i < n; i++) {
    const float2 en = tab[i];
    const float sc = en.x;
    const int p2 = __float_as_int(en.y);
    const int part = (p2 >> 2) & 7;
    if (!(sc > 0.0f)) continue;            // cannot match (matching.cu:317-321)
#pragma unroll
    for (int p = 0; p < 8; p++)
      if (p == part) {
        if (sc > pmx[p]) { if (p == 0) psec0 = pmx[0]; pmx[p] = sc; pidx[p] = p2; }
        else {
          if (sc == pmx[p]) pidx[p] = min(pidx[p], p2);
          if (p == 0) psec0 = fmaxf(psec0, sc);
        }
      }
  }
  float mx = pmx[0], sec = psec0;
  int idx = pidx[0];
#pragma unroll
  for (int y = 0; y < 8; y++)
    if (idx != pidx[y]) {
      if (pmx[y] > mx) { sec = fmaxf(mx, sec); mx = pmx[y]; idx = pidx[y]; }
      else if (pmx[y] > sec) sec = pmx[y];
    }
  SiftPoint *o = sift1 + row;
  o->score = mx;
  o->match = idx;
  o->match_xpos = idx >= 0 ? sift2[idx].xpos : 0.0f;
  o->match_ypos = idx >= 0 ? sift2[idx].ypos : 0.0f;
  o->ambiguity = __fdiv_rn(sec, __fadd_rn(mx, 1e-6f));
}

// exact SIMT scan of the rows the tensor path could not certify (match.cu)
int match_exact_rows(SiftPoint *s1, const SiftPoint *s2, int n2, const int *rows, const unsigned int *nrows,
                     cudaStream_t st);

// ------------------------------------------------------------------------------ host
struct TcWorkspace {
  TcBuffers bf = {};
  size_t cap[12] = {};
  unsigned int *h_counters = nullptr;
  bool configured = false;
};
static TcWorkspace g_ws[16];

static int ensure(void **p, size_t *cap, size_t bytes)
{
  if (*cap >= bytes) return 0;
  if (*p) cudaFree(*p);
  *p = nullptr; *cap = 0;
  CS_CUDA(cudaMalloc(p, bytes));
  *cap = bytes;
  return 0;
}

bool match_tensor_supported()
{
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  return major == 10;
}

int match_tensor(SiftPoint *s1, int n1, const SiftPoint *s2, int n2, cudaStream_t st, unsigned long long stats[4])
{
  stats[0] = stats[1] = stats[2] = 0;
  const int n2v = (n2 / 32) * 32;                      // matching.cu:325 (quirk Q7)
  if (n1 <= 0) return 0;
  if (n2v == 0 || !match_tensor_supported()) return match_exact(s1, n1, s2, n2, st);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  TcWorkspace &ws = g_ws[dev & 15];
  TcPlan pl;
  pl.n1 = n1; pl.n2v = n2v;
  pl.n_mt = idivup(n1, TC_MT); pl.n_nt = idivup(n2v, TC_N);
  pl.total = pl.n_mt * pl.n_nt;
  pl.grid = pl.total < sms ? pl.total : sms;
  pl.U = idivup(pl.total, pl.grid);
  pl.grid = idivup(pl.total, pl.U);
  pl.runs = idivup(pl.U, pl.n_nt) + 1;
  { const char *e = getenv("CS_TC_DBG"); pl.dbg = e ? atoi(e) : 0; }
  const size_t slots = (size_t)pl.grid * pl.runs * TC_MT;
  const size_t rowsPad = (size_t)pl.n_mt * TC_MT;
  const unsigned int qcap = (unsigned int)(16 * rowsPad + 65536);
  TcBuffers &bf = ws.bf;
  int r;
  if ((r = ensure((void **)&bf.a16, &ws.cap[0], (size_t)pl.n_mt * TC_A_BYTES)) < 0) return r;
  if ((r = ensure((void **)&bf.b16, &ws.cap[1], (size_t)pl.n_nt * TC_B_BYTES)) < 0) return r;
  if ((r = ensure((void **)&bf.normA, &ws.cap[2], rowsPad * sizeof(float))) < 0) return r;
  if ((r = ensure((void **)&bf.pm, &ws.cap[3], slots * TC_PM * sizeof(float))) < 0) return r;
  if ((r = ensure((void **)&bf.rowthr, &ws.cap[4], rowsPad * 8 * sizeof(float))) < 0) return r;
  if ((r = ensure((void **)&bf.rcnt, &ws.cap[5], rowsPad * sizeof(unsigned))) < 0) return r;
  if ((r = ensure((void **)&bf.qv, &ws.cap[6], (size_t)qcap * sizeof(float4))) < 0) return r;
  if ((r = ensure((void **)&bf.rtab, &ws.cap[7], rowsPad * TC_RT * sizeof(float2))) < 0) return r;
  if ((r = ensure((void **)&bf.qgid, &ws.cap[8], (size_t)qcap * sizeof(unsigned))) < 0) return r;
  if ((r = ensure((void **)&bf.qrow, &ws.cap[9], (size_t)qcap * sizeof(int))) < 0) return r;
  if ((r = ensure((void **)&bf.fbRows, &ws.cap[11], rowsPad * sizeof(int))) < 0) return r;
  bf.qcap = qcap;
  if (!bf.bmax) {
    size_t dummy = 0;
    if ((r = ensure((void **)&bf.bmax, &dummy, 64)) < 0) return r;
    bf.counters = reinterpret_cast<unsigned int *>(bf.bmax) + 4;
    CS_CUDA(cudaMallocHost((void **)&ws.h_counters, 64));
  }
  if (!ws.configured) {
    CS_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    CS_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    ws.configured = true;
  }
  CS_CUDA(cudaMemsetAsync(bf.bmax, 0, 64, st));
  CS_CUDA(cudaMemsetAsync(bf.rcnt, 0, (size_t)n1 * sizeof(unsigned), st));
  const int blocksA = pl.n_mt * (TC_MT / 32), blocksB = pl.n_nt * (TC_N / 32);
  tc_prep_kernel<<<blocksA + blocksB, 512, 0, st>>>(s1, n1, bf.a16, bf.normA, blocksA, s2, n2v, bf.b16, bf.bmax);
  tc_gemm_kernel<1><<<pl.grid, TC_THREADS, TC_SMEM_BYTES, st>>>(pl, bf);
  tc_bound_kernel<<<idivup(n1, 128), 128, 0, st>>>(pl, bf);
  tc_gemm_kernel<2><<<pl.grid, TC_THREADS, TC_SMEM_BYTES, st>>>(pl, bf);
  tc_chain_kernel<<<sms * 8, 256, 0, st>>>(bf, s1, s2);
  tc_final_kernel<<<idivup(n1, 64), 64, 0, st>>>(pl, bf, s1, s2);
  count_launch(6);
  CS_CUDA(cudaGetLastError());
  if ((r = match_exact_rows(s1, s2, n2, bf.fbRows, bf.counters, st)) < 0) return r;
  // inputs FP16 cannot bound (|x| >= 32768, NaN, Inf) or a queue overflow: redo everything exactly
  CS_CUDA(cudaMemcpyAsync(ws.h_counters, bf.bmax, 64, cudaMemcpyDeviceToHost, st));
  CS_CUDA(cudaStreamSynchronize(st));
  if (ws.h_counters[1] != 0 || ws.h_counters[4 + 3] != 0) return match_exact(s1, n1, s2, n2, st);
  stats[0] = ws.h_counters[4 + 1];
  stats[1] = ws.h_counters[4 + 2];
  stats[2] = ws.h_counters[4 + 0];
  return 0;
}

}  // namespace cs
