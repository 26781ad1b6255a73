// match_tc.cu -- tcgen05 tensor-core matcher with exact FP32 re-score.
//
// Behavioural spec: the results are those of reference FindMaxCorr10 / MatchSiftData
// (matching.cu:301-397, 1090-1206), bit for bit -- see match.cu for the semantics.
// The reference evaluates all N1 x N2 x 128 products on the FP32 SIMT pipe; here the dense
// contraction runs on the 5th-generation tensor cores and only the handful of candidates
// that can decide a row are re-scored with the reference's sequential FMA chain.
//
// Pipeline (3 launches + the exact fallback for rows that cannot be certified):
//   1. prep      both descriptor sets -> split FP16 (hi + lo of the 2^12-scaled value), written
//                directly in the UMMA "interleaved" (no-swizzle, K-major) core-matrix layout so that an
//                operand tile is one contiguous blob = one cp.async.bulk; row norms for the error bound.
//   2. gemm      persistent warp-specialised kernel: cp.async.bulk producer warp, one elected MMA lane
//                (tcgen05.mma kind::f16, M=128 N=128 K=16, A operand in tensor memory, three MMAs per K step:
//                hi*hi + hi*lo + lo*hi, FP32 accumulators in TMEM), 8 epilogue warps read
//                the accumulators with tcgen05.ld and keep, per (row, partition), the three largest
//                4-candidate group maxima.
//   3. resolve   eight lanes per row: merge the segments, certify, exact k=0..127 FMA chains for the
//                surviving candidates, the reference's per-partition update rule and 8-way merge.
// An earlier version (git history) used plain FP16 operands and needed two GEMM passes (maxima, then
// emission of everything within the 27x wider error bound).
#include "common.cuh"

#include <cuda_fp16.h>

namespace cs {

#define TC_KCH 16           // 16-byte chunks (8 halves) per descriptor

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
// one lane of a converged warp (elect.sync): lets ptxas keep the tcgen05 operands in uniform registers
__device__ __forceinline__ bool elect_one()
{
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
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
// A operand taken from tensor memory (128 lanes x 8 columns per K=16 step): the tensor core then fetches
// only B from shared memory, which is what bounds an MMA with a short N (fetch and math do not overlap)
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// shared memory (UMMA descriptor, 128 rows x 32 bytes) -> tensor memory (128 lanes x 8 columns)
__device__ __forceinline__ void tc_cp_128x256b(uint32_t taddr, uint64_t sdesc)
{
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
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
// kind::f16 instruction descriptor (T3_IDESC below): D=F32, A=B=F16, both K-major
// (cute::UMMA::InstrDescriptor: c_format [4,6), n_dim>>3 at [17,23), m_dim>>4 at [24,29)).

// ------------------------------------------------------------------------------ workspace
struct TcPlan {
  int n1, n2v, n_mt, n_nt, total, U, runs, grid;
  // experiments (env CS_TC_DBG, results are wrong when set): 1 = B tiles are not reloaded, 2 = TMEM reads without
  // the top-3 update, 4 = no TMEM reads
  int dbg;
};
// slot of (cta, run, row-in-tile)
__device__ __forceinline__ size_t tc_slot(const TcPlan &pl, int cta, int run, int rowInTile)
{
  return ((size_t)cta * pl.runs + run) * 256 + rowInTile;   // 256 = rows per CTA tile (T3_MT)
}

// ------------------------------------------------------------------------------ single-pass matcher
// Split-FP16 operands (K_eff = 384), top-3 tracking, fused resolve.
// Every descriptor x (scaled by 2^12 so that the low parts stay FP16-normal) is split into
// hi = fp16(x) and lo = fp16(x - hi); the tensor cores accumulate hi*hi + hi*lo + lo*hi in FP32
// (three tcgen05.mma per K=16 step on the same accumulator), which differs from the reference's
// FP32 chain by at most eps = T3_C1*|a|*max|b| -- 27x tighter than plain FP16.  That makes a
// single GEMM pass sufficient: the epilogue keeps, per (row, partition), the three largest
// 4-candidate group maxima (with the ids of the first two); a partition is certified when the
// third lies more than delta = 2*eps below what can still matter, and then only the groups of the
// first (and, if within delta, the second) maximum are re-scored with the exact FMA chain by the
// resolve kernel, which also applies the reference's update rule and 8-way merge.  With K_eff = 384
// the kernel is tensor-pipe bound (12*N cycles of MMA against 8*N cycles of TMEM reads per tile).
#define T3_M 128                    // rows per accumulator (UMMA M)
#define T3_MT 256                   // rows per CTA tile (two accumulators)
#define T3_N 128                    // candidates per stage (UMMA N)
#define T3_A_PART (T3_MT * 256)     // 64 KB: hi (or lo) halves of an A tile, [half][chunk][128 rows][16 B]
#define T3_A_BYTES (2 * T3_A_PART)  // 128 KB
#define T3_B_PART (T3_N * 256)      // 32 KB: [chunk][128 rows][16 B]
#define T3_B_BYTES (2 * T3_B_PART)  // 64 KB
#define T3_STAGES 3                 // B stages; the A tile passes through stages 0+1 on its way to tensor memory
#define T3_THREADS 384
#define T3_SMEM_BYTES (T3_STAGES * T3_B_BYTES + 1024)
static_assert(T3_A_BYTES == 2 * T3_B_BYTES, "the A tile is staged in two B stages");
#define T3_PM 40                    // words per (row, segment): g1[8] g2[8] g3[8] i1[8] i2[8]
#define T3_SCALE 4096.0f            // 2^12 per operand -> scores carry 2^24
#define T3_UNSCALE (1.0f / 16777216.0f)
#define T3_LIMIT 8.0f               // |x| < 8 keeps 2^12*x inside FP16
// dropped lo*lo and residuals 3*2^-22, tensor accumulation 2^-15, reference chain 2^-17
#define T3_C1 4.0e-5f
#define T3_C2 1.0e-9f
#define T3_IDESC ((1u << 4) | ((uint32_t)(T3_N >> 3) << 17) | ((uint32_t)(T3_M >> 4) << 24))

struct T3Buffers {
  __half *a16, *b16;          // packed hi/lo operand tiles
  float *normA;               // per row of set 1
  float *bmax;                // [0] max norm of set 2 (float bits), [1] bad-input flag
  float *pm;                  // per (row, segment) top-3 tables
  int *fbRows;                // fallback row list
  unsigned int *counters;     // [0] fallback rows, [1] candidate groups re-scored, [2] chains
};

// block = 32 rows x 16 chunks; one launch converts both sets (CTAs [0, blocksA) take set 1)
__global__ void __launch_bounds__(512)
t3_prep_kernel(const SiftPoint *__restrict__ ptsA, int nA, __half *__restrict__ outA, float *__restrict__ norms,
               int blocksA, const SiftPoint *__restrict__ ptsB, int nB, __half *__restrict__ outB,
               float *__restrict__ bmax, float *__restrict__ nextFlags)
{
  __shared__ float s_sq[16][33];
  __shared__ int s_bad;
  if (blockIdx.x == 0 && threadIdx.x < 16) nextFlags[threadIdx.x] = 0.0f;   // the next call's flag area
  const int isB = (int)blockIdx.x >= blocksA;
  const SiftPoint *__restrict__ pts = isB ? ptsB : ptsA;
  const int nvalid = isB ? nB : nA;
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
  __half hi[8], lo[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    sq = fmaf(v[i], v[i], sq);
    bad = bad || !(fabsf(v[i]) < T3_LIMIT);     // also catches NaN/Inf
    const float x = v[i] * T3_SCALE;            // exact
    hi[i] = __float2half_rn(x);
    lo[i] = __float2half_rn(x - __half2float(hi[i]));   // the difference is exact in FP32
  }
  if (bad) s_bad = 1;
  uint4 ph, pl;
  ph.x = (uint32_t)__half_as_ushort(hi[0]) | ((uint32_t)__half_as_ushort(hi[1]) << 16);
  ph.y = (uint32_t)__half_as_ushort(hi[2]) | ((uint32_t)__half_as_ushort(hi[3]) << 16);
  ph.z = (uint32_t)__half_as_ushort(hi[4]) | ((uint32_t)__half_as_ushort(hi[5]) << 16);
  ph.w = (uint32_t)__half_as_ushort(hi[6]) | ((uint32_t)__half_as_ushort(hi[7]) << 16);
  pl.x = (uint32_t)__half_as_ushort(lo[0]) | ((uint32_t)__half_as_ushort(lo[1]) << 16);
  pl.y = (uint32_t)__half_as_ushort(lo[2]) | ((uint32_t)__half_as_ushort(lo[3]) << 16);
  pl.z = (uint32_t)__half_as_ushort(lo[4]) | ((uint32_t)__half_as_ushort(lo[5]) << 16);
  pl.w = (uint32_t)__half_as_ushort(lo[6]) | ((uint32_t)__half_as_ushort(lo[7]) << 16);
  if (!isB) {
    const int tile = row / T3_MT, r = row - tile * T3_MT, h = r / T3_M, rr = r - h * T3_M;
    uint8_t *base = reinterpret_cast<uint8_t *>(outA) + (size_t)tile * T3_A_BYTES + ((size_t)(h * TC_KCH + c) * T3_M + rr) * 16;
    *reinterpret_cast<uint4 *>(base) = ph;
    *reinterpret_cast<uint4 *>(base + T3_A_PART) = pl;
  } else {
    const int tile = row / T3_N, r = row - tile * T3_N;
    uint8_t *base = reinterpret_cast<uint8_t *>(outB) + (size_t)tile * T3_B_BYTES + ((size_t)c * T3_N + r) * 16;
    *reinterpret_cast<uint4 *>(base) = ph;
    *reinterpret_cast<uint4 *>(base + T3_B_PART) = pl;
  }
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

__global__ void __launch_bounds__(T3_THREADS, 1)
t3_gemm_kernel(const TcPlan pl, const T3Buffers bf)
{
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *sB = smem;                  // T3_STAGES x 64 KB; stages 0+1 double as the A staging area
  __shared__ uint64_t bar_a_full, bar_a_empty, bar_b_full[T3_STAGES], bar_b_empty[T3_STAGES];
  __shared__ uint64_t bar_acc_full[2], bar_acc_empty[2];      // [row half]
  __shared__ uint64_t bar_drain;
  __shared__ uint32_t s_tmem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x;
  const int u_begin = min(cta * pl.U, pl.total), u_end = min(u_begin + pl.U, pl.total);

  if (threadIdx.x == 0) {
    mbar_init(&bar_a_full, 1); mbar_init(&bar_a_empty, 1); mbar_init(&bar_drain, 1);
    for (int s = 0; s < T3_STAGES; s++) { mbar_init(&bar_b_full[s], 1); mbar_init(&bar_b_empty[s], 1); }
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
  // tensor memory map: columns [0,128) accumulator of rows 0..127, [128,256) of rows 128..255,
  // [256,512) the A tile: [row half][hi|lo][8 K steps][8 columns]

  if (warp == 0) {
    // ============================ producer: bulk copies of whole operand tiles ============================
    if (lane == 0) {
      int runIdx = 0, prevMt = -1;
      int uses[T3_STAGES];
      for (int s = 0; s < T3_STAGES; s++) uses[s] = 0;
      for (int u = u_begin, i = 0; u < u_end; u++, i++) {
        const int mt = u / pl.n_nt, nt = u - mt * pl.n_nt;
        if (mt != prevMt) {
          // the A tile goes through the B stages: they must all have been consumed ...
          for (int s = 0; s < T3_STAGES; s++)
            if (uses[s] > 0) mbar_wait(&bar_b_empty[s], (uses[s] - 1) & 1);
          mbar_expect_tx(&bar_a_full, T3_A_BYTES);
          const uint8_t *srcA = reinterpret_cast<const uint8_t *>(bf.a16) + (size_t)mt * T3_A_BYTES;
          bulk_g2s(sB, srcA, T3_A_PART, &bar_a_full);
          bulk_g2s(sB + T3_A_PART, srcA + T3_A_PART, T3_A_PART, &bar_a_full);
          // ... and are free again once the MMA warp has copied the tile into tensor memory
          mbar_wait(&bar_a_empty, runIdx & 1);
          prevMt = mt; runIdx++;
        }
        const int s = i % T3_STAGES, it = i / T3_STAGES;
        if (it > 0) mbar_wait(&bar_b_empty[s], (it - 1) & 1);
        uses[s] = it + 1;
        if ((pl.dbg & 1) && it > 0) { mbar_arrive(&bar_b_full[s]); continue; }   // experiment: no B traffic
        mbar_expect_tx(&bar_b_full[s], T3_B_BYTES);
        bulk_g2s(sB + s * T3_B_BYTES, reinterpret_cast<const uint8_t *>(bf.b16) + (size_t)nt * T3_B_BYTES, T3_B_BYTES, &bar_b_full[s]);
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    // The whole warp walks the loop (uniform control flow, so descriptors live in uniform registers and
    // an MMA costs a handful of instructions); one elected lane issues the tcgen05 instructions.
    {
      int runIdx = 0, prevMt = -1;
      const uint32_t bBase = smem_u32(sB);
      for (int u = u_begin, i = 0; u < u_end; u++, i++) {
        const int mt = u / pl.n_nt;
        if (mt != prevMt) {
          mbar_wait(&bar_a_full, runIdx & 1);
          tc_fence_after();
          // A tile -> tensor memory.  The MMAs of the previous tile must have drained before their operand is
          // overwritten.
          if (runIdx > 0) {
            if (elect_one()) tc_commit(&bar_drain);
            __syncwarp();
            mbar_wait(&bar_drain, (runIdx - 1) & 1);
            tc_fence_after();
          }
          if (elect_one()) {
#pragma unroll 1
            for (int e = 0; e < 32; e++) {          // e = (half, part, K step)
              const int hh = e >> 4, part = (e >> 3) & 1, k = e & 7;
              const uint64_t sd = umma_desc(bBase + part * T3_A_PART + (uint32_t)(hh * TC_KCH + 2 * k) * (T3_M * 16), T3_M * 16, 128);
              tc_cp_128x256b(tmem + 256 + hh * 128 + part * 64 + k * 8, sd);
            }
            tc_commit(&bar_a_empty);                // the staging area is free as soon as the copies are done
          }
          __syncwarp();
          prevMt = mt; runIdx++;
        }
        const int s = i % T3_STAGES, it = i / T3_STAGES;
        mbar_wait(&bar_b_full[s], it & 1);
        const uint64_t bhi0 = umma_desc(bBase + s * T3_B_BYTES, T3_N * 16, 128);
#pragma unroll
        for (int h = 0; h < 2; h++) {
          // single-buffered accumulators: the epilogue of half h reads while the MMAs of the other half run
          mbar_wait(&bar_acc_empty[h], (i & 1) ^ 1);              // first use passes on a fresh barrier
          tc_fence_after();
          if (elect_one()) {
            const uint32_t d = tmem + (uint32_t)(h * T3_N);
            const uint32_t ta = tmem + 256 + h * 128;             // hi at +0, lo at +64, K step k at +8k
#pragma unroll
            for (int k = 0; k < 8; k++) {
              // K=16 step k = chunks 2k, 2k+1 of the B part [chunk][128 rows][16 B]; advancing the 14-bit
              // start-address field (units of 16 bytes) moves the operand window
              const uint64_t kb = (uint64_t)((2 * k * (T3_N * 16)) >> 4);
              const uint64_t bhi = bhi0 + kb, blo = bhi + (uint64_t)(T3_B_PART >> 4);
              tc_mma_f16_ts(d, ta + 8 * k, bhi, T3_IDESC, k > 0 ? 1u : 0u);
              tc_mma_f16_ts(d, ta + 8 * k, blo, T3_IDESC, 1u);
              tc_mma_f16_ts(d, ta + 64 + 8 * k, bhi, T3_IDESC, 1u);
            }
            tc_commit(&bar_acc_full[h]);
          }
          __syncwarp();
        }
        if (elect_one()) tc_commit(&bar_b_empty[s]);
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ============================ epilogue: 8 warps, one TMEM lane (= row) per thread ============================
    const int h = (warp - 4) >> 2, q = warp & 3;
    const int rowInTile = h * T3_M + q * 32 + lane;
    float g1[8], g2[8], g3[8];
    int i1[8], i2[8];
    int prevMt = -1, runIdx = 0;
    size_t slot = 0;
    auto process = [&](const uint32_t (&r)[32], int gid0) {
#pragma unroll
      for (int j = 0; j < 8; j++) {             // group j (columns 4j..4j+3) belongs to partition j
        const float v = fmaxf(fmaxf(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1])),
                              fmaxf(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])));
        const int id = gid0 + j;
        const bool t = v > g1[j], u2 = v > g2[j];
        g3[j] = fmaxf(g3[j], fminf(g2[j], v));
        i2[j] = u2 ? (t ? i1[j] : id) : i2[j];
        g2[j] = fmaxf(g2[j], fminf(g1[j], v));
        i1[j] = t ? id : i1[j];
        g1[j] = fmaxf(g1[j], v);
      }
    };
    const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * T3_N);
    for (int u = u_begin, i = 0; u < u_end; u++, i++) {
      const int mt = u / pl.n_nt, nt = u - mt * pl.n_nt;
      if (mt != prevMt) {
        slot = tc_slot(pl, cta, runIdx, rowInTile);
        prevMt = mt; runIdx++;
#pragma unroll
        for (int p = 0; p < 8; p++) { g1[p] = 0.0f; g2[p] = 0.0f; g3[p] = 0.0f; i1[p] = -1; i2[p] = -1; }
      }
      mbar_wait(&bar_acc_full[h], i & 1);
      tc_fence_after();
      uint32_t ra[32], rb[32], rc[32], rd[32];
      if (!(pl.dbg & 4)) {                           // experiment (bit 2): no TMEM reads at all
        tc_ld32_issue(tbase, ra);
        tc_ld32_issue(tbase + 32, rb);
        tc_ld32_issue(tbase + 64, rc);
        tc_ld32_issue(tbase + 96, rd);
        tc_ld_wait(ra);
        tc_ld_wait(rb);
        tc_ld_wait(rc);
        tc_ld_wait(rd);
      } else {
#pragma unroll
        for (int e = 0; e < 32; e++) { ra[e] = 0; rb[e] = 0; rc[e] = 0; rd[e] = 0; }
      }
      tc_fence_before();
      mbar_arrive(&bar_acc_empty[h]);                // the accumulator is in registers: hand it back early
      if (!(pl.dbg & 2)) {                           // experiment (bit 1): TMEM reads without the top-3 update
        process(ra, nt * (T3_N / 4));
        process(rb, nt * (T3_N / 4) + 8);
        process(rc, nt * (T3_N / 4) + 16);
        process(rd, nt * (T3_N / 4) + 24);
      }
      const bool lastOfRun = (u + 1 == u_end) || ((u + 1) / pl.n_nt != mt);
      if (lastOfRun) {
        float4 *q4 = reinterpret_cast<float4 *>(bf.pm + slot * T3_PM);
        q4[0] = make_float4(g1[0], g1[1], g1[2], g1[3]); q4[1] = make_float4(g1[4], g1[5], g1[6], g1[7]);
        q4[2] = make_float4(g2[0], g2[1], g2[2], g2[3]); q4[3] = make_float4(g2[4], g2[5], g2[6], g2[7]);
        q4[4] = make_float4(g3[0], g3[1], g3[2], g3[3]); q4[5] = make_float4(g3[4], g3[5], g3[6], g3[7]);
        int4 *qi = reinterpret_cast<int4 *>(q4 + 6);
        qi[0] = make_int4(i1[0], i1[1], i1[2], i1[3]); qi[1] = make_int4(i1[4], i1[5], i1[6], i1[7]);
        qi[2] = make_int4(i2[0], i2[1], i2[2], i2[3]); qi[3] = make_int4(i2[4], i2[5], i2[6], i2[7]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// ---- resolve ---------------------------------------------------------------------------------
// Eight lanes per row (four rows per warp): merge the segments' top-3 tables (lane = partition),
// decide which partitions can matter, re-score their first (and, within delta, second) groups
// exactly (lane = candidate, eight per round), then the reference's per-partition rule
// (matching.cu:354-359, stated order-independently) and 8-way merge (:378-390).
#define T3_RW 4                       // warps per CTA (16 rows)
#define T3_TAB 64                     // exact scores a row can collect (8 partitions x 2 groups x 4)

struct T3ResolveSmem {
  float4 a[T3_RW][4][33];             // the rows' own descriptors (+1: the four rows of a warp hit different banks)
  float sc[T3_RW][4][T3_TAB];         // exact scores
  int p2[T3_RW][4][T3_TAB];           // their indices
  float2 xy[T3_RW][4][T3_TAB];        // their positions (match_xpos / match_ypos)
  int grp[T3_RW][4][16];              // candidate group list
};

__global__ void __launch_bounds__(32 * T3_RW)
t3_resolve_kernel(const TcPlan pl, const T3Buffers bf, SiftPoint *__restrict__ sift1, const SiftPoint *__restrict__ sift2)
{
  __shared__ T3ResolveSmem sm;
  __shared__ unsigned int s_chains;      // one global atomic per CTA: same-address atomics serialise in L2
  if (threadIdx.x == 0) s_chains = 0;
  __syncthreads();
  const unsigned int FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int sub = lane >> 3, p = lane & 7;                   // row within the warp; partition / candidate slot
  const int row = (blockIdx.x * T3_RW + w) * 4 + sub;
  const bool valid = row < pl.n1;
  const int rrow = valid ? row : pl.n1 - 1;                   // clamped: padding lanes mirror the last row
  const int mt = rrow / T3_MT, rowInTile = rrow - mt * T3_MT;
  // the row's descriptor (each lane four float4), needed by every chain
#pragma unroll
  for (int j = 0; j < 4; j++) sm.a[w][sub][p + 8 * j] = __ldg(reinterpret_cast<const float4 *>(sift1[rrow].data) + p + 8 * j);
  const float bmax = __int_as_float(*reinterpret_cast<const int *>(bf.bmax));
  const float normA = bf.normA[rrow];
  // merged top-3 of partition p.  The segments' tables are fetched together (independent loads, one
  // round trip) and inserted afterwards; T3_SEG bounds the unrolled part, more segments take the loop.
  float G = 0.0f, S = 0.0f, Th = 0.0f;
  int I1 = -1, I2 = -1;
  {
    const int u0 = mt * pl.n_nt, u1 = u0 + pl.n_nt - 1;
    const int c0 = u0 / pl.U, c1 = u1 / pl.U;
    auto insert = [&](float v, int id) {
      const bool t = v > G, u2 = v > S;
      Th = fmaxf(Th, fminf(S, v));
      I2 = u2 ? (t ? I1 : id) : I2;
      S = fmaxf(S, fminf(G, v));
      I1 = t ? id : I1;
      G = fmaxf(G, v);
    };
    constexpr int T3_SEG = 6;
    float v[T3_SEG][3];
    int id[T3_SEG][2];
#pragma unroll
    for (int k = 0; k < T3_SEG; k++) {
      const int c = c0 + k;
      if (c <= c1) {
        const int run = mt - (c * pl.U) / pl.n_nt;
        const float *q = bf.pm + tc_slot(pl, c, run, rowInTile) * T3_PM;
        v[k][0] = q[p]; v[k][1] = q[8 + p]; v[k][2] = q[16 + p];
        id[k][0] = __float_as_int(q[24 + p]); id[k][1] = __float_as_int(q[32 + p]);
      } else {
        v[k][0] = v[k][1] = v[k][2] = 0.0f; id[k][0] = id[k][1] = -1;
      }
    }
#pragma unroll
    for (int k = 0; k < T3_SEG; k++) { insert(v[k][0], id[k][0]); insert(v[k][1], id[k][1]); insert(v[k][2], -1); }
    for (int c = c0 + T3_SEG; c <= c1; c++) {
      const int run = mt - (c * pl.U) / pl.n_nt;
      const float *q = bf.pm + tc_slot(pl, c, run, rowInTile) * T3_PM;
      insert(q[p], __float_as_int(q[24 + p])); insert(q[8 + p], __float_as_int(q[32 + p])); insert(q[16 + p], -1);
    }
    G *= T3_UNSCALE; S *= T3_UNSCALE; Th *= T3_UNSCALE;
  }
  const float delta = 2.0f * (T3_C1 * normA * bmax + T3_C2 * (normA + bmax));
  // L = second largest element of the pool {G[0..7], S[0]} (S[0] bounds partition 0's second best, quirk Q9)
  float m1 = __shfl_sync(FULL, S, 0, 8), m2 = 0.0f;
#pragma unroll
  for (int y = 0; y < 8; y++) {
    const float gy = __shfl_sync(FULL, G, y, 8);
    m2 = fmaxf(m2, fminf(m1, gy));
    m1 = fmaxf(m1, gy);
  }
  const float L = m2;
  const bool rowOk = (L > 2.0f * delta) && (L < 3.0e38f);       // the deciding elements are clearly positive
  const bool rel = rowOk && (G >= L - delta);
  const float low = (p == 0 ? fminf(L, G) : G) - delta;        // partition 0 also supplies its second best
  const bool cert = !rel || Th <= 0.0f || Th < low;
  const bool use1 = rel && I1 >= 0;
  const bool use2 = rel && I2 >= 0 && S >= low;
  const unsigned int certMask = (__ballot_sync(FULL, cert) >> (8 * sub)) & 0xffu;
  const bool ok = valid && rowOk && certMask == 0xffu;
  if (valid && !ok && p == 0) {
    // every row lists itself at most once, so the index stays below n1 when the counter started at zero; the clamp
    // keeps a stale counter (a previous call that failed half-way) from writing past the list
    const unsigned at = atomicAdd(&bf.counters[0], 1u);
    if (at < (unsigned)pl.n_mt * T3_MT) bf.fbRows[at] = row;
  }
  // candidate group list (first groups, then second groups)
  const unsigned int b1 = (__ballot_sync(FULL, ok && use1) >> (8 * sub)) & 0xffu;
  const unsigned int b2 = (__ballot_sync(FULL, ok && use2) >> (8 * sub)) & 0xffu;
  const int n1g = __popc(b1), ncand = 4 * (n1g + __popc(b2));
  if (ok && use1) sm.grp[w][sub][__popc(b1 & ((1u << p) - 1))] = I1;
  if (ok && use2) sm.grp[w][sub][n1g + __popc(b2 & ((1u << p) - 1))] = I2;
  __syncwarp();
  int maxc = ncand;
#pragma unroll
  for (int o = 16; o >= 8; o >>= 1) maxc = max(maxc, __shfl_xor_sync(FULL, maxc, o));
  for (int c0 = 0; c0 < maxc; c0 += 8) {
    const int c = c0 + p;
    if (c < ncand) {
      const int p2 = sm.grp[w][sub][c >> 2] * 4 + (c & 3);
      const float4 *bp = reinterpret_cast<const float4 *>(sift2[p2].data);
      const float2 xy = make_float2(sift2[p2].xpos, sift2[p2].ypos);
      // the reference's score: sequential k = 0..127 FMA chain starting from 0 (matching.cu:338-351).
      // Software pipeline: the next eight float4 of the candidate are in flight while eight are consumed.
      float acc = 0.0f;
      float4 bv[8], bn[8];
#pragma unroll
      for (int j = 0; j < 8; j++) bv[j] = __ldg(bp + j);
#pragma unroll
      for (int blk = 0; blk < 4; blk++) {
        if (blk < 3) {
#pragma unroll
          for (int j = 0; j < 8; j++) bn[j] = __ldg(bp + 8 * (blk + 1) + j);
        }
        float4 av[8];
#pragma unroll
        for (int j = 0; j < 8; j++) av[j] = sm.a[w][sub][8 * blk + j];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          acc = __fmaf_rn(av[j].x, bv[j].x, acc);
          acc = __fmaf_rn(av[j].y, bv[j].y, acc);
          acc = __fmaf_rn(av[j].z, bv[j].z, acc);
          acc = __fmaf_rn(av[j].w, bv[j].w, acc);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) bv[j] = bn[j];
      }
      sm.sc[w][sub][c] = acc;
      sm.p2[w][sub][c] = p2;
      sm.xy[w][sub][c] = xy;
    }
  }
  __syncwarp();
  // per-partition rule: lane p scans its row's table for partition p ((p2 >> 2) & 7)
  float pmx = 0.0f, psec = 0.0f;
  int pidx = -1, pent = -1;
  for (int e = 0; e < maxc; e++) {
    if (e < ncand) {
      const int p2 = sm.p2[w][sub][e];
      const float sc = sm.sc[w][sub][e];
      if (((p2 >> 2) & 7) == p && sc > 0.0f) {       // a non-positive score cannot match (matching.cu:317-321)
        if (sc > pmx) { psec = pmx; pmx = sc; pidx = p2; pent = e; }
        else {
          if (sc == pmx && p2 < pidx) { pidx = p2; pent = e; }
          psec = fmaxf(psec, sc);
        }
      }
    }
  }
  // 8-way merge, matching.cu:378-390
  float mx = __shfl_sync(FULL, pmx, 0, 8), sec = __shfl_sync(FULL, psec, 0, 8);
  int idx = __shfl_sync(FULL, pidx, 0, 8), ent = __shfl_sync(FULL, pent, 0, 8);
#pragma unroll
  for (int y = 0; y < 8; y++) {
    const float my = __shfl_sync(FULL, pmx, y, 8);
    const int iy = __shfl_sync(FULL, pidx, y, 8), ey = __shfl_sync(FULL, pent, y, 8);
    if (idx != iy) {
      if (my > mx) { sec = fmaxf(mx, sec); mx = my; idx = iy; ent = ey; }
      else if (my > sec) sec = my;
    }
  }
  if (ok && p == 0) {
    SiftPoint *o = sift1 + row;
    const float2 xy = ent >= 0 ? sm.xy[w][sub][ent] : make_float2(0.0f, 0.0f);
    o->score = mx;
    o->match = idx;
    o->match_xpos = xy.x;
    o->match_ypos = xy.y;
    o->ambiguity = __fdiv_rn(sec, __fadd_rn(mx, 1e-6f));
    atomicAdd(&s_chains, (unsigned)ncand);
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_chains) { atomicAdd(&bf.counters[1], s_chains >> 2); atomicAdd(&bf.counters[2], s_chains); }
}

// exact SIMT scan of the rows the tensor path could not certify (match.cu)
int match_exact_fallback(SiftPoint *s1, int n1, const SiftPoint *s2, int n2, const int *rows, const unsigned int *nrows,
                         const unsigned int *gate, cudaStream_t st);

// ------------------------------------------------------------------------------ host
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

// One workspace per device, used on the caller's stream: like the reference (function-local static caches,
// cudaSiftH.cu:310,409) the matcher is single-threaded per device -- two host threads matching on the same device at
// the same time would share the tables and the flag areas.
struct T3Workspace {
  T3Buffers bf = {};
  size_t cap[8] = {};
  unsigned int *h_counters = nullptr;
  float *flags = nullptr;         // 2 x 16 words: [0] max norm of set 2, [1] bad-input flag, [4..7] counters
  unsigned long long calls = 0;
  bool dirty = false;             // a call returned with an error after it had started to use the current flag area
  bool configured = false;
  bool pendingStats = false;
};
static T3Workspace g_ws3[16];

// stats: [0] candidate groups re-scored, [1] exact chains, [2] rows that fell back to the exact scan
int match_tensor(SiftPoint *s1, int n1, const SiftPoint *s2, int n2, cudaStream_t st, unsigned long long stats[4])
{
  stats[0] = stats[1] = stats[2] = 0;
  const int n2v = (n2 / 32) * 32;                      // matching.cu:325 (quirk Q7)
  if (n1 <= 0) return 0;
  if (n2v == 0 || !match_tensor_supported()) return match_exact(s1, n1, s2, n2, st);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  T3Workspace &ws = g_ws3[dev & 15];
  TcPlan pl;
  pl.n1 = n1; pl.n2v = n2v;
  pl.n_mt = idivup(n1, T3_MT); pl.n_nt = idivup(n2v, T3_N);
  pl.total = pl.n_mt * pl.n_nt;
  pl.grid = pl.total < sms ? pl.total : sms;
  pl.U = idivup(pl.total, pl.grid);
  pl.grid = idivup(pl.total, pl.U);
  pl.runs = idivup(pl.U, pl.n_nt) + 1;
  { const char *e = getenv("CS_TC_DBG"); pl.dbg = e ? atoi(e) : 0; }    // experiments, results are wrong when set
  const size_t slots = (size_t)pl.grid * pl.runs * T3_MT;
  const size_t rowsPad = (size_t)pl.n_mt * T3_MT;
  T3Buffers &bf = ws.bf;
  int r;
  if ((r = ensure((void **)&bf.a16, &ws.cap[0], (size_t)pl.n_mt * T3_A_BYTES)) < 0) return r;
  if ((r = ensure((void **)&bf.b16, &ws.cap[1], (size_t)pl.n_nt * T3_B_BYTES)) < 0) return r;
  if ((r = ensure((void **)&bf.normA, &ws.cap[2], rowsPad * sizeof(float))) < 0) return r;
  if ((r = ensure((void **)&bf.pm, &ws.cap[3], slots * T3_PM * sizeof(float))) < 0) return r;
  if ((r = ensure((void **)&bf.fbRows, &ws.cap[4], rowsPad * sizeof(int))) < 0) return r;
  // flags/counters: two 64-byte areas used alternately; a call's prep kernel clears the other one for the
  // next call, which saves a memset node per call
  if (!ws.flags) {
    size_t dummy = 0;
    if ((r = ensure((void **)&ws.flags, &dummy, 128)) < 0) return r;
    CS_CUDA(cudaMemsetAsync(ws.flags, 0, 128, st));
    CS_CUDA(cudaMallocHost((void **)&ws.h_counters, 64));
  }
  bf.bmax = ws.flags + 16 * (ws.calls & 1);
  float *nextFlags = ws.flags + 16 * ((ws.calls + 1) & 1);
  if (ws.dirty) CS_CUDA(cudaMemsetAsync(ws.flags, 0, 128, st));     // a previous call failed between its launches
  ws.dirty = true;                                                  // until this call's launches are all in
  bf.counters = reinterpret_cast<unsigned int *>(bf.bmax) + 4;
  if (!ws.configured) {
    CS_CUDA(cudaFuncSetAttribute(t3_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T3_SMEM_BYTES));
    ws.configured = true;
  }
  // CUDASIFT_MATCH_TIMING=1: CUDA events between the kernels, printed to stderr (warm caches, unlike ncu)
  static int timing = -1;
  static cudaEvent_t tev[6];
  if (timing < 0) {
    const char *e = getenv("CUDASIFT_MATCH_TIMING");
    timing = (e && atoi(e)) ? 1 : 0;
    if (timing) for (int i = 0; i < 6; i++) cudaEventCreate(&tev[i]);
  }
  if (timing) cudaEventRecord(tev[0], st);
  const int blocksA = pl.n_mt * (T3_MT / 32), blocksB = pl.n_nt * (T3_N / 32);
  if (timing) cudaEventRecord(tev[1], st);
  t3_prep_kernel<<<blocksA + blocksB, 512, 0, st>>>(s1, n1, bf.a16, bf.normA, blocksA, s2, n2v, bf.b16, bf.bmax, nextFlags);
  if (timing) cudaEventRecord(tev[2], st);
  t3_gemm_kernel<<<pl.grid, T3_THREADS, T3_SMEM_BYTES, st>>>(pl, bf);
  if (timing) cudaEventRecord(tev[3], st);
  t3_resolve_kernel<<<idivup(n1, 4 * T3_RW), 32 * T3_RW, 0, st>>>(pl, bf, s1, s2);
  if (timing) cudaEventRecord(tev[4], st);
  count_launch(3);
  CS_CUDA(cudaGetLastError());
  // rows that could not be certified -- or, for inputs the split cannot bound (|x| >= 8, NaN, Inf), every row --
  // go through the exact kernel; which of the two is decided on the device
  if ((r = match_exact_fallback(s1, n1, s2, n2, bf.fbRows, bf.counters, reinterpret_cast<const unsigned int *>(bf.bmax) + 1, st)) < 0) return r;
  if (timing) {
    cudaEventRecord(tev[5], st);
    cudaEventSynchronize(tev[5]);
    float ms[5];
    for (int i = 0; i < 5; i++) cudaEventElapsedTime(&ms[i], tev[i], tev[i + 1]);
    fprintf(stderr, "match timing %dx%d [us]: memset %.1f prep %.1f gemm %.1f resolve %.1f exact-fallback %.1f\n", n1, n2,
            ms[0] * 1e3f, ms[1] * 1e3f, ms[2] * 1e3f, ms[3] * 1e3f, ms[4] * 1e3f);
  }
  CS_CUDA(cudaMemcpyAsync(ws.h_counters, bf.bmax, 64, cudaMemcpyDeviceToHost, st));
  ws.calls++;                           // the prep kernel of this call cleared the other flag area for the next one
  ws.dirty = false;
  ws.pendingStats = true;               // read by match_tensor_stats() after the caller's synchronize
  return 0;
}

// statistics of the last match_tensor on this device (call after the stream was synchronized)
void match_tensor_stats(unsigned long long stats[4])
{
  int dev = 0;
  cudaGetDevice(&dev);
  T3Workspace &ws = g_ws3[dev & 15];
  if (!ws.pendingStats) return;
  ws.pendingStats = false;
  if (ws.h_counters[1] != 0) { stats[0] = stats[1] = 0; stats[2] = ~0ull; return; }   // everything went the exact way
  stats[0] = ws.h_counters[4 + 1];
  stats[1] = ws.h_counters[4 + 2];
  stats[2] = ws.h_counters[4 + 0];
}

}  // namespace cs
