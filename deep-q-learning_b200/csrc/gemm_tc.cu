// gemm_tc.cu -- tcgen05 (5th-gen tensor core) GEMMs for the large-batch path, 3xTF32 error-compensated.
//
//   C[M,N] = op(A) * op(B)   with  a = a_hi + a_lo (a_hi = tf32(a), a_lo = a - a_hi exactly), same for b:
//   C += a_hi*b_lo + a_lo*b_hi + a_hi*b_hi   (three kind::tf32 MMAs per k-step, fp32 accumulation in TMEM)
//
// which restores ~2^-21 relative accuracy per product -- the fp32 parity bar -- at 1/3 of the TF32 rate.
// One CTA computes a 128 x 128 tile with one tcgen05.mma (M=128, N=128, K=8) per 8 reduction elements:
//   warps 0-3  producers: global fp32 (coalesced) -> registers -> {hi, lo} split -> shared memory in the UMMA
//              canonical layout (K-major no-swizzle / MN-major SWIZZLE_128B_BASE32B) or TMEM, 3- or 4-slot mbarrier
//              ring; the split IS the reason operands are staged by threads instead of TMA.  Up to three stages of
//              loads are in flight in registers (setmaxnreg: 240 regs for this warpgroup)
//   warps 4-7  epilogue: tcgen05.ld of the 128x128 fp32 accumulator (one TMEM lane per output row),
//              fused bias+relu / relu'-mask / split-K partial store
//   warp 8     TMEM allocation + single-thread MMA issue, tcgen05.commit onto the stage / accumulator barriers
//   warps 9-11 idle: they only complete the third warpgroup so that it can give its registers away (setmaxnreg 40)
// K-major A operands never touch shared memory: each producer thread owns one row of the A tile and writes its
// hi / lo values straight into TMEM (tcgen05.st, lane = row), and the MMA takes A from TMEM (.ts form).  A 128x128x8
// tf32 MMA with both operands in shared memory reads 8 KB per 64 cycles = the whole 128 B/clk shared-memory
// bandwidth of the SM (ncu: l1tex 76 %, tensor pipe 17 %); with A in TMEM only B streams through shared memory.
// Operand layouts follow the reference data as it lies in HBM, so no transposes are materialised:
//   NN  h2 = h1 . W2        A K-major  (h1 [M][K]),   B MN-major (W2 [K][N])
//   NT  dh1 = dh2 . W2^T    A K-major  (dh2 [M][K]),  B K-major  (W2 [N][K])
//   TN  dW2 = h1^T . dh2    A MN-major (h1 [K][M]),   B MN-major (dh2 [K][N]), split-K over the batch
#include <cuda.h>
#include <stdio.h>

#include "common.cuh"
#include "kernels.h"
#include "large.h"

namespace dqn {

namespace {

constexpr int TM = 128, TN = 128, TK = 32;     // CTA tile; TK fp32 per pipeline stage (4 MMA k-steps of 8)
constexpr int kTileK = TM * TK * 4;            // bytes of a K-major operand tile: 16 (8-row groups) x 8 (k core matrices) x 128 B
// MN-major fp32/tf32 operands must use the SWIZZLE_128B_BASE32B canonical layout (the only MN-major layout the
// tensor core accepts for 32-bit types): atoms of 4 k-rows x 128 B (32 MN elements), 32-byte chunks XOR-swizzled
// by the k-row.  Tile = [8 k-atoms][4 mn-atoms][512 B].
constexpr int kLboMN = 512;                    // stride between atoms along MN
constexpr int kSboMN = 4 * kLboMN;             // stride between atoms along K (2048 B)
constexpr int kTileMN = (TK / 4) * kSboMN;     // 16384 B
constexpr int kNumThreads = 384;             // 3 warpgroups: producers | epilogue | MMA issuer (+3 idle warps)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// UMMA shared-memory descriptor, version 1 (Blackwell).  layout_type 0 = no swizzle, 1 = SWIZZLE_128B_BASE32B
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout_type << 61);
}
// instruction descriptor: D = F32, A = B = TF32, M = 128, N = 128, majorness per operand
__host__ __device__ constexpr uint32_t make_idesc(bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(TN >> 3) << 17) |
         ((uint32_t)(TM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand in tensor memory (.ts form): D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// registers -> 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,"
      "%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
      "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
      "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
      "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void split_tf32(const float4& x, float4& hi, float4& lo) {
  uint32_t h0, h1, h2, h3;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h0) : "f"(x.x));
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h1) : "f"(x.y));
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h2) : "f"(x.z));
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h3) : "f"(x.w));
  hi = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(h2), __uint_as_float(h3));
  lo = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
}

// Fill one operand tile (hi and lo copies) for the reduction range [k0, k0 + TK).
//   KMAJOR:  G[mn][k] row-major (ld); smem byte (mn, k) = (mn/8)*1024 + (k/4)*128 + (mn%8)*16 + (k%4)*4
//   MNMAJOR: G[k][mn] row-major (ld); smem byte (k, mn) = (k/4)*2048 + (mn/32)*512 + (k%4)*128 + swz32((mn%32)*4, k%4)
// Operand staging is split into a load half (global -> registers) and a store half (registers -> {hi, lo} ->
// TMEM / shared memory) so that the loads of stage kt+1 are in flight while the producer waits for its slot.
//   K-major:  G[mn][k] row-major (ld); thread p owns row mn0 + p (32 consecutive k values = 8 float4)
//             smem byte (mn, k) = (mn/8)*1024 + (k/4)*128 + (mn%8)*16 + (k%4)*4;  TMEM: lane = row, column = k
//   MN-major: G[k][mn] row-major (ld); a warp reads one full 512-byte row segment per k (kb + 4 i)
//             smem byte (k, mn) = (k/4)*2048 + (mn/32)*512 + (k%4)*128 + swz32((mn%32)*4, k%4)
template <bool MNMAJOR>
__device__ __forceinline__ void ld_tile(const float* __restrict__ G, int ld, int mn0, int k0, int p, float4 (&v)[8]) {
  if (!MNMAJOR) {
    // coalesced: one instruction = 4 rows x 128 B (8 lanes per row); row_transpose() turns this into "thread = row"
    const int w = p >> 5, lane = p & 31;
    const float* src = G + (size_t)(mn0 + 32 * w + (lane >> 3)) * ld + k0 + 4 * (lane & 7);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const float4*>(src + (size_t)(4 * i) * ld);
  } else {
    const int mq = p & 31, kb = p >> 5;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const float4*>(G + (size_t)(k0 + kb + 4 * i) * ld + mn0 + 4 * mq);
  }
}

// K-major tiles are loaded coalesced (lane = (row % 4, 16-byte chunk)), but both consumers want "thread p owns row p":
// tcgen05.st.32x32b writes one TMEM lane per thread, and the UMMA K-major layout is written conflict-free that way.
// A row-per-thread global load touches 32 different 128-byte lines per instruction (32 L1 wavefronts instead of 4) and
// made the kernel l1tex-bound (ncu: 70-88 % l1tex).  The transpose goes through a 4 KB per-warp staging area in shared
// memory, 16-byte chunks XOR-swizzled by the row so that both the write and the read are conflict-free.
__device__ __forceinline__ void row_transpose(float4 (&v)[8], uint8_t* stage_warp, int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int R = 4 * i + (lane >> 3), c = lane & 7;
    *reinterpret_cast<float4*>(stage_warp + R * 128 + ((c ^ (R & 7)) << 4)) = v[i];
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 8; ++q) v[q] = *reinterpret_cast<const float4*>(stage_warp + lane * 128 + ((q ^ (lane & 7)) << 4));
  __syncwarp();
}

template <bool MNMAJOR>
__device__ __forceinline__ void st_tile(const float4 (&v)[8], uint8_t* s_hi, uint8_t* s_lo, int p) {
  if (!MNMAJOR) {
    const int base = (p >> 3) * 1024 + (p & 7) * 16;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float4 hi, lo;
      split_tf32(v[q], hi, lo);
      *reinterpret_cast<float4*>(s_hi + base + q * 128) = hi;
      *reinterpret_cast<float4*>(s_lo + base + q * 128) = lo;
    }
  } else {
    const int mq = p & 31, kb = p >> 5;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = kb + 4 * i, u = mq & 7;           // u = 16-byte unit inside the 128-byte row of atom (mq >> 3)
      const int off = (k >> 2) * kSboMN + (mq >> 3) * kLboMN + (k & 3) * 128 + ((((u >> 1) ^ (k & 3)) << 5) | ((u & 1) << 4));
      float4 hi, lo;
      split_tf32(v[i], hi, lo);
      *reinterpret_cast<float4*>(s_hi + off) = hi;
      *reinterpret_cast<float4*>(s_lo + off) = lo;
    }
  }
}

// K-major A rows straight into tensor memory, 16 columns at a time to keep the register peak low
__device__ __forceinline__ void st_a_tmem(const float4 (&v)[8], uint32_t t_hi, uint32_t t_lo, int p) {
  const uint32_t lane_base = (uint32_t)(p & ~31) << 16;     // warp w writes lanes 32w .. 32w+31
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t rh[16], rl[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 h, l;
      split_tf32(v[4 * half + q], h, l);
      rh[4 * q] = __float_as_uint(h.x); rh[4 * q + 1] = __float_as_uint(h.y); rh[4 * q + 2] = __float_as_uint(h.z); rh[4 * q + 3] = __float_as_uint(h.w);
      rl[4 * q] = __float_as_uint(l.x); rl[4 * q + 1] = __float_as_uint(l.y); rl[4 * q + 2] = __float_as_uint(l.z); rl[4 * q + 3] = __float_as_uint(l.w);
    }
    tmem_st16(t_hi + lane_base + 16u * half, rh);
    tmem_st16(t_lo + lane_base + 16u * half, rl);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// tcgen05.ld of 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"
      "%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// The tensor core accumulates in fp32 with truncation, so the error of a long reduction grows linearly
// with K (measured: ~5e-9 * K * |C|).  The accumulator is therefore promoted every kChunkK reduction
// elements: two TMEM accumulators (2 x 128 columns) alternate; while the MMAs of chunk c+1 run, the
// epilogue warps pull chunk c out of TMEM and add it to fp32 registers with round-to-nearest.
constexpr int kChunkK = 128;
constexpr int kStagesPerChunk = kChunkK / TK;

template <bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_tc_kernel(int M, int N, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc,
               const float* __restrict__ aux, int ldaux, int k_per_split) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr bool A_TMEM = !A_MN;                       // K-major A lives in tensor memory, not shared memory
  constexpr int kStages = A_TMEM ? 4 : 3;
  constexpr int kATile = A_TMEM ? 0 : kTileMN, kBTile = B_MN ? kTileMN : kTileK;
  constexpr int kStageBytes = 2 * kATile + 2 * kBTile;
  constexpr uint32_t kTmemCols = A_TMEM ? 512u : 256u; // 2 accumulators x 128 (+ 4 stages x (32 hi + 32 lo) columns of A)
  uint8_t* const tiles = smem;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);   // full[S], empty[S], tfull[2], tempty[2]
  uint64_t* const full = bars, *const empty = bars + kStages, *const tfull = bars + 2 * kStages, *const tempty = bars + 2 * kStages + 2;
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int kbeg = blockIdx.z * k_per_split;
  const int nk = k_per_split / TK;
  const int nchunks = nk / kStagesPerChunk;
  if (EPI == kEpiSplitK) C += (size_t)blockIdx.z * M * ldc;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&full[s]), 128);              // every producer thread arrives
      mbar_init(smem_u32(&empty[s]), 1);               // one tcgen05.commit
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&tfull[b]), 1);               // accumulator chunk complete (tcgen05.commit)
      mbar_init(smem_u32(&tempty[b]), 128);            // every epilogue thread has read it
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {                                     // TMEM: 2 accumulators x 128 fp32 columns x 128 lanes (+ the A stages)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp < 4) {
    // ===================== producers =====================
    // Three register sets rotate: the global loads of stage kt+3 are issued as soon as stage kt has been staged, so
    // ~3 stages (96 KB per SM) of loads are in flight and their latency (~1 us under load) overlaps the split / store
    // work.  The registers come from the other warpgroups (setmaxnreg): the MMA warpgroup needs almost none.
    asm volatile("setmaxnreg.inc.sync.aligned.u32 240;");
    float4 va0[8], vb0[8], va1[8], vb1[8], va2[8], vb2[8];
    uint8_t* const xpose = reinterpret_cast<uint8_t*>(tmem_slot) + 64 + warp * 4096;   // per-warp transpose staging
    auto stage_out = [&](int kt, float4 (&va)[8], float4 (&vb)[8]) {
      const int s = kt % kStages;
      if constexpr (!A_MN) row_transpose(va, xpose, lane);
      if constexpr (!B_MN) row_transpose(vb, xpose, lane);
      if (kt >= kStages) mbar_wait(smem_u32(&empty[s]), ((kt / kStages) - 1) & 1);   // slot drained by the MMAs
      uint8_t* st = tiles + s * kStageBytes;
      if constexpr (A_TMEM) st_a_tmem(va, tmem_d + 256u + (uint32_t)(s * 64), tmem_d + 256u + (uint32_t)(s * 64 + 32), tid);
      else st_tile<true>(va, st, st + kATile, tid);
      st_tile<B_MN>(vb, st + 2 * kATile, st + 2 * kATile + kBTile, tid);
      fence_async_smem();                              // generic-proxy stores -> visible to the tensor core (async proxy)
      tc_fence_before();                               // orders the tcgen05.st of A before the barrier hand-off
      mbar_arrive(smem_u32(&full[s]));
    };
    ld_tile<A_MN>(A, lda, m0, kbeg, tid, va0);
    ld_tile<B_MN>(B, ldb, n0, kbeg, tid, vb0);
    if (nk > 1) { ld_tile<A_MN>(A, lda, m0, kbeg + TK, tid, va1); ld_tile<B_MN>(B, ldb, n0, kbeg + TK, tid, vb1); }
    if constexpr (A_TMEM) {            // 4 shared-memory slots: three stages of loads in flight
      if (nk > 2) { ld_tile<A_MN>(A, lda, m0, kbeg + 2 * TK, tid, va2); ld_tile<B_MN>(B, ldb, n0, kbeg + 2 * TK, tid, vb2); }
      for (int kt = 0; kt < nk; kt += 3) {
        stage_out(kt, va0, vb0);
        if (kt + 3 < nk) { ld_tile<A_MN>(A, lda, m0, kbeg + (kt + 3) * TK, tid, va0); ld_tile<B_MN>(B, ldb, n0, kbeg + (kt + 3) * TK, tid, vb0); }
        if (kt + 1 < nk) {
          stage_out(kt + 1, va1, vb1);
          if (kt + 4 < nk) { ld_tile<A_MN>(A, lda, m0, kbeg + (kt + 4) * TK, tid, va1); ld_tile<B_MN>(B, ldb, n0, kbeg + (kt + 4) * TK, tid, vb1); }
        }
        if (kt + 2 < nk) {
          stage_out(kt + 2, va2, vb2);
          if (kt + 5 < nk) { ld_tile<A_MN>(A, lda, m0, kbeg + (kt + 5) * TK, tid, va2); ld_tile<B_MN>(B, ldb, n0, kbeg + (kt + 5) * TK, tid, vb2); }
        }
      }
    } else {                           // both operands in shared memory (3 slots): two stages in flight measure faster
      for (int kt = 0; kt < nk; kt += 2) {
        stage_out(kt, va0, vb0);
        if (kt + 2 < nk) { ld_tile<A_MN>(A, lda, m0, kbeg + (kt + 2) * TK, tid, va0); ld_tile<B_MN>(B, ldb, n0, kbeg + (kt + 2) * TK, tid, vb0); }
        if (kt + 1 < nk) {
          stage_out(kt + 1, va1, vb1);
          if (kt + 3 < nk) { ld_tile<A_MN>(A, lda, m0, kbeg + (kt + 3) * TK, tid, va1); ld_tile<B_MN>(B, ldb, n0, kbeg + (kt + 3) * TK, tid, vb1); }
        }
      }
    }
  } else if (warp >= 8) {
    // ===================== MMA issuer (one thread of warp 8; warps 9-11 only donate their registers) =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 8 && lane == 0) {
      constexpr uint32_t idesc = make_idesc(A_MN, B_MN);
      for (int kt = 0; kt < nk; ++kt) {
        const int s = kt % kStages;
        const int chunk = kt / kStagesPerChunk, b = chunk & 1;
        const bool chunk_start = (kt % kStagesPerChunk) == 0;
        if (chunk_start && chunk >= 2) {               // the epilogue must have drained this accumulator
          mbar_wait(smem_u32(&tempty[b]), ((chunk >> 1) - 1) & 1);
          tc_fence_after();
        }
        mbar_wait(smem_u32(&full[s]), (kt / kStages) & 1);
        tc_fence_after();
        const uint32_t acc = tmem_d + (uint32_t)(b * TN);
        const uint32_t a_hi = smem_u32(tiles + s * kStageBytes), a_lo = a_hi + kATile;
        const uint32_t b_hi = a_hi + 2 * kATile, b_lo = b_hi + kBTile;
        const uint32_t ta_hi = tmem_d + 256u + (uint32_t)(s * 64), ta_lo = ta_hi + 32u;
#pragma unroll
        for (int ks = 0; ks < TK / 8; ++ks) {
          // K-major B (no swizzle): 2 core matrices (256 B) per k-step, LBO = 128 (next k core matrix), SBO = 1024 (next 8 rows)
          // MN-major (128B_BASE32B): 2 k-atoms (4096 B) per k-step, LBO = 512 (next mn atom), SBO = 2048 (next k atom)
          const uint32_t boff = B_MN ? ks * 2 * kSboMN : ks * 256;
          const uint64_t dbh = B_MN ? make_desc(b_hi + boff, kLboMN, kSboMN, 1) : make_desc(b_hi + boff, 128, 1024, 0);
          const uint64_t dbl = B_MN ? make_desc(b_lo + boff, kLboMN, kSboMN, 1) : make_desc(b_lo + boff, 128, 1024, 0);
          const uint32_t first = (chunk_start && ks == 0) ? 0u : 1u;
          if constexpr (A_TMEM) {                      // A from tensor memory: 8 columns per k-step
            umma_tf32_ts(acc, ta_hi + 8u * ks, dbl, idesc, first);                  // small cross terms first
            umma_tf32_ts(acc, ta_lo + 8u * ks, dbh, idesc, 1u);
            umma_tf32_ts(acc, ta_hi + 8u * ks, dbh, idesc, 1u);
          } else {
            const uint32_t aoff = ks * 2 * kSboMN;
            const uint64_t dah = make_desc(a_hi + aoff, kLboMN, kSboMN, 1), dal = make_desc(a_lo + aoff, kLboMN, kSboMN, 1);
            umma_tf32(acc, dah, dbl, idesc, first);
            umma_tf32(acc, dal, dbh, idesc, 1u);
            umma_tf32(acc, dah, dbh, idesc, 1u);
          }
        }
        umma_commit(smem_u32(&empty[s]));                               // frees the smem slot when these MMAs retire
        if ((kt % kStagesPerChunk) == kStagesPerChunk - 1) umma_commit(smem_u32(&tfull[b]));   // chunk accumulator complete
      }
    }
  } else {
    // ===================== epilogue (warps 4..7 own TMEM lanes 32*(warp-4) .. +31) =====================
    const int q = warp - 4;
    const int m = m0 + 32 * q + lane;
    float acc[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[j] = 0.f;
    for (int chunk = 0; chunk < nchunks; ++chunk) {
      const int b = chunk & 1;
      mbar_wait(smem_u32(&tfull[b]), (chunk >> 1) & 1);
      tc_fence_after();
      const uint32_t tbase = tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)(b * TN);
#pragma unroll
      for (int c0 = 0; c0 < TN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tbase + (uint32_t)c0, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(r[j]);      // fp32 round-to-nearest promotion
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&tempty[b]));
    }
    float* crow = C + (size_t)m * ldc + n0;
#pragma unroll
    for (int j = 0; j < TN; j += 4) {
      float4 v = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
      if (EPI == kEpiBiasRelu) {
        const float4 bb = *reinterpret_cast<const float4*>(aux + n0 + j);
        v.x = fmaxf(v.x + bb.x, 0.f); v.y = fmaxf(v.y + bb.y, 0.f); v.z = fmaxf(v.z + bb.z, 0.f); v.w = fmaxf(v.w + bb.w, 0.f);
      } else if (EPI == kEpiReluMask) {
        const float4 h = *reinterpret_cast<const float4*>(aux + (size_t)m * ldaux + n0 + j);
        v.x = h.x > 0.f ? v.x : 0.f; v.y = h.y > 0.f ? v.y : 0.f; v.z = h.z > 0.f ? v.z : 0.f; v.w = h.w > 0.f ? v.w : 0.f;
      }
      *reinterpret_cast<float4*>(crow + j) = v;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kTmemCols) : "memory");
  }
}

// =====================================================================================================================
// cta_group::2 form of the NN products (A K-major in tensor memory, B MN-major in shared memory): two CTAs of a cluster
// -- an SM pair -- compute a 256 x 128 tile with ONE tcgen05.mma stream issued by the leader CTA (M = 256).  Each CTA
// stages its own 128 rows of A into its own tensor memory and only HALF of the B tile (64 of the 128 columns) into its own
// shared memory; the tensor cores of both SMs read both halves.  Per SM that halves the producers' B work (loads, hi / lo
// split, shared-memory stores) and the shared-memory bytes each MMA reads -- the single-CTA kernel is bound by exactly
// those (ncu: l1tex 55 %, tensor pipe 42 %, the MMA warp waiting for the producers).
//   full[s]   (leader's)  one arrival per producer warp of BOTH CTAs (remote mbarrier.arrive for the peer's): count 8
//   empty[s]  (both)      tcgen05.commit.cta_group::2 ... multicast::cluster -> the slot is free in both CTAs
//   tfull[b]  (both)      same multicast commit: accumulator chunk b complete in both CTAs' tensor memory
//   tempty[b] (leader's)  one arrival per epilogue warp of both CTAs: count 8
// =====================================================================================================================
constexpr int TN2 = TN / 2;                    // B columns staged per CTA
constexpr int kSbo2 = (TN2 / 32) * kLboMN;     // stride between k atoms of a 64-column MN-major tile: 1024 B
constexpr int kTileMN2 = (TK / 4) * kSbo2;     // 8192 B

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cta_addr, uint32_t rank) {   // arrive on CTA `rank`'s copy of the barrier
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar_cta_addr), "r"(rank) : "memory");
}
__device__ __forceinline__ void umma2_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc),
      "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}
__device__ __forceinline__ void umma2_commit(uint32_t bar) {      // arrives on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc2(bool b_mn) {   // D = F32, A = B = TF32, M = 256, N = 128, A K-major
  return (1u << 4) | (2u << 7) | (2u << 10) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

// half B tile (32 k x 64 columns) of this CTA: G[k][mn] row-major; thread p: float4 column mq = p & 15, k = (p >> 4) + 8 i
__device__ __forceinline__ void ld_b_half(const float* __restrict__ G, int ld, int n0, int k0, int p, float4 (&v)[4]) {
  const int mq = p & 15, kb = p >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = *reinterpret_cast<const float4*>(G + (size_t)(k0 + kb + 8 * i) * ld + n0 + 4 * mq);
}
__device__ __forceinline__ void st_b_half(const float4 (&v)[4], uint8_t* s_hi, uint8_t* s_lo, int p) {
  const int mq = p & 15, kb = p >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = kb + 8 * i, u = mq & 7;
    const int off = (k >> 2) * kSbo2 + (mq >> 3) * kLboMN + (k & 3) * 128 + ((((u >> 1) ^ (k & 3)) << 5) | ((u & 1) << 4));
    float4 hi, lo;
    split_tf32(v[i], hi, lo);
    *reinterpret_cast<float4*>(s_hi + off) = hi;
    *reinterpret_cast<float4*>(s_lo + off) = lo;
  }
}

template <int EPI>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_tc2_kernel(int M, int N, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc,
                const float* __restrict__ aux, int ldaux, int K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int kStages = 4;
  constexpr int kStageBytes = 2 * kTileMN2;            // b_hi | b_lo of this CTA's 64 columns
  constexpr uint32_t kTmemCols = 512u;                 // 2 accumulators x 128 + 4 stages x (32 hi + 32 lo) columns of A
  uint8_t* const tiles = smem;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* const full = bars, *const empty = bars + kStages, *const tfull = bars + 2 * kStages, *const tempty = bars + 2 * kStages + 2;
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();             // 0 = leader (issues the MMAs)
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;   // grid.x runs over the row tiles: a cluster (2, 1, 1) is two consecutive row tiles
  const int nb0 = n0 + (int)rank * TN2;                // this CTA's half of the B tile
  const int nk = K / TK;
  const int nchunks = nk / kStagesPerChunk;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&full[s]), 8);                // 4 producer warps x 2 CTAs (only the leader's copy is used)
      mbar_init(smem_u32(&empty[s]), 1);               // one multicast tcgen05.commit
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&tfull[b]), 1);               // one multicast tcgen05.commit
      mbar_init(smem_u32(&tempty[b]), 8);              // 4 epilogue warps x 2 CTAs (leader's copy)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {                                     // the same warp of both CTAs allocates collectively
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");   // both CTAs' barriers and TMEM exist
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp < 4) {
    // ===================== producers (both CTAs): own 128 rows of A -> TMEM, own 64 columns of B -> shared memory =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 240;");
    float4 va0[8], va1[8], va2[8], vb0[4], vb1[4], vb2[4];
    uint8_t* const xpose = reinterpret_cast<uint8_t*>(tmem_slot) + 64 + warp * 4096;
    auto stage_out = [&](int kt, float4 (&va)[8], float4 (&vb)[4]) {
      const int s = kt % kStages;
      row_transpose(va, xpose, lane);
      if (kt >= kStages) mbar_wait(smem_u32(&empty[s]), ((kt / kStages) - 1) & 1);
      uint8_t* st = tiles + s * kStageBytes;
      st_a_tmem(va, tmem_d + 256u + (uint32_t)(s * 64), tmem_d + 256u + (uint32_t)(s * 64 + 32), tid);
      st_b_half(vb, st, st + kTileMN2, tid);
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(smem_u32(&full[s]), 0u);
    };
    ld_tile<false>(A, lda, m0, 0, tid, va0); ld_b_half(B, ldb, nb0, 0, tid, vb0);
    if (nk > 1) { ld_tile<false>(A, lda, m0, TK, tid, va1); ld_b_half(B, ldb, nb0, TK, tid, vb1); }
    if (nk > 2) { ld_tile<false>(A, lda, m0, 2 * TK, tid, va2); ld_b_half(B, ldb, nb0, 2 * TK, tid, vb2); }
    for (int kt = 0; kt < nk; kt += 3) {
      stage_out(kt, va0, vb0);
      if (kt + 3 < nk) { ld_tile<false>(A, lda, m0, (kt + 3) * TK, tid, va0); ld_b_half(B, ldb, nb0, (kt + 3) * TK, tid, vb0); }
      if (kt + 1 < nk) {
        stage_out(kt + 1, va1, vb1);
        if (kt + 4 < nk) { ld_tile<false>(A, lda, m0, (kt + 4) * TK, tid, va1); ld_b_half(B, ldb, nb0, (kt + 4) * TK, tid, vb1); }
      }
      if (kt + 2 < nk) {
        stage_out(kt + 2, va2, vb2);
        if (kt + 5 < nk) { ld_tile<false>(A, lda, m0, (kt + 5) * TK, tid, va2); ld_b_half(B, ldb, nb0, (kt + 5) * TK, tid, vb2); }
      }
    }
  } else if (warp >= 8) {
    // ===================== MMA issuer: one thread of the LEADER CTA =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (rank == 0 && warp == 8 && lane == 0) {
      constexpr uint32_t idesc = make_idesc2(true);
      for (int kt = 0; kt < nk; ++kt) {
        const int s = kt % kStages;
        const int chunk = kt / kStagesPerChunk, b = chunk & 1;
        const bool chunk_start = (kt % kStagesPerChunk) == 0;
        if (chunk_start && chunk >= 2) {
          mbar_wait(smem_u32(&tempty[b]), ((chunk >> 1) - 1) & 1);
          tc_fence_after();
        }
        mbar_wait(smem_u32(&full[s]), (kt / kStages) & 1);
        tc_fence_after();
        const uint32_t acc = tmem_d + (uint32_t)(b * TN);
        const uint32_t b_hi = smem_u32(tiles + s * kStageBytes), b_lo = b_hi + kTileMN2;
        const uint32_t ta_hi = tmem_d + 256u + (uint32_t)(s * 64), ta_lo = ta_hi + 32u;
#pragma unroll
        for (int ks = 0; ks < TK / 8; ++ks) {
          const uint32_t boff = ks * 2 * kSbo2;          // two k atoms per k-step of 8
          const uint64_t dbh = make_desc(b_hi + boff, kLboMN, kSbo2, 1), dbl = make_desc(b_lo + boff, kLboMN, kSbo2, 1);
          const uint32_t first = (chunk_start && ks == 0) ? 0u : 1u;
          umma2_tf32_ts(acc, ta_hi + 8u * ks, dbl, idesc, first);
          umma2_tf32_ts(acc, ta_lo + 8u * ks, dbh, idesc, 1u);
          umma2_tf32_ts(acc, ta_hi + 8u * ks, dbh, idesc, 1u);
        }
        umma2_commit(smem_u32(&empty[s]));
        if ((kt % kStagesPerChunk) == kStagesPerChunk - 1) umma2_commit(smem_u32(&tfull[b]));
      }
    }
  } else {
    // ===================== epilogue (both CTAs): own 128 rows x 128 columns =====================
    const int q = warp - 4;
    const int m = m0 + 32 * q + lane;
    float acc[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[j] = 0.f;
    for (int chunk = 0; chunk < nchunks; ++chunk) {
      const int b = chunk & 1;
      mbar_wait(smem_u32(&tfull[b]), (chunk >> 1) & 1);
      tc_fence_after();
      const uint32_t tbase = tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)(b * TN);
#pragma unroll
      for (int c0 = 0; c0 < TN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tbase + (uint32_t)c0, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(r[j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(smem_u32(&tempty[b]), 0u);
    }
    float* crow = C + (size_t)m * ldc + n0;
#pragma unroll
    for (int j = 0; j < TN; j += 4) {
      float4 v = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
      if (EPI == kEpiBiasRelu) {
        const float4 bb = *reinterpret_cast<const float4*>(aux + n0 + j);
        v.x = fmaxf(v.x + bb.x, 0.f); v.y = fmaxf(v.y + bb.y, 0.f); v.z = fmaxf(v.z + bb.z, 0.f); v.w = fmaxf(v.w + bb.w, 0.f);
      } else if (EPI == kEpiReluMask) {
        const float4 h = *reinterpret_cast<const float4*>(aux + (size_t)m * ldaux + n0 + j);
        v.x = h.x > 0.f ? v.x : 0.f; v.y = h.y > 0.f ? v.y : 0.f; v.z = h.z > 0.f ? v.z : 0.f; v.w = h.w > 0.f ? v.w : 0.f;
      }
      *reinterpret_cast<float4*>(crow + j) = v;
    }
  }

  tc_fence_before();
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");   // nobody frees TMEM / leaves while the pair still works
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kTmemCols) : "memory");
  }
}

template <int EPI>
cudaError_t launch_tc2(cudaStream_t st, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                       const float* aux, int ldaux) {
  constexpr int smem = 4 * (2 * kTileMN2) + 256 + 4 * 4096;   // stages | barriers | transpose staging
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc2_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(M / TM, N / TN, 1);                      // cluster (2, 1, 1): CTAs (2j, y) and (2j + 1, y) are a pair
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (getenv("DQN_B200_GEMM_DEBUG")) {
    int ncl = -1;
    cudaError_t qe = cudaOccupancyMaxActiveClusters(&ncl, gemm_tc2_kernel<EPI>, &cfg);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, gemm_tc2_kernel<EPI>);
    fprintf(stderr, "gemm_tc2: max active clusters %d (%s), regs %d, static smem %zu, max dyn smem %d, grid %u x %u\n", ncl, cudaGetErrorString(qe),
            fa.numRegs, fa.sharedSizeBytes, fa.maxDynamicSharedSizeBytes, cfg.gridDim.x, cfg.gridDim.y);
  }
  return cudaLaunchKernelEx(&cfg, gemm_tc2_kernel<EPI>, M, N, A, lda, B, ldb, C, ldc, aux, ldaux, K);
}

// =====================================================================================================================
// The same SM-pair kernel with the RAW fp32 tiles brought in by TMA (cp.async.bulk.tensor) into a deep shared-memory ring:
//   warp 9      one thread per CTA issues, per k-stage, two tensor-map loads completing on raw_full[rs] (24 KB):
//                 A  box {32 k, 128 rows} fp32, SWIZZLE_128B  -- lands in exactly the XOR-swizzled row layout from which a
//                    thread reads "its" row conflict-free (what row_transpose() built with 8 stores + a warp barrier);
//                 B  box {64 n, 32 k} fp32, no swizzle        -- row-major, read back as float4 (k, 4 columns);
//   warps 0-3   "transformers": raw tile -> registers -> {hi, lo} tf32 split -> tensor memory (A) / UMMA shared layout (B).
// With the loads off the producer threads' registers, six stages (144 KB per SM) are in flight instead of three register
// sets (72 KB): on the step's real operands (537 MB of activations streamed from HBM, not L2-resident like the 64 MB of the
// stand-alone check) the producers' wait for global loads was what the MMA warp waited for.
// =====================================================================================================================
constexpr int kRawStages = 6;
constexpr int kRawA = TM * TK * 4;             // 16384 B
constexpr int kRawB = TK * TN2 * 4;            // 8192 B
constexpr int kRawBytes = kRawA + kRawB;       // 24576 B per stage

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}

// HEAD (with EPI = bias + relu): the epilogue also evaluates the dueling head's partial sums over the tile's 128 columns --
// hp[(column tile * M + row) * 8 + c] = sum_j relu(...)[row][j] * Wh[j][c], c = 0 (V), 1..A (advantages) -- so that h2 is
// never read back for the head (dddqn.py:29-31; mlp_large.cu sums the column tiles in order), and only rows < store_rows
// of C are written at all (h2 of the s' rows feeds nothing but the head).
template <int EPI, bool HEAD>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_tc3_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int M, int N, float* __restrict__ C, int ldc,
                const float* __restrict__ aux, int ldaux, int K, const float* __restrict__ head_w, int head_a, float* __restrict__ hp,
                int store_rows) {
  // PERSISTENT: cluster c of NC walks the 256 x 128 tiles t = c, c + NC, ... with t = (row pair) * (N / 128) + (column tile), so
  // the clusters running at the same time share rows of A (L2 hits) and the k-stage / accumulator-chunk counters simply
  // run on across tiles: while the epilogue warps drain the last chunk of a tile and store C, the TMA thread, the
  // transformers and the MMA thread are already two chunks into the next tile (the other accumulator buffer).
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int kStages = 4;
  constexpr int kStageBytes = 2 * kTileMN2;            // b_hi | b_lo of this CTA's 64 columns
  constexpr uint32_t kTmemCols = 512u;
  uint8_t* const raw = smem;                                                  // [kRawStages][A raw 16 KB | B raw 8 KB], 1 KB aligned
  uint8_t* const tiles = smem + kRawStages * kRawBytes;                       // [kStages][b_hi | b_lo]
  uint64_t* const bars = reinterpret_cast<uint64_t*>(tiles + kStages * kStageBytes);
  uint64_t* const full = bars, *const empty = bars + kStages, *const tfull = bars + 2 * kStages, *const tempty = bars + 2 * kStages + 2;
  uint64_t* const raw_full = bars + 2 * kStages + 4, *const raw_empty = raw_full + kRawStages;
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(raw_empty + kRawStages);
  float* const headw_s = reinterpret_cast<float*>(tiles + kStages * kStageBytes + 512);     // HEAD: [128 columns][8] head weights of the tile

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int nk = K / TK;
  const int cpt = nk / kStagesPerChunk;                // accumulator chunks per tile
  const int ntn = N / TN;
  const int ntiles = (M / (2 * TM)) * ntn;
  const int nclusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
  const int my_tiles = cluster_id < ntiles ? (ntiles - cluster_id + nclusters - 1) / nclusters : 0;
  const int total_kt = my_tiles * nk;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(smem_u32(&full[s]), 8); mbar_init(smem_u32(&empty[s]), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&tfull[b]), 1); mbar_init(smem_u32(&tempty[b]), 8); }
    for (int r = 0; r < kRawStages; ++r) { mbar_init(smem_u32(&raw_full[r]), 1); mbar_init(smem_u32(&raw_empty[r]), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp < 4) {
    // ===================== transformers =====================
    for (int g = 0; g < total_kt; ++g) {
      const int rs = g % kRawStages, s = g % kStages;
      mbar_wait(smem_u32(&raw_full[rs]), (g / kRawStages) & 1);
      const uint8_t* ra = raw + rs * kRawBytes;
      const uint8_t* rb = ra + kRawA;
      float4 va[8], vb[4];
#pragma unroll
      for (int q = 0; q < 8; ++q) va[q] = *reinterpret_cast<const float4*>(ra + tid * 128 + ((q ^ (tid & 7)) << 4));     // row tid, SWIZZLE_128B
      {
        const int mq = tid & 15, kb = tid >> 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) vb[i] = *reinterpret_cast<const float4*>(rb + (kb + 8 * i) * (TN2 * 4) + mq * 16);
      }
      if (g >= kStages) mbar_wait(smem_u32(&empty[s]), ((g / kStages) - 1) & 1);      // UMMA slot drained by the MMAs
      uint8_t* st = tiles + s * kStageBytes;
      st_a_tmem(va, tmem_d + 256u + (uint32_t)(s * 64), tmem_d + 256u + (uint32_t)(s * 64 + 32), tid);
      st_b_half(vb, st, st + kTileMN2, tid);
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&raw_empty[rs]));                 // this warp has read its part of the raw stage (values are consumed above)
        mbar_arrive_cluster(smem_u32(&full[s]), 0u);
      }
    }
  } else if (warp == 8) {
    // ===================== MMA issuer: one thread of the LEADER CTA =====================
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = make_idesc2(true);
      for (int g = 0; g < total_kt; ++g) {
        const int s = g % kStages;
        const int chunk = g / kStagesPerChunk, b = chunk & 1;
        const bool chunk_start = (g % kStagesPerChunk) == 0;
        if (chunk_start && chunk >= 2) {
          mbar_wait(smem_u32(&tempty[b]), ((chunk >> 1) - 1) & 1);
          tc_fence_after();
        }
        mbar_wait(smem_u32(&full[s]), (g / kStages) & 1);
        tc_fence_after();
        const uint32_t acc = tmem_d + (uint32_t)(b * TN);
        const uint32_t b_hi = smem_u32(tiles + s * kStageBytes), b_lo = b_hi + kTileMN2;
        const uint32_t ta_hi = tmem_d + 256u + (uint32_t)(s * 64), ta_lo = ta_hi + 32u;
#pragma unroll
        for (int ks = 0; ks < TK / 8; ++ks) {
          const uint32_t boff = ks * 2 * kSbo2;
          const uint64_t dbh = make_desc(b_hi + boff, kLboMN, kSbo2, 1), dbl = make_desc(b_lo + boff, kLboMN, kSbo2, 1);
          const uint32_t first = (chunk_start && ks == 0) ? 0u : 1u;
          umma2_tf32_ts(acc, ta_hi + 8u * ks, dbl, idesc, first);
          umma2_tf32_ts(acc, ta_lo + 8u * ks, dbh, idesc, 1u);
          umma2_tf32_ts(acc, ta_hi + 8u * ks, dbh, idesc, 1u);
        }
        umma2_commit(smem_u32(&empty[s]));
        if ((g % kStagesPerChunk) == kStagesPerChunk - 1) umma2_commit(smem_u32(&tfull[b]));
      }
    }
  } else if (warp == 9) {
    // ===================== TMA issuer: one thread of EACH CTA (own rows of A, own half of B) =====================
    if (lane == 0) {
      int g = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int t = cluster_id + i * nclusters;
        const int m0 = ((t / ntn) * 2 + (int)rank) * TM, nb0 = (t % ntn) * TN + (int)rank * TN2;
        for (int kt = 0; kt < nk; ++kt, ++g) {
          const int rs = g % kRawStages;
          if (g >= kRawStages) mbar_wait(smem_u32(&raw_empty[rs]), ((g / kRawStages) - 1) & 1);
          const uint32_t bar = smem_u32(&raw_full[rs]);
          const uint32_t dst = smem_u32(raw + rs * kRawBytes);
          mbar_arrive_expect_tx(bar, (uint32_t)kRawBytes);
          tma_load_2d(dst, &mapA, kt * TK, m0, bar);                 // {k, row}
          tma_load_2d(dst + kRawA, &mapB, nb0, kt * TK, bar);        // {column, k}
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== epilogue (both CTAs): own 128 rows x 128 columns of every tile =====================
    const int q = warp - 4;
    int chunk = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int t = cluster_id + i * nclusters;
      const int m = ((t / ntn) * 2 + (int)rank) * TM + 32 * q + lane, n0 = (t % ntn) * TN;
      float acc[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[j] = 0.f;
      for (int cc = 0; cc < cpt; ++cc, ++chunk) {
        const int b = chunk & 1;
        mbar_wait(smem_u32(&tfull[b]), (chunk >> 1) & 1);
        tc_fence_after();
        const uint32_t tbase = tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)(b * TN);
#pragma unroll
        for (int c0 = 0; c0 < TN; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tbase + (uint32_t)c0, r);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(r[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(smem_u32(&tempty[b]), 0u);
      }
      float* crow = C + (size_t)m * ldc + n0;
      if constexpr (HEAD) {
        // the tile's head weights [column][V, A_1..A_a, 0..] into shared memory (the four epilogue warps, named barrier 1)
        asm volatile("bar.sync 1, 128;" ::: "memory");                     // everybody is done with the previous tile's copy
        {
          const int j = tid - 128;
          const float* wv = head_w;                                         // flat layout: Wv [N] | bv | Wa [N][a] | ba [a]
          const float* wa = head_w + N + 1;
          float w8[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) w8[c] = 0.f;
          w8[0] = wv[n0 + j];
          for (int c = 0; c < head_a; ++c) w8[1 + c] = wa[(size_t)(n0 + j) * head_a + c];
          *reinterpret_cast<float4*>(headw_s + 8 * j) = make_float4(w8[0], w8[1], w8[2], w8[3]);
          *reinterpret_cast<float4*>(headw_s + 8 * j + 4) = make_float4(w8[4], w8[5], w8[6], w8[7]);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      float hacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const bool store = !HEAD || m < store_rows;
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        float4 v = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        if (EPI == kEpiBiasRelu) {
          const float4 bb = *reinterpret_cast<const float4*>(aux + n0 + j);
          v.x = fmaxf(v.x + bb.x, 0.f); v.y = fmaxf(v.y + bb.y, 0.f); v.z = fmaxf(v.z + bb.z, 0.f); v.w = fmaxf(v.w + bb.w, 0.f);
        } else if (EPI == kEpiReluMask) {
          const float4 h = *reinterpret_cast<const float4*>(aux + (size_t)m * ldaux + n0 + j);
          v.x = h.x > 0.f ? v.x : 0.f; v.y = h.y > 0.f ? v.y : 0.f; v.z = h.z > 0.f ? v.z : 0.f; v.w = h.w > 0.f ? v.w : 0.f;
        }
        if constexpr (HEAD) {
          const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float4 w0 = *reinterpret_cast<const float4*>(headw_s + 8 * (j + e)), w1 = *reinterpret_cast<const float4*>(headw_s + 8 * (j + e) + 4);
            hacc[0] = fmaf(vv[e], w0.x, hacc[0]); hacc[1] = fmaf(vv[e], w0.y, hacc[1]); hacc[2] = fmaf(vv[e], w0.z, hacc[2]); hacc[3] = fmaf(vv[e], w0.w, hacc[3]);
            hacc[4] = fmaf(vv[e], w1.x, hacc[4]); hacc[5] = fmaf(vv[e], w1.y, hacc[5]); hacc[6] = fmaf(vv[e], w1.z, hacc[6]); hacc[7] = fmaf(vv[e], w1.w, hacc[7]);
          }
        }
        if (store) *reinterpret_cast<float4*>(crow + j) = v;
      }
      if constexpr (HEAD) {
        float* dst = hp + ((size_t)(t % ntn) * M + m) * 8;
        *reinterpret_cast<float4*>(dst) = make_float4(hacc[0], hacc[1], hacc[2], hacc[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(hacc[4], hacc[5], hacc[6], hacc[7]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kTmemCols) : "memory");
  }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}
// fp32 matrix [rows][cols] with row stride ld (floats): boxes of box_cols x box_rows
bool make_map_2d(CUtensorMap* map, const float* base, int rows, int cols, int ld, int box_cols, int box_rows, bool swizzle128) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int EPI, bool HEAD = false>
cudaError_t launch_tc3(cudaStream_t st, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                       const float* aux, int ldaux, const float* head_w = nullptr, int head_a = 0, float* hp = nullptr, int store_rows = 0) {
  constexpr int smem = kRawStages * kRawBytes + 4 * (2 * kTileMN2) + 512 + (HEAD ? 128 * 8 * 4 : 0);   // raw ring | UMMA B stages | barriers | head weights
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc3_kernel<EPI, HEAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  CUtensorMap mapA, mapB;
  if (!make_map_2d(&mapA, A, M, K, lda, TK, TM, true) || !make_map_2d(&mapB, B, K, N, ldb, TN2, TK, false)) return cudaErrorNotSupported;
  static int sm_pairs = 0;
  if (!sm_pairs) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    sm_pairs = sms / 2 > 0 ? sms / 2 : 1;
  }
  const int ntiles = (M / (2 * TM)) * (N / TN);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * (ntiles < sm_pairs ? ntiles : sm_pairs), 1, 1);      // persistent: one cluster per SM pair
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, gemm_tc3_kernel<EPI, HEAD>, mapA, mapB, M, N, C, ldc, aux, ldaux, K, head_w, head_a, hp, store_rows);
}

// =====================================================================================================================
// TN product (dW2 = h1^T . dh2: both operands MN-major, reduction over the batch rows, split-K) on SM pairs, persistent,
// TMA-fed.  Work item = (256 x 128 tile, k-split); item w = cluster + i * clusters, split-major (w = z * tiles + t), so the
// clusters running together cover all tiles of the same k rows and share them through L2.  A's own 128 columns of h1 and half
// of dh2's 128 columns go raw into a 3-stage ring by TMA ({128 | 64 columns, 32 rows} boxes, row-major); the transformers
// split them into the MN-major SWIZZLE_128B_BASE32B layouts (hi, lo); the leader issues SS-form tcgen05.mma.cta_group::2.
// =====================================================================================================================
constexpr int kRaw4Stages = 3, kUmma4Stages = 3;
constexpr int kRaw4A = TK * TM * 4;                       // 16384 B: [32 k][128 m]
constexpr int kRaw4Bytes = kRaw4A + kRawB;                 // + [32 k][64 n] = 24576 B
constexpr int kStage4Bytes = 2 * kTileMN + 2 * kTileMN2;   // a_hi | a_lo | b_hi | b_lo = 49152 B

__device__ __forceinline__ void umma2_tf32_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc),
      "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}

__global__ void __launch_bounds__(kNumThreads, 1)
gemm_tc4_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int M, int N, float* __restrict__ C, int ldc,
                int k_per_split, int splitk) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr uint32_t kTmemCols = 256u;                 // two accumulators
  uint8_t* const raw = smem;                                                  // [3][A raw 16 KB | B raw 8 KB]
  uint8_t* const tiles = smem + kRaw4Stages * kRaw4Bytes;                     // [3][a_hi | a_lo | b_hi | b_lo]
  uint64_t* const bars = reinterpret_cast<uint64_t*>(tiles + kUmma4Stages * kStage4Bytes);
  uint64_t* const full = bars, *const empty = bars + kUmma4Stages, *const tfull = bars + 2 * kUmma4Stages, *const tempty = tfull + 2;
  uint64_t* const raw_full = tempty + 2, *const raw_empty = raw_full + kRaw4Stages;
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(raw_empty + kRaw4Stages);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int nk = k_per_split / TK;
  const int cpi = nk / kStagesPerChunk;                // accumulator chunks per work item
  const int ntn = N / TN;
  const int ntiles = (M / (2 * TM)) * ntn;
  const int nitems = ntiles * splitk;
  const int nclusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
  const int my_items = cluster_id < nitems ? (nitems - cluster_id + nclusters - 1) / nclusters : 0;
  const int total_kt = my_items * nk;

  if (tid == 0) {
    for (int s = 0; s < kUmma4Stages; ++s) { mbar_init(smem_u32(&full[s]), 8); mbar_init(smem_u32(&empty[s]), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&tfull[b]), 1); mbar_init(smem_u32(&tempty[b]), 8); }
    for (int r = 0; r < kRaw4Stages; ++r) { mbar_init(smem_u32(&raw_full[r]), 1); mbar_init(smem_u32(&raw_empty[r]), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp < 4) {
    // ===================== transformers =====================
    for (int g = 0; g < total_kt; ++g) {
      const int rs = g % kRaw4Stages, s = g % kUmma4Stages;
      mbar_wait(smem_u32(&raw_full[rs]), (g / kRaw4Stages) & 1);
      const uint8_t* ra = raw + rs * kRaw4Bytes;
      const uint8_t* rb = ra + kRaw4A;
      float4 va[8], vb[4];
      {
        const int mq = tid & 31, kb = tid >> 5;          // the (column, k) ownership st_tile<true> expects
#pragma unroll
        for (int i = 0; i < 8; ++i) va[i] = *reinterpret_cast<const float4*>(ra + (kb + 4 * i) * (TM * 4) + mq * 16);
      }
      {
        const int mq = tid & 15, kb = tid >> 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) vb[i] = *reinterpret_cast<const float4*>(rb + (kb + 8 * i) * (TN2 * 4) + mq * 16);
      }
      if (g >= kUmma4Stages) mbar_wait(smem_u32(&empty[s]), ((g / kUmma4Stages) - 1) & 1);
      uint8_t* st = tiles + s * kStage4Bytes;
      st_tile<true>(va, st, st + kTileMN, tid);
      st_b_half(vb, st + 2 * kTileMN, st + 2 * kTileMN + kTileMN2, tid);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&raw_empty[rs]));
        mbar_arrive_cluster(smem_u32(&full[s]), 0u);
      }
    }
  } else if (warp == 8) {
    // ===================== MMA issuer: one thread of the LEADER CTA =====================
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = make_idesc2(true) | (1u << 15);      // A is MN-major too
      for (int g = 0; g < total_kt; ++g) {
        const int s = g % kUmma4Stages;
        const int chunk = g / kStagesPerChunk, b = chunk & 1;
        const bool chunk_start = (g % kStagesPerChunk) == 0;
        if (chunk_start && chunk >= 2) {
          mbar_wait(smem_u32(&tempty[b]), ((chunk >> 1) - 1) & 1);
          tc_fence_after();
        }
        mbar_wait(smem_u32(&full[s]), (g / kUmma4Stages) & 1);
        tc_fence_after();
        const uint32_t acc = tmem_d + (uint32_t)(b * TN);
        const uint32_t a_hi = smem_u32(tiles + s * kStage4Bytes), a_lo = a_hi + kTileMN;
        const uint32_t b_hi = a_hi + 2 * kTileMN, b_lo = b_hi + kTileMN2;
#pragma unroll
        for (int ks = 0; ks < TK / 8; ++ks) {
          const uint32_t aoff = ks * 2 * kSboMN, boff = ks * 2 * kSbo2;
          const uint64_t dah = make_desc(a_hi + aoff, kLboMN, kSboMN, 1), dal = make_desc(a_lo + aoff, kLboMN, kSboMN, 1);
          const uint64_t dbh = make_desc(b_hi + boff, kLboMN, kSbo2, 1), dbl = make_desc(b_lo + boff, kLboMN, kSbo2, 1);
          const uint32_t first = (chunk_start && ks == 0) ? 0u : 1u;
          umma2_tf32_ss(acc, dah, dbl, idesc, first);
          umma2_tf32_ss(acc, dal, dbh, idesc, 1u);
          umma2_tf32_ss(acc, dah, dbh, idesc, 1u);
        }
        umma2_commit(smem_u32(&empty[s]));
        if ((g % kStagesPerChunk) == kStagesPerChunk - 1) umma2_commit(smem_u32(&tfull[b]));
      }
    }
  } else if (warp == 9) {
    // ===================== TMA issuer: one thread of each CTA =====================
    if (lane == 0) {
      int g = 0;
      for (int i = 0; i < my_items; ++i) {
        const int w = cluster_id + i * nclusters;
        const int z = w / ntiles, t = w - z * ntiles;
        const int m0 = ((t / ntn) * 2 + (int)rank) * TM, nb0 = (t % ntn) * TN + (int)rank * TN2;
        const int kbeg = z * k_per_split;
        for (int kt = 0; kt < nk; ++kt, ++g) {
          const int rs = g % kRaw4Stages;
          if (g >= kRaw4Stages) mbar_wait(smem_u32(&raw_empty[rs]), ((g / kRaw4Stages) - 1) & 1);
          const uint32_t bar = smem_u32(&raw_full[rs]);
          const uint32_t dst = smem_u32(raw + rs * kRaw4Bytes);
          mbar_arrive_expect_tx(bar, (uint32_t)kRaw4Bytes);
          tma_load_2d(dst, &mapA, m0, kbeg + kt * TK, bar);              // {column of h1, batch row}
          tma_load_2d(dst + kRaw4A, &mapB, nb0, kbeg + kt * TK, bar);    // {column of dh2, batch row}
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== epilogue: split-K partial of this CTA's 128 x 128 block =====================
    const int q = warp - 4;
    int chunk = 0;
    for (int i = 0; i < my_items; ++i) {
      const int w = cluster_id + i * nclusters;
      const int z = w / ntiles, t = w - z * ntiles;
      const int m = ((t / ntn) * 2 + (int)rank) * TM + 32 * q + lane, n0 = (t % ntn) * TN;
      float acc[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[j] = 0.f;
      for (int cc = 0; cc < cpi; ++cc, ++chunk) {
        const int b = chunk & 1;
        mbar_wait(smem_u32(&tfull[b]), (chunk >> 1) & 1);
        tc_fence_after();
        const uint32_t tbase = tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)(b * TN);
#pragma unroll
        for (int c0 = 0; c0 < TN; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tbase + (uint32_t)c0, r);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(r[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(smem_u32(&tempty[b]), 0u);
      }
      float* crow = C + (size_t)z * M * ldc + (size_t)m * ldc + n0;
#pragma unroll
      for (int j = 0; j < TN; j += 4) *reinterpret_cast<float4*>(crow + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
    }
  }

  tc_fence_before();
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kTmemCols) : "memory");
  }
}

cudaError_t launch_tc4(cudaStream_t st, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, int splitk) {
  constexpr int smem = kRaw4Stages * kRaw4Bytes + kUmma4Stages * kStage4Bytes + 512;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  CUtensorMap mapA, mapB;     // A = h1 [K rows][M columns], B = dh2 [K rows][N columns]
  if (!make_map_2d(&mapA, A, K, M, lda, TM, TK, false) || !make_map_2d(&mapB, B, K, N, ldb, TN2, TK, false)) return cudaErrorNotSupported;
  static int sm_pairs = 0;
  if (!sm_pairs) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    sm_pairs = sms / 2 > 0 ? sms / 2 : 1;
  }
  const int nitems = (M / (2 * TM)) * (N / TN) * splitk;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * (nitems < sm_pairs ? nitems : sm_pairs), 1, 1);
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, gemm_tc4_kernel, mapA, mapB, M, N, C, ldc, K / splitk, splitk);
}

template <bool A_MN, bool B_MN, int EPI>
cudaError_t launch_tc(cudaStream_t st, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                      const float* aux, int ldaux, int splitk) {
  constexpr int kATile = A_MN ? kTileMN : 0, kBTile = B_MN ? kTileMN : kTileK;
  constexpr int smem = (A_MN ? 3 : 4) * (2 * kATile + 2 * kBTile) + 256 + 4 * 4096;   // stages | barriers | transpose staging
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<A_MN, B_MN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid(N / TN, M / TM, splitk);
  gemm_tc_kernel<A_MN, B_MN, EPI><<<grid, kNumThreads, smem, st>>>(M, N, A, lda, B, ldb, C, ldc, aux, ldaux, K / splitk);
  return cudaGetLastError();
}

}  // namespace

cudaError_t lb_gemm_tc_nn_head(cudaStream_t st, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                               const float* bias, const float* head_w, int head_a, float* hp, int store_rows) {
  if (M % (2 * TM) || N % TN || K % kChunkK || lda % 4 || ldb % 4 || (((uintptr_t)A | (uintptr_t)B) & 15) || encode_tiled_fn() == nullptr)
    return cudaErrorNotSupported;
  return launch_tc3<kEpiBiasRelu, true>(st, M, N, K, A, lda, B, ldb, C, ldc, bias, 0, head_w, head_a, hp, store_rows);
}

cudaError_t lb_gemm_tc(cudaStream_t st, int kind, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                       float* C, int ldc, const float* aux, int ldaux, int splitk, const LbWorkspace& ws) {
  (void)ws;
  if (M % TM || N % TN || (K / splitk) % kChunkK) return cudaErrorInvalidValue;
  // NN products on SM pairs (cta_group::2) when the row tiles pair up; DQN_B200_GEMM_CG=1 keeps the single-CTA kernel
  // DQN_B200_GEMM_CG (experiment knob): 1 = single-CTA kernel, 2 = SM pairs with thread-loaded operands, default 3 = SM pairs
  // with the raw tiles brought in by TMA
  static const int cg = [] { const char* e = getenv("DQN_B200_GEMM_CG"); return e ? atoi(e) : 3; }();
  if (cg >= 2 && (M / TM) % 2 == 0 && (kind == kGemmNN_BiasRelu || kind == kGemmNN_ReluMask)) {
    const bool tma_ok = cg >= 3 && lda % 4 == 0 && ldb % 4 == 0 && (((uintptr_t)A | (uintptr_t)B) & 15) == 0 && encode_tiled_fn() != nullptr;
    if (kind == kGemmNN_BiasRelu)
      return tma_ok ? launch_tc3<kEpiBiasRelu>(st, M, N, K, A, lda, B, ldb, C, ldc, aux, ldaux) : launch_tc2<kEpiBiasRelu>(st, M, N, K, A, lda, B, ldb, C, ldc, aux, ldaux);
    return tma_ok ? launch_tc3<kEpiReluMask>(st, M, N, K, A, lda, B, ldb, C, ldc, aux, ldaux) : launch_tc2<kEpiReluMask>(st, M, N, K, A, lda, B, ldb, C, ldc, aux, ldaux);
  }
  if (cg >= 3 && kind == kGemmTN_SplitK && (M / TM) % 2 == 0 && lda % 4 == 0 && ldb % 4 == 0 && (((uintptr_t)A | (uintptr_t)B) & 15) == 0 &&
      encode_tiled_fn() != nullptr)
    return launch_tc4(st, M, N, K, A, lda, B, ldb, C, ldc, splitk);
  switch (kind) {
    case kGemmNN_BiasRelu: return launch_tc<false, true, kEpiBiasRelu>(st, M, N, K, A, lda, B, ldb, C, ldc, aux, ldaux, 1);
    case kGemmNT_ReluMask: return launch_tc<false, false, kEpiReluMask>(st, M, N, K, A, lda, B, ldb, C, ldc, aux, ldaux, 1);
    case kGemmNN_ReluMask: return launch_tc<false, true, kEpiReluMask>(st, M, N, K, A, lda, B, ldb, C, ldc, aux, ldaux, 1);
    case kGemmTN_SplitK: return launch_tc<true, true, kEpiSplitK>(st, M, N, K, A, lda, B, ldb, C, ldc, aux, ldaux, splitk);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace dqn
