// tile_ops.cuh -- device helpers shared by the fused train-step kernels (train_fused.cu, train_cluster.cu):
// cp.async staging, packed fp32x2 FMAs (FFMA2), outer-product / dot-form register tiles over shared memory.
#pragma once
#include "common.cuh"

#ifndef DQN_FFMA2
#define DQN_FFMA2 1
#endif
#ifndef DQN_OP_UNROLL
#define DQN_OP_UNROLL 8
#endif

namespace dqn {
namespace tile {

constexpr int kOpUnroll = DQN_OP_UNROLL;

__device__ __forceinline__ void cp_async16(float* smem_dst, const void* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}

// ---- packed fp32x2 arithmetic (Blackwell FFMA2: two IEEE fp32 FMAs per lane per issue slot) ----
typedef unsigned long long u64;
struct u64x2 { u64 lo, hi; };
__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void ffma2(u64& d, u64 a, u64 b) {
#if DQN_FFMA2
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
#else
  float dl, dh, al, ah, bl, bh;
  unpack2(d, dl, dh); unpack2(a, al, ah); unpack2(b, bl, bh);
  d = pack2(fmaf(al, bl, dl), fmaf(ah, bh, dh));
#endif
}
__device__ __forceinline__ u64x2 ld2x64(const float* p) {   // one LDS.128 into two 64-bit register pairs
  const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p);
  return u64x2{v.x, v.y};
}
__device__ __forceinline__ u64 ld64(const float* p) { return *reinterpret_cast<const u64*>(p); }

// Outer-product-form tile (forward GEMMs):  C[m0+i][n0+j] (+)= sum_k A[k*lda + i] * Bw[k*ldb + j],
// i < 4, j < TN.  Accumulators are packed pairs along j.  Software-pipelined: the operands of step
// k+1 are loaded before the FMAs of step k issue (row K of both operands is always mapped smem).
template <int TN>
__device__ __forceinline__ void op_tile4(const float* __restrict__ A, int lda, const float* __restrict__ Bw, int ldb,
                                         int K, u64 (&acc)[4][TN / 2]) {
  float4 an = ld4(A);
  u64 bn[TN / 2];
  if constexpr (TN == 2) bn[0] = ld64(Bw);
  else {
#pragma unroll
    for (int q = 0; q < TN / 4; ++q) { const u64x2 v = ld2x64(Bw + 4 * q); bn[2 * q] = v.lo; bn[2 * q + 1] = v.hi; }
  }
#pragma unroll kOpUnroll
  for (int k = 0; k < K; ++k) {
    const float4 a = an;
    u64 b[TN / 2];
#pragma unroll
    for (int q = 0; q < TN / 2; ++q) b[q] = bn[q];
    an = ld4(A + (k + 1) * lda);
    if constexpr (TN == 2) bn[0] = ld64(Bw + (k + 1) * ldb);
    else {
#pragma unroll
      for (int q = 0; q < TN / 4; ++q) { const u64x2 v = ld2x64(Bw + (k + 1) * ldb + 4 * q); bn[2 * q] = v.lo; bn[2 * q + 1] = v.hi; }
    }
    const u64 ad[4] = {pack2(a.x, a.x), pack2(a.y, a.y), pack2(a.z, a.z), pack2(a.w, a.w)};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int q = 0; q < TN / 2; ++q) ffma2(acc[i][q], ad[i], b[q]);
  }
}

// bias-initialised accumulators, relu, k-major store of a 4 x TN tile
template <int TN>
__device__ __forceinline__ void op_init(const float* __restrict__ bias, u64 (&acc)[4][TN / 2]) {
#pragma unroll
  for (int q = 0; q < TN / 2; ++q) {
    const u64 b = ld64(bias + 2 * q);
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][q] = b;
  }
}
template <int TN>
__device__ __forceinline__ void op_store_relu(float* __restrict__ C, int ldc, const u64 (&acc)[4][TN / 2]) {
#pragma unroll
  for (int q = 0; q < TN / 2; ++q) {
    float lo[4], hi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) unpack2(acc[i][q], lo[i], hi[i]);
    st4(C + (2 * q) * ldc, fmaxf(lo[0], 0.f), fmaxf(lo[1], 0.f), fmaxf(lo[2], 0.f), fmaxf(lo[3], 0.f));
    st4(C + (2 * q + 1) * ldc, fmaxf(hi[0], 0.f), fmaxf(hi[1], 0.f), fmaxf(hi[2], 0.f), fmaxf(hi[3], 0.f));
  }
}

// Dot-form tile (gradient GEMMs) over NCH chunks of 4 reduction elements:
//   out[i][j] = sum over chunks c < NCH of dot4(ap[i] + CSTEP*c, bp[j] + CSTEP*c)
// ap / bp are per-thread row base pointers, so every load is [register + immediate].  The packed
// accumulator holds the (even, odd) partial sums; loads of chunk c+1 are issued before the FMAs of c.
template <int TM, int TN, int NCH, int CSTEP = 4>
__device__ __forceinline__ void dot_tile(const float* const (&ap)[TM], const float* const (&bp)[TN], float (&out)[TM][TN]) {
  u64 acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0ull;
  u64x2 an[TM], bn[TN];
#pragma unroll
  for (int i = 0; i < TM; ++i) an[i] = ld2x64(ap[i]);
#pragma unroll
  for (int j = 0; j < TN; ++j) bn[j] = ld2x64(bp[j]);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    u64x2 a[TM], b[TN];
#pragma unroll
    for (int i = 0; i < TM; ++i) a[i] = an[i];
#pragma unroll
    for (int j = 0; j < TN; ++j) b[j] = bn[j];
    if (c + 1 < NCH) {
#pragma unroll
      for (int i = 0; i < TM; ++i) an[i] = ld2x64(ap[i] + CSTEP * (c + 1));
#pragma unroll
      for (int j = 0; j < TN; ++j) bn[j] = ld2x64(bp[j] + CSTEP * (c + 1));
    }
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) ffma2(acc[i][j], a[i].lo, b[j].lo);
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) ffma2(acc[i][j], a[i].hi, b[j].hi);
  }
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) { float lo, hi; unpack2(acc[i][j], lo, hi); out[i][j] = lo + hi; }
}

__device__ __forceinline__ float fast_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}


}  // namespace tile
}  // namespace dqn
