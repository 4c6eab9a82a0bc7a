// mlp_large.cu -- large-batch / wide-hidden dueling double-DQN step (BASELINE configs[3]: B = 65536,
// hidden 1024x1024), data-parallel over GPUs.  Same arithmetic as the fused small-batch kernel
// (q_learning_functions.py:14-64, dddqn.py:24-34), but here the layers are real dense contractions, so
// the step is a sequence of kernels over HBM-resident activations:
//
//   gather (replay.cu) -> layer1 (K = D, elementwise-bound) -> GEMM NN + bias + relu (h2)
//   -> head + dueling -> targets / Huber / d(head) -> fused {dH2, dWh, db2 partials}
//   -> GEMM TN split-K (dW2) -> GEMM NT + relu' mask (dH1) -> fused {dW1, db1 partials}
//   -> [NCCL all-reduce of the flat gradient + loss, done by the caller] -> Adam
//
// The three big GEMM shapes go through launch_gemm(), which dispatches to the fp32 FFMA2 tile kernel in
// this file (exact fp32, the parity baseline) or to the tcgen05 3xTF32 kernel (gemm_tc.cu) when enabled.
#include "common.cuh"
#include "kernels.h"
#include "large.h"

namespace dqn {

namespace {

typedef unsigned long long u64;
__device__ __forceinline__ void ffma2(u64& d, u64 a, u64 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }

// ------------------------------------------------------------------------------------------------
// layer 1: H1[row][:] = relu(x[row] . W1 + b1).  rows [0,B): (theta, s); [B,2B): (theta, s'); [2B,3B): (theta^-, s')
// ------------------------------------------------------------------------------------------------
template <int DT>   // DT = 8: the observation width as a constant (two 16-byte loads per row); 0: run-time D
__global__ void __launch_bounds__(256)
lb_layer1_kernel(const float* __restrict__ s, const float* __restrict__ s2, const float* __restrict__ theta,
                 const float* __restrict__ theta_t, float* __restrict__ H1, int B, int D, int H1n) {
  // A pure HBM writer (3 B x H1 x 4 bytes).  The block's 1024-column slice of [W1; b1] sits in shared memory, so a
  // thread needs ~30 registers and eight blocks stay resident per SM: enough independent 16-byte stores in flight.
  extern __shared__ __align__(16) float w1s[];                 // [D + 1][1024]
  const int col0 = blockIdx.x * 1024;
  const int pass = blockIdx.z;                                 // 0: theta on s, 1: theta on s', 2: theta^- on s'
  const float* W1 = pass == 2 ? theta_t : theta;
  const float* x = pass == 0 ? s : s2;
  const int ncol = min(1024, H1n - col0);
  for (int i = threadIdx.x; i < (D + 1) * (ncol >> 2); i += blockDim.x) {
    const int d = i / (ncol >> 2), c4 = i - d * (ncol >> 2);
    reinterpret_cast<float4*>(w1s)[d * 256 + c4] = *reinterpret_cast<const float4*>(W1 + (size_t)d * H1n + col0 + 4 * c4);   // row D = b1
  }
  __syncthreads();
  const int j4 = threadIdx.x;
  if (4 * j4 >= ncol) return;
  const float4 bias = reinterpret_cast<const float4*>(w1s)[D * 256 + j4];
  // four rows per pass: every 16-byte weight load from shared memory feeds four rows (one row per pass made the kernel
  // LSU-bound at 2.5 TB/s of what is a pure 805 MB write)
  constexpr int RP = 4;
  for (int r0 = blockIdx.y * RP; r0 < B; r0 += gridDim.y * RP) {
    float4 acc[RP];
#pragma unroll
    for (int i = 0; i < RP; ++i) acc[i] = bias;
    if constexpr (DT == 8) {
      float xr[RP][8];
#pragma unroll
      for (int i = 0; i < RP; ++i) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 x0 = r0 + i < B ? __ldg(reinterpret_cast<const float4*>(x + (size_t)(r0 + i) * 8)) : z;
        const float4 x1 = r0 + i < B ? __ldg(reinterpret_cast<const float4*>(x + (size_t)(r0 + i) * 8 + 4)) : z;
        xr[i][0] = x0.x; xr[i][1] = x0.y; xr[i][2] = x0.z; xr[i][3] = x0.w; xr[i][4] = x1.x; xr[i][5] = x1.y; xr[i][6] = x1.z; xr[i][7] = x1.w;
      }
#pragma unroll
      for (int d = 0; d < 8; ++d) {
        const float4 w = reinterpret_cast<const float4*>(w1s)[d * 256 + j4];
#pragma unroll
        for (int i = 0; i < RP; ++i) {
          const float xv = xr[i][d];
          acc[i].x = fmaf(xv, w.x, acc[i].x); acc[i].y = fmaf(xv, w.y, acc[i].y); acc[i].z = fmaf(xv, w.z, acc[i].z); acc[i].w = fmaf(xv, w.w, acc[i].w);
        }
      }
    } else {
      for (int d = 0; d < D; ++d) {
        const float4 w = reinterpret_cast<const float4*>(w1s)[d * 256 + j4];
#pragma unroll
        for (int i = 0; i < RP; ++i) {
          const float xv = r0 + i < B ? __ldg(x + (size_t)(r0 + i) * D + d) : 0.f;
          acc[i].x = fmaf(xv, w.x, acc[i].x); acc[i].y = fmaf(xv, w.y, acc[i].y); acc[i].z = fmaf(xv, w.z, acc[i].z); acc[i].w = fmaf(xv, w.w, acc[i].w);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < RP; ++i) {
      if (r0 + i >= B) break;
      float4 v = acc[i];
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
      *reinterpret_cast<float4*>(H1 + ((size_t)pass * B + r0 + i) * H1n + col0 + 4 * j4) = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 tile GEMM (FFMA2): C[M,N] = op(A) * op(B) with fused epilogues.  128x128x16 tiles, 256 threads,
// 8x8 outputs per thread (as 2x2 blocks of 4x4), double-buffered shared memory.
//   TA = false: A[M][K] row-major      TA = true: A[K][M] row-major
//   TB = false: B[K][N] row-major      TB = true: B[N][K] row-major
// M, N multiples of 128; K (per split) a multiple of 16.
// ------------------------------------------------------------------------------------------------
constexpr int BM = 128, BN = 128, BK = 16;

template <bool TA, bool TB, int EPI>
__global__ void __launch_bounds__(256)
sgemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb,
             float* __restrict__ C, int ldc, const float* __restrict__ aux, int ldaux, int k_per_split) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int nk = k_per_split / BK;
  if (EPI == kEpiSplitK) C += (size_t)blockIdx.z * M * ldc;

  float4 ra[2], rb[2];
  auto gload = [&](int kt) {
    const int k0 = kbeg + kt * BK;
    if (!TA) {        // A[m][k]: thread -> row m = tid & 127, k-octet = tid >> 7
      const float* p = A + (size_t)(m0 + (tid & 127)) * lda + k0 + 8 * (tid >> 7);
      ra[0] = *reinterpret_cast<const float4*>(p); ra[1] = *reinterpret_cast<const float4*>(p + 4);
    } else {          // A[k][m]: thread -> k = tid >> 5 (+8), m4 = tid & 31
      const float* p = A + (size_t)(k0 + (tid >> 5)) * lda + m0 + 4 * (tid & 31);
      ra[0] = *reinterpret_cast<const float4*>(p); ra[1] = *reinterpret_cast<const float4*>(p + (size_t)8 * lda);
    }
    if (TB) {         // B[n][k]
      const float* p = Bm + (size_t)(n0 + (tid & 127)) * ldb + k0 + 8 * (tid >> 7);
      rb[0] = *reinterpret_cast<const float4*>(p); rb[1] = *reinterpret_cast<const float4*>(p + 4);
    } else {          // B[k][n]
      const float* p = Bm + (size_t)(k0 + (tid >> 5)) * ldb + n0 + 4 * (tid & 31);
      rb[0] = *reinterpret_cast<const float4*>(p); rb[1] = *reinterpret_cast<const float4*>(p + (size_t)8 * ldb);
    }
  };
  auto sstore = [&](int buf) {
    if (!TA) {
      const int m = tid & 127, kq = 8 * (tid >> 7);
      As[buf][kq + 0][m] = ra[0].x; As[buf][kq + 1][m] = ra[0].y; As[buf][kq + 2][m] = ra[0].z; As[buf][kq + 3][m] = ra[0].w;
      As[buf][kq + 4][m] = ra[1].x; As[buf][kq + 5][m] = ra[1].y; As[buf][kq + 6][m] = ra[1].z; As[buf][kq + 7][m] = ra[1].w;
    } else {
      *reinterpret_cast<float4*>(&As[buf][tid >> 5][4 * (tid & 31)]) = ra[0];
      *reinterpret_cast<float4*>(&As[buf][(tid >> 5) + 8][4 * (tid & 31)]) = ra[1];
    }
    if (TB) {
      const int n = tid & 127, kq = 8 * (tid >> 7);
      Bs[buf][kq + 0][n] = rb[0].x; Bs[buf][kq + 1][n] = rb[0].y; Bs[buf][kq + 2][n] = rb[0].z; Bs[buf][kq + 3][n] = rb[0].w;
      Bs[buf][kq + 4][n] = rb[1].x; Bs[buf][kq + 5][n] = rb[1].y; Bs[buf][kq + 6][n] = rb[1].z; Bs[buf][kq + 7][n] = rb[1].w;
    } else {
      *reinterpret_cast<float4*>(&Bs[buf][tid >> 5][4 * (tid & 31)]) = rb[0];
      *reinterpret_cast<float4*>(&Bs[buf][(tid >> 5) + 8][4 * (tid & 31)]) = rb[1];
    }
  };

  u64 acc[8][4];     // rows: 4*ty + i (i<4) and 64 + 4*ty + i;  column pairs: 4*tx + {0,1},{2,3} and 64 + ...
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0ull;

  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload(kt + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][4 * ty]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + 4 * ty]);
      const ulonglong2 b0 = *reinterpret_cast<const ulonglong2*>(&Bs[buf][k][4 * tx]);
      const ulonglong2 b1 = *reinterpret_cast<const ulonglong2*>(&Bs[buf][k][64 + 4 * tx]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const u64 aa = pack2(av[i], av[i]);
        ffma2(acc[i][0], aa, b0.x); ffma2(acc[i][1], aa, b0.y); ffma2(acc[i][2], aa, b1.x); ffma2(acc[i][3], aa, b1.y);
      }
    }
    if (kt + 1 < nk) sstore(buf ^ 1);
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? 4 * ty + i : 64 + 4 * ty + (i - 4));
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + 64 * h + 4 * tx;
      float4 v;
      unpack2(acc[i][2 * h], v.x, v.y);
      unpack2(acc[i][2 * h + 1], v.z, v.w);
      if (EPI == kEpiBiasRelu) {
        const float4 b = *reinterpret_cast<const float4*>(aux + n);
        v.x = fmaxf(v.x + b.x, 0.f); v.y = fmaxf(v.y + b.y, 0.f); v.z = fmaxf(v.z + b.z, 0.f); v.w = fmaxf(v.w + b.w, 0.f);
      } else if (EPI == kEpiReluMask) {
        const float4 h1 = *reinterpret_cast<const float4*>(aux + (size_t)m * ldaux + n);
        v.x = h1.x > 0.f ? v.x : 0.f; v.y = h1.y > 0.f ? v.y : 0.f; v.z = h1.z > 0.f ? v.z : 0.f; v.w = h1.w > 0.f ? v.w : 0.f;
      }
      *reinterpret_cast<float4*>(C + (size_t)m * ldc + n) = v;
    }
  }
}

// out[i] = sum_z part[z][i]  (fixed order -> deterministic), i < n
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ part, float* __restrict__ out, long long n, int nz, long long zstride) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int z = 0; z < nz; ++z) s += part[(size_t)z * zstride + i];
  out[i] = s;
}

// Same for few columns and many partials (the column-sum passes: n ~ 9k, nz = B / 128): 32 columns x 8 z-groups per
// block, z-group g sums z = g, g + 8, ..., the 8 group sums are added in group order.  Fixed order -> deterministic.
__global__ void __launch_bounds__(256)
reduce_partials_wide_kernel(const float* __restrict__ part, float* __restrict__ out, long long n, int nz, long long zstride) {
  __shared__ float red[8][33];
  const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
  const long long i = (long long)blockIdx.x * 32 + c;
  float s = 0.f;
  if (i < n) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;          // four independent chains: loads in flight
    int z = g;
    for (; z + 24 < nz; z += 32) {
      s0 += part[(size_t)z * zstride + i];
      s1 += part[(size_t)(z + 8) * zstride + i];
      s2 += part[(size_t)(z + 16) * zstride + i];
      s3 += part[(size_t)(z + 24) * zstride + i];
    }
    for (; z < nz; z += 8) s0 += part[(size_t)z * zstride + i];
    s = (s0 + s1) + (s2 + s3);
  }
  red[g][c] = s;
  __syncthreads();
  if (g == 0 && i < n) {
    float t = red[0][c];
#pragma unroll
    for (int q = 1; q < 8; ++q) t += red[q][c];
    out[i] = t;
  }
}

cudaError_t launch_reduce_partials(cudaStream_t st, const float* part, float* out, long long n, int nz, long long zstride) {
  if (nz >= 64) reduce_partials_wide_kernel<<<(unsigned)((n + 31) / 32), 256, 0, st>>>(part, out, n, nz, zstride);
  else reduce_partials_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(part, out, n, nz, zstride);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// head: Q = V + Adv - mean(Adv)  (dddqn.py:29-31).  One warp per row, rows [0,3B).
// ------------------------------------------------------------------------------------------------
constexpr int kHeadRows = 4;     // rows per warp: every head-weight load is reused for four activation rows (eight measured slower: 128 registers)

__global__ void __launch_bounds__(256)
lb_head_kernel(const float* __restrict__ H2, const float* __restrict__ theta, const float* __restrict__ theta_t,
               float* __restrict__ Q, int B, int H2n, int A, int offWv) {
  const int lane = threadIdx.x & 31;
  const int row0 = kHeadRows * (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));   // 2B and 3B are multiples of 4
  if (row0 >= 3 * B) return;
  const float* th = row0 >= 2 * B ? theta_t : theta;
  const float* Wv = th + offWv;
  const float* bv = Wv + H2n;
  const float* Wa = bv + 1;
  const float* ba = Wa + (size_t)H2n * A;
  const float* h = H2 + (size_t)row0 * H2n;
  float acc[kHeadRows][1 + kMaxA];
#pragma unroll
  for (int r = 0; r < kHeadRows; ++r)
#pragma unroll
    for (int c = 0; c <= kMaxA; ++c) acc[r][c] = 0.f;
  // lane l takes k = l, l + 32, ...: activation rows are read in 128-byte warp loads (16 of them in flight per lane);
  // the head weights (20 KB, L1-resident) are loaded once per k and used for the four rows -- with one row per warp the
  // five weight loads per activation load made the kernel L1-bound at 1.5 TB/s of what is a pure HBM read
  // (3 B x H2 x 4 bytes once).
  for (int k0 = lane; k0 < H2n; k0 += 128) {               // H2n is a multiple of 256
    float hv[kHeadRows][4];
#pragma unroll
    for (int r = 0; r < kHeadRows; ++r)
#pragma unroll
      for (int u = 0; u < 4; ++u) hv[r][u] = h[(size_t)r * H2n + k0 + 32 * u];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = k0 + 32 * u;
      const float wv = __ldg(Wv + k);
#pragma unroll
      for (int r = 0; r < kHeadRows; ++r) acc[r][0] = fmaf(hv[r][u], wv, acc[r][0]);
#pragma unroll
      for (int j = 0; j < kMaxA; ++j) if (j < A) {
        const float wa = __ldg(Wa + (size_t)k * A + j);
#pragma unroll
        for (int r = 0; r < kHeadRows; ++r) acc[r][1 + j] = fmaf(hv[r][u], wa, acc[r][1 + j]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kHeadRows; ++r)
#pragma unroll
    for (int c = 0; c <= kMaxA; ++c)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[r][c] += __shfl_xor_sync(0xffffffffu, acc[r][c], o);
  if (lane < kHeadRows) {
    float a[1 + kMaxA];
#pragma unroll
    for (int c = 0; c <= kMaxA; ++c) {                       // row `lane` of the four (static indexing only)
      a[c] = acc[0][c];
#pragma unroll
      for (int r = 1; r < kHeadRows; ++r) if (lane == r) a[c] = acc[r][c];
    }
    const float val = a[0] + bv[0];
    float msum = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxA; ++j) if (j < A) { a[1 + j] += ba[j]; msum += a[1 + j]; }
    const float mean = msum / (float)A;
#pragma unroll
    for (int j = 0; j < kMaxA; ++j) if (j < A) Q[(size_t)(row0 + lane) * A + j] = val + a[1 + j] - mean;
  }
}

// The same head from the partial sums the tensor-core GEMM's epilogue left behind (lb_gemm_tc_nn_head): hp1 covers the 2B theta
// rows, hp2 the B theta^- rows, each [column tile][row][8]; the tiles are summed in order, then bias and the dueling combine.
__global__ void __launch_bounds__(256)
lb_head_reduce_kernel(const float* __restrict__ hp1, const float* __restrict__ hp2, const float* __restrict__ theta,
                      const float* __restrict__ theta_t, float* __restrict__ Q, int B, int H2n, int A, int offWv, int ntn) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= 3 * B) return;
  const bool tm = row >= 2 * B;
  const float* hp = tm ? hp2 : hp1;
  const int M = tm ? B : 2 * B, r = tm ? row - 2 * B : row;
  const float* bv = (tm ? theta_t : theta) + offWv + H2n;
  const float* ba = bv + 1 + (size_t)H2n * A;
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int tn = 0; tn < ntn; ++tn) {
    const float4 p0 = *reinterpret_cast<const float4*>(hp + ((size_t)tn * M + r) * 8), p1 = *reinterpret_cast<const float4*>(hp + ((size_t)tn * M + r) * 8 + 4);
    a[0] += p0.x; a[1] += p0.y; a[2] += p0.z; a[3] += p0.w; a[4] += p1.x; a[5] += p1.y; a[6] += p1.z; a[7] += p1.w;
  }
  const float val = a[0] + bv[0];
  float msum = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxA; ++j) if (j < A) { a[1 + j] += ba[j]; msum += a[1 + j]; }
  const float mean = msum / (float)A;
#pragma unroll
  for (int j = 0; j < kMaxA; ++j) if (j < A) Q[(size_t)row * A + j] = val + a[1 + j] - mean;
}

// ------------------------------------------------------------------------------------------------
// targets + Huber + d(head)   (q_learning_functions.py:55-59, :35-36).  One thread per sample.
// dhd[i][0] = dV, dhd[i][1+j] = dAdv_j;  per-block partial sums of loss and of dhd columns (-> d head bias).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lb_targets_kernel(const float* __restrict__ Q, const long long* __restrict__ act, const float* __restrict__ rew,
                  const uint8_t* __restrict__ done, float* __restrict__ dhd, float* __restrict__ partial,
                  int B, int A, float gamma, float inv_global_batch, int loss_kind, LbTaps taps) {
  __shared__ float red[8][2 + kMaxA];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float vals[2 + kMaxA];
#pragma unroll
  for (int c = 0; c < 2 + kMaxA; ++c) vals[c] = 0.f;
  if (i < B) {
    float q[kMaxA], nq[kMaxA], nqt[kMaxA];
#pragma unroll
    for (int j = 0; j < kMaxA; ++j) if (j < A) {
      q[j] = Q[(size_t)i * A + j]; nq[j] = Q[(size_t)(B + i) * A + j]; nqt[j] = Q[(size_t)(2 * B + i) * A + j];
    }
    int astar = 0; float best = nq[0];
#pragma unroll
    for (int j = 1; j < kMaxA; ++j) if (j < A && nq[j] > best) { best = nq[j]; astar = j; }
    long long al = act[i];
    const int a = al < 0 ? 0 : (al >= A ? A - 1 : (int)al);
    float qa = q[0], nt = nqt[0];
#pragma unroll
    for (int j = 1; j < kMaxA; ++j) if (j < A) { if (j == a) qa = q[j]; if (j == astar) nt = nqt[j]; }
    const float d = done[i] ? 1.f : 0.f;
    const float tv = rew[i] + (1.0f - d) * (gamma * nt - qa);
    const float tgt = qa + tv;
    const float e = qa - tgt;
    const float ae = fabsf(e), quad = fminf(ae, 1.0f);
    const bool l2loss = loss_kind == kLossL2;
    vals[0] = (l2loss ? 0.5f * e * e : 0.5f * quad * quad + (ae - quad)) * inv_global_batch;
    const float gi = (l2loss ? e : fminf(fmaxf(e, -1.0f), 1.0f)) * inv_global_batch;
    dhd[(size_t)i * 8 + 0] = gi;
    vals[1] = gi;
#pragma unroll
    for (int j = 0; j < kMaxA; ++j) {
      const float dadv = j < A ? (j == a ? gi : 0.f) - gi / (float)A : 0.f;
      dhd[(size_t)i * 8 + 1 + j] = dadv;
      vals[2 + j] = dadv;
    }
    if (taps.enabled) {
      for (int j = 0; j < A; ++j) taps.targets[(size_t)i * A + j] = j == a ? tgt : q[j];
      taps.max_actions[i] = astar;
    }
  }
#pragma unroll
  for (int c = 0; c < 2 + kMaxA; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vals[c] += __shfl_xor_sync(0xffffffffu, vals[c], o);
    if (lane == 0) red[warp][c] = vals[c];
  }
  __syncthreads();
  if (threadIdx.x < 2 + kMaxA) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    partial[(size_t)blockIdx.x * 16 + threadIdx.x] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// fused pass over H2(theta, s):  dH2 = relu'(h2) * (dhd . Wh^T);  partial column sums for dWh and db2.
// thread = one hidden unit j, block = 256 units x RC rows.   part[chunk][c][j], c: 0..A = dWh cols, A+1 = db2
// ------------------------------------------------------------------------------------------------
template <int RC>
__global__ void __launch_bounds__(256)
lb_dh2_kernel(const float* __restrict__ H2s, const float* __restrict__ dhd, const float* __restrict__ theta,
              float* __restrict__ dH2, float* __restrict__ part, int B, int H2n, int A, int offWv) {
  // thread = FOUR adjacent hidden units (16-byte loads of h2, 16-byte stores of dh2)
  __shared__ __align__(16) float sd[RC][8];
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int r0 = blockIdx.y * RC;
  for (int q = threadIdx.x; q < RC * 8; q += blockDim.x) sd[q >> 3][q & 7] = dhd[(size_t)r0 * 8 + q];
  __syncthreads();
  if (j >= H2n) return;
  const float* Wv = theta + offWv;
  const float* Wa = Wv + H2n + 1;
  float4 wh[1 + kMaxA], acc[2 + kMaxA];
  wh[0] = *reinterpret_cast<const float4*>(Wv + j);
#pragma unroll
  for (int c = 0; c < kMaxA; ++c)
    wh[1 + c] = c < A ? make_float4(Wa[(size_t)j * A + c], Wa[(size_t)(j + 1) * A + c], Wa[(size_t)(j + 2) * A + c], Wa[(size_t)(j + 3) * A + c])
                      : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int c = 0; c < 2 + kMaxA; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int r = 0; r < RC; ++r) {
    const float4 h = *reinterpret_cast<const float4*>(H2s + (size_t)(r0 + r) * H2n + j);
    const float4 d0 = *reinterpret_cast<const float4*>(&sd[r][0]), d1 = *reinterpret_cast<const float4*>(&sd[r][4]);
    const float dd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int c = 0; c <= kMaxA; ++c) {
      v.x = fmaf(dd[c], wh[c].x, v.x); v.y = fmaf(dd[c], wh[c].y, v.y); v.z = fmaf(dd[c], wh[c].z, v.z); v.w = fmaf(dd[c], wh[c].w, v.w);
      acc[c].x = fmaf(h.x, dd[c], acc[c].x); acc[c].y = fmaf(h.y, dd[c], acc[c].y); acc[c].z = fmaf(h.z, dd[c], acc[c].z); acc[c].w = fmaf(h.w, dd[c], acc[c].w);
    }
    v.x = h.x > 0.f ? v.x : 0.f; v.y = h.y > 0.f ? v.y : 0.f; v.z = h.z > 0.f ? v.z : 0.f; v.w = h.w > 0.f ? v.w : 0.f;
    *reinterpret_cast<float4*>(dH2 + (size_t)(r0 + r) * H2n + j) = v;
    acc[1 + kMaxA].x += v.x; acc[1 + kMaxA].y += v.y; acc[1 + kMaxA].z += v.z; acc[1 + kMaxA].w += v.w;
  }
#pragma unroll
  for (int c = 0; c < 2 + kMaxA; ++c) *reinterpret_cast<float4*>(part + ((size_t)blockIdx.y * (2 + kMaxA) + c) * H2n + j) = acc[c];
}

// fused pass over dH1: partial sums for dW1[d][k] = sum_r x[r][d] dH1[r][k] and db1[k].  part[chunk][d][k], d = D -> db1.
// DV = the obs dim padded to 8 or 16 (the padding of x is zero): the inner loop is DV / 4 broadcast 16-byte shared loads and DV
// FMAs per element of dH1, eight rows of loads in flight -- the pass is a pure HBM read of dH1 (the round-1 form, with a
// run-time D inside the row loop, spent ~40 instructions per element and ran at 1.8 TB/s).
template <int DV, int RC>
__global__ void __launch_bounds__(256)
lb_dw1_kernel(const float* __restrict__ s, const float* __restrict__ dH1, float* __restrict__ part, int B, int D, int H1n) {
  // thread = FOUR adjacent hidden units (16-byte loads of dH1: a quarter of the load instructions of the one-column form)
  __shared__ __align__(16) float sx[RC][DV];
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int r0 = blockIdx.y * RC;
  for (int q = threadIdx.x; q < RC * DV; q += blockDim.x) {
    const int r = q / DV, d = q - r * DV;
    sx[r][d] = d < D ? s[(size_t)(r0 + r) * D + d] : 0.f;
  }
  __syncthreads();
  if (k >= H1n) return;
  float4 acc[DV + 1];
#pragma unroll
  for (int d = 0; d <= DV; ++d) acc[d] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* g0 = dH1 + (size_t)r0 * H1n + k;
  for (int r = 0; r < RC; r += 8) {
    float4 g[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = *reinterpret_cast<const float4*>(g0 + (size_t)(r + i) * H1n);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int q = 0; q < DV / 4; ++q) {
        const float4 x = *reinterpret_cast<const float4*>(&sx[r + i][4 * q]);
        const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float4& a = acc[4 * q + e];
          a.x = fmaf(xs[e], g[i].x, a.x); a.y = fmaf(xs[e], g[i].y, a.y); a.z = fmaf(xs[e], g[i].z, a.z); a.w = fmaf(xs[e], g[i].w, a.w);
        }
      }
      acc[DV].x += g[i].x; acc[DV].y += g[i].y; acc[DV].z += g[i].z; acc[DV].w += g[i].w;
    }
  }
#pragma unroll
  for (int d = 0; d < DV; ++d) if (d < D) *reinterpret_cast<float4*>(part + ((size_t)blockIdx.y * (D + 1) + d) * H1n + k) = acc[d];
  *reinterpret_cast<float4*>(part + ((size_t)blockIdx.y * (D + 1) + D) * H1n + k) = acc[DV];
}

// scatter the reduced head partials [c][j] into the flat gradient: c = 0 -> dWv[j], 1..A -> dWa[j][c-1], A+1 (index 1+kMaxA) -> db2[j]
__global__ void __launch_bounds__(256)
lb_head_grads_kernel(const float* __restrict__ red, float* __restrict__ grads, int H2n, int A, int offb2, int offWv) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= H2n) return;
  grads[offWv + j] = red[j];
  float* gWa = grads + offWv + H2n + 1;
  for (int c = 0; c < A; ++c) gWa[(size_t)j * A + c] = red[(size_t)(1 + c) * H2n + j];
  grads[offb2 + j] = red[(size_t)(1 + kMaxA) * H2n + j];
}

// final reduction of the targets-kernel partials: loss -> grads[P] (extra slot), d(head bias) -> grads.  One warp per
// column (0 = loss, 1 = d bv, 2.. = d ba): lane l sums blocks l, l + 32, ... in order, then a fixed shuffle tree.
__global__ void __launch_bounds__(32 * (2 + kMaxA))
lb_finish_kernel(const float* __restrict__ partial, int nblk, float* __restrict__ grads, int P, int A, int offbv, int offba) {
  const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float s = 0.f;
  for (int b = lane; b < nblk; b += 32) s += partial[(size_t)b * 16 + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane != 0) return;
  if (c == 0) grads[P] = s;                       // loss (local share of the global mean), extra slot after the P gradients
  else if (c == 1) grads[offbv] = s;              // d bv
  else if (c - 2 < A) grads[offba + (c - 2)] = s; // d ba
}

// optax adam / adamw, exact fp32 (memory-bound: IEEE division and sqrt are free here)
__global__ void __launch_bounds__(256)
lb_adam_kernel(float* __restrict__ theta, float* __restrict__ mu, float* __restrict__ nu, const float* __restrict__ g,
               int P, float b1, float b2, float c1, float c2, float eps, float eps_root, float lr, float wd,
               const unsigned* __restrict__ comm_error) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  if (comm_error && *comm_error) return;   // the gradient exchange failed (comm_p2p.cu): g is not the global sum -- leave theta / mu / nu alone
  const float gr = g[p];
  const float m = b1 * mu[p] + (1.0f - b1) * gr;
  const float v = b2 * nu[p] + (1.0f - b2) * (gr * gr);
  mu[p] = m; nu[p] = v;
  float u = (m / c1) / (sqrtf(v / c2 + eps_root) + eps);
  const float th = theta[p];
  if (wd != 0.f) u = u + wd * th;
  theta[p] = th + (-lr) * u;
}

template <bool TA, bool TB, int EPI>
cudaError_t launch_sgemm(cudaStream_t st, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                         float* C, int ldc, const float* aux, int ldaux, int splitk) {
  dim3 grid(N / BN, M / BM, splitk);
  sgemm_kernel<TA, TB, EPI><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, aux, ldaux, K / splitk);
  return cudaGetLastError();
}

}  // namespace

cudaError_t lb_gemm_ffma(cudaStream_t st, int kind, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                         float* C, int ldc, const float* aux, int ldaux, int splitk) {
  switch (kind) {
    case kGemmNN_BiasRelu: return launch_sgemm<false, false, kEpiBiasRelu>(st, M, N, K, A, lda, B, ldb, C, ldc, aux, ldaux, 1);
    case kGemmNT_ReluMask: return launch_sgemm<false, true, kEpiReluMask>(st, M, N, K, A, lda, B, ldb, C, ldc, aux, ldaux, 1);
    case kGemmTN_SplitK: return launch_sgemm<true, false, kEpiSplitK>(st, M, N, K, A, lda, B, ldb, C, ldc, aux, ldaux, splitk);
    default: return cudaErrorInvalidValue;
  }
}

#define LBCHK(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return _e; } while (0)

// out[c][r] = in[r][c]  (32 x 32 tiles through shared memory); rows, cols multiples of 32
__global__ void __launch_bounds__(256)
lb_transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 4; ++j) tile[ty + 8 * j][tx] = in[(size_t)(r0 + ty + 8 * j) * cols + c0 + tx];
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) out[(size_t)(c0 + ty + 8 * j) * rows + r0 + tx] = tile[tx][ty + 8 * j];
}

// One forward/backward pass over the local batch already gathered into ws.s/a/r/s2/done.
cudaError_t lb_forward_backward(cudaStream_t st, const LbDims& d, const LbWorkspace& ws, float gamma, float inv_global_batch,
                                int gemm_mode, int loss_kind, const LbTaps& taps, cudaEvent_t after_dw2) {
  const int B = d.B, H1n = d.H1, H2n = d.H2, D = d.D, A = d.A;
  const int offb1 = D * H1n, offW2 = offb1 + H1n, offb2 = offW2 + H1n * H2n, offWv = offb2 + H2n;
  const int offbv = offWv + H2n, offba = offbv + 1 + H2n * A;
  // ---- forward ----
  {
    dim3 grid((H1n + 1023) / 1024, B < 512 ? B : 512, 3);
    const int smem = (D + 1) * 1024 * 4;                       // <= 68 KB
    static bool configured = false;
    if (!configured) {
      LBCHK(cudaFuncSetAttribute(lb_layer1_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (kMaxD + 1) * 1024 * 4));
      LBCHK(cudaFuncSetAttribute(lb_layer1_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (kMaxD + 1) * 1024 * 4));
      configured = true;
    }
    if (D == 8) lb_layer1_kernel<8><<<grid, 256, smem, st>>>(ws.s, ws.s2, ws.theta, ws.theta_t, ws.H1, B, D, H1n);
    else lb_layer1_kernel<0><<<grid, 256, smem, st>>>(ws.s, ws.s2, ws.theta, ws.theta_t, ws.H1, B, D, H1n);
    LBCHK(cudaGetLastError());
  }
  bool head_done = false;
  static const bool fuse_head = [] { const char* e = getenv("DQN_B200_LB_FUSE_HEAD"); return !e || atoi(e) != 0; }();   // experiment knob
  if (gemm_mode == kGemmModeTC3xTF32 && fuse_head) {
    // The head rides in the epilogue of the two h2 products: h2 (805 MB at the BASELINE shape) is not read back, and the rows
    // that feed nothing but the head (s' under theta and theta^-) are not even written.  Partial sums per 128-column tile go
    // through the dH2 buffer (free until the backward pass): 3B x H2/128 x 8 floats.
    const int ntn = H2n / 128;
    float* hp1 = ws.dH2;
    float* hp2 = hp1 + (size_t)ntn * 2 * B * 8;
    cudaError_t e = lb_gemm_tc_nn_head(st, 2 * B, H2n, H1n, ws.H1, H1n, ws.theta + offW2, H2n, ws.H2, H2n, ws.theta + offb2,
                                       ws.theta + offWv, A, hp1, B);
    if (e == cudaSuccess)
      e = lb_gemm_tc_nn_head(st, B, H2n, H1n, ws.H1 + (size_t)2 * B * H1n, H1n, ws.theta_t + offW2, H2n, ws.H2 + (size_t)2 * B * H2n, H2n,
                             ws.theta_t + offb2, ws.theta_t + offWv, A, hp2, 0);
    if (e == cudaSuccess) {
      lb_head_reduce_kernel<<<(3 * B + 255) / 256, 256, 0, st>>>(hp1, hp2, ws.theta, ws.theta_t, ws.Q, B, H2n, A, offWv, ntn);
      LBCHK(cudaGetLastError());
      head_done = true;
    } else if (e != cudaErrorNotSupported) {
      return e;
    }
  }
  if (!head_done) {
    LBCHK(lb_gemm(st, gemm_mode, kGemmNN_BiasRelu, 2 * B, H2n, H1n, ws.H1, H1n, ws.theta + offW2, H2n, ws.H2, H2n, ws.theta + offb2, 0, 1, ws));
    LBCHK(lb_gemm(st, gemm_mode, kGemmNN_BiasRelu, B, H2n, H1n, ws.H1 + (size_t)2 * B * H1n, H1n, ws.theta_t + offW2, H2n,
                  ws.H2 + (size_t)2 * B * H2n, H2n, ws.theta_t + offb2, 0, 1, ws));
    lb_head_kernel<<<(3 * B / kHeadRows + 7) / 8, 256, 0, st>>>(ws.H2, ws.theta, ws.theta_t, ws.Q, B, H2n, A, offWv);
    LBCHK(cudaGetLastError());
  }
  const int nblk = (B + 255) / 256;
  lb_targets_kernel<<<nblk, 256, 0, st>>>(ws.Q, ws.a, ws.r, ws.done, ws.dhd, ws.partial, B, A, gamma, inv_global_batch, loss_kind, taps);
  LBCHK(cudaGetLastError());
  lb_finish_kernel<<<1, 32 * (2 + kMaxA), 0, st>>>(ws.partial, nblk, ws.grads, d.P, A, offbv, offba);
  LBCHK(cudaGetLastError());
  // ---- backward ----
  const int rc = lb_row_chunk(B), nchunk = B / rc;
  {
    dim3 grid((H2n + 1023) / 1024, nchunk);
    if (rc == 32) lb_dh2_kernel<32><<<grid, 256, 0, st>>>(ws.H2, ws.dhd, ws.theta, ws.dH2, ws.colpart, B, H2n, A, offWv);
    else lb_dh2_kernel<128><<<grid, 256, 0, st>>>(ws.H2, ws.dhd, ws.theta, ws.dH2, ws.colpart, B, H2n, A, offWv);
    LBCHK(cudaGetLastError());
    const long long n = (long long)(2 + kMaxA) * H2n;
    LBCHK(launch_reduce_partials(st, ws.colpart, ws.colred, n, nchunk, n));
    lb_head_grads_kernel<<<(H2n + 255) / 256, 256, 0, st>>>(ws.colred, ws.grads, H2n, A, offb2, offWv);
    LBCHK(cudaGetLastError());
  }
  {   // dW2 = H1s^T . dH2  (reduction over the batch rows): split-K partials, fixed-order reduce
    int splitk = 1;
    if (gemm_mode == kGemmModeTC3xTF32) {
      // persistent SM-pair kernel: many (tile, split) work items per cluster keep the last round full; a split is >= 256 rows
      while (splitk < 16 && (B / (splitk * 2)) % 128 == 0 && B / (splitk * 2) >= 256) splitk *= 2;
    } else {
      while (splitk < 16 && (H1n / BM) * (H2n / BN) * splitk < 296 && (B / (splitk * 2)) % BK == 0 && B / (splitk * 2) >= 256) splitk *= 2;
    }
    LBCHK(lb_gemm(st, gemm_mode, kGemmTN_SplitK, H1n, H2n, B, ws.H1, H1n, ws.dH2, H2n, ws.gemmpart, H2n, nullptr, 0, splitk, ws));
    const long long n = (long long)H1n * H2n;
    LBCHK(launch_reduce_partials(st, ws.gemmpart, ws.grads + offW2, n, splitk, n));
    LBCHK(cudaGetLastError());
    if (after_dw2) LBCHK(cudaEventRecord(after_dw2, st));
  }
  if (gemm_mode == kGemmModeTC3xTF32) {
    // dh1 = relu'(h1) * (dh2 . W2^T): the tensor-core kernel's NT form (K-major B) runs at 126 TF against 151 TF for the NN
    // form, so W2 (4 MB) is transposed once per step and the product runs as NN
    lb_transpose_kernel<<<dim3(H2n / 32, H1n / 32), 256, 0, st>>>(ws.theta + offW2, ws.tc_scratch, H1n, H2n);
    LBCHK(cudaGetLastError());
    LBCHK(lb_gemm(st, gemm_mode, kGemmNN_ReluMask, B, H1n, H2n, ws.dH2, H2n, ws.tc_scratch, H1n, ws.dH1, H1n, ws.H1, H1n, 1, ws));
  } else {
    LBCHK(lb_gemm(st, gemm_mode, kGemmNT_ReluMask, B, H1n, H2n, ws.dH2, H2n, ws.theta + offW2, H2n, ws.dH1, H1n, ws.H1, H1n, 1, ws));
  }
  {
    dim3 grid((H1n + 1023) / 1024, nchunk);
    if (D <= 8 && rc == 32) lb_dw1_kernel<8, 32><<<grid, 256, 0, st>>>(ws.s, ws.dH1, ws.colpart, B, D, H1n);
    else if (D <= 8) lb_dw1_kernel<8, 128><<<grid, 256, 0, st>>>(ws.s, ws.dH1, ws.colpart, B, D, H1n);
    else if (rc == 32) lb_dw1_kernel<16, 32><<<grid, 256, 0, st>>>(ws.s, ws.dH1, ws.colpart, B, D, H1n);
    else lb_dw1_kernel<16, 128><<<grid, 256, 0, st>>>(ws.s, ws.dH1, ws.colpart, B, D, H1n);
    LBCHK(cudaGetLastError());
    const long long n = (long long)(D + 1) * H1n;                 // [d][k] == flat [W1 | b1]
    LBCHK(launch_reduce_partials(st, ws.colpart, ws.grads, n, nchunk, n));
  }
  return cudaSuccess;
}

// theta^- := tau theta + (1 - tau) theta^-, products and sum rounded separately (see dqn_polyak_target_kernel, act.cu)
__global__ void __launch_bounds__(256) lb_polyak_kernel(const float* __restrict__ theta, float* __restrict__ theta_t, int P, float tau) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  theta_t[p] = __fadd_rn(__fmul_rn(tau, theta[p]), __fmul_rn(__fsub_rn(1.0f, tau), theta_t[p]));
}
cudaError_t lb_polyak(cudaStream_t st, const LbDims& d, const LbWorkspace& ws, float tau) {
  lb_polyak_kernel<<<(d.P + 255) / 256, 256, 0, st>>>(ws.theta, ws.theta_t, d.P, tau);
  return cudaGetLastError();
}

cudaError_t lb_adam(cudaStream_t st, const LbDims& d, const LbWorkspace& ws, float b1, float b2, float c1, float c2,
                    float eps, float eps_root, float lr, float wd, const unsigned* comm_error) {
  lb_adam_kernel<<<(d.P + 255) / 256, 256, 0, st>>>(ws.theta, ws.mu, ws.nu, ws.grads, d.P, b1, b2, c1, c2, eps, eps_root, lr, wd, comm_error);
  return cudaGetLastError();
}

}  // namespace dqn
