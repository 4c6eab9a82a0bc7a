// train_fused.cu -- the fused, persistent dueling double-DQN train step (sm_100a, fp32 FFMA).
//
// One CTA = one agent.  For K consecutive train steps it does everything Agent._step does
// (General/QLearning/q_agent.py:146-169) without leaving the SM:
//   sample_batch        replay_buffer.py:68-85        Philox (or supplied) indices, cp.async gather of
//                                                      whole AoS records into shared memory, one step ahead
//   preprocessing       q_learning_functions.py:76-85  done -> f32, action -> i32 while unpacking
//   compute_q_targets   q_learning_functions.py:42-64  3 forwards, first-max argmax, F5-quirk TD target
//   compute_loss        q_learning_functions.py:31-39  Huber(delta=1) summed over actions, mean over B
//   train_step          q_learning_functions.py:14-28  hand-derived backward, Adam / AdamW, in place
// theta, theta^- and the gradient accumulator live in shared memory for the whole launch, Adam's
// mu/nu live in registers (fixed thread<->parameter ownership); HBM traffic per step is the
// gathered records (one 96-byte record per sample for D <= 10) and one 4-byte loss.
//
// Shared-memory layouts
//   weights ("smem layout"): [W1;b1] (D+1 x 32) | [W2;b2] (33 x 64) | [Wv Wa | bias row] (65 x 8),
//       i.e. bias = last row of each augmented matrix; head columns: 0 = V, 1..A = advantage, rest 0.
//   activations: k-major ("transposed") [feature][row] with row strides 132 (two 64-row halves) or
//       68; strides are = 4 (mod 32) words so 16-byte accesses of 8 consecutive rows hit 8 distinct
//       bank groups.  Forward GEMMs run in outer-product form on these (both operands contiguous
//       along the output tile), weight-gradient GEMMs in dot form (both operands contiguous along
//       the reduction = batch row), so no explicit transposes are needed.
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace dqn {

namespace {

constexpr int NT = 256;          // threads per CTA
constexpr int BT = 64;           // batch rows per tile
constexpr int RS2 = 2 * BT + 4;  // 132
constexpr int RS1 = BT + 4;      // 68
constexpr int HC = 8;            // padded head width
constexpr int NPT = 13;          // ceil(max smem-layout params / NT), D = 16: (17*32 + 33*64 + 65*8) = 3176

struct Lay {   // offsets in floats
  int pW2, pWh, PS;
  int oW, oWt, oG, oX, oH1, oH2, oDh2T, oDh2R, oDh1T, oDhdT, oQB, oScr, oAct, oRew, oDone, oRed, oStage, total;
};

__host__ __device__ inline Lay make_layout(int D, int recw) {
  Lay L;
  L.pW2 = (D + 1) * kH1;
  L.pWh = L.pW2 + (kH1 + 1) * kH2;
  L.PS = L.pWh + (kH2 + 1) * HC;
  const int PSa = (L.PS + 3) & ~3;
  int o = 0;
  L.oW = o; o += PSa;
  L.oWt = o; o += PSa;
  L.oG = o; o += PSa;
  L.oX = o; o += (D + 1) * RS2;
  L.oH1 = o; o += kH1 * RS2;
  L.oH2 = o; o += (kH2 + 1) * RS2;
  L.oDh2T = o; o += kH2 * RS1;
  L.oDh2R = o; o += BT * RS1;
  L.oDh1T = o; o += kH1 * RS1;
  L.oDhdT = o; o += HC * RS1;
  L.oQB = o; o += BT * HC;
  L.oScr = o; o += 3 * BT * 2 * HC;
  L.oAct = o; o += BT;
  L.oRew = o; o += BT;
  L.oDone = o; o += BT;
  L.oRed = o; o += 32;
  L.oStage = o; o += BT * recw;
  L.total = o;
  return L;
}

// smem-layout index -> flat-layout index (-1 = padding)
__device__ __forceinline__ int smem_to_flat(int p, int D, int A, const Lay& L) {
  if (p < L.pWh) return p;   // [W1;b1] and [W2;b2] are contiguous in both layouts
  const int q = p - L.pWh;
  const int k = q >> 3, c = q & 7;
  const int offWv = L.pWh;                 // flat offset of Wv = D*32+32+32*64+64
  const int offbv = offWv + kH2;
  const int offWa = offbv + 1;
  const int offba = offWa + kH2 * A;
  if (c > A) return -1;
  if (k < kH2) return c == 0 ? offWv + k : offWa + k * A + (c - 1);
  return c == 0 ? offbv : offba + (c - 1);
}

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

// Outer-product-form tile: acc[i][j] (+)= sum_k A[k*lda + i] * Bw[k*ldb + j], i < 4, j < TN.
template <int TN>
__device__ __forceinline__ void op_tile4(const float* __restrict__ A, int lda, const float* __restrict__ Bw, int ldb,
                                         int K, float (&acc)[4][TN]) {
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float4 a4 = ld4(A + k * lda);
    const float a[4] = {a4.x, a4.y, a4.z, a4.w};
    float b[TN];
    if constexpr (TN == 2) {
      const float2 b2 = *reinterpret_cast<const float2*>(Bw + k * ldb);
      b[0] = b2.x; b[1] = b2.y;
    } else {
#pragma unroll
      for (int j4 = 0; j4 < TN / 4; ++j4) {
        const float4 b4 = ld4(Bw + k * ldb + 4 * j4);
        b[4 * j4 + 0] = b4.x; b[4 * j4 + 1] = b4.y; b[4 * j4 + 2] = b4.z; b[4 * j4 + 3] = b4.w;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

// Dot-form tile over 64 reduction elements (16 chunks of 4):
// acc[i][j] += sum_r A[(m0 + i*ms)*lda + r] * Bm[(n0 + j*ns)*ldb + r]
template <int TM, int TN>
__device__ __forceinline__ void dot_tile(const float* __restrict__ A, int lda, int m0, int ms,
                                         const float* __restrict__ Bm, int ldb, int n0, int ns,
                                         int chunk0, int nchunks, float (&acc)[TM][TN]) {
#pragma unroll 2
  for (int c = chunk0; c < chunk0 + nchunks; ++c) {
    float4 a[TM], b[TN];
#pragma unroll
    for (int i = 0; i < TM; ++i) a[i] = ld4(A + (m0 + i * ms) * lda + 4 * c);
#pragma unroll
    for (int j = 0; j < TN; ++j) b[j] = ld4(Bm + (n0 + j * ns) * ldb + 4 * c);
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        float v = acc[i][j];
        v = fmaf(a[i].x, b[j].x, v);
        v = fmaf(a[i].y, b[j].y, v);
        v = fmaf(a[i].z, b[j].z, v);
        v = fmaf(a[i].w, b[j].w, v);
        acc[i][j] = v;
      }
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int A>
__global__ void __launch_bounds__(NT, 1) dqn_train_fused_kernel(const TrainArgs args) {
  extern __shared__ __align__(16) float sm[];
  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  const int sel = blockIdx.x;
  const int agent = args.agent_begin + sel;
  const int D = args.dims.D;
  const int recw = args.dims.recw;
  const int PF = args.dims.PF;
  const Lay L = make_layout(D, recw);

  float* const W = sm + L.oW;       // theta      (smem layout)
  float* const Wt = sm + L.oWt;     // theta^-
  float* const G = sm + L.oG;       // gradient accumulator
  float* const X = sm + L.oX;       // [D+1][132]  cols 0..63 = s rows, 64..127 = s' rows; row D = ones
  float* const H1 = sm + L.oH1;     // [32][132]
  float* const H2 = sm + L.oH2;     // [65][132]  row 64 = ones
  float* const Dh2T = sm + L.oDh2T; // [64 j][68]
  float* const Dh2R = sm + L.oDh2R; // [64 r][68]
  float* const Dh1T = sm + L.oDh1T; // [32 k][68]
  float* const DhdT = sm + L.oDhdT; // [8 c][68]
  float* const QB = sm + L.oQB;     // [64][8]   Q(theta^-, s')
  float* const Scr = sm + L.oScr;   // head split-K partials
  int* const Act = reinterpret_cast<int*>(sm + L.oAct);
  float* const Rew = sm + L.oRew;
  float* const Done = sm + L.oDone;
  float* const Red = sm + L.oRed;   // [0..1] loss partials, [8..11] Adam bias corrections (double-buffered)
  float* const Stage = sm + L.oStage;

  float* const gW = args.params + (size_t)agent * 4 * PF;
  float* const gWt = gW + PF;
  float* const gM = gW + 2 * PF;
  float* const gV = gW + 3 * PF;
  AgentCtl* const ctl = args.ctl + agent;
  const uint32_t* const ring = args.rings + (size_t)agent * args.dims.N * recw;

  // ---- one-time: parameters HBM -> smem / registers ---------------------------------------
  float mreg[NPT], vreg[NPT];
#pragma unroll
  for (int i = 0; i < NPT; ++i) {
    const int p = t + i * NT;
    mreg[i] = 0.f; vreg[i] = 0.f;
    if (p < L.PS) {
      const int f = smem_to_flat(p, D, A, L);
      W[p] = f >= 0 ? gW[f] : 0.f;
      Wt[p] = f >= 0 ? gWt[f] : 0.f;
      G[p] = 0.f;
      if (f >= 0) { mreg[i] = gM[f]; vreg[i] = gV[f]; }
    }
  }
  for (int r = t; r < RS2; r += NT) { X[D * RS2 + r] = 1.f; H2[kH2 * RS2 + r] = 1.f; }

  const float gamma = ctl->gamma, lr = ctl->lr, b1 = ctl->b1, b2 = ctl->b2;
  const float eps = ctl->eps, eps_root = ctl->eps_root, wd = ctl->wd;
  const int B = ctl->batch_size;
  const long long step0 = ctl->train_steps;
  const int count0 = ctl->adam_count;
  const long long rc = ctl->ring_counter;
  const long long size = rc < args.dims.N ? rc : args.dims.N;
  const int ntiles = (B + BT - 1) / BT;
  const float fB = (float)B;
  const int cpr = recw >> 2;

  // gather of (step kstep, tile) into the staging buffer; 4 lanes per row, 16-byte cp.async each
  auto prefetch = [&](int kstep, int tile) {
    const int row = t >> 2, l4 = t & 3;
    const int i = tile * BT + row;
    float* dst = Stage + row * recw;
    if (i < B) {
      long long slot;
      if (args.idx) slot = args.idx[((size_t)sel * args.K + kstep) * B + i];
      else slot = philox_index(args.seed, args.agent_id_base + agent, step0 + kstep, i, size);
      const uint32_t* src = ring + slot * recw;
      for (int c = l4; c < cpr; c += 4) cp_async16(dst + 4 * c, src + 4 * c);
      if (args.taps.enabled && l4 == 0 && args.taps.indices) args.taps.indices[i] = slot;
    } else {
      for (int c = l4; c < cpr; c += 4) st4(dst + 4 * c, 0.f, 0.f, 0.f, 0.f);
    }
    cp_async_commit();
  };

  prefetch(0, 0);

  for (int kstep = 0; kstep < args.K; ++kstep) {
    const int count = (count0 > 0x7fffffff - 1 - kstep) ? 0x7fffffff : count0 + kstep + 1;   // safe_int32_increment
    if (t == 0) {
      // optax bias correction 1 - decay**count, decay**count correctly rounded to fp32 (see oracle pow_f32)
      // (double-buffered by step parity: slower warps may still be reading the previous step's pair)
      Red[8 + 2 * (kstep & 1)] = 1.0f - (float)pow((double)b1, (double)count);
      Red[9 + 2 * (kstep & 1)] = 1.0f - (float)pow((double)b2, (double)count);
    }
    float loss_acc = 0.f;   // meaningful in warps 0,1

    for (int tile = 0; tile < ntiles; ++tile) {
      // ---- unpack staged records: preprocessing (q_learning_functions.py:76-85) ------------
      cp_async_wait_all();
      __syncthreads();
      {
        const int nw = 2 * D + 4;
        for (int rr = 0; rr < 8; ++rr) {
          const int row = warp * 8 + rr;
          for (int wd = lane; wd < nw; wd += 32) {
            const float v = Stage[row * recw + wd];
            if (wd < D) X[wd * RS2 + row] = v;
            else if (wd < 2 * D) X[(wd - D) * RS2 + BT + row] = v;
            else if (wd == 2 * D) {
              int a = __float_as_int(v);
              Act[row] = a < 0 ? 0 : (a >= A ? A - 1 : a);   // jax clamps out-of-range gather indices
            } else if (wd == 2 * D + 2) Rew[row] = v;
            else if (wd == 2 * D + 3) Done[row] = __float_as_uint(v) ? 1.f : 0.f;   // dones.astype(float32)
          }
        }
      }
      __syncthreads();
      if (tile + 1 < ntiles) prefetch(kstep, tile + 1);
      else if (kstep + 1 < args.K) prefetch(kstep + 1, 0);

      // ================= batch B: Q(theta^-, s')  (q_learning_functions.py:54) ==============
      {  // layer 1: rows = s' (X cols 64..127) -> H1 cols 0..63
        const int mt = t & 15, nt = t >> 4;
        const int m0 = 4 * mt, n0 = 2 * nt;
        float acc[4][2];
#pragma unroll
        for (int j = 0; j < 2; ++j) { const float bias = Wt[D * kH1 + n0 + j];
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i][j] = bias; }
        op_tile4<2>(X + BT + m0, RS2, Wt + n0, kH1, D, acc);
#pragma unroll
        for (int j = 0; j < 2; ++j)
          st4(H1 + (n0 + j) * RS2 + m0, fmaxf(acc[0][j], 0.f), fmaxf(acc[1][j], 0.f), fmaxf(acc[2][j], 0.f), fmaxf(acc[3][j], 0.f));
      }
      __syncthreads();
      {  // layer 2
        const int mt = t & 15, nt = t >> 4;
        const int m0 = 4 * mt, n0 = 4 * nt;
        float acc[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float bias = Wt[L.pW2 + kH1 * kH2 + n0 + j];
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i][j] = bias; }
        op_tile4<4>(H1 + m0, RS2, Wt + L.pW2 + n0, kH2, kH1, acc);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st4(H2 + (n0 + j) * RS2 + m0, fmaxf(acc[0][j], 0.f), fmaxf(acc[1][j], 0.f), fmaxf(acc[2][j], 0.f), fmaxf(acc[3][j], 0.f));
      }
      __syncthreads();
      {  // head, split-K over 4 thread groups
        const int i = t & 63, part = t >> 6;
        float acc[1 + A];
#pragma unroll
        for (int c = 0; c <= A; ++c) acc[c] = part == 0 ? Wt[L.pWh + kH2 * HC + c] : 0.f;
        const float* wh = Wt + L.pWh;
#pragma unroll 4
        for (int k = 16 * part; k < 16 * part + 16; ++k) {
          const float h = H2[k * RS2 + i];
#pragma unroll
          for (int c = 0; c <= A; ++c) acc[c] = fmaf(h, wh[k * HC + c], acc[c]);
        }
        if (part > 0) {
#pragma unroll
          for (int c = 0; c <= A; ++c) Scr[((part - 1) * BT + i) * HC + c] = acc[c];
        }
        __syncthreads();
        if (part == 0) {
#pragma unroll
          for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int c = 0; c <= A; ++c) acc[c] += Scr[(p * BT + i) * HC + c];
          float msum = 0.f;
#pragma unroll
          for (int j = 1; j <= A; ++j) msum += acc[j];
          const float mean = msum / (float)A;
#pragma unroll
          for (int j = 0; j < A; ++j) QB[i * HC + j] = acc[0] + acc[1 + j] - mean;   // dddqn.py:31
        }
      }
      // (no barrier needed here: the next phase writes H1 only; QB/Scr are re-read after later barriers)

      // ================= batch A: Q(theta, s) and Q(theta, s')  (:52-53) ====================
      {  // layer 1: 128 rows
        const int mt = t & 31, nt = t >> 5;
        const int m0 = 4 * mt, n0 = 4 * nt;
        float acc[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float bias = W[D * kH1 + n0 + j];
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i][j] = bias; }
        op_tile4<4>(X + m0, RS2, W + n0, kH1, D, acc);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st4(H1 + (n0 + j) * RS2 + m0, fmaxf(acc[0][j], 0.f), fmaxf(acc[1][j], 0.f), fmaxf(acc[2][j], 0.f), fmaxf(acc[3][j], 0.f));
      }
      __syncthreads();
      {  // layer 2: 128 rows x 64
        const int mt = t & 31, nt = t >> 5;
        const int m0 = 4 * mt, n0 = 8 * nt;
        float acc[4][8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float bias = W[L.pW2 + kH1 * kH2 + n0 + j];
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i][j] = bias; }
        op_tile4<8>(H1 + m0, RS2, W + L.pW2 + n0, kH2, kH1, acc);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st4(H2 + (n0 + j) * RS2 + m0, fmaxf(acc[0][j], 0.f), fmaxf(acc[1][j], 0.f), fmaxf(acc[2][j], 0.f), fmaxf(acc[3][j], 0.f));
      }
      __syncthreads();
      {  // head for rows i (s) and 64+i (s'), split-K; then targets / loss / d(head) on part 0
        const int i = t & 63, part = t >> 6;
        float as[1 + A], an[1 + A];
#pragma unroll
        for (int c = 0; c <= A; ++c) as[c] = an[c] = part == 0 ? W[L.pWh + kH2 * HC + c] : 0.f;
        const float* wh = W + L.pWh;
#pragma unroll 4
        for (int k = 16 * part; k < 16 * part + 16; ++k) {
          const float hs = H2[k * RS2 + i];
          const float hn = H2[k * RS2 + BT + i];
#pragma unroll
          for (int c = 0; c <= A; ++c) { const float w = wh[k * HC + c]; as[c] = fmaf(hs, w, as[c]); an[c] = fmaf(hn, w, an[c]); }
        }
        if (part > 0) {
#pragma unroll
          for (int c = 0; c <= A; ++c) { Scr[(((part - 1) * BT + i) * 2 + 0) * HC + c] = as[c]; Scr[(((part - 1) * BT + i) * 2 + 1) * HC + c] = an[c]; }
        }
        __syncthreads();
        if (part == 0) {
#pragma unroll
          for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int c = 0; c <= A; ++c) { as[c] += Scr[((p * BT + i) * 2 + 0) * HC + c]; an[c] += Scr[((p * BT + i) * 2 + 1) * HC + c]; }
          float ms = 0.f, mn = 0.f;
#pragma unroll
          for (int j = 1; j <= A; ++j) { ms += as[j]; mn += an[j]; }
          ms = ms / (float)A; mn = mn / (float)A;
          float q[A], nq[A];
#pragma unroll
          for (int j = 0; j < A; ++j) { q[j] = as[0] + as[1 + j] - ms; nq[j] = an[0] + an[1 + j] - mn; }
          // ---- compute_q_targets (q_learning_functions.py:55-59) ----
          int astar = 0; float best = nq[0];
#pragma unroll
          for (int j = 1; j < A; ++j) if (nq[j] > best) { best = nq[j]; astar = j; }   // first max wins
          const int a = Act[i];
          float qa = q[0], nqt = QB[i * HC + 0];
#pragma unroll
          for (int j = 1; j < A; ++j) { if (j == a) qa = q[j]; if (j == astar) nqt = QB[i * HC + j]; }
          const float tv = Rew[i] + (1.0f - Done[i]) * (gamma * nqt - qa);      // :58 (F5 quirk kept)
          const float tgt = qa + tv;                                            // :59
          // ---- compute_loss (:35-36) with pred == q ----
          const float e = qa - tgt;
          const float ae = fabsf(e);
          const float quad = fminf(ae, 1.0f);
          const bool valid = tile * BT + i < B;
          const float l = valid ? 0.5f * quad * quad + (ae - quad) : 0.f;
          const float gi = valid ? fminf(fmaxf(e, -1.0f), 1.0f) / fB : 0.f;     // d mean_i sum_j huber / d pred[i,a]
          // ---- backward through the dueling head: dV = sum_j dQ_j, dAdv = dQ - dV/A ----
          const float dval = gi;
          DhdT[0 * RS1 + i] = dval;
          float dsum[1 + A];
          dsum[0] = dval;
#pragma unroll
          for (int j = 0; j < A; ++j) {
            const float dadv = (j == a ? gi : 0.f) - dval / (float)A;
            DhdT[(1 + j) * RS1 + i] = dadv;
            dsum[1 + j] = dadv;
          }
          // head-bias gradient = column sums of d(head): warp shuffle + one shared atomic per warp
#pragma unroll
          for (int c = 0; c <= A; ++c) {
            const float s = warp_sum(dsum[c]);
            if (lane == 0) atomicAdd(&G[L.pWh + kH2 * HC + c], s);
          }
          loss_acc += warp_sum(l);
          if (args.taps.enabled && valid) {
            const int gi_row = tile * BT + i;
#pragma unroll
            for (int j = 0; j < A; ++j) {
              if (args.taps.q) args.taps.q[gi_row * A + j] = q[j];
              if (args.taps.next_q) args.taps.next_q[gi_row * A + j] = nq[j];
              if (args.taps.next_q_tm) args.taps.next_q_tm[gi_row * A + j] = QB[i * HC + j];
              if (args.taps.targets) args.taps.targets[gi_row * A + j] = (j == a) ? tgt : q[j];
            }
            if (args.taps.max_actions) args.taps.max_actions[gi_row] = astar;
          }
        }
      }
      __syncthreads();

      // ================= backward (jax.grad(compute_loss), :23) =============================
      {  // dh2[r][j] = relu'(h2) * sum_c dhd[r][c] * Wh[j][c]   -> Dh2T (k-major) and Dh2R (row-major)
        const int rt = t & 15, jt = t >> 4;
        float wh[4][1 + A];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
          for (int c = 0; c <= A; ++c) wh[jj][c] = W[L.pWh + (4 * jt + jj) * HC + c];
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          const int r = rt + 16 * ii;
          float dh[1 + A];
#pragma unroll
          for (int c = 0; c <= A; ++c) dh[c] = DhdT[c * RS1 + r];
          float o[4];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            float v = 0.f;
#pragma unroll
            for (int c = 0; c <= A; ++c) v = fmaf(dh[c], wh[jj][c], v);
            o[jj] = H2[(4 * jt + jj) * RS2 + r] > 0.f ? v : 0.f;
            Dh2T[(4 * jt + jj) * RS1 + r] = o[jj];
          }
          st4(Dh2R + r * RS1 + 4 * jt, o[0], o[1], o[2], o[3]);
        }
      }
      __syncthreads();
      {  // (a) dW2[k][j] += sum_r h1[r][k] * dh2[r][j]
        const int mt = t & 15, nt = t >> 4;
        float acc[2][4] = {};
        dot_tile<2, 4>(H1, RS2, mt, 16, Dh2T, RS1, nt, 16, 0, 16, acc);
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) G[L.pW2 + (mt + 16 * i) * kH2 + nt + 16 * j] += acc[i][j];
      }
      {  // (a') db2[j] += sum_r dh2[r][j]  (4-way split over r, shared atomics)
        const int j = t & 63, part = t >> 6;
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) { const float4 v = ld4(Dh2T + j * RS1 + 16 * part + 4 * c); s += (v.x + v.y) + (v.z + v.w); }
        atomicAdd(&G[L.pW2 + kH1 * kH2 + j], s);
      }
      {  // (b) dWh[j][c] += sum_r h2[r][j] * dhd[r][c]   (4-way split over r, shared atomics)
        const int j = t & 63, part = t >> 6;
        float acc[1][1 + A] = {};
        dot_tile<1, 1 + A>(H2, RS2, j, 0, DhdT, RS1, 0, 1, 4 * part, 4, acc);
#pragma unroll
        for (int c = 0; c <= A; ++c) atomicAdd(&G[L.pWh + j * HC + c], acc[0][c]);
      }
      {  // (c) dh1[r][k] = relu'(h1) * sum_j dh2[r][j] * W2[k][j]  -> Dh1T
        const int mt = t & 15, nt = t >> 4;
        float acc[4][2] = {};
        dot_tile<4, 2>(Dh2R, RS1, mt, 16, W + L.pW2, kH2, nt, 16, 0, 16, acc);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int r = mt + 16 * i, k = nt + 16 * j;
            Dh1T[k * RS1 + r] = H1[k * RS2 + r] > 0.f ? acc[i][j] : 0.f;
          }
      }
      __syncthreads();
      {  // (d) [dW1;db1][d][h] += sum_r x[r][d] * dh1[r][h]   (row D of X is ones -> db1)
        const int hcol = t & 31, mt = t >> 5;
        float acc[3][1] = {};
        // rows d = mt, mt+8, mt+16 (<= D); out-of-range rows are clamped for the loads and not stored
        const int d0 = mt, d1 = mt + 8 <= D ? mt + 8 : D, d2 = mt + 16 <= D ? mt + 16 : D;
#pragma unroll 2
        for (int c = 0; c < 16; ++c) {
          const float4 b = ld4(Dh1T + hcol * RS1 + 4 * c);
          const float4 a0 = ld4(X + d0 * RS2 + 4 * c), a1 = ld4(X + d1 * RS2 + 4 * c), a2 = ld4(X + d2 * RS2 + 4 * c);
          acc[0][0] = fmaf(a0.w, b.w, fmaf(a0.z, b.z, fmaf(a0.y, b.y, fmaf(a0.x, b.x, acc[0][0]))));
          acc[1][0] = fmaf(a1.w, b.w, fmaf(a1.z, b.z, fmaf(a1.y, b.y, fmaf(a1.x, b.x, acc[1][0]))));
          acc[2][0] = fmaf(a2.w, b.w, fmaf(a2.z, b.z, fmaf(a2.y, b.y, fmaf(a2.x, b.x, acc[2][0]))));
        }
        if (mt <= D) G[mt * kH1 + hcol] += acc[0][0];
        if (mt + 8 <= D) G[(mt + 8) * kH1 + hcol] += acc[1][0];
        if (mt + 16 <= D) G[(mt + 16) * kH1 + hcol] += acc[2][0];
      }
      // the barrier at the top of the next tile / before Adam orders these G updates
    }  // tiles

    if ((warp == 0 || warp == 1) && lane == 0) Red[warp] = loss_acc;
    __syncthreads();

    if (args.taps.enabled && args.taps.grads) {
      for (int p = t; p < L.PS; p += NT) { const int f = smem_to_flat(p, D, A, L); if (f >= 0) args.taps.grads[f] = G[p]; }
    }
    // ================= optimiser: optax adam / adamw (q_learning_functions.py:24-25) ==========
    {
      const float c1 = Red[8 + 2 * (kstep & 1)], c2 = Red[9 + 2 * (kstep & 1)];
      const float omb1 = 1.0f - b1, omb2 = 1.0f - b2;
#pragma unroll
      for (int i = 0; i < NPT; ++i) {
        const int p = t + i * NT;
        if (p < L.PS) {
          const float g = G[p];
          G[p] = 0.f;
          const float m = b1 * mreg[i] + omb1 * g;
          const float v = b2 * vreg[i] + omb2 * (g * g);
          mreg[i] = m; vreg[i] = v;
          float u = (m / c1) / (sqrtf(v / c2 + eps_root) + eps);
          const float th = W[p];
          if (wd != 0.f) u = u + wd * th;       // add_decayed_weights (adamw)
          W[p] = th + (-lr) * u;                 // scale(-lr); apply_updates
        }
      }
    }
    if (t == 0) {
      const float loss = (Red[0] + Red[1]) / fB;
      args.loss_ring[(size_t)agent * kLossCap + (size_t)((step0 + kstep) % kLossCap)] = loss;
      if (args.taps.enabled && args.taps.loss) args.taps.loss[0] = loss;
    }
    // next step's first barrier (top of tile loop) orders W/G/Red before reuse
  }  // steps

  __syncthreads();
  // ---- write back theta, mu, nu (theta^- is unchanged) --------------------------------------
#pragma unroll
  for (int i = 0; i < NPT; ++i) {
    const int p = t + i * NT;
    if (p < L.PS) {
      const int f = smem_to_flat(p, D, A, L);
      if (f >= 0) { gW[f] = W[p]; gM[f] = mreg[i]; gV[f] = vreg[i]; }
    }
  }
  if (t == 0) {
    ctl->train_steps = step0 + args.K;
    const long long c = (long long)count0 + args.K;
    ctl->adam_count = c > 0x7fffffffLL ? 0x7fffffff : (int)c;
  }
}

typedef void (*TrainKernel)(const TrainArgs);
TrainKernel pick_kernel(int A) {
  switch (A) {
    case 2: return dqn_train_fused_kernel<2>;
    case 3: return dqn_train_fused_kernel<3>;
    case 4: return dqn_train_fused_kernel<4>;
    case 5: return dqn_train_fused_kernel<5>;
    case 6: return dqn_train_fused_kernel<6>;
    case 7: return dqn_train_fused_kernel<7>;
    default: return nullptr;
  }
}

}  // namespace

size_t train_fused_smem_bytes(const Dims& d) { return (size_t)make_layout(d.D, d.recw).total * sizeof(float); }

cudaError_t train_fused_prepare(const Dims& d) {
  TrainKernel k = pick_kernel(d.A);
  if (!k) return cudaErrorInvalidValue;
  return cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)train_fused_smem_bytes(d));
}

cudaError_t launch_train_fused(cudaStream_t st, const TrainArgs& args) {
  TrainKernel k = pick_kernel(args.dims.A);
  if (!k) return cudaErrorInvalidValue;
  k<<<args.n_sel, NT, train_fused_smem_bytes(args.dims), st>>>(args);
  return cudaGetLastError();
}

}  // namespace dqn
