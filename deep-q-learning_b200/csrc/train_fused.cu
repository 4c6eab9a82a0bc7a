// train_fused.cu -- the fused, persistent dueling double-DQN train step (sm_100a, fp32 FFMA).
//
// One CTA = one agent.  For K consecutive train steps it does everything Agent._step does
// (General/QLearning/q_agent.py:146-169) without leaving the SM:
//   sample_batch        replay_buffer.py:68-85        Philox (or supplied) indices; every sampled record is ONE bulk
//                                                      async copy (cp.async.bulk global -> shared, the TMA engine's
//                                                      non-tensor form) completing on an mbarrier, one step ahead
//   preprocessing       q_learning_functions.py:76-85  done -> f32, action -> i32 while unpacking
//   compute_q_targets   q_learning_functions.py:42-64  3 forwards, first-max argmax, F5-quirk TD target
//   compute_loss        q_learning_functions.py:31-39  Huber(delta=1) summed over actions, mean over B
//   train_step          q_learning_functions.py:14-28  hand-derived backward, Adam / AdamW, in place
// theta, theta^- and the gradient accumulator live in shared memory for the whole launch, Adam's
// mu/nu live in registers (fixed thread<->parameter ownership); HBM traffic per step is the
// gathered records (one 96-byte record per sample for D <= 10) and one 4-byte loss.
//
// Shared-memory layouts
//   weights ("smem layout"): [W1;b1] (D+1 x 32) | [W2;b2] (33 x 64, row stride 68) | [Wv Wa | bias row] (65 x 8),
//       i.e. bias = last row of each augmented matrix; head columns: 0 = V, 1..A = advantage, rest 0.
//   activations: k-major ("transposed") [feature][row] with row strides 132 (two 64-row halves) or
//       68; strides are = 4 (mod 32) words so 16-byte accesses of 8 consecutive rows hit 8 distinct
//       bank groups.  Forward GEMMs run in outer-product form on these (both operands contiguous
//       along the output tile), weight-gradient GEMMs in dot form (both operands contiguous along
//       the reduction = batch row), so no explicit transposes are needed.
#include <math.h>

#include "common.cuh"
#include "kernels.h"
#include "tile_ops.cuh"
#include "tile16.cuh"

namespace dqn {

namespace {

using namespace tile;

constexpr int NT = 256;          // threads per CTA
constexpr int BT = 64;           // batch rows per tile
constexpr int RS2 = 2 * BT + 4;  // 132
constexpr int RS1 = BT + 4;      // 68
constexpr int HC = kHeadCols;    // padded head width
constexpr int WS2 = kW2Stride;   // row stride of the [W2;b2] block (68: rows 4 banks apart)
constexpr int NPT = 13;
          // ceil(max smem-layout params / NT), D = 16: (17*32 + 33*64 + 65*8) = 3176

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
// one whole record: global -> shared by the bulk-copy engine, bytes counted on the mbarrier
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct Lay {   // offsets in floats
  int pW2, pWh, PS;
  int oW, oWt, oG, oX, oH1, oH2, oDh2T, oDh2R, oDh1T, oDhdT, oQB, oScr, oMeta, oDummy, oRed, oBar, oStage, total;
};

__host__ __device__ inline Lay make_layout(int D, int recw) {
  Lay L;
  L.pW2 = packed_w2(D);
  L.pWh = packed_head(D);
  L.PS = packed_count(D);          // the shared-memory weight layout IS the packed HBM layout (common.cuh)
  const int PSa = L.PS;
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
  L.oMeta = o; o += BT * 4;
  L.oDummy = o; o += 4;
  L.oRed = o; o += 32;
  L.oBar = o; o += 4;               // mbarrier of the record gather
  L.oStage = o; o += BT * recw;
  L.total = o;
  return L;
}

template <int A>
__global__ void __launch_bounds__(NT, 1) dqn_train_fused_kernel(const TrainArgs args) {
  extern __shared__ __align__(16) float sm[];
  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  // OP-form tile coordinates: a half-warp is 8 m-tiles x 2 n-tiles (128-bit smem loads are served per
  // half-warp, so this keeps every operand load at one wavefront per half-warp)
  const int tmi = (lane & 7) + 8 * (lane >> 4), tni = (lane >> 3) & 1;
  const int sel = blockIdx.x;
  const int agent = args.agent_begin + sel;
  const int D = args.dims.D;
  const int recw = args.dims.recw;
  const int PK = args.dims.PK;
  const Lay L = make_layout(D, recw);
  if (args.gate && !args.gate[agent].train_flag) return;   // episode gate closed (q_agent.py:186): uniform over the CTA / cluster

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
  float* const QB = sm + L.oQB;     // [8 j][64 i]   Q(theta^-, s')
  float* const Scr = sm + L.oScr;   // head split-K partials, [part][s|s'][c][i]
  float* const Meta = sm + L.oMeta; // [64][4]  raw action lo, action hi, reward, done
  float* const Red = sm + L.oRed;   // [0..1] loss partials, [8..11] Adam bias corrections, [16..31] head-bias partials
  float* const Stage = sm + L.oStage;

  float* const gW = args.params + (size_t)agent * 4 * PK;
  float* const gWt = gW + PK;
  float* const gM = gW + 2 * PK;
  float* const gV = gW + 3 * PK;
  AgentCtl* const ctl = args.ctl + agent;
  const uint32_t* const ring = args.rings + (size_t)agent * args.dims.N * recw;

  // ---- one-time: parameters HBM -> smem / registers ---------------------------------------
  float mreg[NPT], vreg[NPT];
#pragma unroll
  for (int i = 0; i < NPT; ++i) {
    const int p = t + i * NT;
    mreg[i] = 0.f; vreg[i] = 0.f;
    if (p < L.PS) { mreg[i] = gM[p]; vreg[i] = gV[p]; }      // padding entries are 0 and stay 0 (zero gradient)
  }
  for (int p4 = t; p4 < (L.PS >> 2); p4 += NT) {   // theta | theta^- : straight 16-byte copies of the packed layout
    const float4 w = reinterpret_cast<const float4*>(gW)[p4], wt = reinterpret_cast<const float4*>(gWt)[p4];
    st4(W + 4 * p4, w.x, w.y, w.z, w.w);
    st4(Wt + 4 * p4, wt.x, wt.y, wt.z, wt.w);
    st4(G + 4 * p4, 0.f, 0.f, 0.f, 0.f);
  }
  for (int r = t; r < RS2; r += NT) { X[D * RS2 + r] = 1.f; H2[kH2 * RS2 + r] = 1.f; }

  const float gamma = ctl->gamma, lr = ctl->lr, b1 = ctl->b1, b2 = ctl->b2;
  const float eps = ctl->eps, eps_root = ctl->eps_root, wd = ctl->wd;
  const int B = ctl->batch_size;
  const bool l2loss = ctl->loss_kind == kLossL2;
  const long long step0 = ctl->train_steps;
  const int count0 = ctl->adam_count;
  const long long rc = ctl->ring_counter;
  const long long size = rc < args.dims.N ? rc : args.dims.N;
  const int ntiles = (B + BT - 1) / BT;
  const float fB = (float)B;
  const int cpr = recw >> 2;
  // Batches of 65..80 rows (the sweep draws up to 70, hyperparameter_optimization.py:121): 64 rows through the 64-row tile
  // code, the rest through the 16-row tile code of tile16.cuh -- a second 64-row tile for six rows doubled the step and, in
  // a one-wave launch (128 agents per GPU at 8 GPUs), the whole launch.  Its buffers alias X | H1 | H2, dead between tiles.
  const bool tail16 = B > BT && B - BT <= t16::R;
  const int nfull = tail16 ? 1 : ntiles;
  t16::Bufs tb;
  tb.W = W; tb.Wt = Wt; tb.G = G; tb.Red = Red; tb.pW2 = L.pW2; tb.pWh = L.pWh; tb.D = D;
  tb.carve(sm + L.oX);

  // Record word -> smem destination of this thread's staged chunks (fixed for the whole launch):
  // thread t owns 16-byte chunks (t&3), (t&3)+4, (t&3)+8 of staged row t>>2.
  const int urow = t >> 2, ul4 = t & 3;
  int udst[12];
#pragma unroll
  for (int q = 0; q < 12; ++q) {
    const int wd = 4 * (ul4 + 4 * (q >> 2)) + (q & 3);
    int o = L.oDummy;
    if (wd < D) o = L.oX + wd * RS2 + urow;
    else if (wd < 2 * D) o = L.oX + (wd - D) * RS2 + BT + urow;
    else if (wd < 2 * D + 4) o = L.oMeta + urow * 4 + (wd - 2 * D);
    udst[q] = o;
  }

  // gather of (step kstep, tile) into the staging buffer: one thread per row issues ONE bulk copy of the whole record
  // (96 B for D <= 10); thread 0 posts the expected byte count of the tile on the mbarrier every thread waits on
  const uint32_t bar = smem_addr(sm + L.oBar);
  uint32_t bar_parity = 0;
  if (t == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  auto prefetch = [&](int kstep, int tile) {
    const int nvalid = B - tile * BT < BT ? B - tile * BT : BT;
    if (t == 0) mbar_arrive_expect_tx(bar, (uint32_t)(nvalid * recw * 4));
    if (ul4 == 0) {
      const int i = tile * BT + urow;
      float* dst = Stage + urow * recw;
      if (i < B) {
        long long slot;
        if (args.idx) slot = args.idx[((size_t)sel * args.K + kstep) * B + i];
        else slot = philox_index(args.seed, args.agent_id_base + agent, step0 + kstep, i, size);
        bulk_load(smem_addr(dst), ring + slot * recw, (uint32_t)(recw * 4), bar);
        if (args.taps.enabled && args.taps.indices) args.taps.indices[i] = slot;
      } else {
        for (int c = 0; c < cpr; ++c) st4(dst + 4 * c, 0.f, 0.f, 0.f, 0.f);
      }
    }
  };
  auto gather_wait = [&]() { mbar_wait(bar, bar_parity); bar_parity ^= 1u; };

  double pb1 = ctl->pb1, pb2 = ctl->pb2;   // b1**count, b2**count carried across launches (thread 0 uses them)

  prefetch(0, 0);

  for (int kstep = 0; kstep < args.K; ++kstep) {
    if (t == 0) {
      // optax bias correction 1 - decay**count with decay**count rounded once to fp32 (oracle pow_f32).
      // decay**count is carried in double across the fused steps (one DMUL per step instead of a pow()).
      // Double-buffered by step parity: slower warps may still be reading the previous step's pair.
      if (count0 + kstep < 0x7fffffff) { pb1 *= (double)b1; pb2 *= (double)b2; }   // safe_int32_increment saturates
      Red[8 + 2 * (kstep & 1)] = 1.0f - (float)pb1;
      Red[9 + 2 * (kstep & 1)] = 1.0f - (float)pb2;
    }
    float loss_acc = 0.f;   // meaningful in warps 0,1

    for (int tile = 0; tile < nfull; ++tile) {
      // ---- unpack staged records (k-major X, raw meta words): part of preprocessing (:76-85) ----
      gather_wait();
      __syncthreads();
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        const int c = ul4 + 4 * cc;
        if (c < cpr) {
          const float4 v = ld4(Stage + urow * recw + 4 * c);
          sm[udst[4 * cc + 0]] = v.x; sm[udst[4 * cc + 1]] = v.y; sm[udst[4 * cc + 2]] = v.z; sm[udst[4 * cc + 3]] = v.w;
        }
      }
      __syncthreads();
      if (tile + 1 < ntiles) prefetch(kstep, tile + 1);
      else if (kstep + 1 < args.K) prefetch(kstep + 1, 0);

      // ================= batch B: Q(theta^-, s')  (q_learning_functions.py:54) ==============
      {  // layer 1: rows = s' (X cols 64..127) -> H1 cols 0..63
        const int m0 = 4 * tmi, n0 = 2 * (tni + 2 * warp);
        u64 acc[4][1];
        op_init<2>(Wt + D * kH1 + n0, acc);
        op_tile4<2>(X + BT + m0, RS2, Wt + n0, kH1, D, acc);
        op_store_relu<2>(H1 + n0 * RS2 + m0, RS2, acc);
      }
      __syncthreads();
      {  // layer 2
        const int m0 = 4 * tmi, n0 = 4 * (tni + 2 * warp);
        u64 acc[4][2];
        op_init<4>(Wt + L.pW2 + kH1 * WS2 + n0, acc);
        op_tile4<4>(H1 + m0, RS2, Wt + L.pW2 + n0, WS2, kH1, acc);
        op_store_relu<4>(H2 + n0 * RS2 + m0, RS2, acc);
      }
      __syncthreads();
      {  // head, split-K over 4 thread groups; packed pairs over head columns (col 0 = V, 1..A = advantage)
        const int i = t & 63, part = t >> 6;
        constexpr int NP = (A + 2) / 2;
        u64 acc2[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) acc2[q] = part == 0 ? ld64(Wt + L.pWh + kH2 * HC + 2 * q) : 0ull;
        const float* wh = Wt + L.pWh;
#pragma unroll 8
        for (int k = 16 * part; k < 16 * part + 16; ++k) {
          const float h = H2[k * RS2 + i];
          const u64 hh = pack2(h, h);
          const u64x2 w0 = ld2x64(wh + k * HC);
          ffma2(acc2[0], hh, w0.lo);
          if constexpr (NP > 1) ffma2(acc2[1], hh, w0.hi);
          if constexpr (NP > 2) { const u64x2 w1 = ld2x64(wh + k * HC + 4); ffma2(acc2[2], hh, w1.lo); if constexpr (NP > 3) ffma2(acc2[3], hh, w1.hi); }
        }
        float acc[2 * NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) unpack2(acc2[q], acc[2 * q], acc[2 * q + 1]);
        if (part > 0) {
#pragma unroll
          for (int c = 0; c <= A; ++c) Scr[((part - 1) * HC + c) * BT + i] = acc[c];
        }
        __syncthreads();
        if (part == 0) {
#pragma unroll
          for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int c = 0; c <= A; ++c) acc[c] += Scr[(p * HC + c) * BT + i];
          float msum = 0.f;
#pragma unroll
          for (int j = 1; j <= A; ++j) msum += acc[j];
          const float mean = msum / (float)A;
#pragma unroll
          for (int j = 0; j < A; ++j) QB[j * BT + i] = acc[0] + acc[1 + j] - mean;   // dddqn.py:31
        }
      }
      // (no barrier: the next phase writes H1 only; QB / Scr are touched again only after later barriers)

      // ================= batch A: Q(theta, s) and Q(theta, s')  (:52-53) ====================
      {  // layer 1: 128 rows
        const int m0 = 4 * (tmi + 16 * (warp & 1)), n0 = 4 * (tni + 2 * (warp >> 1));
        u64 acc[4][2];
        op_init<4>(W + D * kH1 + n0, acc);
        op_tile4<4>(X + m0, RS2, W + n0, kH1, D, acc);
        op_store_relu<4>(H1 + n0 * RS2 + m0, RS2, acc);
      }
      __syncthreads();
      {  // layer 2: 128 rows x 64; a warp covers 64 rows x 16 columns (4 smem wavefronts per k-step)
        const int m0 = 4 * (tmi + 16 * (warp & 1)), n0 = 8 * (tni + 2 * (warp >> 1));
        u64 acc[4][4];
        op_init<8>(W + L.pW2 + kH1 * WS2 + n0, acc);
        op_tile4<8>(H1 + m0, RS2, W + L.pW2 + n0, WS2, kH1, acc);
        op_store_relu<8>(H2 + n0 * RS2 + m0, RS2, acc);
      }
      __syncthreads();
      {  // head for rows i (s) and 64+i (s'), split-K; then targets / loss / d(head) on part 0
        const int i = t & 63, part = t >> 6;
        constexpr int NP = (A + 2) / 2;
        u64 as2[NP], an2[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) as2[q] = an2[q] = part == 0 ? ld64(W + L.pWh + kH2 * HC + 2 * q) : 0ull;
        const float* wh = W + L.pWh;
#pragma unroll 8
        for (int k = 16 * part; k < 16 * part + 16; ++k) {
          const float hs = H2[k * RS2 + i];
          const float hn = H2[k * RS2 + BT + i];
          const u64 hs2 = pack2(hs, hs), hn2 = pack2(hn, hn);
          const u64x2 w0 = ld2x64(wh + k * HC);
          ffma2(as2[0], hs2, w0.lo); ffma2(an2[0], hn2, w0.lo);
          if constexpr (NP > 1) { ffma2(as2[1], hs2, w0.hi); ffma2(an2[1], hn2, w0.hi); }
          if constexpr (NP > 2) {
            const u64x2 w1 = ld2x64(wh + k * HC + 4);
            ffma2(as2[2], hs2, w1.lo); ffma2(an2[2], hn2, w1.lo);
            if constexpr (NP > 3) { ffma2(as2[3], hs2, w1.hi); ffma2(an2[3], hn2, w1.hi); }
          }
        }
        float as[2 * NP], an[2 * NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) { unpack2(as2[q], as[2 * q], as[2 * q + 1]); unpack2(an2[q], an[2 * q], an[2 * q + 1]); }
        if (part > 0) {
#pragma unroll
          for (int c = 0; c <= A; ++c) {
            Scr[(((part - 1) * 2 + 0) * HC + c) * BT + i] = as[c];
            Scr[(((part - 1) * 2 + 1) * HC + c) * BT + i] = an[c];
          }
        }
        __syncthreads();
        if (part == 0) {
#pragma unroll
          for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int c = 0; c <= A; ++c) { as[c] += Scr[((p * 2 + 0) * HC + c) * BT + i]; an[c] += Scr[((p * 2 + 1) * HC + c) * BT + i]; }
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
          const float4 meta = ld4(Meta + 4 * i);
          int a = __float_as_int(meta.x);
          a = a < 0 ? 0 : (a >= A ? A - 1 : a);                 // jax clamps out-of-range gather indices
          const float rew = meta.z;
          const float done = __float_as_uint(meta.w) ? 1.f : 0.f;   // dones.astype(float32), :84
          float qa = q[0], nqt = QB[i];
#pragma unroll
          for (int j = 1; j < A; ++j) { if (j == a) qa = q[j]; if (j == astar) nqt = QB[j * BT + i]; }
          const float tv = rew + (1.0f - done) * (gamma * nqt - qa);              // :58 (F5 quirk kept)
          const float tgt = qa + tv;                                            // :59
          // ---- compute_loss (:35-36) with pred == q ----
          const float e = qa - tgt;
          const float ae = fabsf(e);
          const float quad = fminf(ae, 1.0f);
          const bool valid = tile * BT + i < B;
          // (l2 loss: 0.5 e^2 = the Huber expression without the clip at delta; not in the reference)
          const float l = valid ? (l2loss ? 0.5f * e * e : 0.5f * quad * quad + (ae - quad)) : 0.f;
          const float gi = valid ? (l2loss ? e : fminf(fmaxf(e, -1.0f), 1.0f)) / fB : 0.f;     // d mean_i sum_j loss / d pred[i,a]
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
          // head-bias gradient = column sums of d(head): warp shuffle, then a fixed-order add of the two
          // warps' sums in the next phase (no atomics anywhere: results are run-to-run bit-identical)
#pragma unroll
          for (int c = 0; c <= A; ++c) {
            const float s = warp_sum(dsum[c]);
            if (lane == 0) Red[16 + warp * 8 + c] = s;
          }
          loss_acc += warp_sum(l);
          if (args.taps.enabled && valid) {
            const int gi_row = tile * BT + i;
#pragma unroll
            for (int j = 0; j < A; ++j) {
              if (args.taps.q) args.taps.q[gi_row * A + j] = q[j];
              if (args.taps.next_q) args.taps.next_q[gi_row * A + j] = nq[j];
              if (args.taps.next_q_tm) args.taps.next_q_tm[gi_row * A + j] = QB[j * BT + i];
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
        if (t <= A) G[L.pWh + kH2 * HC + t] += Red[16 + t] + Red[24 + t];     // d(head bias), fixed order
        float wh[4][HC];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const float4 w0 = ld4(W + L.pWh + (4 * jt + jj) * HC), w1 = ld4(W + L.pWh + (4 * jt + jj) * HC + 4);
          wh[jj][0] = w0.x; wh[jj][1] = w0.y; wh[jj][2] = w0.z; wh[jj][3] = w0.w;
          wh[jj][4] = w1.x; wh[jj][5] = w1.y; wh[jj][6] = w1.z; wh[jj][7] = w1.w;
        }
        float bsum[4] = {0.f, 0.f, 0.f, 0.f};
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
            bsum[jj] += o[jj];
          }
          st4(Dh2R + r * RS1 + 4 * jt, o[0], o[1], o[2], o[3]);
        }
        // db2[j] += sum_r dh2[r][j]: butterfly over the 16 lanes that share jt (fixed order)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) bsum[jj] += __shfl_xor_sync(0xffffffffu, bsum[jj], o);
        }
        if (rt == 0) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) G[L.pW2 + kH1 * WS2 + 4 * jt + jj] += bsum[jj];
        }
      }
      __syncthreads();
      // (a) and (c) below are smem-bandwidth critical: 4x4 register tiles with a 2-way split of the
      // reduction across adjacent lanes (lane&1 takes the even / odd 16-byte chunks) halve the operand
      // bytes per FMA; rows inside a half-warp are 2 apart so the 8 (row, chunk-parity) pairs of every
      // 128-bit load hit 8 distinct bank groups (one wavefront per half-warp).
      const int kh = lane & 1, ntl = (lane >> 1) & 3, mtl = (lane >> 3) & 1, hwid = t >> 4;
      {  // (a) dW2[k][j] += sum_r h1[r][k] * dh2[r][j]
        const int k0 = 16 * ((hwid >> 1) & 1) + (hwid & 1) + 2 * mtl;          // rows k0 + 4 i
        const int j0 = 32 * (hwid >> 3) + ((hwid >> 2) & 1) + 2 * ntl;         // rows j0 + 8 jj
        const float* ap[4];
        const float* bp[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { ap[i] = H1 + (k0 + 4 * i) * RS2 + 4 * kh; bp[i] = Dh2T + (j0 + 8 * i) * RS1 + 4 * kh; }
        float acc[4][4];
        dot_tile<4, 4, 8, 8>(ap, bp, acc);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] += __shfl_xor_sync(0xffffffffu, acc[i][j], 1);
#pragma unroll
        for (int ii = 0; ii < 2; ++ii) {      // lane kh finishes rows i = kh, kh + 2
          const int i = kh + 2 * ii;
#pragma unroll
          for (int j = 0; j < 4; ++j) G[L.pW2 + (k0 + 4 * i) * WS2 + j0 + 8 * j] += kh ? acc[1 + 2 * ii][j] : acc[2 * ii][j];
        }
      }
      {  // (b) dWh[j][c] += sum_r h2[r][j] * dhd[r][c]: 4-way split over r inside a warp, shuffle-reduced
        const int j = warp * 8 + (lane & 7), part = lane >> 3;
        const float* ap[1] = {H2 + j * RS2 + 16 * part};
        const float* bp[1 + A];
#pragma unroll
        for (int c = 0; c <= A; ++c) bp[c] = DhdT + c * RS1 + 16 * part;
        float acc[1][1 + A];
        dot_tile<1, 1 + A, 4>(ap, bp, acc);
#pragma unroll
        for (int c = 0; c <= A; ++c) {
          float v = acc[0][c];
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (part == 0) G[L.pWh + j * HC + c] += v;
        }
      }
      {  // (c) dh1[r][k] = relu'(h1) * sum_j dh2[r][j] * W2[k][j]  -> Dh1T
        const int r0 = 16 * (hwid >> 2) + ((hwid >> 1) & 1) + 2 * mtl;           // rows r0 + 4 i
        const int k0 = (hwid & 1) + 2 * ntl;                                      // units k0 + 8 jj
        const float* ap[4];
        const float* bp[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { ap[i] = Dh2R + (r0 + 4 * i) * RS1 + 4 * kh; bp[i] = W + L.pW2 + (k0 + 8 * i) * WS2 + 4 * kh; }
        float acc[4][4];
        dot_tile<4, 4, 8, 8>(ap, bp, acc);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] += __shfl_xor_sync(0xffffffffu, acc[i][j], 1);
#pragma unroll
        for (int ii = 0; ii < 2; ++ii) {
          const int r = r0 + 4 * (kh + 2 * ii);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = k0 + 8 * j;
            const float v = kh ? acc[1 + 2 * ii][j] : acc[2 * ii][j];
            Dh1T[k * RS1 + r] = H1[k * RS2 + r] > 0.f ? v : 0.f;
          }
        }
      }
      __syncthreads();
      {  // (d) [dW1;db1][d][h] += sum_r x[r][d] * dh1[r][h]   (row D of X is ones -> db1)
        const int hcol = t & 31, mt = t >> 5;
        // rows d = mt, mt+8, mt+16 (<= D); out-of-range rows are clamped for the loads and not stored
        const int d1 = mt + 8 <= D ? mt + 8 : D, d2 = mt + 16 <= D ? mt + 16 : D;
        const float* bp[1] = {Dh1T + hcol * RS1};
        if (D <= 7) {
          const float* ap[1] = {X + mt * RS2};
          float acc[1][1];
          dot_tile<1, 1, 16>(ap, bp, acc);
          if (mt <= D) G[mt * kH1 + hcol] += acc[0][0];
        } else if (D <= 15) {
          const float* ap[2] = {X + mt * RS2, X + d1 * RS2};
          float acc[2][1];
          dot_tile<2, 1, 16>(ap, bp, acc);
          G[mt * kH1 + hcol] += acc[0][0];
          if (mt + 8 <= D) G[(mt + 8) * kH1 + hcol] += acc[1][0];
        } else {
          const float* ap[3] = {X + mt * RS2, X + d1 * RS2, X + d2 * RS2};
          float acc[3][1];
          dot_tile<3, 1, 16>(ap, bp, acc);
          G[mt * kH1 + hcol] += acc[0][0];
          G[(mt + 8) * kH1 + hcol] += acc[1][0];
          if (mt + 16 <= D) G[(mt + 16) * kH1 + hcol] += acc[2][0];
        }
      }
      // the barrier at the top of the next tile / before Adam orders these G updates
    }  // tiles

    if (tail16) {   // rows 64 .. B-1 (staged by the prefetch of "tile 1": rows >= B are zero-filled)
      gather_wait();
      __syncthreads();                       // also: every thread is done with tile 0 (its buffers are about to be reused)
      if (t < 4 * t16::R) {
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          const int c = ul4 + 4 * cc;
          if (c < cpr) {
            const float4 v = ld4(Stage + urow * recw + 4 * c);
            const float w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int wd = 4 * c + j;
              if (wd < D) tb.X[wd * t16::XS + urow] = w4[j];
              else if (wd < 2 * D) tb.X[(wd - D) * t16::XS + t16::R + urow] = w4[j];
              else if (wd < 2 * D + 4) tb.Meta[urow * 4 + (wd - 2 * D)] = w4[j];
            }
          }
        }
      }
      for (int r = t; r < t16::HS; r += NT) { if (r < t16::XS) tb.X[D * t16::XS + r] = 1.f; tb.H2[kH2 * t16::HS + r] = 1.f; }
      __syncthreads();
      if (kstep + 1 < args.K) prefetch(kstep + 1, 0);
      const float lp = t16::step<A>(tb, BT, B, fB, gamma, l2loss, args.taps);
      if (t < t16::R) loss_acc += lp;        // warp 0 (its lane 0 publishes the sum below)
      __syncthreads();
      for (int r = t; r < RS2; r += NT) { X[D * RS2 + r] = 1.f; H2[kH2 * RS2 + r] = 1.f; }   // the 64-row layout's ones rows were overwritten
    }

    if ((warp == 0 || warp == 1) && lane == 0) Red[warp] = loss_acc;
    __syncthreads();

    if (args.taps.enabled && args.taps.grads) {
      for (int p = t; p < L.PS; p += NT) args.taps.grads[p] = G[p];
    }
    // ================= optimiser: optax adam / adamw (q_learning_functions.py:24-25) ==========
    // mu/nu are updated in full precision; the bias-corrected quotient uses the SFU reciprocal / sqrt
    // (<= 2 ulp each, i.e. ~1e-7 relative on an update that is itself lr ~ 1e-4 of the weight scale).
    {
      const float c1 = Red[8 + 2 * (kstep & 1)], c2 = Red[9 + 2 * (kstep & 1)];
      const float rc1 = 1.0f / c1, rc2 = 1.0f / c2;
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
          const float u = (m * rc1) * fast_rcp(fast_sqrt(v * rc2 + eps_root) + eps);
          const float th = W[p];
          W[p] = th - lr * (u + wd * th);        // add_decayed_weights (wd = 0 for adam); scale(-lr); apply_updates
        }
      }
    }
    if (t == 0) {
      const float loss = (Red[0] + Red[1]) / fB;
      args.loss_ring[(size_t)agent * kLossCap + (size_t)((step0 + kstep) % kLossCap)] = loss;
      if (kstep == args.K - 1)   // host-visible without a D2H copy; the step count makes it pollable
        args.loss_mailbox[agent] = ((unsigned long long)(uint32_t)(step0 + args.K) << 32) | __float_as_uint(loss);
      if (args.taps.enabled && args.taps.loss) args.taps.loss[0] = loss;
    }
    // next step's first barrier (top of tile loop) orders W/G/Red before reuse
  }  // steps

  __syncthreads();
  // ---- write back theta, mu, nu (theta^- is unchanged) --------------------------------------
#pragma unroll
  for (int i = 0; i < NPT; ++i) {
    const int p = t + i * NT;
    if (p < L.PS) { gW[p] = W[p]; gM[p] = mreg[i]; gV[p] = vreg[i]; }
  }
  if (t == 0) {
    ctl->train_steps = step0 + args.K;
    const long long c = (long long)count0 + args.K;
    ctl->adam_count = c > 0x7fffffffLL ? 0x7fffffff : (int)c;
    ctl->pb1 = pb1; ctl->pb2 = pb2;
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
