// train_fused_tc_tail.cuh -- body of the tensor-core population kernel (train_fused_tc.cu) for batches of 65..80 rows: the
// structure of train_fused_tc_main.cuh (read its header first) with wider buffers.  Such a batch (the sweep draws up to 70,
// hyperparameter_optimization.py:121) stays ONE tile: the up to 16 extra "tail" rows ride in lanes and reduction steps the
// 64-row tile leaves idle -- their h1 in the lower half of P1's A tile, with P1's B operand widened to [W2^- | W2] (N = 128:
// same number of tcgen05.mma), their dh2^T columns as two more reduction steps of the backward product (zeros in the
// matching columns of the dh1 half) and their 16 x 32 dh1 on the CUDA cores under that product.
#pragma once
#include <math.h>

#include "common.cuh"
#include "kernels.h"
#include "tile_ops.cuh"
#include "tc_prims.cuh"

namespace dqn {
namespace tct {

using namespace tile;
using namespace tcp;


constexpr int NT = 512;          // threads per CTA
constexpr int NW = NT / 32;
constexpr int BT = 64;           // batch rows per tile
constexpr int TL = 16;           // + up to 16 "tail" rows riding in the same tile (batches of 65..80 rows)
constexpr int NR = BT + TL;      // 80
constexpr int XS = 2 * NR + 4;   // 164: row stride of X [d][s 0..63 | s' 64..127 | tail s 128..143 | tail s' 144..159]
constexpr int RS1 = NR + 4;      // 84: row stride of the [unit][batch row] buffers (rows 64..79 = tail)
constexpr int HC = kHeadCols;
constexpr int WS2 = kW2Stride;
constexpr int NPP = 4;           // parameter PAIRS per thread: pair i of thread t = packed entries 2 (t + i NT), + 1 (3176 entries at D = 16;
                                 // the fourth pair is only there when the agent has more than 6 NT entries)

// Shared-memory B operands, K-major no-swizzle canonical layout (cute: ((8,n),(T,2)):((1T,SBO),(1,LBO)), T = 4 tf32):
//   byte(mn, k) = (mn / 8) * SBO + (k / 4) * LBO + (mn % 8) * 16 + (k % 4) * 4        one k-step of 8 = two 16-byte chunks
//   B2 / Bt2  W2 / W2^- as (mn = j 64, K = k 32): LBO 128, SBO 1024 (dense; written as 16-byte chunks)
//   B4        W2 as (mn = k 32, K = j 64),  B3  h1 of the s rows as (mn = k 32, K = r 64): LBO 144, SBO 2304 -- the 16 spare
//             bytes per chunk column put the 4-byte stores of 32 consecutive r (or the chunks of 8 consecutive k) on
//             32 different banks
constexpr int kLbo2 = 128, kSbo2 = 1024, kB2Bytes = 8 * kSbo2;          // 8192
constexpr int kLbo4 = 144, kSbo4 = (NR / 4) * kLbo4, kB4Bytes = 4 * kSbo4;    // 20 chunk columns (K up to 80): 2880, 11520
// Tensor-memory columns (512 allocated; lane = row)
// D1 is 128 wide: for a batch with tail rows P1's B operand is [W2^- | W2] (see below); the backward accumulator reuses D2's columns
constexpr uint32_t kA2Hi = 0, kA2Lo = 32, kA1Hi = 64, kA1Lo = 96, kD2 = 128, kD1 = 192, kAbHi = 320, kAbLo = 400, kD4 = 128, kD3 = 160;
// B4 and B3 are adjacent (hi: B4 | B3, then lo: B4 | B3), so ONE operand of mn = 64 (SBO groups 0..3 = W2, 4..7 = h1) feeds a single
// N = 64 product whose accumulator columns are [D4 | D3]: lanes 0..63 of D4 and lanes 64..127 of D3 are the useful halves
constexpr int kBbHalf = 2 * kB4Bytes;   // bytes from a hi buffer to its lo twin

struct Lay {   // offsets in floats
  int pW2, pWh, PS;
  int oW, oWt, oG, oX, oH2, oDh1T, oDhdT, oHP, oWhT, oDh2Tl, oMeta, oDummy, oRed, oBar, oStage, oBt2, oB2, oB4, oB3, total;
};
constexpr int kBfHalf = 2 * kB2Bytes;   // forward B operands: hi = [W2^- | W2] (one mn = 128 operand), then lo = [W2^- | W2]
constexpr int kRedLoss = 32, kRedDb2 = kRedLoss + NR;

__host__ __device__ inline Lay make_layout(int D, int recw) {
  Lay L;
  L.pW2 = packed_w2(D);
  L.pWh = packed_head(D);
  L.PS = packed_count(D);          // the shared-memory weight layout IS the packed HBM layout (common.cuh)
  int o = 0;
  L.oW = o; o += L.PS;
  L.oWt = o; o += L.PS;
  L.oG = o; o += L.PS;
  L.oX = o; o += (D + 1) * XS;        // [D+1][164]  cols: s | s' | tail s | tail s'; row D = ones
  L.oH2 = o; o += kH2 * RS1;          // [64 j][84]   h2(theta, s) incl. tail rows: dWh, relu' of the per-unit dh2 pass
  L.oDh1T = o; o += kH1 * RS1;        // [32 k][84]
  L.oDhdT = o; o += HC * RS1;         // [8 c][84]
  L.oHP = o; o += 3 * 4 * HC * NR;    // head partials [forward set][column quarter][c][row]
  L.oWhT = o; o += HC * kH2;          // [8 c][64 j]  head weights of theta, transposed (rebuilt every step)
  L.oDh2Tl = o; o += TL * (kH2 + 4) + TL;  // [16 tail rows][68] dh2 of the tail rows (their dh1 is computed on the CUDA cores) | [16] relu' bits of their h1
  L.oMeta = o; o += NR * 4;
  L.oDummy = o; o += 4;
  L.oRed = o; o += kRedDb2 + 4 * kH2; // [0..1] loss, [8..11] Adam bias corrections, [32..111] per-row loss, then db2 partials [quarter][j]
  L.oBar = o; o += 8;                 // [0,1] mbarrier of the record gather, [2,3] mbarrier of the MMAs, [4] TMEM base address, [6,7] mbarrier of P1
  L.oStage = o; o += NR * recw;
  L.oBt2 = o; o += kB2Bytes / 4;      // hi: W2^- | W2, lo: W2^- | W2
  L.oB2 = o; o += 3 * kB2Bytes / 4;
  L.oB4 = o; o += kB4Bytes / 4;       // hi: B4 | B3, lo: B4 | B3
  L.oB3 = o; o += 3 * kB4Bytes / 4;
  L.total = o;
  return L;
}

template <int A>
__device__ __forceinline__ void step_body(const TrainArgs& args, const int sel) {
  extern __shared__ __align__(16) float sm[];
  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  const int agent = args.agent_begin + sel;
  const int D = args.dims.D;
  const int recw = args.dims.recw;
  const int PK = args.dims.PK;
  const Lay L = make_layout(D, recw);
  if (args.gate && !args.gate[agent].train_flag) return;   // episode gate closed (q_agent.py:186): uniform over the CTA

  float* const W = sm + L.oW;       // theta      (packed layout)
  float* const Wt = sm + L.oWt;     // theta^-
  float* const G = sm + L.oG;       // gradient accumulator
  float* const X = sm + L.oX;
  float* const H2 = sm + L.oH2;
  float* const Dh1T = sm + L.oDh1T;
  float* const DhdT = sm + L.oDhdT;
  float* const HP = sm + L.oHP;
  float* const WhT = sm + L.oWhT;
  float* const Dh2Tl = sm + L.oDh2Tl;
  uint8_t* const M1T = reinterpret_cast<uint8_t*>(Dh2Tl + TL * (kH2 + 4));   // [16 tail rows][4 column quarters]
  float* const Qs = Dh1T;                                   // [3 forward sets][8][80 rows] Q-values (Dh1T is dead until the backward epilogue)
  float* const DW1P = sm + L.oB3;                           // [5 row groups][17 * 32] partial dW1 (the h1 operand is dead after the backward product)
  float* const Meta = sm + L.oMeta; // [80][4]  raw action lo, action hi, reward, done
  float* const Red = sm + L.oRed;
  float* const Stage = sm + L.oStage;
  const int wq = warp & 3, wc = warp >> 2;          // TMEM quadrant (lanes 32 wq ..), column quarter
  const int row2 = 32 * wq + lane;                  // forward row of batch A: < 64 = s row, else s' row (row2 - 64); backward: r | 64 + j
  const uint32_t tlane = (uint32_t)(32 * wq) << 16;
  const uint32_t sBt2 = smem_addr(sm + L.oBt2), sB2 = smem_addr(sm + L.oB2), sB4 = smem_addr(sm + L.oB4);
  const uint32_t bar = smem_addr(sm + L.oBar), mbar = smem_addr(sm + L.oBar + 2), mbar1 = smem_addr(sm + L.oBar + 6);

  float* const gW = args.params + (size_t)agent * 4 * PK;
  float* const gWt = gW + PK;
  float* const gM = gW + 2 * PK;
  float* const gV = gW + 3 * PK;
  AgentCtl* const ctl = args.ctl + agent;
  const uint32_t* const ring = args.rings + (size_t)agent * args.dims.N * recw;

  // ---- one-time: parameters HBM -> smem / registers ---------------------------------------
  float2 mreg[NPP], vreg[NPP];
  const bool pair4 = L.PS > 6 * NT;    // uniform over the CTA
#pragma unroll
  for (int i = 0; i < NPP; ++i) {
    const int p = 2 * (t + i * NT);
    mreg[i] = make_float2(0.f, 0.f); vreg[i] = make_float2(0.f, 0.f);
    if (p < L.PS) { mreg[i] = *reinterpret_cast<const float2*>(gM + p); vreg[i] = *reinterpret_cast<const float2*>(gV + p); }   // padding entries are 0 and stay 0 (zero gradient)
  }
  for (int p4 = t; p4 < (L.PS >> 2); p4 += NT) {
    const float4 w = reinterpret_cast<const float4*>(gW)[p4], wt = reinterpret_cast<const float4*>(gWt)[p4];
    st4(W + 4 * p4, w.x, w.y, w.z, w.w);
    st4(Wt + 4 * p4, wt.x, wt.y, wt.z, wt.w);
    st4(G + 4 * p4, 0.f, 0.f, 0.f, 0.f);
  }
  for (int r = t; r < XS; r += NT) X[D * XS + r] = 1.f;
  // the K = 64..79 chunk columns of the W2 operand of the backward product stay zero (hi and lo): with them a batch with tail
  // rows extends the reduction of the dW2 half without touching the dh1 half
  for (int q = t; q < 2 * 4 * 144; q += NT) {      // 2 halves (hi, lo) x 4 groups of 8 units x 4 chunk columns of 144 bytes
    const int half = q / 576, rem = q - half * 576, g = rem / 144, w = rem - g * 144;
    reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(sm + L.oB4) + half * kBbHalf + g * kSbo4 + 16 * kLbo4)[w] = 0u;
  }

  const float gamma = ctl->gamma, lr = ctl->lr, b1 = ctl->b1, b2 = ctl->b2;
  const float eps = ctl->eps, eps_root = ctl->eps_root, wd = ctl->wd;
  const int B = ctl->batch_size;
  const bool l2loss = ctl->loss_kind == kLossL2;
  const long long step0 = ctl->train_steps;
  const int count0 = ctl->adam_count;
  const long long rc = ctl->ring_counter;
  const long long size = rc < args.dims.N ? rc : args.dims.N;
  const float fB = (float)B;
  const int cpr = recw >> 2;
  // Batches of 65..80 rows (the sweep draws up to 70, hyperparameter_optimization.py:121) are ONE tile: 64 rows + up to 16
  // "tail" rows that ride in lanes the 64-row tile leaves idle (P1's lower half) and in extra reduction steps of the
  // backward product.  Larger batches are tiles of 64 rows.
  constexpr bool tail = true;   // this body only runs batches of 65..80 rows (dispatch in train_fused_tc.cu)
  const int ntiles = tail ? 1 : (B + BT - 1) / BT;
  const int nslice = tail ? 5 : 4;                  // row groups of the dW1 partial sums

  // record word -> smem destination of this thread's staged chunks: thread t owns 16-byte chunks (t & 7), (t & 7) + 8 of row t >> 3
  // (and, for tail rows, of row 64 + (t >> 3), t < 128)
  const int urow = t >> 3, u8 = t & 7;
  int udst[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int wdx = 4 * (u8 + 8 * (q >> 2)) + (q & 3);
    int o = L.oDummy;
    if (wdx < D) o = L.oX + wdx * XS + urow;
    else if (wdx < 2 * D) o = L.oX + (wdx - D) * XS + BT + urow;
    else if (wdx < 2 * D + 4) o = L.oMeta + urow * 4 + (wdx - 2 * D);
    udst[q] = o;
  }

  uint32_t bar_parity = 0, mma_parity = 0, p1_parity = 0;
  if (t == 0) { mbar_init(bar, 1); mbar_init(mbar, 1); mbar_init(mbar1, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {          // all 512 columns of the SM's tensor memory (one CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(sm + L.oBar + 4)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<const uint32_t*>(sm + L.oBar + 4);
  if (wq < 2) {   // backward A tile, lanes 0..63 (dh2 rows), reduction columns 64..79: zero for good (see the B4 note above)
    uint32_t z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0u;
    if (wc == 0) tmem_st16(tmem + tlane + kAbHi + 64u, z);
    if (wc == 1) tmem_st16(tmem + tlane + kAbLo + 64u, z);
    tmem_wait_st();
  }

  // gather of (step kstep, tile): one thread per row issues ONE bulk copy of the whole record; thread 0 posts the byte count
  auto prefetch = [&](int kstep, int tile) {
    const int nrows = tail ? B : (B - tile * BT < BT ? B - tile * BT : BT);
    if (t == 0) mbar_arrive_expect_tx(bar, (uint32_t)(nrows * recw * 4));
    // four lanes of every warp (rows 0..63) + four more of warps 0..3 (tail rows 64..79): all warps reach the MMA wait together
    const int prow = (t & 7) == 0 ? (t >> 3) : ((t & 7) == 4 && t < 8 * TL ? BT + (t >> 3) : -1);
    if (prow >= 0) {
      const int row = prow;
      const int i = tile * BT + row;
      float* dst = Stage + row * recw;
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
  // W2 as a (mn = j, K = k) operand, hi | lo: 16-byte chunk (j, kc) at (j / 8) * SBO + kc * LBO + (j % 8) * 16
  auto lay_b2 = [&](const float* w2, float* dstf) {
    for (int ci = t; ci < 512; ci += NT) {
      const int j = ci & 63, kc = ci >> 6;
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) split_tf32(w2[(4 * kc + e) * WS2 + j], hi[e], lo[e]);
      uint8_t* dst = reinterpret_cast<uint8_t*>(dstf) + (j >> 3) * kSbo2 + kc * kLbo2 + (j & 7) * 16;
      *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(dst + kBfHalf) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  };
  auto dw1_fold = [&](int p) {   // sum of the row groups' partial dW1 sums, fixed order
    float v = ((DW1P[p] + DW1P[17 * kH1 + p]) + DW1P[2 * 17 * kH1 + p]) + DW1P[3 * 17 * kH1 + p];
    if (nslice == 5) v += DW1P[4 * 17 * kH1 + p];
    return v;
  };
  lay_b2(Wt + L.pW2, sm + L.oBt2);     // theta^- does not change inside a launch

  double pb1 = ctl->pb1, pb2 = ctl->pb2;   // b1**count, b2**count carried across launches (thread 0 uses them)

  prefetch(0, 0);

  for (int kstep = 0; kstep < args.K; ++kstep) {
    if (t == 0) {
      // optax bias correction 1 - decay**count with decay**count rounded once to fp32 (oracle pow_f32), double-buffered by step parity
      if (count0 + kstep < 0x7fffffff) { pb1 *= (double)b1; pb2 *= (double)b2; }   // safe_int32_increment saturates
      Red[8 + 2 * (kstep & 1)] = 1.0f - (float)pb1;
      Red[9 + 2 * (kstep & 1)] = 1.0f - (float)pb2;
    }
    float loss_acc = 0.f;   // meaningful in warp 0

    for (int tile = 0; tile < ntiles; ++tile) {
      // ---- unpack staged records (k-major X, raw meta words): part of preprocessing (:76-85) ----
      gather_wait();
      __syncthreads();
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = u8 + 8 * cc;
        if (c < cpr) {
          const float4 v = ld4(Stage + urow * recw + 4 * c);
          sm[udst[4 * cc + 0]] = v.x; sm[udst[4 * cc + 1]] = v.y; sm[udst[4 * cc + 2]] = v.z; sm[udst[4 * cc + 3]] = v.w;
          if (tail && urow < TL) {    // tail row 64 + urow: columns 128 + urow (s) / 144 + urow (s'), meta row 64 + urow
            const float4 u = ld4(Stage + (BT + urow) * recw + 4 * c);
            const float uv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int wdx = 4 * c + e;
              if (wdx < D) X[wdx * XS + 2 * BT + urow] = uv[e];
              else if (wdx < 2 * D) X[(wdx - D) * XS + 2 * BT + TL + urow] = uv[e];
              else if (wdx < 2 * D + 4) Meta[(BT + urow) * 4 + (wdx - 2 * D)] = uv[e];
            }
          }
        }
      }
      if (tile > 0) {    // fold the previous tile's partial dW1 (Adam folds the last tile's)
        for (int p = t; p < (D + 1) * kH1; p += NT) G[p] += dw1_fold(p);
      }
      if (tile == 0) {   // this step's W2 as tensor-core operands: B2 (mn = j, K = k) and B4 (mn = k, K = j); head weights transposed
        WhT[t] = W[L.pWh + (t & 63) * HC + (t >> 6)];
        lay_b2(W + L.pW2, sm + L.oB2);
        {
          const int k = (t & 7) + 8 * (t >> 7), jc = (t >> 3) & 15;
          const float4 w = ld4(W + L.pW2 + k * WS2 + 4 * jc);
          uint32_t hi[4], lo[4];
          split_tf32(w.x, hi[0], lo[0]); split_tf32(w.y, hi[1], lo[1]); split_tf32(w.z, hi[2], lo[2]); split_tf32(w.w, hi[3], lo[3]);
          uint8_t* dst = reinterpret_cast<uint8_t*>(sm + L.oB4) + (k >> 3) * kSbo4 + jc * kLbo4 + (k & 7) * 16;
          *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(dst + kBbHalf) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      __syncthreads();

      // relu' of this thread's 8 h1 / 16 h2 entries, kept from the forward for the backward (t: of its tail row)
      uint32_t mask1 = 0, mask2 = 0;
      {  // ---- layer 1, one batch row per thread, 8 units: theta on (s | s') = 128 rows; theta^- on s' (lanes 64..127) ----
        u64 acc[4], acct[4];
        const float* w1 = W + 8 * wc;
        const float* w1t = Wt + 8 * wc;
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[i] = ld64(w1 + D * kH1 + 2 * i); acct[i] = ld64(w1t + D * kH1 + 2 * i); }
        if (wq < 2) {
#pragma unroll 4
          for (int d = 0; d < D; ++d) {
            const float x = X[d * XS + row2];
            const u64 xx = pack2(x, x);
            const u64x2 a0 = ld2x64(w1 + d * kH1), a1 = ld2x64(w1 + d * kH1 + 4);
            ffma2(acc[0], xx, a0.lo); ffma2(acc[1], xx, a0.hi); ffma2(acc[2], xx, a1.lo); ffma2(acc[3], xx, a1.hi);
          }
        } else {
#pragma unroll 4
          for (int d = 0; d < D; ++d) {
            const float x = X[d * XS + row2];
            const u64 xx = pack2(x, x);
            const u64x2 a0 = ld2x64(w1 + d * kH1), a1 = ld2x64(w1 + d * kH1 + 4);
            const u64x2 c0 = ld2x64(w1t + d * kH1), c1 = ld2x64(w1t + d * kH1 + 4);
            ffma2(acc[0], xx, a0.lo); ffma2(acc[1], xx, a0.hi); ffma2(acc[2], xx, a1.lo); ffma2(acc[3], xx, a1.hi);
            ffma2(acct[0], xx, c0.lo); ffma2(acct[1], xx, c0.hi); ffma2(acct[2], xx, c1.lo); ffma2(acct[3], xx, c1.hi);
          }
        }
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float a, b;
          unpack2(acc[i], a, b);
          a = fmaxf(a, 0.f); b = fmaxf(b, 0.f);
          mask_push(mask1, a); mask_push(mask1, b);
          split_tf32(a, hi[2 * i], lo[2 * i]); split_tf32(b, hi[2 * i + 1], lo[2 * i + 1]);
        }
        mask1 = mask_finish(mask1, 8);
        tmem_st8(tmem + tlane + kA2Hi + 8u * wc, hi);
        tmem_st8(tmem + tlane + kA2Lo + 8u * wc, lo);
        if (wq < 2) {   // h1 of the s rows is also the B operand of P3 (mn = k, K = r): 4-byte stores, conflict-free (LBO = 144)
          uint8_t* b3 = reinterpret_cast<uint8_t*>(sm + L.oB3) + wc * kSbo4 + (row2 >> 2) * kLbo4 + (row2 & 3) * 4;   // k = 8 wc + i
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            *reinterpret_cast<uint32_t*>(b3 + i * 16) = hi[i];
            *reinterpret_cast<uint32_t*>(b3 + i * 16 + kBbHalf) = lo[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float a, b;
            unpack2(acct[i], a, b);
            split_tf32(fmaxf(a, 0.f), hi[2 * i], lo[2 * i]); split_tf32(fmaxf(b, 0.f), hi[2 * i + 1], lo[2 * i + 1]);
          }
          tmem_st8(tmem + tlane + kA1Hi + 8u * wc, hi);
          tmem_st8(tmem + tlane + kA1Lo + 8u * wc, lo);
        }
        if (tail && wq < 2) {
          // tail rows in the lower half of P1's A tile: lanes 0..15 h1(theta; tail s), 16..31 h1(theta; tail s'), 32..47
          // h1(theta^-; tail s') (lanes 48..63 repeat them: tcgen05.st is warp-wide).  X columns 128 + lane / 144 + (lane & 15).
          const float* wl = wq == 0 ? w1 : w1t;
          const int xcol = wq == 0 ? 2 * BT + lane : 2 * BT + TL + (lane & 15);
          uint32_t mask1t = 0;
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] = ld64(wl + D * kH1 + 2 * i);
#pragma unroll 4
          for (int d = 0; d < D; ++d) {
            const float x = X[d * XS + xcol];
            const u64 xx = pack2(x, x);
            const u64x2 a0 = ld2x64(wl + d * kH1), a1 = ld2x64(wl + d * kH1 + 4);
            ffma2(acc[0], xx, a0.lo); ffma2(acc[1], xx, a0.hi); ffma2(acc[2], xx, a1.lo); ffma2(acc[3], xx, a1.hi);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float a, b;
            unpack2(acc[i], a, b);
            a = fmaxf(a, 0.f); b = fmaxf(b, 0.f);
            mask_push(mask1t, a); mask_push(mask1t, b);
            split_tf32(a, hi[2 * i], lo[2 * i]); split_tf32(b, hi[2 * i + 1], lo[2 * i + 1]);
          }
          mask1t = mask_finish(mask1t, 8);
          tmem_st8(tmem + tlane + kA1Hi + 8u * wc, hi);
          tmem_st8(tmem + tlane + kA1Lo + 8u * wc, lo);
          if (wq == 0 && lane < TL) {      // h1 of the tail s rows: reduction index r = 64 + lane of the P3 operand, and its relu' bits
            M1T[4 * lane + wc] = (uint8_t)mask1t;
            uint8_t* b3 = reinterpret_cast<uint8_t*>(sm + L.oB3) + wc * kSbo4 + ((BT + lane) >> 2) * kLbo4 + (lane & 3) * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              *reinterpret_cast<uint32_t*>(b3 + i * 16) = hi[i];
              *reinterpret_cast<uint32_t*>(b3 + i * 16 + kBbHalf) = lo[i];
            }
          }
        }
        tmem_wait_st();
        fence_async_smem();     // this thread's operand stores (B3 here, B2 / B4 in the unpack phase) -> async proxy
        tc_fence_before();
      }
      __syncthreads();
      if (t == 0) {   // ---- forward products: h2 pre-activations of all forward rows ----
        tc_fence_after();
        umma_3x<kH1 / 8>(tmem + kD2, tmem + kA2Hi, tmem + kA2Lo, sB2, sB2 + kBfHalf, kLbo2, kSbo2, make_idesc(kH2));
        umma_commit(mbar);        // P2 alone: its epilogue pass runs while the tensor core is still on P1
        // P1: B = W2^- (64 columns), or with tail rows [W2^- | W2] (128 columns: the tail's theta rows take the right half)
        if (tail) umma_3x<kH1 / 8>(tmem + kD1, tmem + kA1Hi, tmem + kA1Lo, sBt2, sBt2 + kBfHalf, kLbo2, kSbo2, make_idesc(2 * kH2));
        else umma_3x<kH1 / 8>(tmem + kD1, tmem + kA1Hi, tmem + kA1Lo, sBt2, sBt2 + kBfHalf, kLbo2, kSbo2, make_idesc(kH2));
        umma_commit(mbar1);
      }
      // the staging buffer is free (unpacked two barriers ago): gather the next tile / step while the tensor core works
      if (tile + 1 < ntiles) prefetch(kstep, tile + 1);
      else if (kstep + 1 < args.K) prefetch(kstep + 1, 0);
      mbar_wait(mbar, mma_parity); mma_parity ^= 1u;
      tc_fence_after();
      {  // ---- epilogue: + b2, relu; the head's partial sums over this thread's 16 units straight from the registers ----
        constexpr int NP = (A + 2) / 2;
        // one pass: 16 accumulator columns of this lane -> relu(. + bias) -> head partial sums into HP[set][wc][c][row];
        // returns the relu' mask; h2dst (or nullptr): where this row's h2 goes ([unit][row] layout, stride RS1)
        auto pass = [&](const uint32_t (&r0)[16], const float* bias, const float* wh, float* hpdst, float* h2dst, bool write) {
          uint32_t mask = 0;
          u64 hp[NP];
#pragma unroll
          for (int q = 0; q < NP; ++q) hp[q] = 0ull;
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const float4 bv = ld4(bias + 4 * i4);
            const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = 4 * i4 + e;
              const float v = fmaxf(__uint_as_float(r0[i]) + bb[e], 0.f);
              mask_push(mask, v);
              if (h2dst) h2dst[i * RS1] = v;
              const u64 vv = pack2(v, v);
              const u64x2 w0 = ld2x64(wh + i * HC);
              ffma2(hp[0], vv, w0.lo);
              if constexpr (NP > 1) ffma2(hp[1], vv, w0.hi);
              if constexpr (NP > 2) { const u64x2 w1 = ld2x64(wh + i * HC + 4); ffma2(hp[2], vv, w1.lo); if constexpr (NP > 3) ffma2(hp[3], vv, w1.hi); }
            }
          }
          if (write) {
#pragma unroll
            for (int q = 0; q < NP; ++q) {
              float a, b;
              unpack2(hp[q], a, b);
              hpdst[(2 * q) * NR] = a;
              if (2 * q + 1 <= A) hpdst[(2 * q + 1) * NR] = b;
            }
          }
          return mask_finish(mask, 16);
        };
        const float* bias = W + L.pW2 + kH1 * WS2 + 16 * wc;
        const float* wh = W + L.pWh + 16 * wc * HC;
        const float* biast = Wt + L.pW2 + kH1 * WS2 + 16 * wc;
        const float* wht = Wt + L.pWh + 16 * wc * HC;
        // main rows: set 0 = (theta, s), 1 = (theta, s'), 2 = (theta^-, s')
        // second accumulator of this warp: quadrants 2, 3 = (theta^-, s') of the same batch row; with tail rows quadrant 0 = their
        // theta rows (right half of D1: lanes 0..15 tail s, 16..31 tail s'), quadrant 1 lanes 0..15 = their theta^- rows (left half)
        const bool second = wq >= 2 || tail;
        uint32_t ra[16], rb[16];
        tmem_ld16(tmem + tlane + kD2 + 16u * wc, ra);
        tmem_wait_ld();
        mask2 = pass(ra, bias, wh, HP + ((wq < 2 ? 0 : 1) * 4 + wc) * HC * NR + (row2 & 63), wq < 2 ? H2 + (16 * wc) * RS1 + row2 : nullptr, true);
        if (second) {
          mbar_wait(mbar1, p1_parity);
          tc_fence_after();
          tmem_ld16(tmem + tlane + kD1 + (wq == 0 ? 64u : 0u) + 16u * wc, rb);
          tmem_wait_ld();
        }
        p1_parity ^= 1u;
        if (wq >= 2) pass(rb, biast, wht, HP + (2 * 4 + wc) * HC * NR + (row2 & 63), nullptr, true);
        else if (tail && wq == 0)
          pass(rb, bias, wh, HP + ((lane < TL ? 0 : 1) * 4 + wc) * HC * NR + BT + (lane & 15),
                        lane < TL ? H2 + (16 * wc) * RS1 + BT + lane : nullptr, true);
        else if (tail)
          pass(rb, biast, wht, HP + (2 * 4 + wc) * HC * NR + BT + (lane & 15), nullptr, lane < TL);
        tc_fence_before();
      }
      __syncthreads();
      if (warp < 8) {   // ---- dueling head: thread (row i, forward set g) sums the four column quarters; dddqn.py:31 ----
        int i = t & 63, g = t >> 6;
        if (g == 3) { g = (t - 192) >> 4; i = BT + (t & 15); if (!tail || t >= 240) g = 3; }   // tail rows: threads 192..239
        if (g < 3) {
          float hd[1 + A];
#pragma unroll
          for (int c = 0; c <= A; ++c) {
            float v = (g < 2 ? W : Wt)[L.pWh + kH2 * HC + c];
#pragma unroll
            for (int p = 0; p < 4; ++p) v += HP[((g * 4 + p) * HC + c) * NR + i];
            hd[c] = v;
          }
          float ms = 0.f;
#pragma unroll
          for (int j = 1; j <= A; ++j) ms += hd[j];
          ms = ms / (float)A;
#pragma unroll
          for (int j = 0; j < A; ++j) Qs[(g * HC + j) * NR + i] = hd[0] + hd[1 + j] - ms;
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
      }
      if (t < (tail ? NR : BT)) {   // ---- targets / loss / d(head) of batch row i (tile-local; tail rows are 64..79) ----
        const int i = t;
        float q[A], nq[A], qb[A];
#pragma unroll
        for (int j = 0; j < A; ++j) { q[j] = Qs[(0 * HC + j) * NR + i]; nq[j] = Qs[(1 * HC + j) * NR + i]; qb[j] = Qs[(2 * HC + j) * NR + i]; }
        // ---- compute_q_targets (q_learning_functions.py:55-59) ----
        int astar = 0; float best = nq[0];
#pragma unroll
        for (int j = 1; j < A; ++j) if (nq[j] > best) { best = nq[j]; astar = j; }   // first max wins
        const float4 meta = ld4(Meta + 4 * i);
        int a = __float_as_int(meta.x);
        a = a < 0 ? 0 : (a >= A ? A - 1 : a);                 // jax clamps out-of-range gather indices
        const float rew = meta.z;
        const float done = __float_as_uint(meta.w) ? 1.f : 0.f;   // dones.astype(float32), :84
        float qa = q[0], nqt = qb[0];
#pragma unroll
        for (int j = 1; j < A; ++j) { if (j == a) qa = q[j]; if (j == astar) nqt = qb[j]; }
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
        DhdT[0 * RS1 + i] = gi;
        const float gia = gi / (float)A;
#pragma unroll
        for (int j = 0; j < A; ++j) DhdT[(1 + j) * RS1 + i] = (j == a ? gi : 0.f) - gia;
        Red[kRedLoss + i] = l;   // summed (like the head-bias gradient = column sums of d(head)) by idle warps under the backward product
        if (args.taps.enabled && valid) {
          const int gi_row = tile * BT + i;
#pragma unroll
          for (int j = 0; j < A; ++j) {
            if (args.taps.q) args.taps.q[gi_row * A + j] = q[j];
            if (args.taps.next_q) args.taps.next_q[gi_row * A + j] = nq[j];
            if (args.taps.next_q_tm) args.taps.next_q_tm[gi_row * A + j] = qb[j];
            if (args.taps.targets) args.taps.targets[gi_row * A + j] = (j == a) ? tgt : q[j];
          }
          if (args.taps.max_actions) args.taps.max_actions[gi_row] = astar;
        }
      }
      __syncthreads();
      float db2_part = 0.f;
      {  // ---- dh2 = relu'(h2) * (d(head) . Wh^T), per row r (lanes 0..63) AND per unit j (lanes 64..127): the two halves of
         //      the backward A tile; the reduction index runs over this warp's 16 columns ----
        uint32_t hi[16], lo[16];
        u64 v2[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v2[q] = 0ull;
        if (wq < 2) {          // lane = batch row r; columns = units j of [16 wc, 16 wc + 16): v[j] = sum_c dhd[r][c] * Wh[j][c]
#pragma unroll
          for (int c = 0; c <= A; ++c) {
            const float dv = DhdT[c * RS1 + row2];
            const u64 dd = pack2(dv, dv);
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const u64x2 w = ld2x64(WhT + c * kH2 + 16 * wc + 4 * q4);
              ffma2(v2[2 * q4], dd, w.lo); ffma2(v2[2 * q4 + 1], dd, w.hi);
            }
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float a, b;
            unpack2(v2[q], a, b);
            a = (mask2 >> (2 * q)) & 1u ? a : 0.f;
            b = (mask2 >> (2 * q + 1)) & 1u ? b : 0.f;
            split_tf32(a, hi[2 * q], lo[2 * q]); split_tf32(b, hi[2 * q + 1], lo[2 * q + 1]);
          }
        } else {               // lane = unit j; columns = batch rows r of [16 wc, 16 wc + 16): the same products in the same order
          const int j = row2 - BT;
#pragma unroll
          for (int c = 0; c <= A; ++c) {
            const float wv = WhT[c * kH2 + j];
            const u64 ww = pack2(wv, wv);
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const u64x2 d = ld2x64(DhdT + c * RS1 + 16 * wc + 4 * q4);
              ffma2(v2[2 * q4], d.lo, ww); ffma2(v2[2 * q4 + 1], d.hi, ww);
            }
          }
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const float4 h = ld4(H2 + j * RS1 + 16 * wc + 4 * q4);
            float a, b, c2, d2;
            unpack2(v2[2 * q4], a, b); unpack2(v2[2 * q4 + 1], c2, d2);
            a = h.x > 0.f ? a : 0.f; b = h.y > 0.f ? b : 0.f; c2 = h.z > 0.f ? c2 : 0.f; d2 = h.w > 0.f ? d2 : 0.f;
            db2_part += (a + b) + (c2 + d2);
            split_tf32(a, hi[4 * q4], lo[4 * q4]); split_tf32(b, hi[4 * q4 + 1], lo[4 * q4 + 1]);
            split_tf32(c2, hi[4 * q4 + 2], lo[4 * q4 + 2]); split_tf32(d2, hi[4 * q4 + 3], lo[4 * q4 + 3]);
          }
        }
        tmem_st16(tmem + tlane + kAbHi + 16u * wc, hi);
        tmem_st16(tmem + tlane + kAbLo + 16u * wc, lo);
        if (tail) {
          if (wq >= 2) {       // the tail rows' columns of dh2^T: reduction indices 64 + 4 wc .. + 3
            const int j = row2 - BT;
            u64 t2[2] = {0ull, 0ull};
#pragma unroll
            for (int c = 0; c <= A; ++c) {
              const float wv = WhT[c * kH2 + j];
              const u64 ww = pack2(wv, wv);
              const u64x2 d = ld2x64(DhdT + c * RS1 + BT + 4 * wc);
              ffma2(t2[0], d.lo, ww); ffma2(t2[1], d.hi, ww);
            }
            const float4 h = ld4(H2 + j * RS1 + BT + 4 * wc);
            float a, b, c2, d2;
            unpack2(t2[0], a, b); unpack2(t2[1], c2, d2);
            a = h.x > 0.f ? a : 0.f; b = h.y > 0.f ? b : 0.f; c2 = h.z > 0.f ? c2 : 0.f; d2 = h.w > 0.f ? d2 : 0.f;
            db2_part += (a + b) + (c2 + d2);
            uint32_t h4[4], l4[4];
            split_tf32(a, h4[0], l4[0]); split_tf32(b, h4[1], l4[1]); split_tf32(c2, h4[2], l4[2]); split_tf32(d2, h4[3], l4[3]);
            tmem_st4(tmem + tlane + kAbHi + 64u + 4u * wc, h4);
            tmem_st4(tmem + tlane + kAbLo + 64u + 4u * wc, l4);
          }
          {   // dh2 of the tail rows in fp32 -> shared memory (their dh1 runs on the CUDA cores): thread = (row r, units j, j + 32)
            const int r = t & 15, j = t >> 4;
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int c = 0; c <= A; ++c) {
              const float dv = DhdT[c * RS1 + BT + r];
              a0 = fmaf(dv, WhT[c * kH2 + j], a0);
              a1 = fmaf(dv, WhT[c * kH2 + j + 32], a1);
            }
            Dh2Tl[r * (kH2 + 4) + j] = H2[j * RS1 + BT + r] > 0.f ? a0 : 0.f;
            Dh2Tl[r * (kH2 + 4) + j + 32] = H2[(j + 32) * RS1 + BT + r] > 0.f ? a1 : 0.f;
          }
        }
        if (wq >= 2 && wc > 0) Red[kRedDb2 + wc * kH2 + (row2 - BT)] = db2_part;
        tmem_wait_st();
        tc_fence_before();
      }
      __syncthreads();
      if (t == 0) {   // ---- backward product: dh1 (lanes 0..63 x cols 0..31) and dW2^T (lanes 64..127 x cols 32..63) from the one A tile ----
        tc_fence_after();
        if (tail) umma_3x<NR / 8>(tmem + kD4, tmem + kAbHi, tmem + kAbLo, sB4, sB4 + kBbHalf, kLbo4, kSbo4, make_idesc(2 * kH1));
        else umma_3x<BT / 8>(tmem + kD4, tmem + kAbHi, tmem + kAbLo, sB4, sB4 + kBbHalf, kLbo4, kSbo4, make_idesc(2 * kH1));
        umma_commit(mbar);
      }
      if (warp == 0) {       // sum of the tile's per-sample losses (fixed order)
        float v = Red[kRedLoss + lane] + Red[kRedLoss + 32 + lane];
        if (tail && lane < TL) v += Red[kRedLoss + BT + lane];
        loss_acc += warp_sum(v);
      } else if (warp == 1) {   // d(head bias)[c] = sum_r dhd[r][c]
        const int c = lane & 7, part = lane >> 3;
        const float4 d0 = ld4(DhdT + c * RS1 + 16 * part), d1 = ld4(DhdT + c * RS1 + 16 * part + 4), d2 = ld4(DhdT + c * RS1 + 16 * part + 8),
                     d3 = ld4(DhdT + c * RS1 + 16 * part + 12);
        float v = (((d0.x + d0.y) + (d0.z + d0.w)) + ((d1.x + d1.y) + (d1.z + d1.w))) + (((d2.x + d2.y) + (d2.z + d2.w)) + ((d3.x + d3.y) + (d3.z + d3.w)));
        if (tail) { const float4 e0 = ld4(DhdT + c * RS1 + BT + 4 * part); v += (e0.x + e0.y) + (e0.z + e0.w); }
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (part == 0 && c <= A) G[L.pWh + kH2 * HC + c] += v;
      }
      if (warp >= NW / 2) {  // dWh[j][c] += sum_r h2[r][j] * dhd[r][c] on the CUDA cores (warps 8..15) while the tensor core works
        const int j = (warp - NW / 2) * 8 + (lane & 7), part = lane >> 3;
        const float* ap[1] = {H2 + j * RS1 + 16 * part};
        const float* bp[1 + A];
#pragma unroll
        for (int c = 0; c <= A; ++c) bp[c] = DhdT + c * RS1 + 16 * part;
        float acc[1][1 + A];
        dot_tile<1, 1 + A, 4>(ap, bp, acc);
        if (tail) {          // + the tail rows, four per lane group
          const float4 h = ld4(H2 + j * RS1 + BT + 4 * part);
#pragma unroll
          for (int c = 0; c <= A; ++c) {
            const float4 d = ld4(DhdT + c * RS1 + BT + 4 * part);
            acc[0][c] += (h.x * d.x + h.y * d.y) + (h.z * d.z + h.w * d.w);
          }
        }
#pragma unroll
        for (int c = 0; c <= A; ++c) {
          float v = acc[0][c];
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (part == 0) G[L.pWh + j * HC + c] += v;
        }
      }
      if (tail) {
        // dh1 of the tail rows: relu'(h1) * (dh2 . W2^T) in fp32, one (row, unit) per thread (16 x 32 x 64: too small for a
        // product of its own); under the backward product
        const int r = t & 15, k = t >> 4;
        const float* dr = Dh2Tl + r * (kH2 + 4);
        const float* wr = W + L.pW2 + k * WS2;
        float o0 = 0.f, o1 = 0.f;
#pragma unroll
        for (int j4 = 0; j4 < kH2 / 4; j4 += 2) {
          const float4 d0 = ld4(dr + 4 * j4), w0 = ld4(wr + 4 * j4), d1 = ld4(dr + 4 * j4 + 4), w1 = ld4(wr + 4 * j4 + 4);
          o0 += (d0.x * w0.x + d0.y * w0.y) + (d0.z * w0.z + d0.w * w0.w);
          o1 += (d1.x * w1.x + d1.y * w1.y) + (d1.z * w1.z + d1.w * w1.w);
        }
        Dh1T[k * RS1 + BT + r] = (M1T[4 * r + (k >> 3)] >> (k & 7)) & 1u ? o0 + o1 : 0.f;
      }
      mbar_wait(mbar, mma_parity); mma_parity ^= 1u;
      tc_fence_after();
      {
        uint32_t r0[8];
        if (wq < 2) {          // dh1[r][k] = relu'(h1) * (dh2 . W2^T)  -> Dh1T (k-major, for dW1)
          tmem_ld8(tmem + tlane + kD4 + 8u * wc, r0);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 8; ++i) Dh1T[(8 * wc + i) * RS1 + row2] = (mask1 >> i) & 1u ? __uint_as_float(r0[i]) : 0.f;
        } else {               // dW2[k][j] += (dh2^T . h1)[j][k];  db2[j] += sum_r dh2[r][j]
          const int j = row2 - BT;
          tmem_ld8(tmem + tlane + kD3 + 8u * wc, r0);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 8; ++i) G[L.pW2 + (8 * wc + i) * WS2 + j] += __uint_as_float(r0[i]);
          if (wc == 0) G[L.pW2 + kH1 * WS2 + j] += ((db2_part + Red[kRedDb2 + kH2 + j]) + Red[kRedDb2 + 2 * kH2 + j]) + Red[kRedDb2 + 3 * kH2 + j];
        }
        tc_fence_before();
      }
      __syncthreads();
      {  // [dW1;db1][d][h] = sum_r x[r][d] * dh1[r][h]  (row D of X is ones -> db1): lane = unit h, one warp per (row group, four
         // input rows d); the groups' partial sums are folded into G by the next tile's unpack phase or by Adam.  Row groups:
         // 0..3 = rows 16 g .. 16 g + 15, 4 = the tail rows (X columns 128..143)
        const int ndg = (D + 4) / 4;                      // groups of four input rows (incl. the ones row)
        for (int task = warp; task < nslice * ndg; task += NW) {
          const int rg = task % nslice, d0 = 4 * (task / nslice);
          const int xoff = rg < 4 ? 16 * rg : 2 * BT, hoff = rg < 4 ? 16 * rg : BT;
          const float* bp[1] = {Dh1T + lane * RS1 + hoff};
          const float* ap[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) ap[i] = X + (d0 + i <= D ? d0 + i : D) * XS + xoff;
          float acc[4][1];
          dot_tile<4, 1, 4>(ap, bp, acc);
#pragma unroll
          for (int i = 0; i < 4; ++i) if (d0 + i <= D) DW1P[rg * 17 * kH1 + (d0 + i) * kH1 + lane] = acc[i][0];
        }
      }
      // the barrier at the top of the next tile / before Adam orders these G updates
    }  // tiles

    if (warp == 0 && lane == 0) Red[0] = loss_acc;
    __syncthreads();

    if (args.taps.enabled && args.taps.grads) {
      for (int p = t; p < L.PS; p += NT) args.taps.grads[p] = G[p] + (p < (D + 1) * kH1 ? dw1_fold(p) : 0.f);
    }
    // ================= optimiser: optax adam / adamw (q_learning_functions.py:24-25) ==========
    {
      const float c1 = Red[8 + 2 * (kstep & 1)], c2 = Red[9 + 2 * (kstep & 1)];
      const float rc1 = 1.0f / c1, rc2 = 1.0f / c2;
      const float omb1 = 1.0f - b1, omb2 = 1.0f - b2;
      // two adjacent packed entries per pass (8-byte loads / stores, one index and one guard for both), branch-free (indices
      // clamped, stores guarded) so that the sqrt / reciprocal chains of a thread overlap
      float2 gv[NPP], th[NPP];
      const int nW1 = (D + 1) * kH1;                 // even, <= 544: only pair 0 can hold layer-1 parameters
#pragma unroll
      for (int i = 0; i < NPP; ++i) {
        if (i == NPP - 1 && !pair4) break;
        const int p = 2 * (t + i * NT), pc = p < L.PS ? p : L.PS - 2;
        gv[i] = *reinterpret_cast<const float2*>(G + pc); th[i] = *reinterpret_cast<const float2*>(W + pc);
        if (i == 0 && p < nW1) {
          gv[0].x += dw1_fold(p); gv[0].y += dw1_fold(p + 1);
        }
      }
#pragma unroll
      for (int i = 0; i < NPP; ++i) {
        if (i == NPP - 1 && !pair4) break;
        const int p = 2 * (t + i * NT);
        const bool live = p < L.PS;
        float gg[2] = {live ? gv[i].x : 0.f, live ? gv[i].y : 0.f};
        const float tt[2] = {th[i].x, th[i].y};
        float mm[2] = {mreg[i].x, mreg[i].y}, vv[2] = {vreg[i].x, vreg[i].y}, ww[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float g = gg[e];
          const float m = b1 * mm[e] + omb1 * g;
          const float v = b2 * vv[e] + omb2 * (g * g);
          mm[e] = m; vv[e] = v;
          const float u = (m * rc1) * fast_rcp(fast_sqrt(v * rc2 + eps_root) + eps);
          ww[e] = tt[e] - lr * (u + wd * tt[e]);  // add_decayed_weights (wd = 0 for adam); scale(-lr); apply_updates
        }
        mreg[i] = make_float2(mm[0], mm[1]); vreg[i] = make_float2(vv[0], vv[1]);
        if (live) {
          *reinterpret_cast<float2*>(G + p) = make_float2(0.f, 0.f);
          *reinterpret_cast<float2*>(W + p) = make_float2(ww[0], ww[1]);
        }
      }
    }
    if (t == 0) {
      const float loss = Red[0] / fB;
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
  for (int i = 0; i < NPP; ++i) {
    const int p = 2 * (t + i * NT);
    if (p < L.PS) {
      *reinterpret_cast<float2*>(gW + p) = *reinterpret_cast<const float2*>(W + p);
      *reinterpret_cast<float2*>(gM + p) = mreg[i]; *reinterpret_cast<float2*>(gV + p) = vreg[i];
    }
  }
  if (t == 0) {
    ctl->train_steps = step0 + args.K;
    const long long c = (long long)count0 + args.K;
    ctl->adam_count = c > 0x7fffffffLL ? 0x7fffffff : (int)c;
    ctl->pb1 = pb1; ctl->pb2 = pb2;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

}  // namespace tct
}  // namespace dqn
