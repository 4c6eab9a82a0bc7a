// train_fused_tc_main.cuh -- body of the tensor-core population kernel (train_fused_tc.cu) for batches of <= 64 rows (one tile)
// and of > 80 rows (tiles of 64 rows).  One CTA of 512 threads = one agent, K steps per launch; the same arithmetic as
// train_fused.cu (General/QLearning/q_agent.py:146-169, q_learning_functions.py:14-64, dddqn.py:24-34), but the four
// products that hold 80 % of the step's flops run as error-compensated 3xTF32 on tcgen05.mma (kind::tf32, M = 128;
// a = a_hi + a_lo:  a_hi*b_lo + a_lo*b_hi + a_hi*b_hi, fp32 accumulation in tensor memory), in two round trips per tile:
//   forward   P2  h2(theta; s | s') [128 x 64] = h1 [128 x 32] . W2       P1  h2(theta^-; s')   (rows in lanes 64..127)
//             (P2 and P1 commit on mbarriers of their own: the P2 epilogue pass runs while the tensor core is on P1)
//   backward  ONE product with B = [W2 | h1]:  dh1 [64 r x 32 k] = dh2 . W2^T  |  dW2^T [64 j x 32 k] = dh2^T . h1
// "thread = TMEM lane = row": a warp reaches the 32 lanes of its quadrant (warp % 4), the four warps of a quadrant
// split the columns.  A operands never touch shared memory: layer 1 is computed one batch row per thread and its
// hi / lo halves go straight into tensor memory (tcgen05.st); dh2 is computed twice -- per row r (lanes 0..63) and
// per unit j (lanes 64..127) -- so the backward product reads ONE A tile whose halves are dh2 and dh2^T.  B operands
// (W2 in both orientations, h1 of the s rows) sit in shared memory in the K-major no-swizzle canonical layout (8 x 16 B
// core matrices).  The head is evaluated from the epilogue's registers (h2 never makes a round trip for it); targets,
// dh2, dWh, dW1 and Adam are fp32 on the CUDA cores.  16 warps (128 registers each) instead of the FFMA kernel's 8:
// every phase is a short dependent chain, so the step is latency-bound and the extra warps are what hides it.
#pragma once
#include <math.h>

#include "common.cuh"
#include "kernels.h"
#include "tile_ops.cuh"
#include "tc_prims.cuh"

namespace dqn {
namespace tcm {

using namespace tile;
using namespace tcp;


constexpr int NT = 512;          // threads per CTA
constexpr int NW = NT / 32;
constexpr int BT = 64;           // batch rows per tile
constexpr int RS2 = 2 * BT + 4;  // 132: row stride of X [d][s rows | s' rows]
constexpr int RS1 = BT + 4;      // 68
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
constexpr int kLbo4 = 144, kSbo4 = 16 * kLbo4, kB4Bytes = 4 * kSbo4;    // 2304, 9216
// Tensor-memory columns (512 allocated; lane = row)
constexpr uint32_t kA2Hi = 0, kA2Lo = 32, kA1Hi = 64, kA1Lo = 96, kD2 = 128, kD1 = 192, kAbHi = 256, kAbLo = 320, kD4 = 384, kD3 = 416;
// B4 and B3 are adjacent (hi: B4 | B3, then lo: B4 | B3), so ONE operand of mn = 64 (SBO groups 0..3 = W2, 4..7 = h1) feeds a single
// N = 64 product whose accumulator columns are [D4 | D3]: lanes 0..63 of D4 and lanes 64..127 of D3 are the useful halves
constexpr int kBbHalf = 2 * kB4Bytes;   // bytes from a hi buffer to its lo twin

struct Lay {   // offsets in floats
  int pW2, pWh, PS;
  int oW, oWt, oG, oX, oH2, oDh1T, oDhdT, oHP, oWhT, oMeta, oDummy, oRed, oBar, oStage, oB2, oBt2, oB4, oB3, total;
};

__host__ __device__ inline Lay make_layout(int D, int recw) {
  Lay L;
  L.pW2 = packed_w2(D);
  L.pWh = packed_head(D);
  L.PS = packed_count(D);          // the shared-memory weight layout IS the packed HBM layout (common.cuh)
  int o = 0;
  L.oW = o; o += L.PS;
  L.oWt = o; o += L.PS;
  L.oG = o; o += L.PS;
  L.oX = o; o += (D + 1) * RS2;       // [D+1][132]  cols 0..63 = s rows, 64..127 = s' rows; row D = ones
  L.oH2 = o; o += kH2 * RS1;          // [64 j][68]   h2(theta, s): dWh, relu' of the per-unit dh2 pass
  L.oDh1T = o; o += kH1 * RS1;        // [32 k][68]
  L.oDhdT = o; o += HC * RS1;         // [8 c][68]
  L.oHP = o; o += 3 * 4 * HC * BT;    // head partials [forward set][column quarter][c][row]
  L.oWhT = o; o += HC * kH2;          // [8 c][64 j]  head weights of theta, transposed (rebuilt every step)
  L.oMeta = o; o += BT * 4;
  L.oDummy = o; o += 4;
  L.oRed = o; o += 32 + 4 * kH2;      // [0..1] loss, [8..11] Adam bias corrections, [16..31] head-bias partials, [32..] db2 partials [quarter][j]
  L.oBar = o; o += 8;                 // [0,1] mbarrier of the record gather, [2,3] mbarrier of the MMAs, [4] TMEM base address, [6,7] mbarrier of P1
  L.oStage = o; o += BT * recw;
  L.oB2 = o; o += 2 * kB2Bytes / 4;   // hi | lo
  L.oBt2 = o; o += 2 * kB2Bytes / 4;
  L.oB4 = o; o += kB4Bytes / 4;       // hi: B4 | B3, lo: B4 | B3
  L.oB3 = o; o += kB4Bytes / 4 + 2 * kB4Bytes / 4;
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
  float* const Qs = Dh1T;                                   // [3 forward sets][8][64 rows] Q-values (Dh1T is dead until the backward epilogue)
  float* const DW1P = sm + L.oB3;                           // [4 row quarters][(D+1) * 32] partial dW1 (the h1 operand is dead after the backward product)
  float* const Meta = sm + L.oMeta; // [64][4]  raw action lo, action hi, reward, done
  float* const Red = sm + L.oRed;
  float* const Stage = sm + L.oStage;
  const int wq = warp & 3, wc = warp >> 2;          // TMEM quadrant (lanes 32 wq ..), column quarter
  const int row2 = 32 * wq + lane;                  // forward row of batch A: < 64 = s row, else s' row (row2 - 64); backward: r | 64 + j
  const uint32_t tlane = (uint32_t)(32 * wq) << 16;
  const uint32_t sB2 = smem_addr(sm + L.oB2), sBt2 = smem_addr(sm + L.oBt2), sB4 = smem_addr(sm + L.oB4);
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
  for (int r = t; r < RS2; r += NT) X[D * RS2 + r] = 1.f;

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
  // (batches of 65..80 rows never come here: train_fused_tc.cu sends them to the tail body)

  // record word -> smem destination of this thread's staged chunks: thread t owns 16-byte chunks (t & 7), (t & 7) + 8 of row t >> 3
  const int urow = t >> 3, u8 = t & 7;
  int udst[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int wdx = 4 * (u8 + 8 * (q >> 2)) + (q & 3);
    int o = L.oDummy;
    if (wdx < D) o = L.oX + wdx * RS2 + urow;
    else if (wdx < 2 * D) o = L.oX + (wdx - D) * RS2 + BT + urow;
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

  // gather of (step kstep, tile): one thread per row issues ONE bulk copy of the whole record; thread 0 posts the tile's byte count
  auto prefetch = [&](int kstep, int tile) {
    const int nvalid = B - tile * BT < BT ? B - tile * BT : BT;
    if (t == 0) mbar_arrive_expect_tx(bar, (uint32_t)(nvalid * recw * 4));
    if (u8 == 0) {      // four lanes of every warp: all warps reach the MMA wait together (two whole warps doing this while 14 spin is slower)
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
  // W2 as a (mn = j, K = k) operand, hi | lo: 16-byte chunk (j, kc) at (j / 8) * SBO + kc * LBO + (j % 8) * 16
  auto lay_b2 = [&](const float* w2, float* dstf) {
    for (int ci = t; ci < 512; ci += NT) {
      const int j = ci & 63, kc = ci >> 6;
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) split_tf32(w2[(4 * kc + e) * WS2 + j], hi[e], lo[e]);
      uint8_t* dst = reinterpret_cast<uint8_t*>(dstf) + (j >> 3) * kSbo2 + kc * kLbo2 + (j & 7) * 16;
      *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(dst + kB2Bytes) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
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
    float loss_acc = 0.f;   // meaningful in warps 0,1

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
        }
      }
      if (tile > 0) {    // fold the previous tile's partial dW1 (Adam folds the last tile's)
        for (int p = t; p < (D + 1) * kH1; p += NT) G[p] += ((DW1P[p] + DW1P[17 * kH1 + p]) + DW1P[2 * 17 * kH1 + p]) + DW1P[3 * 17 * kH1 + p];
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

      uint32_t mask1 = 0, mask2 = 0;   // relu' of this thread's 8 h1 / 16 h2 entries (kept from the forward for the backward)
      {  // ---- layer 1, one batch row per thread, 8 units: theta on (s | s') = 128 rows; theta^- on s' (lanes 64..127) ----
        u64 acc[4], acct[4];
        const float* w1 = W + 8 * wc;
        const float* w1t = Wt + 8 * wc;
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[i] = ld64(w1 + D * kH1 + 2 * i); acct[i] = ld64(w1t + D * kH1 + 2 * i); }
        if (wq < 2) {
#pragma unroll 4
          for (int d = 0; d < D; ++d) {
            const float x = X[d * RS2 + row2];
            const u64 xx = pack2(x, x);
            const u64x2 a0 = ld2x64(w1 + d * kH1), a1 = ld2x64(w1 + d * kH1 + 4);
            ffma2(acc[0], xx, a0.lo); ffma2(acc[1], xx, a0.hi); ffma2(acc[2], xx, a1.lo); ffma2(acc[3], xx, a1.hi);
          }
        } else {
#pragma unroll 4
          for (int d = 0; d < D; ++d) {
            const float x = X[d * RS2 + row2];
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
        tmem_wait_st();
        fence_async_smem();     // this thread's operand stores (B3 here, B2 / B4 in the unpack phase) -> async proxy
        tc_fence_before();
      }
      __syncthreads();
      if (t == 0) {   // ---- forward products: h2 pre-activations of all 192 forward rows ----
        tc_fence_after();
        umma_3x<kH1 / 8>(tmem + kD2, tmem + kA2Hi, tmem + kA2Lo, sB2, sB2 + kB2Bytes, kLbo2, kSbo2, make_idesc(kH2));
        umma_commit(mbar);        // P2 alone: its epilogue pass runs while the tensor core is still on P1
        umma_3x<kH1 / 8>(tmem + kD1, tmem + kA1Hi, tmem + kA1Lo, sBt2, sBt2 + kB2Bytes, kLbo2, kSbo2, make_idesc(kH2));
        umma_commit(mbar1);
      }
      // the staging buffer is free (unpacked two barriers ago): gather the next tile / step while the tensor core works
      if (tile + 1 < ntiles) prefetch(kstep, tile + 1);
      else if (kstep + 1 < args.K) prefetch(kstep + 1, 0);
      mbar_wait(mbar, mma_parity); mma_parity ^= 1u;
      tc_fence_after();
      {  // ---- epilogue: + b2, relu; the head's partial sums over this thread's 16 units straight from the registers ----
        constexpr int NP = (A + 2) / 2;
        uint32_t r0[16];
        tmem_ld16(tmem + tlane + kD2 + 16u * wc, r0);
        tmem_wait_ld();
        const float* bias = W + L.pW2 + kH1 * WS2 + 16 * wc;
        const float* wh = W + L.pWh + 16 * wc * HC;
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
            mask_push(mask2, v);
            if (wq < 2) H2[(16 * wc + i) * RS1 + row2] = v;
            const u64 vv = pack2(v, v);
            const u64x2 w0 = ld2x64(wh + i * HC);
            ffma2(hp[0], vv, w0.lo);
            if constexpr (NP > 1) ffma2(hp[1], vv, w0.hi);
            if constexpr (NP > 2) { const u64x2 w1 = ld2x64(wh + i * HC + 4); ffma2(hp[2], vv, w1.lo); if constexpr (NP > 3) ffma2(hp[3], vv, w1.hi); }
          }
        }
        mask2 = mask_finish(mask2, 16);
        {
          float* dst = HP + ((wq < 2 ? 0 : 1) * 4 + wc) * HC * BT + (row2 & 63);
#pragma unroll
          for (int q = 0; q < NP; ++q) {
            float a, b;
            unpack2(hp[q], a, b);
            dst[(2 * q) * BT] = a;
            if (2 * q + 1 <= A) dst[(2 * q + 1) * BT] = b;
          }
        }
        if (wq >= 2) {   // h2(theta^-, s') of the same batch row (the barrier below keeps the other warps behind P1 as well)
          mbar_wait(mbar1, p1_parity);
          tc_fence_after();
          uint32_t r1[16];
          tmem_ld16(tmem + tlane + kD1 + 16u * wc, r1);
          tmem_wait_ld();
          const float* biast = Wt + L.pW2 + kH1 * WS2 + 16 * wc;
          const float* wht = Wt + L.pWh + 16 * wc * HC;
#pragma unroll
          for (int q = 0; q < NP; ++q) hp[q] = 0ull;
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const float4 bv = ld4(biast + 4 * i4);
            const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = 4 * i4 + e;
              const float v = fmaxf(__uint_as_float(r1[i]) + bb[e], 0.f);
              const u64 vv = pack2(v, v);
              const u64x2 w0 = ld2x64(wht + i * HC);
              ffma2(hp[0], vv, w0.lo);
              if constexpr (NP > 1) ffma2(hp[1], vv, w0.hi);
              if constexpr (NP > 2) { const u64x2 w1 = ld2x64(wht + i * HC + 4); ffma2(hp[2], vv, w1.lo); if constexpr (NP > 3) ffma2(hp[3], vv, w1.hi); }
            }
          }
          float* dst = HP + (2 * 4 + wc) * HC * BT + (row2 & 63);
#pragma unroll
          for (int q = 0; q < NP; ++q) {
            float a, b;
            unpack2(hp[q], a, b);
            dst[(2 * q) * BT] = a;
            if (2 * q + 1 <= A) dst[(2 * q + 1) * BT] = b;
          }
        }
        p1_parity ^= 1u;
        tc_fence_before();
      }
      __syncthreads();
      if (warp < 8) {   // ---- dueling head: thread (row i, forward set g) sums the four column quarters; dddqn.py:31 ----
        const int i = t & 63, g = t >> 6;
        if (g < 3) {
          float hd[1 + A];
#pragma unroll
          for (int c = 0; c <= A; ++c) {
            float v = (g < 2 ? W : Wt)[L.pWh + kH2 * HC + c];
#pragma unroll
            for (int p = 0; p < 4; ++p) v += HP[((g * 4 + p) * HC + c) * BT + i];
            hd[c] = v;
          }
          float ms = 0.f;
#pragma unroll
          for (int j = 1; j <= A; ++j) ms += hd[j];
          ms = ms / (float)A;
#pragma unroll
          for (int j = 0; j < A; ++j) Qs[(g * HC + j) * BT + i] = hd[0] + hd[1 + j] - ms;
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
      }
      if (t < BT) {   // ---- targets / loss / d(head) of batch row i ----
        const int i = t;
        float q[A], nq[A], qb[A];
#pragma unroll
        for (int j = 0; j < A; ++j) { q[j] = Qs[(0 * HC + j) * BT + i]; nq[j] = Qs[(1 * HC + j) * BT + i]; qb[j] = Qs[(2 * HC + j) * BT + i]; }
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
        Red[32 + i] = l;   // summed (like the head-bias gradient = column sums of d(head)) by idle warps under the backward product
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
          if (wc > 0) Red[32 + wc * kH2 + j] = db2_part;
        }
        tmem_st16(tmem + tlane + kAbHi + 16u * wc, hi);
        tmem_st16(tmem + tlane + kAbLo + 16u * wc, lo);
        tmem_wait_st();
        tc_fence_before();
      }
      __syncthreads();
      if (t == 0) {   // ---- backward products: dh1 (lanes 0..63) and dW2^T (lanes 64..127) from the one A tile ----
        tc_fence_after();
        umma_3x<BT / 8>(tmem + kD4, tmem + kAbHi, tmem + kAbLo, sB4, sB4 + kBbHalf, kLbo4, kSbo4, make_idesc(2 * kH1));
        umma_commit(mbar);
      }
      if (warp == 0) {       // sum of the tile's per-sample losses (fixed order)
        loss_acc += warp_sum(Red[32 + lane] + Red[32 + 32 + lane]);
      } else if (warp == 1) {   // d(head bias)[c] = sum_r dhd[r][c]
        const int c = lane & 7, part = lane >> 3;
        const float4 d0 = ld4(DhdT + c * RS1 + 16 * part), d1 = ld4(DhdT + c * RS1 + 16 * part + 4), d2 = ld4(DhdT + c * RS1 + 16 * part + 8),
                     d3 = ld4(DhdT + c * RS1 + 16 * part + 12);
        float v = (((d0.x + d0.y) + (d0.z + d0.w)) + ((d1.x + d1.y) + (d1.z + d1.w))) + (((d2.x + d2.y) + (d2.z + d2.w)) + ((d3.x + d3.y) + (d3.z + d3.w)));
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
#pragma unroll
        for (int c = 0; c <= A; ++c) {
          float v = acc[0][c];
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (part == 0) G[L.pWh + j * HC + c] += v;
        }
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
          if (wc == 0) G[L.pW2 + kH1 * WS2 + j] += ((db2_part + Red[32 + kH2 + j]) + Red[32 + 2 * kH2 + j]) + Red[32 + 3 * kH2 + j];
        }
        tc_fence_before();
      }
      __syncthreads();
      {  // [dW1;db1][d][h] = sum_r x[r][d] * dh1[r][h]  (row D of X is ones -> db1): lane = unit h, warp = (row quarter, four
         // input rows d); the four quarters' partial sums are folded into G by the next tile's unpack phase or by Adam
        const int rq = warp & 3;
        const float* bp[1] = {Dh1T + lane * RS1 + 16 * rq};
        for (int d0 = 4 * (warp >> 2); d0 <= D; d0 += 16) {
          const float* ap[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) ap[i] = X + (d0 + i <= D ? d0 + i : D) * RS2 + 16 * rq;
          float acc[4][1];
          dot_tile<4, 1, 4>(ap, bp, acc);
#pragma unroll
          for (int i = 0; i < 4; ++i) if (d0 + i <= D) DW1P[rq * 17 * kH1 + (d0 + i) * kH1 + lane] = acc[i][0];
        }
      }
      // the barrier at the top of the next tile / before Adam orders these G updates
    }  // tiles

    if ((warp == 0 || warp == 1) && lane == 0) Red[warp] = loss_acc;   // (warp 1 contributes 0)
    __syncthreads();

    if (args.taps.enabled && args.taps.grads) {
      for (int p = t; p < L.PS; p += NT)
        args.taps.grads[p] = G[p] + (p < (D + 1) * kH1 ? ((DW1P[p] + DW1P[17 * kH1 + p]) + DW1P[2 * 17 * kH1 + p]) + DW1P[3 * 17 * kH1 + p] : 0.f);
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
          const float2 q0 = *reinterpret_cast<const float2*>(DW1P + p), q1 = *reinterpret_cast<const float2*>(DW1P + 17 * kH1 + p),
                       q2 = *reinterpret_cast<const float2*>(DW1P + 2 * 17 * kH1 + p), q3 = *reinterpret_cast<const float2*>(DW1P + 3 * 17 * kH1 + p);
          gv[0].x += ((q0.x + q1.x) + q2.x) + q3.x; gv[0].y += ((q0.y + q1.y) + q2.y) + q3.y;
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

}  // namespace tcm
}  // namespace dqn
