// train_cluster.cu -- the fused train step for ONE agent spread over a 4-CTA thread-block cluster.
//
// The single-CTA kernel (train_fused.cu) is latency-bound: one agent's step is a serial chain of small phases on one
// SM.  Steps cannot overlap (step t+1 needs theta_t), so the only parallelism left is inside the minibatch: here the
// 64 rows of each tile are split over the 4 CTAs of a cluster (16 rows each, data-parallel inside the cluster):
//
//   every CTA   gathers its 16 transitions (cp.async, one step ahead), runs the three forwards on its rows
//               (48 "forward rows": (theta,s) | (theta,s') | (theta^-,s')), targets / Huber, and the whole backward,
//               producing a FULL-SIZE partial gradient in its own shared memory;
//   cluster     reduce-scatter + all-gather over distributed shared memory as PUSHES by the bulk-copy engine
//               (cp.async.bulk shared::cta -> shared::cluster), each completing on an mbarrier of the receiving CTA:
//               every CTA sends slice r of its partial gradient to CTA r; CTA r waits for the three slices, sums the four
//               in rank order 0..3 (deterministic), applies Adam to its slice (mu / nu of the slice live in its
//               registers), and sends the new slice of theta to the three peers' replicas; everybody waits for the three
//               slices it does not own -> next step.  No cluster barrier and no cluster-scope fence inside the step loop:
//               data movement and its completion signal are one operation of the async proxy, which is what the PTX
//               memory model sanctions for cross-CTA shared memory (the previous pull form needed a release at cluster
//               scope = MEMBAR.ALL.GPU twice per step, 12 % of the step, or relied on unstated hardware behaviour).
// theta and theta^- stay replicated in every CTA's shared memory; HBM sees only the gathered records and the loss.
// Same arithmetic and same Philox sampling as train_fused.cu; summation orders differ, results agree to fp32 round-off.
#include <cooperative_groups.h>
#include <math.h>

#include "act_device.cuh"
#include "common.cuh"
#include "kernels.h"
#include "tile_ops.cuh"
#include "tile16.cuh"

namespace cg = cooperative_groups;

namespace dqn {

namespace {

using namespace tile;

constexpr int NT = 256;
constexpr int CS = 4;             // CTAs per cluster
constexpr int BT = 64;            // rows per tile (whole cluster)
using t16::R;                     // 16 rows per CTA per tile (= BT / CS); strides and buffers of the 16-row tile: tile16.cuh
using t16::FR;
using t16::XS;
using t16::HS;
using t16::RS;
using t16::DRS;
using t16::HC;
static_assert(R * CS == BT, "the cluster splits a 64-row tile into four 16-row tiles");
constexpr int NPC = 4;            // slice parameters per thread: one 16-byte chunk (slice <= 828 floats = 207 chunks <= 256 threads)

struct CLay {
  int pW2, pWh, PS, SL;
  int oW, oWt, oG, oGin, oX, oH1, oH2, oDh2T, oDh2R, oDh1T, oDhdT, oScr, oMeta, oDummy, oRed, oCmd, oBar, oStage, oAct, total;
};

__host__ __device__ inline CLay make_clayout(int D, int recw) {
  CLay L;
  L.pW2 = packed_w2(D);
  L.pWh = packed_head(D);
  L.PS = packed_count(D);          // the shared-memory weight layout IS the packed HBM layout (common.cuh)
  L.SL = (((L.PS + CS - 1) / CS) + 3) & ~3;
  const int PSa = L.PS;
  int o = 0;
  L.oW = o; o += PSa;
  L.oWt = o; o += PSa;
  L.oG = o; o += PSa;
  L.oGin = o; o += (CS - 1) * L.SL;      // gradient slices pushed in by the three peers
  L.oX = o; o += (D + 1) * XS;
  L.oH1 = o; o += kH1 * HS;
  L.oH2 = o; o += (kH2 + 1) * HS;
  L.oDh2T = o; o += kH2 * RS;
  L.oDh2R = o; o += R * DRS;
  L.oDh1T = o; o += kH1 * RS;
  L.oDhdT = o; o += HC * RS;
  L.oScr = o; o += 4 * HC * FR;
  L.oMeta = o; o += R * 4;
  L.oDummy = o; o += 4;
  L.oRed = o; o += 64;
  L.oCmd = o; o += 8;
  L.oBar = o; o += 4;                     // two mbarriers: [0] peers' gradient slices have landed, [1] peers' weight slices have landed
  L.oStage = o; o += R * recw;
  L.oAct = o; o += kMaxD;                 // session: the state of an ACT command (the record staging buffer may hold a speculative gather)
  L.total = o;
  return L;
}

#ifdef DQN_PHASE_CLOCKS   // profiling variant only (profiles/phase_clocks.py): SM-cycle stamps of rank 0, thread 0
__device__ long long g_phase_clock[16];
#define PHASE_CLOCK(i) do { if (rank == 0 && t == 0) g_phase_clock[i] = clock64(); } while (0)
#else
#define PHASE_CLOCK(i) do { } while (0)
#endif

// Cluster barrier with release / acquire semantics (MEMBAR.ALL.GPU): used outside the step loop only -- the session's
// SYNC command and the exit.  The per-step exchange needs none (push + mbarrier, see the file header).
__device__ __forceinline__ void cluster_barrier_after_local_stores() { cg::this_cluster().sync(); }

// ---- mbarrier / bulk-copy primitives of the push exchange ----
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, int rank) {      // shared::cta address -> shared::cluster address of CTA `rank`
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
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
// this CTA's shared memory -> a peer's shared memory, completion (bytes) signalled on the PEER's mbarrier
__device__ __forceinline__ void bulk_push(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster) : "memory");
}
// generic-proxy stores of this thread -> visible to the async proxy (the bulk copy that will read them)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- session ("serve") mode: the kernel stays resident and takes commands from mapped host memory (kernels.h) ----
__device__ __forceinline__ unsigned long long ld_sys_u64(const volatile unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_sys_u32(const volatile uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys_u64(volatile unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
constexpr long long kIdleCycles = 60000000;   // ~30 ms at 1.965 GHz without a command: write back and exit on its own

// SERVE = false: K steps then write back (dqn_train_step*).  SERVE = true: session mode, commands until EXIT / idle.
// Two instantiations so that the K-step form carries none of the command loop (it costs ~3 % of a fused step).
template <int A, bool SERVE>
__global__ void __launch_bounds__(NT, 1) dqn_train_cluster_kernel(const TrainArgs args, const InlineStore ist) {
  extern __shared__ __align__(16) float sm[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int sel = blockIdx.x / CS;
  const int agent = args.agent_begin + sel;
  const int D = args.dims.D, recw = args.dims.recw, PK = args.dims.PK;
  const CLay L = make_clayout(D, recw);
  if (args.gate && !args.gate[agent].train_flag) return;   // episode gate closed (q_agent.py:186): uniform over the CTA / cluster
  PHASE_CLOCK(0);
  if (threadIdx.x == 0) {
    mbar_init(smem_addr(sm + L.oBar), 1);          // one arrive.expect_tx by thread 0 per phase + the peers' complete_tx bytes
    mbar_init(smem_addr(sm + L.oBar) + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (SERVE) *reinterpret_cast<volatile unsigned long long*>(sm + L.oCmd) = 0ull;   // session command word: exists before rank 0 may store into it
  }
  // every CTA's mbarriers exist before any peer may push into it: arrive here, wait just before the first step
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");

  float* const W = sm + L.oW;
  float* const Wt = sm + L.oWt;
  float* const G = sm + L.oG;
  float* const X = sm + L.oX;         // [D+1][36]
  float* const H1 = sm + L.oH1;       // [32][52]
  float* const H2 = sm + L.oH2;       // [65][52]   row 64 = ones
  float* const Dh2T = sm + L.oDh2T;   // [64 j][20]
  float* const Dh2R = sm + L.oDh2R;   // [16 r][68]
  float* const Dh1T = sm + L.oDh1T;   // [32 k][20]
  float* const DhdT = sm + L.oDhdT;   // [8 c][20]
  float* const Scr = sm + L.oScr;     // [4 parts][8 c][48 forward rows]
  float* const Meta = sm + L.oMeta;   // [16][4]
  float* const Red = sm + L.oRed;     // [0] this CTA's loss share, [8..11] Adam bias corrections, [16..31] head-bias partials
  volatile unsigned long long* const CmdWord = reinterpret_cast<volatile unsigned long long*>(sm + L.oCmd);   // session: (seq << 16 | op << 8 | n), written by rank 0
  volatile int* const Cmd = reinterpret_cast<volatile int*>(sm + L.oCmd + 2);                                // session: op, n for the whole CTA
  float* const Stage = sm + L.oStage;
  float* const ActS = sm + L.oAct;
  t16::Bufs tb;                         // the 16-row tile code (tile16.cuh) works on these buffers
  tb.W = W; tb.Wt = Wt; tb.G = G; tb.X = X; tb.H1 = H1; tb.H2 = H2; tb.Dh2T = Dh2T; tb.Dh2R = Dh2R; tb.Dh1T = Dh1T; tb.DhdT = DhdT;
  tb.Scr = Scr; tb.Meta = Meta; tb.Red = Red; tb.pW2 = L.pW2; tb.pWh = L.pWh; tb.D = D;

  float* const gW = args.params + (size_t)agent * 4 * PK;
  float* const gWt = gW + PK;
  float* const gM = gW + 2 * PK;
  float* const gV = gW + 3 * PK;
  AgentCtl* const ctl = args.ctl + agent;
  uint32_t* const ring = args.rings + (size_t)agent * args.dims.N * recw;

  // ---- one-time: full theta / theta^- replicas into smem; mu / nu of this CTA's slice into registers ----
  for (int p4 = t; p4 < (L.PS >> 2); p4 += NT) {   // packed layout: asynchronous 16-byte copies, awaited with the first gather
    cp_async16(W + 4 * p4, gW + 4 * p4);
    cp_async16(Wt + 4 * p4, gWt + 4 * p4);
    st4(G + 4 * p4, 0.f, 0.f, 0.f, 0.f);
  }
  // Adam ownership: thread t owns the 16-byte chunk [rank * SL + 4 t, +4) of the packed parameters (mu / nu in registers)
  const int pown = rank * L.SL + 4 * t;
  const bool owner = 4 * t < L.SL && pown < L.PS;
  float mreg[NPC] = {0.f, 0.f, 0.f, 0.f}, vreg[NPC] = {0.f, 0.f, 0.f, 0.f};
  if (owner) {
    const float4 m4 = *reinterpret_cast<const float4*>(gM + pown), v4 = *reinterpret_cast<const float4*>(gV + pown);
    mreg[0] = m4.x; mreg[1] = m4.y; mreg[2] = m4.z; mreg[3] = m4.w;
    vreg[0] = v4.x; vreg[1] = v4.y; vreg[2] = v4.z; vreg[3] = v4.w;
  }
  for (int r = t; r < XS; r += NT) X[D * XS + r] = 1.f;
  for (int r = t; r < HS; r += NT) H2[kH2 * HS + r] = 1.f;

  const float gamma = ctl->gamma, lr = ctl->lr, b1 = ctl->b1, b2 = ctl->b2;
  const float eps = ctl->eps, eps_root = ctl->eps_root, wd = ctl->wd;
  const int B = ctl->batch_size;
  const bool l2loss = ctl->loss_kind == kLossL2;
  const long long step0 = ctl->train_steps;
  const int count0 = ctl->adam_count;
  // ReplayBuffer.add x ist.n (replay_buffer.py:58-65) for transitions that arrived in the parameter buffer.  Every
  // CTA of the cluster writes the same few hundred bytes (identical values, so the race is benign): its own gather
  // below is then ordered behind its own writes by a fence + CTA barrier, without a cluster barrier.
  const long long rc0 = ctl->ring_counter;
  if (ist.n > 0) {
    for (int w = t; w < ist.n * recw; w += NT) {
      const int i = w / recw, c = w - i * recw;
      ring[(size_t)((rc0 + i) % args.dims.N) * recw + c] = ist.rec[w];
    }
    __threadfence();
    __syncthreads();
  }
  long long rc = rc0 + ist.n;
  long long size = rc < args.dims.N ? rc : args.dims.N;
  SessionCtl* const sess = args.sess;
  constexpr bool serve = SERVE;
  const int ntiles = (B + BT - 1) / BT;
  const float fB = (float)B;
  const int cpr = recw >> 2;

  // push exchange: slice c of the packed parameter vector is [c * SL, c * SL + slice_floats(c)), owned by CTA c
  float* const Gin = sm + L.oGin;      // [3][SL]: slot of source rank s is s - (s > rank)
  const uint32_t bar_g = smem_addr(sm + L.oBar), bar_w = bar_g + 8;
  auto slice_bytes = [&](int c) { const int n = L.PS - c * L.SL; return (uint32_t)(4 * (n < L.SL ? n : L.SL)); };
  uint32_t bytes_w_in = 0;
#pragma unroll
  for (int c = 0; c < CS; ++c) if (c != rank) bytes_w_in += slice_bytes(c);
  const int pLoss = L.pW2 + kH2;       // a padding slot of the packed layout inside slice 0 (row 0 of [W2;b2], column 64): carries
                                       //   each CTA's loss share through the gradient exchange; masked out before Adam
  uint32_t par_g = 0, par_w = 0;       // phase parities of the two mbarriers

  // staged-record word -> smem destination (threads 0..63: 4 lanes per row)
  const int urow = t >> 2, ul4 = t & 3;
  int udst[12];
#pragma unroll
  for (int q = 0; q < 12; ++q) {
    const int wd = 4 * (ul4 + 4 * (q >> 2)) + (q & 3);
    int o = L.oDummy;
    if (t < 4 * R) {
      if (wd < D) o = L.oX + wd * XS + urow;
      else if (wd < 2 * D) o = L.oX + (wd - D) * XS + R + urow;
      else if (wd < 2 * D + 4) o = L.oMeta + urow * 4 + (wd - 2 * D);
    }
    udst[q] = o;
  }

  long long my_slot = -1;        // ring slot of the row this thread gathered last (threads 0..63)
  auto prefetch = [&](int kstep, int tile) {
    if (t < 4 * R) {
      const int i = tile * BT + rank * R + urow;
      float* dst = Stage + urow * recw;
      my_slot = -1;
      if (i < B) {
        long long slot;
        if (args.idx) slot = args.idx[((size_t)sel * args.K + kstep) * B + i];
        else slot = philox_index(args.seed, args.agent_id_base + agent, step0 + kstep, i, size);
        my_slot = slot;
        const uint32_t* src = ring + slot * recw;
        for (int c = ul4; c < cpr; c += 4) cp_async16(dst + 4 * c, src + 4 * c);
        if (args.taps.enabled && ul4 == 0 && args.taps.indices) args.taps.indices[i] = slot;
      } else {
        for (int c = ul4; c < cpr; c += 4) st4(dst + 4 * c, 0.f, 0.f, 0.f, 0.f);
      }
    }
    cp_async_commit();
  };

  // staged records -> k-major X / raw meta words; every thread reads back exactly the 16-byte chunks it copied itself
  auto unpack = [&]() {
    if (t < 4 * R) {
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        const int c = ul4 + 4 * cc;
        if (c < cpr) {
          const float4 v = ld4(Stage + urow * recw + 4 * c);
          sm[udst[4 * cc + 0]] = v.x; sm[udst[4 * cc + 1]] = v.y; sm[udst[4 * cc + 2]] = v.z; sm[udst[4 * cc + 3]] = v.w;
        }
      }
    }
  };
  bool unpacked = false;                    // K-step launches, one tile: the next step's rows were unpacked under the all-gather
  double pb1 = ctl->pb1, pb2 = ctl->pb2;    // b1**count, b2**count carried across launches (thread 0 uses them)

  PHASE_CLOCK(1);
  if (!serve) prefetch(0, 0);
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");   // pairs with the arrive at the top: all mbarriers are initialised
  PHASE_CLOCK(2);

  // ---- session mode -------------------------------------------------------------------------------------------------
  // Rank 0 polls the doorbell in mapped host memory and forwards each command word to the four CTAs' shared memory, so
  // the whole cluster takes the same decision (including the idle time-out).  Commands: STEP = n ReplayBuffer.add +
  // one Agent._step (records read from the host slot), ACT = greedy action for one state (rank 0), SYNC = theta^- :=
  // theta, EXIT.  Command seq lives in slot seq % 2 (doorbell, payload, response); the host has at most two in flight.
  unsigned long long next_seq = args.sess_first_seq;
  bool wt_dirty = false;
  // The host may have published the NEXT command (other slot) while this one runs: rank 0's polling warp issues the loads of
  // its doorbell and first 128 payload units at the tail of a step, under the all-gather, and looks at them first.
  unsigned long long pre_w = 0ull, pre_u[4] = {0ull, 0ull, 0ull, 0ull};
  bool have_pre = false;
  // Speculative gather: once the ring is full its size no longer depends on the next command, so the NEXT step's minibatch
  // indices are known before that command arrives.  The rows are gathered at the tail of a step, under the all-gather and
  // the command intake; when the command is there, a row whose slot its add()s overwrote is gathered again.
  bool spec = false;
  // payload of command `w` (polling warp of rank 0): STEP = its n records into the ring (ReplayBuffer.add x n,
  // replay_buffer.py:58-65), ACT = the state into shared memory; units whose stamp is not the command's are read again
  auto intake = [&](unsigned long long w, const unsigned long long (&u)[4], unsigned long long seq) {
    volatile unsigned long long* const units = sess->stamped[seq & 1];
    const int wop = (int)((w >> 8) & 0xff), wn = (int)(w & 0xff);
    const int need = wop == kOpStep ? wn * recw : (wop == kOpAct ? D : 0);
    for (int base = 0; base < need; base += 128) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int idx = base + lane + 32 * q;
        if (idx < need) {
          unsigned long long x = base == 0 ? u[q] : ld_sys_u64(&units[idx]);
          while ((uint32_t)(x >> 32) != (uint32_t)seq) x = ld_sys_u64(&units[idx]);
          if (wop == kOpStep) {
            const int i = idx / recw, c = idx - i * recw;
            ring[(size_t)((rc + i) % args.dims.N) * recw + c] = (uint32_t)x;
          } else {
            ActS[idx] = __uint_as_float((uint32_t)x);
          }
        }
      }
    }
    __threadfence();                       // the records are in the ring before any CTA is told about the step
    __syncwarp();
  };
  unsigned long long early_w = 0ull;       // a STEP command whose records went into the ring under the previous step's all-gather
  bool early_done = false;
  volatile unsigned long long* CmdWordR[CS];
#pragma unroll
  for (int c = 0; c < CS; ++c) CmdWordR[c] = cluster.map_shared_rank(const_cast<unsigned long long*>(CmdWord), c);

  int kstep = 0;
  for (;; ++kstep) {
    if (!serve && kstep >= args.K) break;
    if (serve) {
      int op, n;
      for (;;) {                         // commands that are not a train step are handled right here
        if (rank == 0) {
          // Warp 0 polls.  Every poll also reads the first 128 stamped payload units (seq << 32 | word), so that when the
          // doorbell shows the command its payload -- up to 128 words: 5 records at D <= 10, or the state of an ACT --
          // has arrived in the same PCIe round trip; a unit whose stamp is not the command's is re-read.
          if (warp == 0) {
            unsigned long long w = early_w, u[4];
            const long long c0 = clock64();
            volatile unsigned long long* const door = &sess->doorbell[next_seq & 1];      // the command's slot: seq % 2
            volatile unsigned long long* const units = sess->stamped[next_seq & 1];
            if (!early_done) {
            for (;;) {
              if (have_pre) {
                w = pre_w;
#pragma unroll
                for (int q = 0; q < 4; ++q) u[q] = pre_u[q];
                have_pre = false;
              } else {
                if (lane == 0) {
                  w = ld_sys_u64(door);
                  if ((w >> 16) != next_seq && clock64() - c0 > kIdleCycles) w = (next_seq << 16) | ((unsigned long long)kOpExit << 8);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) u[q] = ld_sys_u64(&units[lane + 32 * q]);
              }
              w = __shfl_sync(0xffffffffu, w, 0);
              if ((w >> 16) == next_seq) break;
            }
            intake(w, u, next_seq);
            }
            early_done = false;
            const int wop = (int)((w >> 8) & 0xff), wn = (int)(w & 0xff);
            if (lane == 0) {
#pragma unroll
              for (int c = 0; c < CS; ++c) *CmdWordR[c] = w;
              Cmd[0] = wop; Cmd[1] = wn; Cmd[2] = 0;
            }
          }
        } else if (t == 0) {
          // rank 0 answers ACT on its own and may already have forwarded a later command: ACTs can be skipped here
          // (every other command holds rank 0 at a cluster barrier until this CTA has taken it)
          unsigned long long w;
          do { w = *CmdWord; } while ((w >> 16) < next_seq);
          Cmd[0] = (int)((w >> 8) & 0xff);
          Cmd[1] = (int)(w & 0xff);
          Cmd[2] = (int)((w >> 16) - next_seq);
        }
        __syncthreads();
        op = Cmd[0]; n = Cmd[1];
        next_seq += (unsigned long long)Cmd[2];
        __syncthreads();
        if (op == kOpAct) {              // compute_action (q_learning_functions.py:67-73) from the resident weights
          if (rank == 0 && warp == 0) {          // the state is already in shared memory (stamped payload)
            const int best = warp_greedy_action(W, D, A, ActS, nullptr);
            if (lane == 0) st_sys_u64(&sess->response[next_seq & 1], (next_seq << 32) | (unsigned long long)(uint32_t)best);
          }
          __syncthreads();
          ++next_seq;
        } else if (op == kOpSync) {      // Agent._update_target_model (q_agent.py:143-144) on every replica
          for (int p4 = t; p4 < (L.PS >> 2); p4 += NT) { const float4 w4 = ld4(W + 4 * p4); st4(Wt + 4 * p4, w4.x, w4.y, w4.z, w4.w); }
          wt_dirty = true;
          cluster_barrier_after_local_stores();     // every CTA has taken the command before rank 0 can forward the next one
          if (rank == 0 && t == 0) st_sys_u64(&sess->response[next_seq & 1], next_seq << 32);
          ++next_seq;
        } else {
          break;
        }
      }
      if (op != kOpStep) break;          // EXIT (or the idle time-out)
      const long long rc_old = rc;
      rc += n;                           // rank 0 has put the n records into the ring (above)
      size = rc < args.dims.N ? rc : args.dims.N;
      if (spec) {                        // rows gathered ahead: only those in the window [rc_old, rc_old + n) (mod N) are stale
        spec = false;
        if (t < 4 * R && my_slot >= 0) {
          long long dlt = my_slot - rc_old % args.dims.N;
          if (dlt < 0) dlt += args.dims.N;
          if (dlt < n) {
            cp_async_wait_all();         // the stale copy has landed before the fresh one is issued
            const uint32_t* src = ring + my_slot * recw;
            float* dst = Stage + urow * recw;
            for (int c = ul4; c < cpr; c += 4) cp_async16(dst + 4 * c, src + 4 * c);
          }
        }
        cp_async_commit();
      } else {
        prefetch(kstep, 0);
      }
    }
    const bool gather_after = serve || kstep + 1 < args.K;      // the last step of a K-step launch needs no all-gather
    if (t == 0) {
      if (count0 + kstep < 0x7fffffff) { pb1 *= (double)b1; pb2 *= (double)b2; }
      Red[8 + 2 * (kstep & 1)] = 1.0f - (float)pb1;
      Red[9 + 2 * (kstep & 1)] = 1.0f - (float)pb2;
      mbar_arrive_expect_tx(bar_g, (CS - 1) * slice_bytes(rank));        // this step's incoming gradient slices
      if (gather_after) mbar_arrive_expect_tx(bar_w, bytes_w_in);         // ... and weight slices
    }
    float loss_acc = 0.f;    // warp 0 only

    for (int tile = 0; tile < ntiles; ++tile) {
      // ---- unpack (already done under the previous step's all-gather when `unpacked`) ----
      if (!(unpacked && tile == 0)) {
        cp_async_wait_all();
        __syncthreads();
        if (kstep == 0 && tile == 0) PHASE_CLOCK(3);
        unpack();
      }
      unpacked = false;
      __syncthreads();
      if (tile + 1 < ntiles) prefetch(kstep, tile + 1);
      else if (!serve && kstep + 1 < args.K) prefetch(kstep + 1, 0);

      // ---- forward, targets / Huber, backward of this CTA's 16 rows (tile16.cuh); partial gradient += into G ----
      {
        const float lp = t16::step<A>(tb, tile * BT + rank * R, B, fB, gamma, l2loss, args.taps);
        if (t < R) loss_acc += lp;
      }
    }  // tiles

    if (kstep == args.K - 1) PHASE_CLOCK(4);
    if (t == 0) G[pLoss] = loss_acc;         // this CTA's share of the loss rides in a padding slot of slice 0
    fence_proxy_async();                     // this thread's stores into G -> visible to the bulk copies issued below
    __syncthreads();                         // the partial gradient of this CTA is complete
    // ---- reduce-scatter (push): slice c of this CTA's partial gradient -> slot of CTA c ----
    if (t < CS && t != rank)
      bulk_push(map_to_rank(smem_addr(Gin + (rank - (rank > t ? 1 : 0)) * L.SL), t), smem_addr(G + t * L.SL), slice_bytes(t), map_to_rank(bar_g, t));
    if (serve && rank == 0 && warp == 0) {   // the NEXT command's doorbell and first payload units: loads in flight under the exchange and Adam
      const unsigned long long ns = next_seq + 1;
      if (lane == 0) pre_w = ld_sys_u64(&sess->doorbell[ns & 1]);
#pragma unroll
      for (int q = 0; q < 4; ++q) pre_u[q] = ld_sys_u64(&sess->stamped[ns & 1][lane + 32 * q]);
      have_pre = true;
    }
    mbar_wait(bar_g, par_g);                 // the three peers' slices of MY slice have landed
    par_g ^= 1u;
    if (kstep == args.K - 1) PHASE_CLOCK(5);

    // ---- Adam on this CTA's slice; the new slice goes to the local replica, then to the peers ----
    if (owner) {
      const float c1 = Red[8 + 2 * (kstep & 1)], c2 = Red[9 + 2 * (kstep & 1)];
      const float rc1 = 1.0f / c1, rc2 = 1.0f / c2;
      const float omb1 = 1.0f - b1, omb2 = 1.0f - b2;
      float4 gs[CS];
#pragma unroll
      for (int c = 0; c < CS; ++c) gs[c] = c == rank ? ld4(G + pown) : ld4(Gin + (c - (c > rank ? 1 : 0)) * L.SL + 4 * t);
      const float4 th4 = ld4(W + pown);
      float g[4] = {((gs[0].x + gs[1].x) + gs[2].x) + gs[3].x, ((gs[0].y + gs[1].y) + gs[2].y) + gs[3].y,      // fixed rank order: deterministic
                    ((gs[0].z + gs[1].z) + gs[2].z) + gs[3].z, ((gs[0].w + gs[1].w) + gs[2].w) + gs[3].w};
      if (rank == 0 && 4 * t == (pLoss & ~3)) {      // the thread whose chunk holds the loss slot (pLoss % 4 == 0: component 0)
        const float loss = g[0] / fB;
        g[0] = 0.f;                                   // padding parameter: no gradient, stays 0
        args.loss_ring[(size_t)agent * kLossCap + (size_t)((step0 + kstep) % kLossCap)] = loss;
        if (serve || kstep == args.K - 1)
          args.loss_mailbox[agent] = ((unsigned long long)(uint32_t)(step0 + kstep + 1) << 32) | __float_as_uint(loss);
        if (serve) st_sys_u64(&sess->response[next_seq & 1], (next_seq << 32) | __float_as_uint(loss));
        if (args.taps.enabled && args.taps.loss) args.taps.loss[0] = loss;
      }
      const float th[4] = {th4.x, th4.y, th4.z, th4.w};
      float nw[4];
#pragma unroll
      for (int i = 0; i < NPC; ++i) {
        const float m = b1 * mreg[i] + omb1 * g[i];
        const float v = b2 * vreg[i] + omb2 * (g[i] * g[i]);
        mreg[i] = m; vreg[i] = v;
        const float u = (m * rc1) * fast_rcp(fast_sqrt(v * rc2 + eps_root) + eps);
        nw[i] = th[i] - lr * (u + wd * th[i]);
      }
      st4(W + pown, nw[0], nw[1], nw[2], nw[3]);
      if (args.taps.enabled && args.taps.grads) st4(args.taps.grads + pown, g[0], g[1], g[2], g[3]);
    }
    if (serve) {
      ++next_seq;
      if (rc >= args.dims.N && !args.idx) { prefetch(kstep + 1, 0); spec = true; }   // (the staging buffer was unpacked long ago)
    }
    if (gather_after) {
      // ---- all-gather (push): my new slice -> the three peers' replicas; wait for theirs ----
      fence_proxy_async();
      __syncthreads();                       // the whole slice is written
      if (t < CS && t != rank)
        bulk_push(map_to_rank(smem_addr(W + rank * L.SL), t), smem_addr(W + rank * L.SL), slice_bytes(rank), map_to_rank(bar_w, t));
      if (serve && rank == 0 && warp == 0 && have_pre) {
        // If the next command is already there and is a train step, its records go into the ring now, while the weight slices
        // travel (this step's gather finished long ago; a speculatively gathered row they overwrite is gathered again).
        const unsigned long long w = __shfl_sync(0xffffffffu, pre_w, 0);
        if ((w >> 16) == next_seq && ((w >> 8) & 0xff) == kOpStep) {
          intake(w, pre_u, next_seq);
          early_w = w; early_done = true; have_pre = false;
          // ... and the peers are told right away (they look at the word at the top of their loop, after their own wait below;
          // everything of THIS step they could disturb -- the gradient slots -- was consumed by the Adam above)
          if (lane == 0) {
#pragma unroll
            for (int c = 1; c < CS; ++c) *CmdWordR[c] = w;
          }
        }
      }
      if (!serve && ntiles == 1 && kstep + 1 < args.K) {
        // the next step's rows (gathered since the top of this step) go to X / Meta while the weight slices travel: X is free
        // since the barrier after the backward, and a thread unpacks only chunks of its own copies (no barrier needed before)
        cp_async_wait_all();
        unpack();
        unpacked = true;
      }
      mbar_wait(bar_w, par_w);
      par_w ^= 1u;
      // Only now may the partial gradient be cleared: a bulk copy signals completion at its DESTINATION, so the sender
      // learns that its pushes out of G were consumed through the peers' weight slices -- each of which was sent behind
      // an Adam step that had waited for exactly those pushes.
      for (int p4 = t; p4 < (L.PS >> 2); p4 += NT) st4(G + 4 * p4, 0.f, 0.f, 0.f, 0.f);
    }
  }  // steps
  PHASE_CLOCK(6);
  cp_async_wait_all();                     // (a speculative gather that no command consumed)
  // No CTA may leave while a bulk copy still reads its shared memory or a peer may still push into it: every push of
  // the last step has completed once all four CTAs are past their last mbarrier wait.
  cluster.sync();

  // ---- write back this CTA's slice (the last cluster barrier above was the last DSMEM access: CTAs may exit freely) ----
  if (owner) {
    if (wt_dirty) *reinterpret_cast<float4*>(gWt + pown) = ld4(Wt + pown);
    *reinterpret_cast<float4*>(gW + pown) = ld4(W + pown);
    *reinterpret_cast<float4*>(gM + pown) = make_float4(mreg[0], mreg[1], mreg[2], mreg[3]);
    *reinterpret_cast<float4*>(gV + pown) = make_float4(vreg[0], vreg[1], vreg[2], vreg[3]);
  }
  if (rank == 0 && t == 0) {
    ctl->train_steps = step0 + kstep;
    if (ist.n > 0 || serve) ctl->ring_counter = rc;
    const long long c = (long long)count0 + kstep;
    ctl->adam_count = c > 0x7fffffffLL ? 0x7fffffff : (int)c;
    ctl->pb1 = pb1; ctl->pb2 = pb2;
    if (serve) { __threadfence_system(); st_sys_u64(&sess->closed, next_seq); }
  }
  PHASE_CLOCK(7);
}

typedef void (*ClusterKernel)(const TrainArgs, const InlineStore);
template <bool SERVE>
ClusterKernel pick_cluster_kernel_t(int A) {
  switch (A) {
    case 2: return dqn_train_cluster_kernel<2, SERVE>;
    case 3: return dqn_train_cluster_kernel<3, SERVE>;
    case 4: return dqn_train_cluster_kernel<4, SERVE>;
    case 5: return dqn_train_cluster_kernel<5, SERVE>;
    case 6: return dqn_train_cluster_kernel<6, SERVE>;
    case 7: return dqn_train_cluster_kernel<7, SERVE>;
    default: return nullptr;
  }
}
ClusterKernel pick_cluster_kernel(int A, bool serve) { return serve ? pick_cluster_kernel_t<true>(A) : pick_cluster_kernel_t<false>(A); }

}  // namespace

#ifdef DQN_PHASE_CLOCKS
extern "C" __attribute__((visibility("default"))) int dqn_debug_phase_clocks(long long* out16) {
  return (int)cudaMemcpyFromSymbol(out16, g_phase_clock, sizeof(long long) * 16);
}
#endif

size_t train_cluster_smem_bytes(const Dims& d) { return (size_t)make_clayout(d.D, d.recw).total * sizeof(float); }

cudaError_t train_cluster_prepare(const Dims& d) {
  for (int serve = 0; serve < 2; ++serve) {
    ClusterKernel k = pick_cluster_kernel(d.A, serve != 0);
    if (!k) return cudaErrorInvalidValue;
    const cudaError_t e = cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)train_cluster_smem_bytes(d));
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_train_cluster(cudaStream_t st, const TrainArgs& args, const InlineStore* ist) {
  static const InlineStore none = {};
  ClusterKernel k = pick_cluster_kernel(args.dims.A, args.sess != nullptr);
  if (!k) return cudaErrorInvalidValue;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(args.n_sel * CS));
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = train_cluster_smem_bytes(args.dims);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k, args, ist ? *ist : none);
}

}  // namespace dqn
