// common.cuh -- shared device/host definitions for libdqn_b200 (sm_100a).
//
// HBM data layout (see DESIGN.md "Data layout"):
//   params   f32 [n_agents][4][PK]        online | target | mu | nu, each in the PACKED layout below (the layout
//                                          the train-step kernels keep in shared memory, so a launch starts with a
//                                          straight 16-byte copy); the C ABI converts to / from the flat layout of
//                                          include/dqn_b200.h in dqn_set_* / dqn_get_*.
//   ctl      AgentCtl [n_agents]          per-agent hyper-parameters and counters.
//   ring     u8  [n_agents][N][REC]       AoS replay records, REC = round_up((2D+4)*4, 32) bytes:
//                                          words [0,D) s | [D,2D) s' | 2D,2D+1 action (i64) |
//                                          2D+2 reward (f32) | 2D+3 done (u32 0/1) | zero padding.
//            One record = one 32-byte-sector-aligned gather (96 B for D <= 10), instead of five
//            scattered sectors for the reference's five SoA arrays (replay_buffer.py:26-30);
//            dqn_buffer_export() de-interleaves back to those arrays bit-exactly.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

namespace dqn {

constexpr int kH1 = 32;        // LunarLander/dddqn.py:19
constexpr int kH2 = 64;        // LunarLander/dddqn.py:20
constexpr int kMaxD = 16;
constexpr int kMaxA = 7;
constexpr int kLossCap = 4096; // per-agent ring of recent losses
enum { kLossHuber = 0, kLossL2 = 1 };

struct AgentCtl {
  float gamma, lr, b1, b2, eps, eps_root, wd;
  int batch_size;
  int loss_kind;           // kLossHuber (the reference, q_learning_functions.py:36) or kLossL2 (0.5 e^2, optax.l2_loss)
  int adam_count;          // optax ScaleByAdamState.count (int32, saturating)
  long long ring_counter;  // ReplayBuffer._counter (total adds)
  long long train_steps;   // number of _step() calls so far == Philox step counter
  double pb1, pb2;         // b1**adam_count, b2**adam_count carried in double (one DMUL per step instead of a pow() per
                           //   launch); the host re-seeds them whenever it sets the count or the decay rates
};

// Per-agent episode-loop state (episode.cu): what Agent._run_episode / training keep in Python attributes and locals
// (General/QLearning/q_agent.py:96-107, :171-222).  epsilon / rewards are doubles like the Python floats they mirror.
constexpr int kRewardWindow = 50;              // q_agent.py:125
constexpr uint32_t kPolicyTag = 0x504F4C49u;   // 'POLI': Philox key tweak of the epsilon-greedy stream
struct EpisodeCtl {
  double epsilon, eps_decay, min_eps, reward_to_reach;
  double epi_reward, last_epi_reward, avg_reward;
  double window[kRewardWindow];
  int window_len, window_pos;
  int max_episodes, max_steps, training_start, train_frequency, replace_frequency;
  int step_in_episode, episode;
  int train_flag, sync_flag, finished;
  long long step_count, policy_calls;
};

struct Dims {
  int D, A;      // obs dim, actions
  int P;         // flat parameter count
  int PK;        // packed (HBM / shared-memory) parameter count, a multiple of 4
  int recw;      // record stride in 32-bit words
  long long N;   // ring slots per agent
};

__host__ __device__ inline int flat_param_count(int D, int A) {
  return D * kH1 + kH1 + kH1 * kH2 + kH2 + kH2 + 1 + kH2 * A + A;
}

// Packed parameter layout ("augmented" matrices: the bias is the last row of each block, padding is 0):
//   [W1;b1]  (D+1) x 32              at 0
//   [W2;b2]  33 x 64, row stride 68  at packed_w2(D)      (68 = 4 mod 32 words: rows 4 banks apart in shared memory)
//   head     65 x 8                  at packed_head(D)    (col 0 = V, cols 1..A = advantage, row 64 = biases)
constexpr int kW2Stride = kH2 + 4;
constexpr int kHeadCols = 8;
__host__ __device__ inline int packed_w2(int D) { return (D + 1) * kH1; }
__host__ __device__ inline int packed_head(int D) { return packed_w2(D) + (kH1 + 1) * kW2Stride; }
__host__ __device__ inline int packed_count(int D) { return (packed_head(D) + (kH2 + 1) * kHeadCols + 3) & ~3; }
// packed index -> flat index of include/dqn_b200.h (-1 = padding)
__host__ __device__ inline int packed_to_flat(int p, int D, int A) {
  const int pW2 = packed_w2(D), pWh = packed_head(D);
  if (p < pW2) return p;                              // [W1;b1] is contiguous in both layouts
  const int offWv = pW2 + (kH1 + 1) * kH2;            // flat offset of Wv = D*32 + 32 + 32*64 + 64
  if (p < pWh) {
    const int q = p - pW2;
    const int row = q / kW2Stride, col = q - row * kW2Stride;
    return col < kH2 ? pW2 + row * kH2 + col : -1;
  }
  const int q = p - pWh;
  const int k = q >> 3, c = q & 7;
  const int offbv = offWv + kH2, offWa = offbv + 1, offba = offWa + kH2 * A;
  if (k > kH2 || c > A) return -1;
  if (k < kH2) return c == 0 ? offWv + k : offWa + k * A + (c - 1);
  return c == 0 ? offbv : offba + (c - 1);
}
__host__ __device__ inline int record_words(int D) { return (((2 * D + 4) * 4 + 31) / 32) * 8; }
// Host side: the record stride a handle is created with.  DQN_B200_RECORD_BYTES (a multiple of 32, experiment knob) raises
// the stride, e.g. to one 128-byte L2 line per record.
inline int record_words_host(int D) {
  int w = record_words(D);
  if (const char* e = getenv("DQN_B200_RECORD_BYTES")) {
    const int b = atoi(e);
    if (b % 32 == 0 && b / 4 >= w && b <= 160) w = b / 4;
  }
  return w;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. SC'11).  Bit-exact twin of oracle/philox.py.
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Slot index for sample `i` of train step `step` of `agent`: uniform in [0,size), with replacement
// (the distribution of numpy.random.randint(0, size, B), replay_buffer.py:77).
__device__ __forceinline__ long long philox_index(uint64_t seed, int agent, long long step, int i, long long size) {
  uint32_t o[4];
  philox4x32_10((uint32_t)i, (uint32_t)step, (uint32_t)((uint64_t)step >> 32), (uint32_t)agent,
                (uint32_t)seed, (uint32_t)(seed >> 32), o);
  const uint64_t x = (uint64_t)o[0] | ((uint64_t)o[1] << 32);
  return (long long)__umul64hi(x, (uint64_t)size);
}

}  // namespace dqn
