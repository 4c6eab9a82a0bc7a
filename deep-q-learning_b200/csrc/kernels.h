// kernels.h -- host-side launchers of the sm_100a kernels (internal; the public surface is
// include/dqn_b200.h).
#pragma once
#include "common.cuh"

namespace dqn {

void set_last_error(const char* msg);   // thread-local message behind dqn_last_error()

// Device-pointer taps of one train step (see dqn_debug_taps in include/dqn_b200.h).
struct TapsDev {
  long long* indices;
  float* q;
  float* next_q;
  float* next_q_tm;
  int* max_actions;
  float* targets;
  float* loss;
  float* grads;   // packed layout, PK floats
  int enabled;
};

// Session mode of the cluster train-step kernel: a block of MAPPED PINNED host memory through which a resident kernel
// takes commands (no launch, no copy per command).  The host writes the payload and then the doorbell of slot seq % 2;
// the kernel answers in that slot's `response`.  Commands are served in sequence order; at most TWO are in flight (one per
// slot), so the host can publish train step i + 1 while the kernel runs step i and read step i's loss afterwards.
enum { kOpStep = 1, kOpAct = 2, kOpSync = 3, kOpExit = 4 };
constexpr int kSessionSlotUnits = 16 * 40;
struct SessionCtl {
  volatile unsigned long long doorbell[2];   // host -> device: (seq << 16) | (op << 8) | n
  volatile unsigned long long response[2];   // device -> host: (seq << 32) | payload (loss bits / action)
  volatile unsigned long long closed;        // device -> host: written once when the kernel leaves (next seq it would have served)
  unsigned long long pad[3];
  // payload, one 8-byte unit per 32-bit word: (low 32 bits of seq) << 32 | word.  STEP: n records in the ring's AoS
  // layout, stride dims.recw words; ACT: the D floats of the state.  A unit is valid when its stamp is the command's.
  volatile unsigned long long stamped[2][kSessionSlotUnits];
};

struct TrainArgs {
  float* params;              // [n_agents][4][PK], packed layout (common.cuh)
  AgentCtl* ctl;              // [n_agents]
  uint32_t* rings;            // [n_agents][N][recw]
  float* loss_ring;           // [n_agents][kLossCap]
  unsigned long long* loss_mailbox;   // [n_agents] zero-copy (mapped pinned host memory): (train_steps << 32) | loss bits
                                      //   of the launch's last step, one 8-byte store -- the host can poll it
  const long long* idx;       // device i64 [n_sel][K][B] or nullptr (Philox)
  const EpisodeCtl* gate;     // nullptr, or per-agent episode state: only agents with gate[agent].train_flag step
  const int* order;           // nullptr, or (population kernel) [n_sel]: CTA b runs selected agent order[b] -- costliest first
  SessionCtl* sess;           // nullptr, or (cluster kernel, one agent) the session block: serve commands until EXIT / idle
  unsigned long long sess_first_seq;   // sequence number of the first command this launch serves
  Dims dims;
  unsigned long long seed;
  int agent_begin;
  int agent_id_base;   // Philox agent id = agent_id_base + agent
  int n_sel;
  int K;
  TapsDev taps;
};

cudaError_t launch_replay_store(cudaStream_t st, uint32_t* ring, const Dims& d, long long counter, long long n,
                                const float* s, const long long* a, const float* r, const float* s2,
                                const uint8_t* done, AgentCtl* ctl);
// out[i] = Philox slot index of sample (first + i) of `step` -- `first` lets a data-parallel rank draw its slice
cudaError_t launch_philox_indices(cudaStream_t st, long long* out, int batch, uint64_t seed, int agent,
                                  long long step, long long size, int first = 0);
// mode: 0 explicit idx, 1 Philox(seed, agent, step), 2 identity (export)
cudaError_t launch_replay_gather(cudaStream_t st, const uint32_t* ring, const Dims& d, int mode, const long long* idx,
                                 uint64_t seed, int agent, long long step, long long size, long long batch,
                                 float* s, long long* a, float* r, float* s2, uint8_t* done);

size_t train_fused_smem_bytes(const Dims& d);
cudaError_t train_fused_prepare(const Dims& d);   // cudaFuncSetAttribute for the instantiation
cudaError_t launch_train_fused(cudaStream_t st, const TrainArgs& args);
// the same one-CTA-per-agent step with the layer-2 products on the tensor cores (train_fused_tc.cu: tcgen05 3xTF32, 512 threads)
size_t train_tc_smem_bytes(const Dims& d);
cudaError_t train_tc_prepare(const Dims& d);
cudaError_t launch_train_tc(cudaStream_t st, const TrainArgs& args);
// the same step with one agent spread over a 4-CTA thread-block cluster (train_cluster.cu)
size_t train_cluster_smem_bytes(const Dims& d);
cudaError_t train_cluster_prepare(const Dims& d);
// Transitions that ride in the kernel-parameter buffer (no H2D copy, no separate store launch): n records in the
// ring's AoS layout with stride dims.recw words, written to slots (ring_counter + i) % N before the first gather.
constexpr int kInlineMax = 16;
constexpr int kInlineWords = 40;      // record words at D = 16
struct InlineStore {
  int n;
  int pad[3];
  uint32_t rec[kInlineMax * kInlineWords];
};
cudaError_t launch_train_cluster(cudaStream_t st, const TrainArgs& args, const InlineStore* ist /* or nullptr */);

cudaError_t launch_act(cudaStream_t st, const float* params, const Dims& d, int agent_begin, int n_sel,
                       const float* states /* device [n_sel][D] */, int* actions_out /* device [n_sel] */,
                       float* q_out /* device [n_sel][A] or nullptr */);
cudaError_t launch_sync_target(cudaStream_t st, float* params, const Dims& d, int agent_begin, int n_sel);
cudaError_t launch_polyak_target(cudaStream_t st, float* params, const Dims& d, int agent_begin, int n_sel, float tau);

// episode-loop control on the device (episode.cu)
cudaError_t launch_policy(cudaStream_t st, const float* params, const Dims& d, EpisodeCtl* ep, int agent_begin, int n_sel,
                          int agent_id_base, unsigned long long seed, const float* states, int* actions);
cudaError_t launch_observe(cudaStream_t st, uint32_t* rings, AgentCtl* ctl, EpisodeCtl* ep, const Dims& d, int agent_begin, int n_sel,
                           const float* s, const int* a, const float* r, const float* s2, const uint8_t* done, uint8_t* episode_end);
cudaError_t launch_episode_post(cudaStream_t st, float* params, const Dims& d, EpisodeCtl* ep, int agent_begin, int n_sel);

// prioritized replay (per.cu)
cudaError_t launch_per_update(cudaStream_t st, float* tree, long long L, int levels, const long long* idx, const float* val, int n,
                              int is_td, float alpha, float eps);
cudaError_t launch_per_rebuild(cudaStream_t st, float* tree, long long L);
cudaError_t launch_per_sample(cudaStream_t st, const float* tree, long long L, int levels, int batch, uint64_t seed, long long step,
                              long long* idx_out, float* prio_out);

}  // namespace dqn
