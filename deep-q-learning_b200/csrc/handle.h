// handle.h -- internals shared by the translation units behind include/dqn_b200.h (api.cu, api_session.cu,
// api_episode.cu): the handle, the error / CUDA-check helpers and the few functions they call across files.
#pragma once
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <string>
#include <vector>

#include "../../include/dqn_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace dqn {
int host_fail(int code, const std::string& msg);     // records the thread-local message behind dqn_last_error()
}
static inline int fail(int code, const std::string& msg) { return dqn::host_fail(code, msg); }

#define CU(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      char _b[512];                                                                                \
      snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return fail(DQN_E_CUDA, _b);                                                                 \
    }                                                                                              \
  } while (0)

namespace dqn {

constexpr size_t kStageBytes = 8u << 20;
constexpr size_t kPinnedBytes = 128u << 10;   // [0,64K): ring of store slots; [64K,128K): bounce for synchronous calls
constexpr size_t kSlotBytes = 2u << 10;
constexpr int kSlots = 32;
constexpr size_t kBounceOff = 64u << 10;
constexpr size_t kBounceBytes = 64u << 10;


inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct Carve {
  size_t params, ctl, ep, rings, loss, stage, taps, total;
};

struct TapsOff {
  size_t idx, q, nq, nqt, maxa, tgt, loss, grads, bytes;
};

}  // namespace dqn

struct dqn_handle {
  dqn_config cfg;
  int step_kernel, sm_count;
  dqn::Dims dims;
  cudaStream_t stream;
  uint8_t* arena;
  bool own_arena;
  dqn::Carve cv;
  dqn::TapsOff to;
  float* params;
  dqn::AgentCtl* ctl;
  uint32_t* rings;
  float* loss_ring;
  uint8_t* stage;
  uint8_t* taps;
  uint8_t* pinned;
  uint8_t* bounce;              // pinned + kBounceOff
  volatile unsigned long long* mailbox;   // mapped pinned host memory, [n_agents]: (train_steps << 32) | loss bits of each
  unsigned long long* mailbox_dev;        //   agent's last launch (one 8-byte store by the kernel); device alias
  int* order;                             // mapped pinned, behind the mailbox: launch order of the population kernel for the agent
  const int* order_dev;                   //   range [order_begin, order_end) (-1: stale; rebuilt when batch sizes change)
  int order_begin, order_end;
  cudaEvent_t slot_ev[dqn::kSlots];  // completion of the H2D copy that last used each pinned store slot
  int slot_next;
  std::vector<dqn::AgentCtl> hctl;   // host mirror of the per-agent control blocks
  // session mode (dqn_set_session): a resident cluster kernel serves STEP / ACT / SYNC commands from mapped host memory
  dqn::SessionCtl* sess;             // mapped pinned host memory (nullptr until enabled)
  dqn::SessionCtl* sess_dev;         // device alias
  bool session_enabled, session_active;
  int session_inflight;         // commands published and not yet collected (0..2): seq - inflight + 1 .. seq, slot = seq % 2
  int session_op[2];            // op of the command last published in each slot
  long long session_step_no[2]; // train-step number (1-based) of the STEP command last published in each slot
  float session_loss[2];        // loss of train step T, kept at [T % 2] once its answer has been collected ...
  long long session_loss_step[2];   // ... together with T
  bool session_launch_blocked;  // the last session launch call returned only after the kernel had left (profiler)
  bool session_no_lease;        // diagnostics (dqn_set_session(h, 2)): never retire the kernel early, rely on the re-send path
  unsigned long long session_seq;       // sequence number of the last command published
  float session_last_loss;
  double session_last_cmd;      // host clock (s) of the last command: the kernel leaves after ~30 ms of silence
  dqn::EpisodeCtl* ep;               // device: per-agent episode-loop state (episode.cu)
  struct HostEpisode { long long step_count; int training_start, train_frequency; bool configured, pending_train; };
  std::vector<HostEpisode> hep; // host mirror of what decides the train gate (it depends on counters only, never on data)
};

namespace dqn {

// session mode (api_session.cu)
int session_stop(dqn_handle* h);                                  // answers in flight collected, EXIT, stream drained
constexpr int kSessionUnavailable = 1;                            // session_prepare: launches are synchronous here, session mode switched off
int session_prepare(dqn_handle* h, int keep = 0);                 // at most `keep` (0 / 1) commands still in flight, a live kernel
void session_publish(dqn_handle* h, int op, int n);               // payload already written (slot (session_seq + 1) % 2) with stamp session_seq + 1
int session_collect(dqn_handle* h, uint32_t* payload_out, int keep = 0);   // wait until at most `keep` commands are in flight
// the train-step launch shared by dqn_train_step*, dqn_store_train_step and dqn_train_flagged (api.cu)
int train_common(dqn_handle* h, int b, int e, int K, const long long* idx_dev, dqn_debug_taps* taps, const InlineStore* ist = nullptr,
                 const EpisodeCtl* gate = nullptr);

inline int check_agent_raw(const dqn_handle* h, int agent) {
  if (!h) return fail(DQN_E_INVALID, "handle is NULL");
  if (agent < 0 || agent >= h->cfg.n_agents) return fail(DQN_E_INVALID, "agent index out of range");
  return DQN_OK;
}
// Every entry point that is not served by a running session first ends it (the resident kernel owns the agent's
// parameters in shared memory and occupies the stream): the state is written back and the stream drained.
inline int check_agent(dqn_handle* h, int agent) {
  if (int rc = check_agent_raw(h, agent)) return rc;
  return h->session_active ? session_stop(h) : DQN_OK;
}
inline int check_range(dqn_handle* h, int b, int e) {
  if (!h) return fail(DQN_E_INVALID, "handle is NULL");
  if (b < 0 || e > h->cfg.n_agents || b >= e) return fail(DQN_E_INVALID, "agent range out of bounds or empty");
  return h->session_active ? session_stop(h) : DQN_OK;
}
inline long long size_of(const dqn_handle* h, int agent) {
  const long long c = h->hctl[agent].ring_counter;
  return c < h->dims.N ? c : h->dims.N;
}

}  // namespace dqn
