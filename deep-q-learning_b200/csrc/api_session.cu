// api_session.cu -- host side of session mode (dqn_set_session): one resident launch of the cluster train-step kernel
// (train_cluster.cu, "serve" loop) is fed commands through a block of mapped pinned host memory (SessionCtl, kernels.h).
#include "handle.h"

using namespace dqn;

// ---------------------------------------------------------------------------------------------------------------
// session mode: one resident launch of the cluster kernel serves the agent's per-env-step calls
// ---------------------------------------------------------------------------------------------------------------
namespace {
double host_now() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
void cpu_relax() {
#if defined(__x86_64__)
  __builtin_ia32_pause();
#endif
}

int session_launch(dqn_handle* h, unsigned long long first_seq) {
  TrainArgs ta;
  memset(&ta, 0, sizeof ta);
  ta.params = h->params; ta.ctl = h->ctl; ta.rings = h->rings; ta.loss_ring = h->loss_ring; ta.loss_mailbox = h->mailbox_dev;
  ta.dims = h->dims; ta.seed = h->cfg.seed; ta.agent_begin = 0; ta.agent_id_base = h->cfg.agent_id_base; ta.n_sel = 1; ta.K = 0;
  ta.sess = h->sess_dev; ta.sess_first_seq = first_seq;
  const double t0 = host_now();
  CU(launch_train_cluster(h->stream, ta, nullptr));
  h->session_active = true;
  h->session_last_cmd = host_now();
  // A launch call that returns only after the kernel's whole idle time-out means launches are synchronous here (a
  // profiler serialising kernels): the resident kernel can never be sent a command.
  h->session_launch_blocked = h->session_last_cmd - t0 > 0.020;
  return DQN_OK;
}

}  // namespace

namespace dqn {
// Wait for the answer to the OLDEST command in flight; a command that a timed-out kernel never saw is served by a fresh
// launch (the kernel starts at that sequence number; a younger command published meanwhile sits in the other slot).
static int session_collect_oldest(dqn_handle* h, uint32_t* payload_out) {
  const unsigned long long seq = h->session_seq - (unsigned long long)h->session_inflight + 1;
  const int slot = (int)(seq & 1);
  const uint32_t want = (uint32_t)seq;
  for (unsigned long spin = 1;; ++spin) {
    const unsigned long long v = h->sess->response[slot];
    if ((uint32_t)(v >> 32) == want) {
      if (h->session_op[slot] == kOpStep) {
        const uint32_t bits = (uint32_t)v;
        const long long T = h->session_step_no[slot];
        memcpy(&h->session_loss[T & 1], &bits, 4);
        h->session_loss_step[T & 1] = T;
        h->session_last_loss = h->session_loss[T & 1];
      }
      if (payload_out) *payload_out = (uint32_t)v;
      h->session_inflight -= 1;
      return DQN_OK;
    }
    if ((spin & 0x3fff) == 0) {
      const cudaError_t e = cudaStreamQuery(h->stream);
      if (e == cudaSuccess) {                       // the kernel has left (idle time-out) ...
        if ((uint32_t)(h->sess->response[slot] >> 32) == want) continue;   // ... after answering
        if (int rc = session_launch(h, seq)) return rc;                    // ... without seeing the command: a fresh launch serves it
      } else if (e != cudaErrorNotReady) {
        h->session_active = false; h->session_inflight = 0;
        return fail(DQN_E_CUDA, std::string("session kernel failed: ") + cudaGetErrorString(e));
      }
    }
    cpu_relax();
  }
}

// wait until at most `keep` commands are in flight; payload_out = the answer of the last one collected
int session_collect(dqn_handle* h, uint32_t* payload_out, int keep) {
  while (h->session_inflight > keep)
    if (int rc = session_collect_oldest(h, payload_out)) return rc;
  return DQN_OK;
}

// Make the session ready for the next command: at most `keep` (0 or 1) earlier commands still in flight -- the slot of
// session_seq + 1 is free -- and a live kernel.  The caller then writes the payload stamped with session_seq + 1 and publishes.
int session_prepare(dqn_handle* h, int keep) {
  if (int rc = session_collect(h, nullptr, keep)) return rc;
  if (h->session_active && !h->session_no_lease && host_now() - h->session_last_cmd > 0.010) {
    // the kernel leaves by itself after ~30 ms of silence; past 10 ms do not race it: retire it and start a fresh one
    if (int rc = session_stop(h)) return rc;
  }
  if (!h->session_active) {
    if (int rc = session_launch(h, h->session_seq + 1)) return rc;
    if (h->session_launch_blocked) {           // fall back to one launch per call for the rest of this handle's life
      h->session_active = false;
      h->session_enabled = false;
      CU(cudaStreamSynchronize(h->stream));
      return kSessionUnavailable;
    }
  }
  return DQN_OK;
}
void session_publish(dqn_handle* h, int op, int n) {
  h->session_seq += 1;
  const int slot = (int)(h->session_seq & 1);
  h->session_op[slot] = op;
  __sync_synchronize();                                      // payload before the doorbell
  h->sess->doorbell[slot] = (h->session_seq << 16) | ((unsigned long long)op << 8) | (unsigned long long)n;
  h->session_inflight += 1;
  h->session_last_cmd = host_now();
}

int session_stop(dqn_handle* h) {
  if (!h->session_active) return DQN_OK;
  CU(cudaSetDevice(h->cfg.device));
  if (int rc = session_collect(h, nullptr, 0)) return rc;    // (train-step answers land in session_last_loss)
  h->session_seq += 1;
  __sync_synchronize();
  h->sess->doorbell[h->session_seq & 1] = (h->session_seq << 16) | ((unsigned long long)kOpExit << 8);
  h->session_active = false;                                 // (a kernel that already timed out never reads the EXIT; harmless)
  CU(cudaStreamSynchronize(h->stream));
  return DQN_OK;
}
}  // namespace dqn

extern "C" {

DQN_API int dqn_set_session(dqn_handle* h, int32_t enable) {
  if (!h) return fail(DQN_E_INVALID, "handle is NULL");
  CU(cudaSetDevice(h->cfg.device));
  if (!enable) {
    if (int rc = session_stop(h)) return rc;
    h->session_enabled = false;
    return DQN_OK;
  }
  if (h->cfg.n_agents != 1) return fail(DQN_E_INVALID, "dqn_set_session: session mode serves a single-agent handle");
  if (!h->sess) {
    CU(cudaHostAlloc((void**)&h->sess, sizeof(SessionCtl), cudaHostAllocMapped));
    memset((void*)h->sess, 0, sizeof(SessionCtl));
    CU(cudaHostGetDevicePointer((void**)&h->sess_dev, (void*)h->sess, 0));
  }
  h->session_enabled = true;
  h->session_no_lease = enable == 2;
  return DQN_OK;
}


}  // extern "C"
