// act.cu -- greedy action selection and hard target sync.
//
// act:  compute_action (General/QLearning/q_learning_functions.py:67-73) as called by Agent._policy
//       (q_agent.py:139) and Agent.evaluate (:229): argmax over the flattened [1,A] output of the
//       dueling MLP (LunarLander/dddqn.py:24-31), first max wins.  One warp per (agent, state):
//       lane = hidden unit, activations broadcast by shuffles, head reduced by shuffles.
// sync: Agent._update_target_model (q_agent.py:143-144): theta^- := theta (exact copy).
#include "common.cuh"
#include "kernels.h"
#include "act_device.cuh"

namespace dqn {

__global__ void __launch_bounds__(128)
dqn_act_kernel(const float* __restrict__ params, Dims d, int agent_begin, int n_sel,
               const float* __restrict__ states, int* __restrict__ actions, float* __restrict__ q_out) {
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (item >= n_sel) return;
  const int best = warp_greedy_action(params + (size_t)(agent_begin + item) * 4 * d.PK, d.D, d.A, states + (size_t)item * d.D,
                                      q_out ? q_out + (size_t)item * d.A : nullptr);
  if ((threadIdx.x & 31) == 0) actions[item] = best;
}

__global__ void __launch_bounds__(256)
dqn_sync_target_kernel(float* __restrict__ params, int PK, int agent_begin) {
  float4* base = reinterpret_cast<float4*>(params + (size_t)(agent_begin + blockIdx.x) * 4 * PK);
  const int n4 = PK >> 2;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) base[n4 + i] = base[i];
}

// Polyak / soft target update (not in the reference; optax.incremental_update semantics):
//   theta^- := tau * theta + (1 - tau) * theta^-   with every product and the sum rounded to fp32 separately (no FMA
// contraction), so that the NumPy oracle reproduces it bit for bit.  tau = 1 is the hard sync.
__global__ void __launch_bounds__(256)
dqn_polyak_target_kernel(float* __restrict__ params, int PK, int agent_begin, float tau) {
  float4* base = reinterpret_cast<float4*>(params + (size_t)(agent_begin + blockIdx.x) * 4 * PK);
  const int n4 = PK >> 2;
  const float omt = __fsub_rn(1.0f, tau);
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    const float4 w = base[i], t = base[n4 + i];
    base[n4 + i] = make_float4(__fadd_rn(__fmul_rn(tau, w.x), __fmul_rn(omt, t.x)), __fadd_rn(__fmul_rn(tau, w.y), __fmul_rn(omt, t.y)),
                               __fadd_rn(__fmul_rn(tau, w.z), __fmul_rn(omt, t.z)), __fadd_rn(__fmul_rn(tau, w.w), __fmul_rn(omt, t.w)));
  }
}

cudaError_t launch_polyak_target(cudaStream_t st, float* params, const Dims& d, int agent_begin, int n_sel, float tau) {
  if (n_sel <= 0) return cudaSuccess;
  dqn_polyak_target_kernel<<<n_sel, 256, 0, st>>>(params, d.PK, agent_begin, tau);
  return cudaGetLastError();
}

cudaError_t launch_act(cudaStream_t st, const float* params, const Dims& d, int agent_begin, int n_sel,
                       const float* states, int* actions_out, float* q_out) {
  if (n_sel <= 0) return cudaSuccess;
  dqn_act_kernel<<<(n_sel + 3) / 4, 128, 0, st>>>(params, d, agent_begin, n_sel, states, actions_out, q_out);
  return cudaGetLastError();
}

cudaError_t launch_sync_target(cudaStream_t st, float* params, const Dims& d, int agent_begin, int n_sel) {
  if (n_sel <= 0) return cudaSuccess;
  dqn_sync_target_kernel<<<n_sel, 256, 0, st>>>(params, d.PK, agent_begin);
  return cudaGetLastError();
}

}  // namespace dqn
