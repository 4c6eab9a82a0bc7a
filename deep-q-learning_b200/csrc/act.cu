// act.cu -- greedy action selection and hard target sync.
//
// act:  compute_action (General/QLearning/q_learning_functions.py:67-73) as called by Agent._policy
//       (q_agent.py:139) and Agent.evaluate (:229): argmax over the flattened [1,A] output of the
//       dueling MLP (LunarLander/dddqn.py:24-31), first max wins.  One warp per (agent, state):
//       lane = hidden unit, activations broadcast by shuffles, head reduced by shuffles.
// sync: Agent._update_target_model (q_agent.py:143-144): theta^- := theta (exact copy).
#include "common.cuh"
#include "kernels.h"

namespace dqn {

__global__ void __launch_bounds__(128)
dqn_act_kernel(const float* __restrict__ params, Dims d, int agent_begin, int n_sel,
               const float* __restrict__ states, int* __restrict__ actions, float* __restrict__ q_out) {
  const int lane = threadIdx.x & 31;
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (item >= n_sel) return;
  const int D = d.D, A = d.A;
  // packed layout (common.cuh): [W1;b1] | [W2;b2] rows of 68 | head 65 x 8 (col 0 = V, 1..A = advantage)
  const float* W1 = params + (size_t)(agent_begin + item) * 4 * d.PK;
  const float* b1 = W1 + D * kH1;
  const float* W2 = W1 + packed_w2(D);
  const float* b2 = W2 + kH1 * kW2Stride;
  const float* Wh = W1 + packed_head(D);
  const float* bh = Wh + kH2 * kHeadCols;
  const float* x = states + (size_t)item * D;

  float z = b1[lane];
  for (int k = 0; k < D; ++k) z = fmaf(x[k], W1[k * kH1 + lane], z);
  const float h1 = fmaxf(z, 0.f);
  float z0 = b2[lane], z1 = b2[lane + 32];
#pragma unroll 8
  for (int k = 0; k < kH1; ++k) {
    const float hk = __shfl_sync(0xffffffffu, h1, k);
    z0 = fmaf(hk, W2[k * kW2Stride + lane], z0);
    z1 = fmaf(hk, W2[k * kW2Stride + lane + 32], z1);
  }
  const float h20 = fmaxf(z0, 0.f), h21 = fmaxf(z1, 0.f);
  float head[1 + kMaxA];
  for (int c = 0; c <= kMaxA; ++c)      // packed head columns > A are zero
    head[c] = h20 * Wh[lane * kHeadCols + c] + h21 * Wh[(lane + 32) * kHeadCols + c];
#pragma unroll
  for (int c = 0; c <= kMaxA; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) head[c] += __shfl_xor_sync(0xffffffffu, head[c], o);
  }
  if (lane == 0) {
    const float val = head[0] + bh[0];
    float msum = 0.f;
    for (int j = 0; j < A; ++j) { head[1 + j] += bh[1 + j]; msum += head[1 + j]; }
    const float mean = msum / (float)A;
    int best = 0; float bq = 0.f;
    for (int j = 0; j < A; ++j) {
      const float q = val + head[1 + j] - mean;
      if (q_out) q_out[(size_t)item * A + j] = q;
      if (j == 0 || q > bq) { bq = q; best = j; }
    }
    actions[item] = best;
  }
}

__global__ void __launch_bounds__(256)
dqn_sync_target_kernel(float* __restrict__ params, int PK, int agent_begin) {
  float4* base = reinterpret_cast<float4*>(params + (size_t)(agent_begin + blockIdx.x) * 4 * PK);
  const int n4 = PK >> 2;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) base[n4 + i] = base[i];
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
