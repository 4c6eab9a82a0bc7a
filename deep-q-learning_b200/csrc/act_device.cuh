// act_device.cuh -- compute_action (General/QLearning/q_learning_functions.py:67-73) for one state by one warp:
// the dueling MLP forward of LunarLander/dddqn.py:24-31 (lane = hidden unit, activations broadcast by shuffles, head
// reduced by shuffles) and the first-max argmax over the flattened [1,A] output.  Shared by the greedy act kernel
// (act.cu) and the epsilon-greedy policy kernel (episode.cu).
#pragma once
#include "common.cuh"

namespace dqn {

// `W` = one agent's online parameters in the packed layout (common.cuh); returns the greedy action in every lane;
// lane 0 optionally writes the A Q-values.
__device__ __forceinline__ int warp_greedy_action(const float* __restrict__ W, int D, int A, const float* __restrict__ x,
                                                  float* __restrict__ q_out) {
  const int lane = threadIdx.x & 31;
  const float* W1 = W;
  const float* b1 = W1 + D * kH1;
  const float* W2 = W1 + packed_w2(D);
  const float* b2 = W2 + kH1 * kW2Stride;
  const float* Wh = W1 + packed_head(D);
  const float* bh = Wh + kH2 * kHeadCols;

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
  // every lane holds the same sums (xor butterfly): finish redundantly, no broadcast needed
  const float val = head[0] + bh[0];
  float msum = 0.f;
  for (int j = 0; j < A; ++j) { head[1 + j] += bh[1 + j]; msum += head[1 + j]; }
  const float mean = msum / (float)A;
  int best = 0; float bq = 0.f;
  for (int j = 0; j < A; ++j) {
    const float q = val + head[1 + j] - mean;
    if (q_out && lane == 0) q_out[j] = q;
    if (j == 0 || q > bq) { bq = q; best = j; }
  }
  return best;
}

}  // namespace dqn
