// episode.cu -- the reference's per-env-step control flow around the train step, batched over agents on the device.
//
//   policy   Agent._policy (General/QLearning/q_agent.py:137-141): epsilon < uniform(0,1) ? greedy argmax : randint(0, A).
//            The reference draws from Python's `random` / numpy's global RNG, both unseeded; here the two
//            draws of policy call c of agent g come from Philox4x32-10 with counter (0, c_lo, c_hi, g) and the handle's
//            seed with the high key word xor kPolicyTag (a stream disjoint from the minibatch sampler's).
//   observe  one iteration of Agent._run_episode (q_agent.py:174-203) after env.step: forced done at max_steps (:179),
//            replay_buffer.add (:182), reward accumulation (:184), the train gate (:186), and at an episode end the
//            hard-sync cadence (:192), epsilon decay (:120-121, :202), the 50-episode reward window (:123-126, :203)
//            and the stop criterion of Agent.training (:211, :219).
//   post     after the (gated) train step: theta^- := theta for agents whose episode just ended on a sync episode
//            (q_agent.py:143-144, :192-193), then the per-step flags are cleared.
// All per-agent state is an EpisodeCtl in HBM; nothing returns to the host between env steps.
#include "act_device.cuh"
#include "common.cuh"
#include "kernels.h"

namespace dqn {

__global__ void __launch_bounds__(128)
dqn_policy_kernel(const float* __restrict__ params, Dims d, EpisodeCtl* __restrict__ ep, int agent_begin, int n_sel,
                  int agent_id_base, unsigned long long seed, const float* __restrict__ states, int* __restrict__ actions) {
  const int lane = threadIdx.x & 31;
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (item >= n_sel) return;
  const int agent = agent_begin + item;
  EpisodeCtl& e = ep[agent];
  const long long calls = e.policy_calls;
  uint32_t o[4];
  philox4x32_10(0u, (uint32_t)calls, (uint32_t)((unsigned long long)calls >> 32), (uint32_t)(agent_id_base + agent),
                (uint32_t)seed, (uint32_t)(seed >> 32) ^ kPolicyTag, o);
  // random.uniform(0, 1) == random.random(): 53 random bits / 2^53 (CPython genrand_res53 construction)
  const double u = ((double)(o[0] >> 5) * 67108864.0 + (double)(o[1] >> 6)) * (1.0 / 9007199254740992.0);
  int act;
  if (e.epsilon < u) act = warp_greedy_action(params + (size_t)agent * 4 * d.PK, d.D, d.A, states + (size_t)item * d.D, nullptr);
  else act = (int)(((unsigned long long)o[2] * (unsigned long long)d.A) >> 32);            // numpy.random.randint(0, A)
  __syncwarp();
  if (lane == 0) {
    actions[item] = act;
    e.policy_calls = calls + 1;
  }
}

__global__ void __launch_bounds__(128)
dqn_observe_kernel(uint32_t* __restrict__ rings, AgentCtl* __restrict__ ctl, EpisodeCtl* __restrict__ ep, Dims d, int agent_begin,
                   int n_sel, const float* __restrict__ s, const int* __restrict__ a, const float* __restrict__ r,
                   const float* __restrict__ s2, const uint8_t* __restrict__ done, uint8_t* __restrict__ episode_end) {
  const int lane = threadIdx.x & 31;
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (item >= n_sel) return;
  const int agent = agent_begin + item;
  EpisodeCtl& e = ep[agent];
  const int D = d.D;
  const int step = e.step_in_episode + 1;                    // `for step in range(1, ...)`, q_agent.py:174
  bool dn = done[item] != 0;
  if (step == e.max_steps) dn = true;                        // q_agent.py:179-180
  // ---- replay_buffer.add(state[0], action, reward, observation[0], done), q_agent.py:182 ----
  const long long rc = ctl[agent].ring_counter;
  uint32_t* rec = rings + ((size_t)agent * d.N + (size_t)(rc % d.N)) * d.recw;
  const int act = a[item];
  for (int w = lane; w < d.recw; w += 32) {
    uint32_t v = 0u;
    if (w < D) v = __float_as_uint(s[(size_t)item * D + w]);
    else if (w < 2 * D) v = __float_as_uint(s2[(size_t)item * D + (w - D)]);
    else if (w == 2 * D) v = (uint32_t)act;
    else if (w == 2 * D + 1) v = act < 0 ? 0xffffffffu : 0u;  // int64 action, sign-extended
    else if (w == 2 * D + 2) v = __float_as_uint(r[item]);
    else if (w == 2 * D + 3) v = dn ? 1u : 0u;
    rec[w] = v;
  }
  if (lane != 0) return;
  ctl[agent].ring_counter = rc + 1;
  e.step_count += 1;                                          // q_agent.py:175
  e.epi_reward += (double)r[item];                            // q_agent.py:184
  const long long size = rc + 1 < d.N ? rc + 1 : d.N;
  e.train_flag = (size >= e.training_start && e.step_count % e.train_frequency == 0) ? 1 : 0;   // q_agent.py:186
  // the episode ends on done (q_agent.py:189) or when the step loop runs out -- bounded by max_episodes, sic (:174)
  const bool ended = dn || step == e.max_episodes;
  if (ended) {
    if (e.episode % e.replace_frequency == 0) e.sync_flag = 1;                      // q_agent.py:192-193 (applied by post)
    e.epsilon = fmax(e.epsilon * e.eps_decay, e.min_eps);                           // q_agent.py:120-121
    if (e.window_len < kRewardWindow) {                                             // q_agent.py:123-126
      e.window[(e.window_pos + e.window_len) % kRewardWindow] = e.epi_reward;
      e.window_len += 1;
    } else {
      e.window[e.window_pos] = e.epi_reward;
      e.window_pos = (e.window_pos + 1) % kRewardWindow;
    }
    double sum = 0.0;
    for (int i = 0; i < e.window_len; ++i) sum += e.window[(e.window_pos + i) % kRewardWindow];   // oldest -> newest
    e.avg_reward = sum / (double)e.window_len;
    e.last_epi_reward = e.epi_reward;
    e.epi_reward = 0.0;
    e.step_in_episode = 0;
    e.episode += 1;
    if (e.avg_reward > e.reward_to_reach || e.episode >= e.max_episodes) e.finished = 1;          // q_agent.py:219, :211
  } else {
    e.step_in_episode = step;
  }
  episode_end[item] = ended ? 1 : 0;
}

__global__ void __launch_bounds__(128)
dqn_episode_post_kernel(float* __restrict__ params, int PK, EpisodeCtl* __restrict__ ep, int agent_begin) {
  const int agent = agent_begin + blockIdx.x;
  EpisodeCtl& e = ep[agent];
  if (e.sync_flag) {                                          // Agent._update_target_model, q_agent.py:143-144
    float4* base = reinterpret_cast<float4*>(params + (size_t)agent * 4 * PK);
    const int n4 = PK >> 2;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) base[n4 + i] = base[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) { e.sync_flag = 0; e.train_flag = 0; }
}

cudaError_t launch_policy(cudaStream_t st, const float* params, const Dims& d, EpisodeCtl* ep, int agent_begin, int n_sel,
                          int agent_id_base, unsigned long long seed, const float* states, int* actions) {
  if (n_sel <= 0) return cudaSuccess;
  dqn_policy_kernel<<<(n_sel + 3) / 4, 128, 0, st>>>(params, d, ep, agent_begin, n_sel, agent_id_base, seed, states, actions);
  return cudaGetLastError();
}

cudaError_t launch_observe(cudaStream_t st, uint32_t* rings, AgentCtl* ctl, EpisodeCtl* ep, const Dims& d, int agent_begin, int n_sel,
                           const float* s, const int* a, const float* r, const float* s2, const uint8_t* done, uint8_t* episode_end) {
  if (n_sel <= 0) return cudaSuccess;
  dqn_observe_kernel<<<(n_sel + 3) / 4, 128, 0, st>>>(rings, ctl, ep, d, agent_begin, n_sel, s, a, r, s2, done, episode_end);
  return cudaGetLastError();
}

cudaError_t launch_episode_post(cudaStream_t st, float* params, const Dims& d, EpisodeCtl* ep, int agent_begin, int n_sel) {
  if (n_sel <= 0) return cudaSuccess;
  dqn_episode_post_kernel<<<n_sel, 128, 0, st>>>(params, d.PK, ep, agent_begin);
  return cudaGetLastError();
}

}  // namespace dqn
