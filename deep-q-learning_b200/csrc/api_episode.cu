// api_episode.cu -- C ABI of the device-side episode loop (episode.cu): configure / policy / observe / gated train.
#include "handle.h"

using namespace dqn;

extern "C" {

// ---------------------------------------------------------------------------------------------------------------
// episode-loop control on the device (episode.cu): Agent._policy / one iteration of Agent._run_episode, batched
// ---------------------------------------------------------------------------------------------------------------
DQN_API int dqn_episode_configure(dqn_handle* h, int32_t agent_begin, int32_t agent_end, const dqn_episode_config* cfgs, int32_t reset_counters) {
  if (int rc = check_range(h, agent_begin, agent_end)) return rc;
  if (!cfgs) return fail(DQN_E_INVALID, "dqn_episode_configure: cfgs is NULL");
  const int n = agent_end - agent_begin;
  if ((size_t)n * sizeof(EpisodeCtl) > kStageBytes) return fail(DQN_E_INVALID, "dqn_episode_configure: too many agents for one call");
  CU(cudaSetDevice(h->cfg.device));
  std::vector<EpisodeCtl> cur(n);
  CU(cudaMemcpyAsync(cur.data(), h->ep + agent_begin, (size_t)n * sizeof(EpisodeCtl), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < n; ++i) {
    const dqn_episode_config& c = cfgs[i];
    if (c.train_frequency < 1 || c.replace_frequency < 1 || c.max_steps < 1 || c.max_episodes < 1 || c.training_start < 0)
      return fail(DQN_E_INVALID, "dqn_episode_configure: train_frequency, replace_frequency, max_steps, max_episodes must be >= 1 and training_start >= 0");
    EpisodeCtl& e = cur[i];
    e.epsilon = c.epsilon; e.eps_decay = c.epsilon_decay_rate; e.min_eps = c.min_epsilon; e.reward_to_reach = c.reward_to_reach;
    e.max_episodes = c.max_episodes; e.max_steps = c.max_steps; e.training_start = c.training_start;
    e.train_frequency = c.train_frequency; e.replace_frequency = c.replace_frequency;
    dqn_handle::HostEpisode& m = h->hep[agent_begin + i];
    if (reset_counters || !m.configured) {        // a fresh Agent.training() call: its locals restart (q_agent.py:210-211, :172-173)
      e.step_count = 0; e.episode = 0; e.step_in_episode = 0; e.epi_reward = 0.0; e.finished = 0; e.train_flag = 0; e.sync_flag = 0;
      m.step_count = 0; m.pending_train = false;
      if (!m.configured) { e.window_len = 0; e.window_pos = 0; e.avg_reward = 0.0; e.last_epi_reward = 0.0; e.policy_calls = 0; }
    }
    m.training_start = c.training_start; m.train_frequency = c.train_frequency; m.configured = true;
  }
  CU(cudaMemcpyAsync(h->ep + agent_begin, cur.data(), (size_t)n * sizeof(EpisodeCtl), cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return DQN_OK;
}

namespace {
int check_configured(const dqn_handle* h, int b, int e, const char* who) {
  for (int ag = b; ag < e; ++ag)
    if (!h->hep[ag].configured) return fail(DQN_E_INVALID, std::string(who) + ": call dqn_episode_configure for these agents first");
  return DQN_OK;
}
}  // namespace

DQN_API int dqn_policy_batch(dqn_handle* h, int32_t agent_begin, int32_t agent_end, const float* states_dev, int32_t* actions_dev) {
  if (int rc = check_range(h, agent_begin, agent_end)) return rc;
  if (int rc = check_configured(h, agent_begin, agent_end, "dqn_policy_batch")) return rc;
  if (!states_dev || !actions_dev) return fail(DQN_E_INVALID, "dqn_policy_batch: NULL argument");
  CU(cudaSetDevice(h->cfg.device));
  CU(launch_policy(h->stream, h->params, h->dims, h->ep, agent_begin, agent_end - agent_begin, h->cfg.agent_id_base, h->cfg.seed,
                   states_dev, actions_dev));
  return DQN_OK;
}

DQN_API int dqn_observe_batch(dqn_handle* h, int32_t agent_begin, int32_t agent_end, const float* s_dev, const int32_t* a_dev,
                              const float* r_dev, const float* s2_dev, const uint8_t* done_dev, uint8_t* episode_end_dev) {
  if (int rc = check_range(h, agent_begin, agent_end)) return rc;
  if (int rc = check_configured(h, agent_begin, agent_end, "dqn_observe_batch")) return rc;
  if (!s_dev || !a_dev || !r_dev || !s2_dev || !done_dev || !episode_end_dev) return fail(DQN_E_INVALID, "dqn_observe_batch: NULL argument");
  CU(cudaSetDevice(h->cfg.device));
  CU(launch_observe(h->stream, h->rings, h->ctl, h->ep, h->dims, agent_begin, agent_end - agent_begin, s_dev, a_dev, r_dev, s2_dev,
                    done_dev, episode_end_dev));
  for (int ag = agent_begin; ag < agent_end; ++ag) {     // the gate is a function of counters only: mirror it without a read-back
    dqn_handle::HostEpisode& m = h->hep[ag];
    h->hctl[ag].ring_counter += 1;
    m.step_count += 1;
    m.pending_train = size_of(h, ag) >= m.training_start && m.step_count % m.train_frequency == 0;
  }
  return DQN_OK;
}

DQN_API int dqn_train_flagged(dqn_handle* h, int32_t agent_begin, int32_t agent_end) {
  if (int rc = check_range(h, agent_begin, agent_end)) return rc;
  if (int rc = check_configured(h, agent_begin, agent_end, "dqn_train_flagged")) return rc;
  CU(cudaSetDevice(h->cfg.device));
  bool any = false;
  for (int ag = agent_begin; ag < agent_end; ++ag) any = any || h->hep[ag].pending_train;
  if (any) if (int rc = train_common(h, agent_begin, agent_end, 1, nullptr, nullptr, nullptr, h->ep)) return rc;
  for (int ag = agent_begin; ag < agent_end; ++ag) h->hep[ag].pending_train = false;
  CU(launch_episode_post(h->stream, h->params, h->dims, h->ep, agent_begin, agent_end - agent_begin));
  return DQN_OK;
}

DQN_API int dqn_episode_get_state(dqn_handle* h, int32_t agent, dqn_episode_state* out) {
  if (int rc = check_agent(h, agent)) return rc;
  if (!out) return fail(DQN_E_INVALID, "dqn_episode_get_state: out is NULL");
  CU(cudaSetDevice(h->cfg.device));
  EpisodeCtl e;
  CU(cudaMemcpyAsync(&e, h->ep + agent, sizeof e, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  out->epsilon = e.epsilon; out->average_reward = e.avg_reward; out->last_episode_reward = e.last_epi_reward;
  out->episode_reward = e.epi_reward; out->step_count = e.step_count; out->policy_calls = e.policy_calls;
  out->episode = e.episode; out->step_in_episode = e.step_in_episode; out->window_len = e.window_len; out->finished = e.finished;
  return DQN_OK;
}

}  // extern "C"
