/*
 * dqn_b200.h -- C ABI of libdqn_b200.so: the B200 (sm_100a) dueling double-DQN train-step +
 * replay path.  Plain pointers and sizes only; no torch / C++ types cross this boundary.
 *
 * The reference (hal9000universe/deep-q-learning) has no FFI layer: its seam is the set of Python
 * callables and the buffer object that `Agent` stores (General/QLearning/q_agent.py:93-95,110-113).
 * Each entry point below names the reference interface it replaces.  All state (online/target
 * parameters, Adam moments, the replay ring, per-agent counters) is device-resident and owned by an
 * opaque handle; a handle holds `n_agents >= 1` independent agents (1 = the reference's `Agent`,
 * >1 = a population of the hyper-parameter sweep) that never communicate.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on failure (DQN_E_*); the message is available from
 *     dqn_last_error() (thread-local, valid until the next failing call on the thread);
 *   - there is NO CPU fallback: without a usable sm_100a device every compute call fails;
 *   - all work is enqueued in order on the handle's stream; calls that return data to host
 *     pointers synchronise that stream before returning, the others do not;
 *   - a handle is not thread-safe (neither is the reference's Agent).
 *
 * Flat parameter layout (P = dqn_param_count floats), the reference checkpoint's leaves in order
 * (Test/lunar_lander/params.pickle; haiku Linear stores w as [in,out], row-major):
 *     W1[D][H1] b1[H1] | W2[H1][H2] b2[H2] | Wv[H2][1] bv[1] | Wa[H2][A] ba[A]
 *     'model/~/linear'   'model/~/linear_1'  'model/~/linear_2'  'model/~/linear_3'
 */
#ifndef DQN_B200_H
#define DQN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DQN_ABI_VERSION 1

#if defined(__GNUC__)
#define DQN_API __attribute__((visibility("default")))
#else
#define DQN_API
#endif

enum {
  DQN_OK = 0,
  DQN_E_INVALID = -1,  /* bad argument / shape / state */
  DQN_E_CUDA = -2,     /* CUDA runtime error (message carries cudaGetErrorString) */
  DQN_E_ARCH = -3,     /* no sm_100a device / library not built for this device */
  DQN_E_NOMEM = -4
};

enum { DQN_OPT_ADAM = 0, DQN_OPT_ADAMW = 1 };
enum { DQN_PARAMS_ONLINE = 0, DQN_PARAMS_TARGET = 1 };
/* Loss of the train step.  HUBER (delta = 1, summed over actions, mean over the batch) is the reference's
 * (q_learning_functions.py:36) and the default; L2 = optax.l2_loss = 0.5 e^2 in its place is an extension the reference
 * does not have: same reduction, gradient e / B instead of clip(e, -1, 1) / B. */
enum { DQN_LOSS_HUBER = 0, DQN_LOSS_L2 = 1 };
/* Train-step kernel: one CTA per agent (throughput form, used for populations) -- CTA = fp32 FFMA throughout, CTA_TC = the
 * layer-2 products (80 % of the flops) on the tensor cores as error-compensated 3xTF32 (tcgen05) -- or one agent spread
 * over a 4-CTA thread-block cluster (latency form, used for a single agent).  AUTO = cluster while 4*agents <= SMs, else
 * CTA_TC.  All compute the same update; summation orders differ (results agree to fp32 round-off). */
enum { DQN_STEP_AUTO = 0, DQN_STEP_CTA = 1, DQN_STEP_CLUSTER = 2, DQN_STEP_CTA_TC = 3 };

/* Construction-time configuration == the part of `Agent.__init__` (q_agent.py:61-118) that shapes
 * the hot path.  `network` becomes (obs_dim, hidden1, hidden2, num_actions) -- the dueling MLP of
 * LunarLander/dddqn.py:17-22; `optimizer` becomes (opt_kind, lr, b1, b2, eps, eps_root,
 * weight_decay) -- optax.adamw(2e-4) in Test/lunar_lander.py:48, optax.adam(1e-4) in
 * Test/lunar_lander_hyper_params.py:41.  Per-agent values can be changed later (dqn_set_hparams). */
typedef struct dqn_config {
  int32_t struct_size;   /* sizeof(dqn_config), for ABI checking */
  int32_t device;        /* CUDA device ordinal */
  int32_t n_agents;      /* independent agents in this handle (>= 1) */
  int32_t obs_dim;       /* D, 1..16  (9 in Test/lunar_lander.py:40, 8 in the synthetic configs) */
  int32_t num_actions;   /* A, 2..7   (env.action_space.n, Test/lunar_lander.py:42) */
  int32_t hidden1;       /* 32  (dddqn.py:19) */
  int32_t hidden2;       /* 64  (dddqn.py:20) */
  int32_t batch_size;    /* B, 1..DQN_MAX_BATCH (q_agent.py:153) */
  int64_t buffer_size;   /* N slots per agent (replay_buffer.py:25) */
  float gamma;           /* q_agent.py:111 */
  int32_t opt_kind;      /* DQN_OPT_ADAM / DQN_OPT_ADAMW */
  float lr, b1, b2, eps, eps_root, weight_decay;
  uint64_t seed;         /* Philox key for minibatch indices */
  int32_t agent_id_base; /* global id of local agent 0: the Philox counter uses (agent_id_base + agent), so a
                          *   sharded population draws the same indices as the unsharded one */
  int32_t step_kernel;   /* DQN_STEP_AUTO / _CTA / _CLUSTER / _CTA_TC: which train-step kernel (see the enum) */
  void* stream;          /* cudaStream_t to enqueue on (NULL = default stream) */
  void* arena;           /* optional caller-allocated device memory (e.g. a torch tensor's  */
  uint64_t arena_bytes;  /*   data_ptr()); NULL => the library allocates dqn_arena_bytes() itself */
} dqn_config;

#define DQN_MAX_BATCH 1024
#define DQN_MAX_OBS_DIM 16
#define DQN_MAX_ACTIONS 7

typedef struct dqn_handle dqn_handle;

/* Per-agent hyper-parameters that the sweep injects (ParamAgent.inject,
 * hyperparameter_optimization.py:76-91) plus the optimiser constants.  Fields < 0 / NaN keep the
 * current value. */
typedef struct dqn_hparams {
  float gamma;
  int32_t batch_size;
  float lr, b1, b2, eps, eps_root, weight_decay;
} dqn_hparams;

/* Optional debug taps of ONE train step (K must be 1): every intermediate the parity tests compare
 * with the oracle.  All pointers are HOST pointers or NULL; shapes use the agent's B, A, P. */
typedef struct dqn_debug_taps {
  int64_t* indices;   /* [B]    sampled slots */
  float* q;           /* [B*A]  Q(theta, s)            q_learning_functions.py:52 */
  float* next_q;      /* [B*A]  Q(theta, s')           :53 */
  float* next_q_tm;   /* [B*A]  Q(theta^-, s')         :54 */
  int32_t* max_actions; /* [B]  argmax_a Q(theta, s')  :55 */
  float* targets;     /* [B*A]  q + tv*onehot(a)       :58-59 */
  float* loss;        /* [1]    mean_i sum_j huber     :36 */
  float* grads;       /* [P]    d loss / d theta, flat layout   :23 */
} dqn_debug_taps;

DQN_API int dqn_abi_version(void);
DQN_API const char* dqn_last_error(void);

/* Bytes of device memory a handle with this configuration needs (for caller-side allocation). */
DQN_API int dqn_arena_bytes(const dqn_config* cfg, uint64_t* bytes_out);

/* Agent.__init__ (q_agent.py:87-113): allocates / carves state; parameters start at zero, target =
 * online, Adam count = 0, ring empty.  ReplayBuffer.__init__ (replay_buffer.py:20-32). */
DQN_API int dqn_create(const dqn_config* cfg, dqn_handle** out);
DQN_API int dqn_destroy(dqn_handle* h);
DQN_API int dqn_param_count(const dqn_handle* h, int32_t* p_out);
DQN_API int dqn_synchronize(dqn_handle* h);

/* `self._params` / `self._target_params` (q_agent.py:88,91) in the flat layout, host pointers. */
DQN_API int dqn_set_params(dqn_handle* h, int32_t agent, int32_t which, const float* host_flat, int32_t n);
DQN_API int dqn_get_params(dqn_handle* h, int32_t agent, int32_t which, float* host_flat, int32_t n);
/* `self._opt_state` = ScaleByAdamState(count, mu, nu) (Test/lunar_lander/opt_state.pickle). */
DQN_API int dqn_set_opt_state(dqn_handle* h, int32_t agent, int32_t count, const float* mu, const float* nu, int32_t n);
DQN_API int dqn_get_opt_state(dqn_handle* h, int32_t agent, int32_t* count, float* mu, float* nu, int32_t n);
DQN_API int dqn_set_hparams(dqn_handle* h, int32_t agent, const dqn_hparams* hp);
DQN_API int dqn_get_hparams(dqn_handle* h, int32_t agent, dqn_hparams* hp);
DQN_API int dqn_set_step_kernel(dqn_handle* h, int32_t step_kernel);   /* DQN_STEP_* */
/* Resume support: what the reference's checkpoint (General/Base/utils.py:21-29) omits.  ring_counter = ReplayBuffer._counter
 * (replay_buffer.py:64), train_steps = number of Agent._step() calls so far (the position of the Philox index stream),
 * adam_count / pb1 / pb2 = optax count and the carried b1**count, b2**count.  A handle restored with dqn_set_params (both
 * networks), dqn_set_opt_state, dqn_store (ring contents in slot order) and dqn_set_counters continues bit for bit. */
typedef struct dqn_counters {
  int64_t ring_counter, train_steps;
  int32_t adam_count, reserved;
  double pb1, pb2;
} dqn_counters;
DQN_API int dqn_get_counters(dqn_handle* h, int32_t agent, dqn_counters* out);
DQN_API int dqn_set_counters(dqn_handle* h, int32_t agent, const dqn_counters* in);

/* ReplayBuffer.add (replay_buffer.py:58-65), vectorised: equivalent to n scalar add() calls in order
 * into agent's ring (slot (counter+i) % N).  Host pointers: s/s2 f32[n*D], a i64[n], r f32[n],
 * done u8[n] (numpy bool).  Asynchronous w.r.t. the host when the pointers are pinned. */
DQN_API int dqn_store(dqn_handle* h, int32_t agent, int64_t n, const float* s, const int64_t* a,
              const float* r, const float* s2, const uint8_t* done);
/* Same with DEVICE pointers (no host copy). */
DQN_API int dqn_store_device(dqn_handle* h, int32_t agent, int64_t n, const float* s, const int64_t* a,
                     const float* r, const float* s2, const uint8_t* done);
/* ReplayBuffer.size / ._counter (replay_buffer.py:34-36,64-65). */
DQN_API int dqn_buffer_state(dqn_handle* h, int32_t agent, int64_t* size_out, int64_t* counter_out);
/* ReplayBuffer.states/.actions/.rewards/.observations/.dones (replay_buffer.py:38-56): the whole
 * ring de-interleaved into the reference's five arrays (f32[N*D], i64[N], f32[N], f32[N*D], u8[N]). */
DQN_API int dqn_buffer_export(dqn_handle* h, int32_t agent, float* s, int64_t* a, float* r, float* s2, uint8_t* done);

/* The index draw of sample_batch (replay_buffer.py:77): B Philox4x32-10 indices in [0,size) for
 * train step `step` of `agent` (the indices dqn_train_step uses when given none). Host pointer. */
DQN_API int dqn_sample_indices(dqn_handle* h, int32_t agent, int64_t step, int32_t batch, int64_t* idx_out);
/* sample_batch (replay_buffer.py:68-85): gather `batch` transitions by `idx` (host i64[batch], or
 * NULL = Philox indices of `step`) into the reference's five output arrays (host pointers). */
DQN_API int dqn_sample_batch(dqn_handle* h, int32_t agent, const int64_t* idx, int64_t step, int32_t batch,
                     float* s, int64_t* a, float* r, float* s2, uint8_t* done);
/* Device-resident variant used for throughput measurement: indices (device i64[batch] or NULL) ->
 * device output arrays.  Enqueue only. */
DQN_API int dqn_sample_batch_device(dqn_handle* h, int32_t agent, const int64_t* idx_dev, int64_t step, int32_t batch,
                            float* s, int64_t* a, float* r, float* s2, uint8_t* done);

/* Agent._step (q_agent.py:146-169), K times back to back for every agent in [agent_begin,
 * agent_end): sample -> preprocessing -> compute_q_targets -> train_step, fused in one persistent
 * kernel.  `idx` is NULL (Philox indices) or a HOST array i64[n_sel*K*B] of explicit slots
 * (agent-major, then step).  `taps` non-NULL requires K == 1 and a single agent.  Enqueue only
 * (unless taps are requested). */
DQN_API int dqn_train_step(dqn_handle* h, int32_t agent_begin, int32_t agent_end, int32_t K,
                   const int64_t* idx, dqn_debug_taps* taps);
/* The reference's inner loop between two train steps (q_agent.py:182-187) as ONE launch: n ReplayBuffer.add calls
 * (host arrays as in dqn_store) followed by K Agent._step()s of `agent`.  For n <= DQN_MAX_INLINE_STORE with the
 * cluster kernel the transitions travel in the kernel's parameter buffer (no H2D copy, no store launch); otherwise
 * this is dqn_store + dqn_train_step.  `loss_out` (host, optional): wait for the launch -- polling the zero-copy
 * loss mailbox, no stream synchronisation -- and return the loss of the last step. */
#define DQN_MAX_INLINE_STORE 16
DQN_API int dqn_store_train_step(dqn_handle* h, int32_t agent, int64_t n, const float* s, const int64_t* a, const float* r,
                                 const float* s2, const uint8_t* done, int32_t K, float* loss_out);
/* Session mode (single-agent handles): ONE resident launch of the cluster kernel keeps theta / theta^- / Adam state in
 * shared memory and serves the agent's per-env-step calls from commands in mapped host memory -- no launch and no copy
 * per call: dqn_store_train_step (n <= 16, K = 1), dqn_act, dqn_sync_target(h, 0, 1), dqn_get_losses(n <= 1),
 * dqn_get_loss_lagged.  This is
 * the reference's env loop (q_agent.py:174-189: _policy -> add -> _step) without a kernel launch on its path.  The
 * kernel is started on demand, leaves by itself after ~30 ms without a command (state written back), and any other
 * entry point ends the session first.  While it is resident the handle's stream is occupied. */
DQN_API int dqn_set_session(dqn_handle* h, int32_t enable);   /* 0 off, 1 on, 2 on without the host-side lease (diagnostics:
                                                                   a command sent to a timed-out kernel is re-sent to a fresh one) */
/* Same, explicit indices already on the device (i64[n_sel*K*B]) -- no host copy, enqueue only. */
DQN_API int dqn_train_step_device_idx(dqn_handle* h, int32_t agent_begin, int32_t agent_end, int32_t K,
                              const int64_t* idx_dev);
/* Loss of the most recent `n` train steps of `agent` (oldest first), host pointer. */
DQN_API int dqn_get_losses(dqn_handle* h, int32_t agent, int32_t n, float* loss_out, int64_t* train_steps_out);
/* Loss of the train step `lag` (0 or 1) steps before the most recent one; lag 0 == dqn_get_losses(h, agent, 1, ...).
 * Waits for THAT step only: in session mode the most recent step may still be in flight (the resident kernel takes two
 * commands, one per slot), so an env loop (q_agent.py:174-189) can publish step i + 1 and then read step i's loss --
 * every loss is read, one step behind -- and the device never waits for the host between two steps. */
DQN_API int dqn_get_loss_lagged(dqn_handle* h, int32_t agent, int32_t lag, float* loss_out);

/* Agent._update_target_model (q_agent.py:143-144): theta^- := theta for agents in the range. */
DQN_API int dqn_sync_target(dqn_handle* h, int32_t agent_begin, int32_t agent_end);
/* Soft (Polyak) target update, an extension: the reference only has the hard copy above (q_agent.py:143-144).
 * theta^- := tau * theta + (1 - tau) * theta^- (optax.incremental_update: two fp32 products and one fp32 sum per
 * element, no fused multiply-add).  tau in [0,1]; tau = 1 reproduces dqn_sync_target bit for bit. */
DQN_API int dqn_polyak_target(dqn_handle* h, int32_t agent_begin, int32_t agent_end, float tau);
/* Select the loss (DQN_LOSS_*) of the agents in the range; takes effect with the next train step. */
DQN_API int dqn_set_loss_kind(dqn_handle* h, int32_t agent_begin, int32_t agent_end, int32_t kind);

/* compute_action (q_learning_functions.py:67-73) as used by Agent._policy (q_agent.py:139):
 * greedy argmax_a Q(theta, state) for ONE state f32[D] (host pointer).  Synchronises. */
DQN_API int dqn_act(dqn_handle* h, int32_t agent, const float* state, int32_t* action_out);
/* Batched: agents [agent_begin, agent_end), one state each (host f32[n_sel*D] -> i32[n_sel]). */
DQN_API int dqn_act_batch(dqn_handle* h, int32_t agent_begin, int32_t agent_end, const float* states, int32_t* actions_out);

/* ------------------------------------------------------------------------------------------------
 * Episode-loop control on the device: the reference's per-env-step work around the train step,
 * batched over the agents [agent_begin, agent_end) with no host round trip.  All array arguments are DEVICE pointers,
 * one entry per agent of the range; every call only enqueues.
 *   dqn_policy_batch   Agent._policy (q_agent.py:137-141): epsilon < uniform(0,1) ? argmax_a Q(theta, state) : randint(0, A);
 *                      the two draws come from the handle's Philox stream (the reference's host RNGs are unseeded).
 *   dqn_observe_batch  one iteration of Agent._run_episode after env.step (q_agent.py:174-203): done forced at max_steps,
 *                      ReplayBuffer.add, reward accumulation, the train gate (size >= training_start and step_count %
 *                      train_frequency == 0), and at an episode end: hard-sync cadence (episode % replace_frequency == 0),
 *                      epsilon decay, the 50-episode reward window, the stop test of Agent.training (q_agent.py:211,:219).
 *                      episode_end[i] = 1 tells the caller to reset environment i.
 *   dqn_train_flagged  Agent._step() for the agents whose gate opened in the last observe (ONE launch, the others' CTAs
 *                      exit at once), then theta^- := theta for agents whose episode ended on a sync episode.
 * A `finished` agent is only reported (dqn_episode_get_state); the caller retires it (the reference returns from training()).
 * ---------------------------------------------------------------------------------------------- */
typedef struct dqn_episode_config {          /* Agent.__init__ kwargs of the same names, q_agent.py:61-86 */
  double epsilon, epsilon_decay_rate, min_epsilon, reward_to_reach;
  int32_t max_episodes, max_steps, training_start, train_frequency, replace_frequency, reserved;
} dqn_episode_config;
typedef struct dqn_episode_state {
  double epsilon, average_reward, last_episode_reward, episode_reward;
  int64_t step_count, policy_calls;
  int32_t episode, step_in_episode, window_len, finished;
} dqn_episode_state;
/* cfgs: HOST array, one per agent of the range.  reset_counters != 0 restarts episode / step counters (a new
 * Agent.training() call); the reward window and the policy RNG position persist like the reference's attributes. */
DQN_API int dqn_episode_configure(dqn_handle* h, int32_t agent_begin, int32_t agent_end, const dqn_episode_config* cfgs, int32_t reset_counters);
DQN_API int dqn_policy_batch(dqn_handle* h, int32_t agent_begin, int32_t agent_end, const float* states_dev, int32_t* actions_dev);
DQN_API int dqn_observe_batch(dqn_handle* h, int32_t agent_begin, int32_t agent_end, const float* s_dev, const int32_t* a_dev,
                              const float* r_dev, const float* s2_dev, const uint8_t* done_dev, uint8_t* episode_end_dev);
DQN_API int dqn_train_flagged(dqn_handle* h, int32_t agent_begin, int32_t agent_end);
DQN_API int dqn_episode_get_state(dqn_handle* h, int32_t agent, dqn_episode_state* out);   /* synchronises */

/* ------------------------------------------------------------------------------------------------
 * Large-batch data-parallel mode (BASELINE configs[3]: batch 65536, hidden 1024x1024).  Same update as
 * Agent._step (q_agent.py:146-169) for ONE agent whose minibatch is split over `world` ranks: every rank
 * holds replicas of theta / theta^- / Adam state and of the replay ring, runs forward+backward on its
 * slice [rank*B_local, (rank+1)*B_local) of the global minibatch with gradients pre-scaled by 1/B_global,
 * the caller all-reduces (sum) the P+1 floats at dqn_lb_grads() -- P gradients + the loss -- over NCCL,
 * then dqn_lb_apply() runs the identical Adam update on every rank.  world = 1 needs no collective.
 * hidden1 / hidden2 / batch_local must be multiples of 128 (hidden >= 128).
 * ---------------------------------------------------------------------------------------------- */
typedef struct dqn_lb_config {
  int32_t struct_size, device;
  int32_t obs_dim, num_actions, hidden1, hidden2;
  int32_t batch_local;    /* rows of the minibatch this rank processes */
  int32_t gemm_mode;      /* 0: fp32 FFMA2 tiles (exact fp32)   1: tcgen05 tensor cores, 3xTF32 split */
  int64_t buffer_size;
  float gamma;
  int32_t opt_kind;
  float lr, b1, b2, eps, eps_root, weight_decay;
  uint64_t seed;
  int32_t rank, world;
  void* stream;
  void* arena;
  uint64_t arena_bytes;
} dqn_lb_config;

typedef struct dqn_lb_handle dqn_lb_handle;
enum { DQN_LB_READ_Q = 0, DQN_LB_READ_TARGETS = 1, DQN_LB_READ_MAX_ACTIONS = 2, DQN_LB_READ_GRADS = 3, DQN_LB_READ_INDICES = 4,
       DQN_LB_READ_H1 = 5, DQN_LB_READ_H2 = 6 };   /* hidden activations of the (theta, s) rows: [B][hidden1], [B][hidden2] */

DQN_API int dqn_lb_arena_bytes(const dqn_lb_config* cfg, uint64_t* bytes_out);
DQN_API int dqn_lb_create(const dqn_lb_config* cfg, dqn_lb_handle** out);
DQN_API int dqn_lb_destroy(dqn_lb_handle* h);
DQN_API int dqn_lb_param_count(const dqn_lb_handle* h, int32_t* p_out);
DQN_API int dqn_lb_set_params(dqn_lb_handle* h, int32_t which, const float* host_flat, int32_t n);
DQN_API int dqn_lb_get_params(dqn_lb_handle* h, int32_t which, float* host_flat, int32_t n);
DQN_API int dqn_lb_set_opt_state(dqn_lb_handle* h, int32_t count, const float* mu, const float* nu, int32_t n);
DQN_API int dqn_lb_get_opt_state(dqn_lb_handle* h, int32_t* count, float* mu, float* nu, int32_t n);
/* ReplayBuffer.add, vectorised (host / device pointers) -- same ring as the small-batch path. */
DQN_API int dqn_lb_store(dqn_lb_handle* h, int64_t n, const float* s, const int64_t* a, const float* r, const float* s2, const uint8_t* done);
DQN_API int dqn_lb_store_device(dqn_lb_handle* h, int64_t n, const float* s, const int64_t* a, const float* r, const float* s2, const uint8_t* done);
DQN_API int dqn_lb_buffer_state(dqn_lb_handle* h, int64_t* size_out, int64_t* counter_out);
/* sample (Philox slice of the global draw, or `idx` = host i64[batch_local]) -> targets -> loss -> backward.
 * Leaves the local gradient (already divided by the global batch) and loss share at dqn_lb_grads(). */
DQN_API int dqn_lb_forward_backward(dqn_lb_handle* h, const int64_t* idx, int32_t debug);
DQN_API int dqn_lb_grads(dqn_lb_handle* h, void** dev_ptr_out, int64_t* count_out);
/* The gradient sum as ONE kernel over NVLink peer memory instead of a library collective (world = 2, 4 or 8 GPUs of a
 * node): comm_init moves the gradient into a cudaMalloc'ed window and returns its 64-byte cudaIpcMemHandle_t; the caller
 * gathers the handles of all ranks (any transport) and passes them to comm_connect -- or, for ranks living in ONE process,
 * the raw window pointers.  dqn_lb_allreduce then enqueues: flag barrier -> rank r sums slice r of every rank's window in
 * rank order -> stores it into every window -> flag barrier.  Deterministic; replicas receive bit-identical sums.  Every
 * rank must enqueue it once per step; a rank that never arrives raises an error (reported by dqn_lb_get_loss), not a hang. */
DQN_API int dqn_lb_comm_init(dqn_lb_handle* h, void* ipc_handle_out /* 64 bytes or NULL */, void** window_out /* or NULL */);
DQN_API int dqn_lb_comm_connect(dqn_lb_handle* h, const void* ipc_handles /* world * 64 bytes, or NULL */,
                                void* const* peer_windows /* world pointers, or NULL */);
DQN_API int dqn_lb_allreduce(dqn_lb_handle* h);
/* optimizer.update + apply_updates (q_learning_functions.py:24-25) with whatever is in the gradient buffer. */
DQN_API int dqn_lb_apply(dqn_lb_handle* h);
/* One whole step = dqn_lb_forward_backward + gradient exchange over the peer-memory windows + dqn_lb_apply, with the
 * exchange of the W2 gradient (the bulk of the vector) issued on a second stream as soon as it is final, so that it
 * overlaps the remaining backward GEMM (q_learning_functions.py:36 makes the shard gradients additive).
 * world = 1: no exchange.  Replicas stay bit-identical (one reducer per element, fixed rank order). */
DQN_API int dqn_lb_train_step(dqn_lb_handle* h, const int64_t* idx_or_null, int32_t debug);
DQN_API int dqn_lb_sync_target(dqn_lb_handle* h);
/* Extensions, as for the small-batch handle: soft target update and the L2 loss. */
DQN_API int dqn_lb_polyak_target(dqn_lb_handle* h, float tau);
DQN_API int dqn_lb_set_loss_kind(dqn_lb_handle* h, int32_t kind);
DQN_API int dqn_lb_get_loss(dqn_lb_handle* h, float* loss_out);
/* parity taps after dqn_lb_forward_backward(debug = 1): `what` = DQN_LB_READ_*; Q is [3*B][A] (q | next_q | next_q_tm). */
DQN_API int dqn_lb_debug_read(dqn_lb_handle* h, int32_t what, void* host_out, uint64_t nbytes);
DQN_API int dqn_lb_synchronize(dqn_lb_handle* h);

/* ------------------------------------------------------------------------------------------------
 * Prioritized replay (BASELINE configs[4]): sum-tree proportional sampler + priority update.  The reference
 * has no prioritized replay (its sampler is uniform, General/Base/replay_buffer.py:68-85), so this is a
 * self-specified extension with its own oracle (oracle/per_oracle.py); indices it returns can be fed to
 * dqn_train_step / dqn_sample_batch as explicit indices.  All array arguments are DEVICE pointers unless
 * the name says host.
 * ---------------------------------------------------------------------------------------------- */
typedef struct dqn_per_handle dqn_per_handle;
DQN_API int dqn_per_arena_bytes(int64_t capacity, uint64_t* bytes_out);
DQN_API int dqn_per_create(int32_t device, int64_t capacity, float alpha, float eps, uint64_t seed, void* stream,
                           void* arena, uint64_t arena_bytes, dqn_per_handle** out);
DQN_API int dqn_per_destroy(dqn_per_handle* h);
/* leaves[idx[i]] = prio[i] (is_td = 0) or (|prio[i]| + eps)^alpha (is_td = 1); parents recomputed. */
DQN_API int dqn_per_update(dqn_per_handle* h, const int64_t* idx_dev, const float* val_dev, int32_t n, int32_t is_td);
DQN_API int dqn_per_update_host(dqn_per_handle* h, const int64_t* idx, const float* val, int32_t n, int32_t is_td);
/* set leaves [0, n) from a device array and rebuild the whole tree (bulk initialisation). */
DQN_API int dqn_per_fill(dqn_per_handle* h, const float* prio_dev, int64_t n);
/* B stratified proportional samples for `step` -> indices (i64[B]) and their priorities (f32[B]). */
DQN_API int dqn_per_sample(dqn_per_handle* h, int64_t step, int32_t batch, int64_t* idx_dev, float* prio_dev);
DQN_API int dqn_per_sample_host(dqn_per_handle* h, int64_t step, int32_t batch, int64_t* idx, float* prio);
DQN_API int dqn_per_total(dqn_per_handle* h, float* total_out);
/* debug: copy tree nodes [first, first + n) to host (node 1 = root, leaves start at dqn_per_leaf_base). */
DQN_API int dqn_per_read_nodes(dqn_per_handle* h, int64_t first, int64_t n, float* host_out);
DQN_API int dqn_per_leaf_base(dqn_per_handle* h, int64_t* leaf_base_out);

#ifdef __cplusplus
}
#endif
#endif /* DQN_B200_H */
