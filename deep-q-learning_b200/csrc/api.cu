// api.cu -- the C ABI of include/dqn_b200.h: handle management, host<->device staging, launches.
// No compute happens on the host; every entry point either moves bytes or launches a kernel.
// (Session mode lives in api_session.cu, the device-side episode loop in api_episode.cu.)
#include "handle.h"
#include <algorithm>
#include <vector>

using namespace dqn;

namespace {
thread_local std::string g_err;
}  // namespace
namespace dqn {
int host_fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
void set_last_error(const char* msg) { g_err = msg; }
}  // namespace dqn

namespace {

TapsOff taps_layout(const Dims& d) {
  TapsOff t;
  size_t o = 0;
  t.idx = o; o += align_up((size_t)DQN_MAX_BATCH * 8, 256);
  t.q = o; o += align_up((size_t)DQN_MAX_BATCH * d.A * 4, 256);
  t.nq = o; o += align_up((size_t)DQN_MAX_BATCH * d.A * 4, 256);
  t.nqt = o; o += align_up((size_t)DQN_MAX_BATCH * d.A * 4, 256);
  t.tgt = o; o += align_up((size_t)DQN_MAX_BATCH * d.A * 4, 256);
  t.maxa = o; o += align_up((size_t)DQN_MAX_BATCH * 4, 256);
  t.loss = o; o += 256;
  t.grads = o; o += align_up((size_t)d.PK * 4, 256);
  t.bytes = o;
  return t;
}

Carve carve(const Dims& d, int n_agents) {
  Carve c;
  size_t o = 0;
  c.params = o; o += align_up((size_t)n_agents * 4 * d.PK * 4, 256);
  c.ctl = o; o += align_up((size_t)n_agents * sizeof(AgentCtl), 256);
  c.ep = o; o += align_up((size_t)n_agents * sizeof(EpisodeCtl), 256);
  c.rings = o; o += align_up((size_t)n_agents * (size_t)d.N * d.recw * 4, 256);
  c.loss = o; o += align_up((size_t)n_agents * kLossCap * 4, 256);
  c.stage = o; o += kStageBytes;
  c.taps = o; o += taps_layout(d).bytes;
  c.total = o;
  return c;
}

int validate(const dqn_config* cfg, Dims* d) {
  if (!cfg) return fail(DQN_E_INVALID, "config is NULL");
  if (cfg->struct_size != (int32_t)sizeof(dqn_config))
    return fail(DQN_E_INVALID, "dqn_config.struct_size mismatch (ABI version skew)");
  if (cfg->n_agents < 1) return fail(DQN_E_INVALID, "n_agents must be >= 1");
  if (cfg->obs_dim < 1 || cfg->obs_dim > DQN_MAX_OBS_DIM) return fail(DQN_E_INVALID, "obs_dim must be in [1,16]");
  if (cfg->num_actions < 2 || cfg->num_actions > DQN_MAX_ACTIONS) return fail(DQN_E_INVALID, "num_actions must be in [2,7]");
  if (cfg->hidden1 != kH1 || cfg->hidden2 != kH2)
    return fail(DQN_E_INVALID, "the fused path implements the reference network only: hidden = (32, 64) (LunarLander/dddqn.py:19-20)");
  if (cfg->batch_size < 1 || cfg->batch_size > DQN_MAX_BATCH) return fail(DQN_E_INVALID, "batch_size must be in [1,1024]");
  if (cfg->buffer_size < 1 || cfg->buffer_size >= (1ll << 31)) return fail(DQN_E_INVALID, "buffer_size must be in [1, 2^31)");
  if (cfg->opt_kind != DQN_OPT_ADAM && cfg->opt_kind != DQN_OPT_ADAMW) return fail(DQN_E_INVALID, "opt_kind must be adam or adamw");
  d->D = cfg->obs_dim;
  d->A = cfg->num_actions;
  d->P = flat_param_count(d->D, d->A);
  d->PK = packed_count(d->D);
  d->recw = record_words_host(d->D);
  d->N = cfg->buffer_size;
  return DQN_OK;
}

}  // namespace

namespace {

uint32_t* ring_of(dqn_handle* h, int agent) { return h->rings + (size_t)agent * (size_t)h->dims.N * h->dims.recw; }
struct SoA {   // carve of the staging buffer into the five transition arrays for n transitions
  float* s; long long* a; float* r; float* s2; uint8_t* done;
};
SoA stage_soa(dqn_handle* h, long long n) {
  const int D = h->dims.D;
  uint8_t* p = h->stage;
  SoA o;
  o.a = (long long*)p; p += align_up((size_t)n * 8, 16);
  o.s = (float*)p; p += align_up((size_t)n * D * 4, 16);
  o.s2 = (float*)p; p += align_up((size_t)n * D * 4, 16);
  o.r = (float*)p; p += align_up((size_t)n * 4, 16);
  o.done = p;
  return o;
}
long long stage_capacity(const dqn_handle* h) { return (long long)((kStageBytes - 128) / (8 * h->dims.D + 13)); }

}  // namespace

extern "C" {

DQN_API int dqn_abi_version(void) { return DQN_ABI_VERSION; }
DQN_API const char* dqn_last_error(void) { return g_err.c_str(); }

DQN_API int dqn_arena_bytes(const dqn_config* cfg, uint64_t* bytes_out) {
  Dims d;
  if (int rc = validate(cfg, &d)) return rc;
  if (!bytes_out) return fail(DQN_E_INVALID, "bytes_out is NULL");
  *bytes_out = carve(d, cfg->n_agents).total;
  return DQN_OK;
}

DQN_API int dqn_create(const dqn_config* cfg, dqn_handle** out) {
  Dims d;
  if (int rc = validate(cfg, &d)) return rc;
  if (!out) return fail(DQN_E_INVALID, "out is NULL");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(DQN_E_ARCH, "no CUDA device visible: libdqn_b200 has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(DQN_E_INVALID, "device ordinal out of range");
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) {
    char b[256];
    snprintf(b, sizeof b, "device %d is sm_%d%d; libdqn_b200 is built for sm_100a (B200) only", cfg->device, prop.major, prop.minor);
    return fail(DQN_E_ARCH, b);
  }
  CU(cudaSetDevice(cfg->device));
  dqn_handle* h = new dqn_handle();
  h->cfg = *cfg;
  h->dims = d;
  h->stream = (cudaStream_t)cfg->stream;
  h->cv = carve(d, cfg->n_agents);
  h->to = taps_layout(d);
  if (cfg->arena) {
    if (cfg->arena_bytes < h->cv.total || ((uintptr_t)cfg->arena & 255)) {
      delete h;
      return fail(DQN_E_INVALID, "arena too small (see dqn_arena_bytes) or not 256-byte aligned");
    }
    h->arena = (uint8_t*)cfg->arena;
    h->own_arena = false;
  } else {
    cudaError_t e = cudaMalloc((void**)&h->arena, h->cv.total);
    if (e != cudaSuccess) { delete h; return fail(DQN_E_NOMEM, std::string("cudaMalloc(arena) failed: ") + cudaGetErrorString(e)); }
    h->own_arena = true;
  }
  h->params = (float*)(h->arena + h->cv.params);
  h->ctl = (AgentCtl*)(h->arena + h->cv.ctl);
  h->ep = (EpisodeCtl*)(h->arena + h->cv.ep);
  h->hep.assign(cfg->n_agents, dqn_handle::HostEpisode{0, 0, 1, false, false});
  h->rings = (uint32_t*)(h->arena + h->cv.rings);
  h->loss_ring = (float*)(h->arena + h->cv.loss);
  h->stage = h->arena + h->cv.stage;
  h->taps = h->arena + h->cv.taps;
  h->pinned = nullptr;
  cudaError_t e = cudaMallocHost((void**)&h->pinned, kPinnedBytes);
  if (e != cudaSuccess) { if (h->own_arena) cudaFree(h->arena); delete h; return fail(DQN_E_NOMEM, "cudaMallocHost failed"); }
  h->bounce = h->pinned + kBounceOff;
  h->mailbox = nullptr; h->mailbox_dev = nullptr;
  // + [n_agents] int: the launch order of the population kernel (costliest agents first), read once per CTA
  e = cudaHostAlloc((void**)&h->mailbox, (sizeof(unsigned long long) + sizeof(int)) * cfg->n_agents, cudaHostAllocMapped);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&h->mailbox_dev, (void*)h->mailbox, 0);
  h->order = (int*)(h->mailbox + cfg->n_agents); h->order_dev = (const int*)(h->mailbox_dev + cfg->n_agents);
  h->order_begin = h->order_end = -1;
  if (e != cudaSuccess) { cudaFreeHost(h->pinned); if (h->own_arena) cudaFree(h->arena); delete h; return fail(DQN_E_NOMEM, "cudaHostAlloc(mapped) failed"); }
  memset((void*)h->mailbox, 0, sizeof(unsigned long long) * cfg->n_agents);
  h->sess = nullptr; h->sess_dev = nullptr;
  h->session_enabled = h->session_active = h->session_no_lease = h->session_launch_blocked = false;
  h->session_inflight = 0; h->session_op[0] = h->session_op[1] = 0; h->session_loss[0] = h->session_loss[1] = 0.f;
  h->session_step_no[0] = h->session_step_no[1] = h->session_loss_step[0] = h->session_loss_step[1] = -1;
  h->session_seq = 0; h->session_last_loss = 0.f; h->session_last_cmd = 0.0;
  h->slot_next = 0;
  for (int i = 0; i < kSlots; ++i) cudaEventCreateWithFlags(&h->slot_ev[i], cudaEventDisableTiming);
  // zero parameters / moments / rings / losses (ReplayBuffer.__init__ zero-fills, replay_buffer.py:26-30)
  e = cudaMemsetAsync(h->arena, 0, h->cv.stage, h->stream);
  if (e == cudaSuccess) e = train_fused_prepare(d);
  if (e == cudaSuccess) e = train_tc_prepare(d);
  if (e == cudaSuccess) e = train_cluster_prepare(d);
  h->step_kernel = cfg->step_kernel;
  {
    int sms = 0;
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device);
    h->sm_count = sms;
  }
  AgentCtl c;
  memset(&c, 0, sizeof c);
  c.gamma = cfg->gamma; c.lr = cfg->lr; c.b1 = cfg->b1; c.b2 = cfg->b2; c.eps = cfg->eps;
  c.eps_root = cfg->eps_root; c.wd = cfg->opt_kind == DQN_OPT_ADAMW ? cfg->weight_decay : 0.f;
  c.batch_size = cfg->batch_size;
  c.pb1 = 1.0; c.pb2 = 1.0;
  h->hctl.assign(cfg->n_agents, c);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(h->ctl, h->hctl.data(), sizeof(AgentCtl) * cfg->n_agents, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) {
    std::string m = std::string("dqn_create: device initialisation failed: ") + cudaGetErrorString(e);
    for (int i = 0; i < kSlots; ++i) cudaEventDestroy(h->slot_ev[i]);
    cudaFreeHost((void*)h->mailbox);
    cudaFreeHost(h->pinned);
    if (h->own_arena) cudaFree(h->arena);
    delete h;
    return fail(e == cudaErrorNoKernelImageForDevice || e == cudaErrorInvalidDeviceFunction ? DQN_E_ARCH : DQN_E_CUDA, m);
  }
  *out = h;
  return DQN_OK;
}

DQN_API int dqn_destroy(dqn_handle* h) {
  if (!h) return DQN_OK;
  cudaSetDevice(h->cfg.device);
  if (h->session_active) session_stop(h);
  cudaStreamSynchronize(h->stream);
  if (h->sess) cudaFreeHost((void*)h->sess);
  for (int i = 0; i < kSlots; ++i) cudaEventDestroy(h->slot_ev[i]);
  if (h->mailbox) cudaFreeHost((void*)h->mailbox);
  if (h->pinned) cudaFreeHost(h->pinned);
  if (h->own_arena) cudaFree(h->arena);
  delete h;
  return DQN_OK;
}

DQN_API int dqn_param_count(const dqn_handle* h, int32_t* p_out) {
  if (!h || !p_out) return fail(DQN_E_INVALID, "NULL argument");
  *p_out = h->dims.P;
  return DQN_OK;
}

DQN_API int dqn_synchronize(dqn_handle* h) {
  if (!h) return fail(DQN_E_INVALID, "handle is NULL");
  CU(cudaSetDevice(h->cfg.device));
  if (h->session_active) if (int rc = session_stop(h)) return rc;
  CU(cudaStreamSynchronize(h->stream));
  return DQN_OK;
}

namespace {
// flat (include/dqn_b200.h) <-> packed (common.cuh) through the pinned bounce buffer; `slot` picks one of the
// PK-float regions of the bounce so that several arrays can be in flight before one synchronisation
float* bounce_slot(dqn_handle* h, int slot) { return (float*)h->bounce + (size_t)slot * h->dims.PK; }
void pack_flat(const dqn_handle* h, const float* flat, float* packed) {
  const int PK = h->dims.PK, D = h->dims.D, A = h->dims.A;
  for (int p = 0; p < PK; ++p) { const int f = packed_to_flat(p, D, A); packed[p] = f >= 0 ? flat[f] : 0.f; }
}
void unpack_flat(const dqn_handle* h, const float* packed, float* flat) {
  const int PK = h->dims.PK, D = h->dims.D, A = h->dims.A;
  for (int p = 0; p < PK; ++p) { const int f = packed_to_flat(p, D, A); if (f >= 0) flat[f] = packed[p]; }
}
}  // namespace

namespace {
// AgentCtl.pb1 / pb2 = b1**count, b2**count (optax's bias-correction powers), re-seeded from the host mirror
int seed_decay_powers(dqn_handle* h, int agent) {
  AgentCtl& c = h->hctl[agent];
  c.pb1 = pow((double)c.b1, (double)c.adam_count);
  c.pb2 = pow((double)c.b2, (double)c.adam_count);
  CU(cudaMemcpyAsync(&h->ctl[agent].pb1, &c.pb1, 2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  return DQN_OK;
}
}  // namespace

DQN_API int dqn_set_params(dqn_handle* h, int32_t agent, int32_t which, const float* host_flat, int32_t n) {
  if (int rc = check_agent(h, agent)) return rc;
  if ((which != DQN_PARAMS_ONLINE && which != DQN_PARAMS_TARGET) || !host_flat || n != h->dims.P)
    return fail(DQN_E_INVALID, "dqn_set_params: bad `which`, NULL pointer or n != dqn_param_count");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaStreamSynchronize(h->stream));     // the bounce buffer may still be in flight
  const int PK = h->dims.PK;
  pack_flat(h, host_flat, bounce_slot(h, 0));
  float* dst = h->params + ((size_t)agent * 4 + which) * PK;
  CU(cudaMemcpyAsync(dst, bounce_slot(h, 0), (size_t)PK * 4, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return DQN_OK;
}

DQN_API int dqn_get_params(dqn_handle* h, int32_t agent, int32_t which, float* host_flat, int32_t n) {
  if (int rc = check_agent(h, agent)) return rc;
  if ((which != DQN_PARAMS_ONLINE && which != DQN_PARAMS_TARGET) || !host_flat || n != h->dims.P)
    return fail(DQN_E_INVALID, "dqn_get_params: bad `which`, NULL pointer or n != dqn_param_count");
  CU(cudaSetDevice(h->cfg.device));
  const int PK = h->dims.PK;
  const float* src = h->params + ((size_t)agent * 4 + which) * PK;
  CU(cudaMemcpyAsync(bounce_slot(h, 0), src, (size_t)PK * 4, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  unpack_flat(h, bounce_slot(h, 0), host_flat);
  return DQN_OK;
}

DQN_API int dqn_set_opt_state(dqn_handle* h, int32_t agent, int32_t count, const float* mu, const float* nu, int32_t n) {
  if (int rc = check_agent(h, agent)) return rc;
  if (!mu || !nu || n != h->dims.P || count < 0) return fail(DQN_E_INVALID, "dqn_set_opt_state: NULL pointer, n != dqn_param_count or count < 0");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaStreamSynchronize(h->stream));
  const int PK = h->dims.PK;
  float* base = h->params + (size_t)agent * 4 * PK;
  pack_flat(h, mu, bounce_slot(h, 0));
  pack_flat(h, nu, bounce_slot(h, 1));
  CU(cudaMemcpyAsync(base + 2 * PK, bounce_slot(h, 0), (size_t)2 * PK * 4, cudaMemcpyHostToDevice, h->stream));   // mu | nu are adjacent
  h->hctl[agent].adam_count = count;
  CU(cudaMemcpyAsync(&h->ctl[agent].adam_count, &h->hctl[agent].adam_count, 4, cudaMemcpyHostToDevice, h->stream));
  if (int rc = seed_decay_powers(h, agent)) return rc;
  CU(cudaStreamSynchronize(h->stream));
  return DQN_OK;
}

DQN_API int dqn_get_opt_state(dqn_handle* h, int32_t agent, int32_t* count, float* mu, float* nu, int32_t n) {
  if (int rc = check_agent(h, agent)) return rc;
  if (n != h->dims.P) return fail(DQN_E_INVALID, "dqn_get_opt_state: n != dqn_param_count");
  CU(cudaSetDevice(h->cfg.device));
  const int PK = h->dims.PK;
  const float* base = h->params + (size_t)agent * 4 * PK;
  CU(cudaMemcpyAsync(bounce_slot(h, 0), base + 2 * PK, (size_t)2 * PK * 4, cudaMemcpyDeviceToHost, h->stream));
  int32_t* pc = (int32_t*)bounce_slot(h, 2);
  CU(cudaMemcpyAsync(pc, &h->ctl[agent].adam_count, 4, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (mu) unpack_flat(h, bounce_slot(h, 0), mu);
  if (nu) unpack_flat(h, bounce_slot(h, 1), nu);
  if (count) *count = *pc;
  return DQN_OK;
}

DQN_API int dqn_set_hparams(dqn_handle* h, int32_t agent, const dqn_hparams* hp) {
  if (int rc = check_agent(h, agent)) return rc;
  if (!hp) return fail(DQN_E_INVALID, "hp is NULL");
  AgentCtl& c = h->hctl[agent];
  if (hp->batch_size > DQN_MAX_BATCH || hp->batch_size == 0) return fail(DQN_E_INVALID, "batch_size must be in [1,1024]");
  if (hp->gamma == hp->gamma && hp->gamma >= 0.f) c.gamma = hp->gamma;
  if (hp->batch_size > 0) { c.batch_size = hp->batch_size; h->order_begin = h->order_end = -1; }
  if (hp->lr == hp->lr && hp->lr >= 0.f) c.lr = hp->lr;
  if (hp->b1 == hp->b1 && hp->b1 >= 0.f) c.b1 = hp->b1;
  if (hp->b2 == hp->b2 && hp->b2 >= 0.f) c.b2 = hp->b2;
  if (hp->eps == hp->eps && hp->eps >= 0.f) c.eps = hp->eps;
  if (hp->eps_root == hp->eps_root && hp->eps_root >= 0.f) c.eps_root = hp->eps_root;
  if (hp->weight_decay == hp->weight_decay && hp->weight_decay >= 0.f) c.wd = hp->weight_decay;
  CU(cudaSetDevice(h->cfg.device));
  // only the hyper-parameter prefix: the counters behind it are owned by the device
  memcpy(h->bounce, &c, offsetof(AgentCtl, adam_count));
  CU(cudaMemcpyAsync(&h->ctl[agent], h->bounce, offsetof(AgentCtl, adam_count), cudaMemcpyHostToDevice, h->stream));
  if (int rc = seed_decay_powers(h, agent)) return rc;
  CU(cudaStreamSynchronize(h->stream));
  return DQN_OK;
}

DQN_API int dqn_get_hparams(dqn_handle* h, int32_t agent, dqn_hparams* hp) {
  if (int rc = check_agent(h, agent)) return rc;
  if (!hp) return fail(DQN_E_INVALID, "hp is NULL");
  const AgentCtl& c = h->hctl[agent];
  hp->gamma = c.gamma; hp->batch_size = c.batch_size; hp->lr = c.lr; hp->b1 = c.b1; hp->b2 = c.b2;
  hp->eps = c.eps; hp->eps_root = c.eps_root; hp->weight_decay = c.wd;
  return DQN_OK;
}

// Resume support: the counters the reference never saves -- ReplayBuffer._counter, the number of
// _step() calls (= Philox stream position) and the carried Adam decay powers -- so that a restored agent continues
// bit for bit.  Parameters / moments / ring contents go through dqn_set_params, dqn_set_opt_state, dqn_store.
DQN_API int dqn_get_counters(dqn_handle* h, int32_t agent, dqn_counters* out) {
  if (int rc = check_agent(h, agent)) return rc;
  if (!out) return fail(DQN_E_INVALID, "dqn_get_counters: out is NULL");
  CU(cudaSetDevice(h->cfg.device));
  AgentCtl c;
  CU(cudaMemcpyAsync(&c, &h->ctl[agent], sizeof c, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  out->ring_counter = c.ring_counter; out->train_steps = c.train_steps; out->adam_count = c.adam_count;
  out->reserved = 0; out->pb1 = c.pb1; out->pb2 = c.pb2;
  return DQN_OK;
}

DQN_API int dqn_set_counters(dqn_handle* h, int32_t agent, const dqn_counters* in) {
  if (int rc = check_agent(h, agent)) return rc;
  if (!in || in->ring_counter < 0 || in->train_steps < 0 || in->adam_count < 0)
    return fail(DQN_E_INVALID, "dqn_set_counters: NULL or negative counter");
  CU(cudaSetDevice(h->cfg.device));
  AgentCtl& c = h->hctl[agent];
  c.ring_counter = in->ring_counter; c.train_steps = in->train_steps; c.adam_count = in->adam_count;
  c.pb1 = in->pb1; c.pb2 = in->pb2;
  CU(cudaMemcpyAsync(&h->ctl[agent].adam_count, &c.adam_count, sizeof(AgentCtl) - offsetof(AgentCtl, adam_count),
                     cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return DQN_OK;
}

DQN_API int dqn_set_step_kernel(dqn_handle* h, int32_t step_kernel) {
  if (!h) return fail(DQN_E_INVALID, "handle is NULL");
  if (h->session_active) if (int rc = session_stop(h)) return rc;
  if (step_kernel < DQN_STEP_AUTO || step_kernel > DQN_STEP_CTA_TC) return fail(DQN_E_INVALID, "dqn_set_step_kernel: unknown kernel id");
  h->step_kernel = step_kernel;
  return DQN_OK;
}

DQN_API int dqn_store_device(dqn_handle* h, int32_t agent, int64_t n, const float* s, const int64_t* a, const float* r,
                     const float* s2, const uint8_t* done) {
  if (int rc = check_agent(h, agent)) return rc;
  if (n < 0 || (n > 0 && (!s || !a || !r || !s2 || !done))) return fail(DQN_E_INVALID, "dqn_store: negative n or NULL array");
  if (n == 0) return DQN_OK;
  CU(cudaSetDevice(h->cfg.device));
  const long long N = h->dims.N;
  long long counter = h->hctl[agent].ring_counter;
  long long skip = 0;
  if (n > N) {   // only the last N transitions survive n scalar add() calls; the counter still advances by n
    skip = n - N;
  }
  const int D = h->dims.D;
  CU(launch_replay_store(h->stream, ring_of(h, agent), h->dims, counter + skip, n - skip, s + skip * D,
                         (const long long*)a + skip, r + skip, s2 + skip * D, done + skip, &h->ctl[agent]));
  h->hctl[agent].ring_counter = counter + n;
  return DQN_OK;
}

DQN_API int dqn_store(dqn_handle* h, int32_t agent, int64_t n, const float* s, const int64_t* a, const float* r,
              const float* s2, const uint8_t* done) {
  if (int rc = check_agent(h, agent)) return rc;
  if (n < 0 || (n > 0 && (!s || !a || !r || !s2 || !done))) return fail(DQN_E_INVALID, "dqn_store: negative n or NULL array");
  CU(cudaSetDevice(h->cfg.device));
  const int D = h->dims.D;
  if (n > 0 && (size_t)n * (8 * D + 13) + 64 <= kSlotBytes) {
    // small store (the reference adds ONE transition per env step): pack the five arrays into a pinned
    // slot in the staging layout and move them with a single asynchronous copy
    const int slot = h->slot_next;
    h->slot_next = (slot + 1) % kSlots;
    CU(cudaEventSynchronize(h->slot_ev[slot]));
    uint8_t* p = h->pinned + (size_t)slot * kSlotBytes;
    SoA d = stage_soa(h, n);
    const size_t bytes = (size_t)(d.done - h->stage) + (size_t)n;
    memcpy(p + ((uint8_t*)d.a - h->stage), a, (size_t)n * 8);
    memcpy(p + ((uint8_t*)d.s - h->stage), s, (size_t)n * D * 4);
    memcpy(p + ((uint8_t*)d.s2 - h->stage), s2, (size_t)n * D * 4);
    memcpy(p + ((uint8_t*)d.r - h->stage), r, (size_t)n * 4);
    memcpy(p + (d.done - h->stage), done, (size_t)n);
    CU(cudaMemcpyAsync(h->stage, p, bytes, cudaMemcpyHostToDevice, h->stream));
    CU(cudaEventRecord(h->slot_ev[slot], h->stream));
    return dqn_store_device(h, agent, n, d.s, (const int64_t*)d.a, d.r, d.s2, d.done);
  }
  const long long cap = stage_capacity(h);
  for (long long off = 0; off < n; off += cap) {
    const long long m = n - off < cap ? n - off : cap;
    SoA d = stage_soa(h, m);
    CU(cudaMemcpyAsync(d.s, s + off * D, (size_t)m * D * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d.a, a + off, (size_t)m * 8, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d.r, r + off, (size_t)m * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d.s2, s2 + off * D, (size_t)m * D * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d.done, done + off, (size_t)m, cudaMemcpyHostToDevice, h->stream));
    if (int rc = dqn_store_device(h, agent, m, d.s, (const int64_t*)d.a, d.r, d.s2, d.done)) return rc;
  }
  return DQN_OK;
}

DQN_API int dqn_buffer_state(dqn_handle* h, int32_t agent, int64_t* size_out, int64_t* counter_out) {
  if (int rc = check_agent(h, agent)) return rc;
  if (size_out) *size_out = size_of(h, agent);
  if (counter_out) *counter_out = h->hctl[agent].ring_counter;
  return DQN_OK;
}

DQN_API int dqn_sample_batch_device(dqn_handle* h, int32_t agent, const int64_t* idx_dev, int64_t step, int32_t batch,
                            float* s, int64_t* a, float* r, float* s2, uint8_t* done) {
  if (int rc = check_agent(h, agent)) return rc;
  if (batch < 0 || !s || !a || !r || !s2 || !done) return fail(DQN_E_INVALID, "dqn_sample_batch: negative batch or NULL output");
  const long long size = size_of(h, agent);
  if (batch > 0 && size == 0) return fail(DQN_E_INVALID, "dqn_sample_batch: the replay ring is empty (randint(0, 0) raises in the reference)");
  CU(cudaSetDevice(h->cfg.device));
  CU(launch_replay_gather(h->stream, ring_of(h, agent), h->dims, idx_dev ? 0 : 1, (const long long*)idx_dev, h->cfg.seed, h->cfg.agent_id_base + agent,
                          step, size, batch, s, (long long*)a, r, s2, done));
  return DQN_OK;
}

namespace {
// gather `count` records (mode 0: staged explicit idx, 1: philox, 2: identity from `first`) to host arrays via staging
int gather_to_host(dqn_handle* h, int agent, int mode, const int64_t* idx_host, long long step, long long first,
                   long long count, float* s, int64_t* a, float* r, float* s2, uint8_t* done) {
  const int D = h->dims.D;
  const long long cap = (long long)((kStageBytes - 256) / (8 * D + 13 + 8));
  const long long size = size_of(h, agent);
  for (long long off = 0; off < count; off += cap) {
    const long long m = count - off < cap ? count - off : cap;
    SoA d = stage_soa(h, m);
    long long* didx = (long long*)(d.done + align_up((size_t)m, 16));
    const uint32_t* ring = ring_of(h, agent);
    if (mode == 0) {
      CU(cudaMemcpyAsync(didx, idx_host + off, (size_t)m * 8, cudaMemcpyHostToDevice, h->stream));
      CU(launch_replay_gather(h->stream, ring, h->dims, 0, didx, 0, agent, 0, size, m, d.s, d.a, d.r, d.s2, d.done));
    } else if (mode == 1) {
      // Philox slot index i is global within the step: draw then gather explicitly for chunks beyond the first
      CU(launch_philox_indices(h->stream, didx, (int)m, h->cfg.seed, h->cfg.agent_id_base + agent, step, size));
      if (off != 0) return fail(DQN_E_INVALID, "dqn_sample_batch: batch too large for one staging chunk");
      CU(launch_replay_gather(h->stream, ring, h->dims, 0, didx, 0, agent, 0, size, m, d.s, d.a, d.r, d.s2, d.done));
    } else {
      CU(launch_replay_gather(h->stream, ring + (size_t)(first + off) * h->dims.recw, h->dims, 2, nullptr, 0, agent, 0, size, m,
                              d.s, d.a, d.r, d.s2, d.done));
    }
    CU(cudaMemcpyAsync(s + off * D, d.s, (size_t)m * D * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(a + off, d.a, (size_t)m * 8, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(r + off, d.r, (size_t)m * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(s2 + off * D, d.s2, (size_t)m * D * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(done + off, d.done, (size_t)m, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  return DQN_OK;
}
}  // namespace

DQN_API int dqn_buffer_export(dqn_handle* h, int32_t agent, float* s, int64_t* a, float* r, float* s2, uint8_t* done) {
  if (int rc = check_agent(h, agent)) return rc;
  if (!s || !a || !r || !s2 || !done) return fail(DQN_E_INVALID, "dqn_buffer_export: NULL output");
  CU(cudaSetDevice(h->cfg.device));
  return gather_to_host(h, agent, 2, nullptr, 0, 0, h->dims.N, s, a, r, s2, done);
}

DQN_API int dqn_sample_indices(dqn_handle* h, int32_t agent, int64_t step, int32_t batch, int64_t* idx_out) {
  if (int rc = check_agent(h, agent)) return rc;
  if (batch < 0 || (size_t)batch * 8 > kStageBytes || !idx_out) return fail(DQN_E_INVALID, "dqn_sample_indices: bad batch or NULL output");
  const long long size = size_of(h, agent);
  if (batch > 0 && size == 0) return fail(DQN_E_INVALID, "dqn_sample_indices: the replay ring is empty");
  if (batch == 0) return DQN_OK;
  CU(cudaSetDevice(h->cfg.device));
  CU(launch_philox_indices(h->stream, (long long*)h->stage, batch, h->cfg.seed, h->cfg.agent_id_base + agent, step, size));
  CU(cudaMemcpyAsync(idx_out, h->stage, (size_t)batch * 8, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return DQN_OK;
}

DQN_API int dqn_sample_batch(dqn_handle* h, int32_t agent, const int64_t* idx, int64_t step, int32_t batch,
                     float* s, int64_t* a, float* r, float* s2, uint8_t* done) {
  if (int rc = check_agent(h, agent)) return rc;
  if (batch < 0 || !s || !a || !r || !s2 || !done) return fail(DQN_E_INVALID, "dqn_sample_batch: negative batch or NULL output");
  const long long size = size_of(h, agent);
  if (batch > 0 && size == 0) return fail(DQN_E_INVALID, "dqn_sample_batch: the replay ring is empty (randint(0, 0) raises in the reference)");
  if (idx) for (int i = 0; i < batch; ++i) if (idx[i] < 0 || idx[i] >= h->dims.N) return fail(DQN_E_INVALID, "dqn_sample_batch: index out of range");
  CU(cudaSetDevice(h->cfg.device));
  return gather_to_host(h, agent, idx ? 0 : 1, idx, step, 0, batch, s, a, r, s2, done);
}

namespace {
// one agent per CTA fills the chip once there are ~SMs agents; below that, spread each agent over a 4-CTA cluster
bool uses_cluster(const dqn_handle* h, int n_sel) {
  return h->step_kernel == DQN_STEP_CLUSTER || (h->step_kernel == DQN_STEP_AUTO && 4 * n_sel <= h->sm_count);
}

}  // namespace
extern "C++" {
namespace dqn {
int train_common(dqn_handle* h, int b, int e, int K, const long long* idx_dev, dqn_debug_taps* taps, const InlineStore* ist,
                 const EpisodeCtl* gate) {
  const int n_sel = e - b;
  for (int ag = b; ag < e; ++ag) {
    if (size_of(h, ag) == 0 && !(ist && ist->n > 0)) return fail(DQN_E_INVALID, "dqn_train_step: replay ring of an agent is empty");
  }
  TrainArgs ta;
  memset(&ta, 0, sizeof ta);
  ta.params = h->params; ta.ctl = h->ctl; ta.rings = h->rings; ta.loss_ring = h->loss_ring; ta.loss_mailbox = h->mailbox_dev;
  ta.idx = idx_dev; ta.gate = gate; ta.dims = h->dims; ta.seed = h->cfg.seed; ta.agent_begin = b; ta.agent_id_base = h->cfg.agent_id_base; ta.n_sel = n_sel; ta.K = K;
  if (taps) {
    if (K != 1 || n_sel != 1) return fail(DQN_E_INVALID, "dqn_train_step: debug taps need K == 1 and a single agent");
    uint8_t* t = h->taps;
    ta.taps.enabled = 1;
    ta.taps.indices = (long long*)(t + h->to.idx);
    ta.taps.q = (float*)(t + h->to.q);
    ta.taps.next_q = (float*)(t + h->to.nq);
    ta.taps.next_q_tm = (float*)(t + h->to.nqt);
    ta.taps.targets = (float*)(t + h->to.tgt);
    ta.taps.max_actions = (int*)(t + h->to.maxa);
    ta.taps.loss = (float*)(t + h->to.loss);
    ta.taps.grads = (float*)(t + h->to.grads);
  }
  // one agent per CTA fills the chip once there are >= ~SMs agents; below that an agent's step is a serial
  // latency chain on one SM, so the minibatch rows are spread over a 4-CTA cluster instead
  const bool cluster = uses_cluster(h, n_sel);
  if (ist && !cluster) return fail(DQN_E_INVALID, "internal: inline store needs the cluster kernel");
  if (cluster) CU(launch_train_cluster(h->stream, ta, ist));
  else if (h->step_kernel == DQN_STEP_CTA) CU(launch_train_fused(h->stream, ta));
  else {                                                                       // AUTO / CTA_TC: tensor-core form
    if (n_sel > h->sm_count && !idx_dev) {
      // more agents than SMs = several waves of one-agent CTAs: start the costly agents (more 64-row tiles, a tail tile) first
      // so that the last wave is made of short ones.  Agents are independent -- the order changes no result.
      if (h->order_begin != b || h->order_end != e) {
        CU(cudaStreamSynchronize(h->stream));                                  // an earlier launch may still be reading the old order
        auto cost = [&](int s) { const int B = h->hctl[b + s].batch_size; return B <= 64 ? 10 : B <= 80 ? 14 : 10 * ((B + 63) / 64); };
        std::vector<int> ord(n_sel);
        for (int s = 0; s < n_sel; ++s) ord[s] = s;
        std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) { return cost(x) > cost(y); });
        memcpy(h->order, ord.data(), sizeof(int) * n_sel);
        h->order_begin = b; h->order_end = e;
      }
      ta.order = h->order_dev;
    }
    CU(launch_train_tc(h->stream, ta));
  }
  if (ist) h->hctl[b].ring_counter += ist->n;
  for (int ag = b; ag < e; ++ag) {
    if (gate && !h->hep[ag].pending_train) continue;      // the device gate is closed for this agent: it does not step
    AgentCtl& c = h->hctl[ag];
    c.train_steps += K;
    const long long cnt = (long long)c.adam_count + K;
    c.adam_count = cnt > 0x7fffffffLL ? 0x7fffffff : (int)cnt;
  }
  if (taps) {
    const int B = h->hctl[b].batch_size, A = h->dims.A;
    uint8_t* t = h->taps;
    if (taps->indices) CU(cudaMemcpyAsync(taps->indices, t + h->to.idx, (size_t)B * 8, cudaMemcpyDeviceToHost, h->stream));
    if (taps->q) CU(cudaMemcpyAsync(taps->q, t + h->to.q, (size_t)B * A * 4, cudaMemcpyDeviceToHost, h->stream));
    if (taps->next_q) CU(cudaMemcpyAsync(taps->next_q, t + h->to.nq, (size_t)B * A * 4, cudaMemcpyDeviceToHost, h->stream));
    if (taps->next_q_tm) CU(cudaMemcpyAsync(taps->next_q_tm, t + h->to.nqt, (size_t)B * A * 4, cudaMemcpyDeviceToHost, h->stream));
    if (taps->targets) CU(cudaMemcpyAsync(taps->targets, t + h->to.tgt, (size_t)B * A * 4, cudaMemcpyDeviceToHost, h->stream));
    if (taps->max_actions) CU(cudaMemcpyAsync(taps->max_actions, t + h->to.maxa, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream));
    if (taps->loss) CU(cudaMemcpyAsync(taps->loss, t + h->to.loss, 4, cudaMemcpyDeviceToHost, h->stream));
    if (taps->grads) CU(cudaMemcpyAsync(bounce_slot(h, 0), t + h->to.grads, (size_t)h->dims.PK * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (taps->grads) unpack_flat(h, bounce_slot(h, 0), taps->grads);
  }
  return DQN_OK;
}
}  // namespace dqn
}  // extern "C++"

namespace {
// Loss of the agent's most recent launch without a stream synchronisation: the kernel's last store is
// (train_steps << 32 | loss bits) into mapped pinned memory, so the host polls for the step count it expects.
int wait_mailbox(dqn_handle* h, int agent, float* loss_out) {
  const uint32_t want = (uint32_t)h->hctl[agent].train_steps;
  volatile unsigned long long* mb = h->mailbox + agent;
  for (unsigned long spin = 1;; ++spin) {
    const unsigned long long v = *mb;
    if ((uint32_t)(v >> 32) == want) {
      const uint32_t bits = (uint32_t)v;
      memcpy(loss_out, &bits, 4);
      return DQN_OK;
    }
    if ((spin & 0xfff) == 0) {     // every 4096 polls make sure the stream is still alive (a faulted kernel never writes)
      const cudaError_t e = cudaStreamQuery(h->stream);
      if (e == cudaSuccess) {
        // stream drained: the launch that produced `want` has completed; if the mailbox still disagrees the handle's
        // state was edited between launches (e.g. counters restored) -- read the loss ring instead
        if ((uint32_t)(*mb >> 32) == want) continue;
        const size_t pos = (size_t)((h->hctl[agent].train_steps - 1) % kLossCap);
        CU(cudaMemcpyAsync(h->bounce, h->loss_ring + (size_t)agent * kLossCap + pos, 4, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        memcpy(loss_out, h->bounce, 4);
        return DQN_OK;
      }
      if (e != cudaErrorNotReady) return fail(DQN_E_CUDA, std::string("dqn_get_losses: ") + cudaGetErrorString(e));
    }
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
}
}  // namespace

DQN_API int dqn_store_train_step(dqn_handle* h, int32_t agent, int64_t n, const float* s, const int64_t* a, const float* r,
                                 const float* s2, const uint8_t* done, int32_t K, float* loss_out) {
  if (int rc = check_agent_raw(h, agent)) return rc;
  if (K < 1) return fail(DQN_E_INVALID, "dqn_store_train_step: K must be >= 1");
  if (n < 0 || (n > 0 && (!s || !a || !r || !s2 || !done))) return fail(DQN_E_INVALID, "dqn_store_train_step: negative n or NULL array");
  CU(cudaSetDevice(h->cfg.device));
  int srv = DQN_OK;
  if (h->session_enabled && K == 1 && n <= kInlineMax) {
    if (size_of(h, agent) + n == 0) return fail(DQN_E_INVALID, "dqn_train_step: replay ring of an agent is empty");
    srv = session_prepare(h, 1);                             // at most ONE earlier command still in flight (this one's slot is free), kernel alive
    if (srv < 0) return srv;
  }
  if (h->session_enabled && K == 1 && n <= kInlineMax && srv == DQN_OK) {
    // served by the resident kernel: records into the mapped slot, ring the doorbell; no launch, no copy
    const int D = h->dims.D, recw = h->dims.recw;
    const unsigned long long stamp = ((h->session_seq + 1) & 0xffffffffull) << 32;     // the command's sequence number
    volatile unsigned long long* slot = h->sess->stamped[(h->session_seq + 1) & 1];
    for (int i = 0; i < (int)n; ++i) {
      uint32_t rec[kInlineWords];
      memset(rec, 0, sizeof rec);
      memcpy(rec, s + (size_t)i * D, (size_t)D * 4);
      memcpy(rec + D, s2 + (size_t)i * D, (size_t)D * 4);
      memcpy(rec + 2 * D, a + i, 8);
      memcpy(rec + 2 * D + 2, r + i, 4);
      rec[2 * D + 3] = done[i] ? 1u : 0u;
      for (int w = 0; w < recw; ++w) slot[(size_t)i * recw + w] = stamp | rec[w];
    }
    AgentCtl& c = h->hctl[agent];
    h->session_step_no[(h->session_seq + 1) & 1] = c.train_steps + 1;
    session_publish(h, kOpStep, (int)n);
    c.ring_counter += n;
    c.train_steps += 1;
    if (c.adam_count < 0x7fffffff) c.adam_count += 1;
    if (loss_out) {
      if (int rc = session_collect(h, nullptr)) return rc;
      *loss_out = h->session_last_loss;
    }
    return DQN_OK;
  }
  if (h->session_active) if (int rc = session_stop(h)) return rc;
  if (n > 0 && n <= kInlineMax && uses_cluster(h, 1)) {
    // the transitions ride in the kernel-parameter buffer: no H2D copy, no store launch
    InlineStore ist;
    ist.n = (int)n;
    const int D = h->dims.D, recw = h->dims.recw;
    memset(ist.rec, 0, (size_t)n * recw * 4);
    for (int i = 0; i < (int)n; ++i) {
      uint32_t* rec = ist.rec + (size_t)i * recw;
      memcpy(rec, s + (size_t)i * D, (size_t)D * 4);
      memcpy(rec + D, s2 + (size_t)i * D, (size_t)D * 4);
      memcpy(rec + 2 * D, a + i, 8);
      memcpy(rec + 2 * D + 2, r + i, 4);
      rec[2 * D + 3] = done[i] ? 1u : 0u;
    }
    if (int rc = train_common(h, agent, agent + 1, K, nullptr, nullptr, &ist)) return rc;
  } else {
    if (n > 0) if (int rc = dqn_store(h, agent, n, s, a, r, s2, done)) return rc;
    if (int rc = train_common(h, agent, agent + 1, K, nullptr, nullptr)) return rc;
  }
  if (loss_out) return wait_mailbox(h, agent, loss_out);
  return DQN_OK;
}

DQN_API int dqn_train_step(dqn_handle* h, int32_t agent_begin, int32_t agent_end, int32_t K, const int64_t* idx, dqn_debug_taps* taps) {
  if (int rc = check_range(h, agent_begin, agent_end)) return rc;
  if (K < 1) return fail(DQN_E_INVALID, "dqn_train_step: K must be >= 1");
  CU(cudaSetDevice(h->cfg.device));
  const long long* idx_dev = nullptr;
  if (idx) {
    // explicit indices need one common batch size across the selected agents (layout [n_sel][K][B])
    const int B = h->hctl[agent_begin].batch_size;
    size_t total = 0;
    for (int ag = agent_begin; ag < agent_end; ++ag) {
      if (h->hctl[ag].batch_size != B) return fail(DQN_E_INVALID, "dqn_train_step: explicit indices need equal batch sizes in the agent range");
      const long long size = size_of(h, ag);
      const int64_t* p = idx + (size_t)(ag - agent_begin) * K * B;
      for (size_t i = 0; i < (size_t)K * B; ++i) if (p[i] < 0 || p[i] >= size) return fail(DQN_E_INVALID, "dqn_train_step: index outside [0, size)");
      total += (size_t)K * B;
    }
    if (total * 8 > kStageBytes) return fail(DQN_E_INVALID, "dqn_train_step: explicit index block exceeds the 8 MiB staging buffer");
    CU(cudaMemcpyAsync(h->stage, idx, total * 8, cudaMemcpyHostToDevice, h->stream));
    idx_dev = (const long long*)h->stage;
  }
  return train_common(h, agent_begin, agent_end, K, idx_dev, taps);
}

DQN_API int dqn_train_step_device_idx(dqn_handle* h, int32_t agent_begin, int32_t agent_end, int32_t K, const int64_t* idx_dev) {
  if (int rc = check_range(h, agent_begin, agent_end)) return rc;
  if (K < 1) return fail(DQN_E_INVALID, "dqn_train_step: K must be >= 1");
  CU(cudaSetDevice(h->cfg.device));
  return train_common(h, agent_begin, agent_end, K, (const long long*)idx_dev, nullptr);
}

DQN_API int dqn_get_losses(dqn_handle* h, int32_t agent, int32_t n, float* loss_out, int64_t* train_steps_out) {
  if (int rc = check_agent_raw(h, agent)) return rc;
  if (h->session_active && n <= 1) {
    if (train_steps_out) *train_steps_out = h->hctl[agent].train_steps;
    if (n == 0) return DQN_OK;
    if (!loss_out || h->hctl[agent].train_steps < 1) return fail(DQN_E_INVALID, "dqn_get_losses: n must be <= min(train steps so far, 4096)");
    if (int rc = session_collect(h, nullptr)) return rc;      // (train-step answers land in session_last_loss)
    if ((uint32_t)(h->mailbox[agent] >> 32) == (uint32_t)h->hctl[agent].train_steps) {   // also covers steps of earlier launches
      const uint32_t bits = (uint32_t)h->mailbox[agent];
      memcpy(loss_out, &bits, 4);
    } else {
      *loss_out = h->session_last_loss;
    }
    return DQN_OK;
  }
  if (int rc = check_agent(h, agent)) return rc;
  const long long ts = h->hctl[agent].train_steps;
  if (train_steps_out) *train_steps_out = ts;
  if (n == 0) return DQN_OK;
  if (n < 0 || n > kLossCap || n > ts || !loss_out) return fail(DQN_E_INVALID, "dqn_get_losses: n must be <= min(train steps so far, 4096)");
  CU(cudaSetDevice(h->cfg.device));
  if (n == 1) return wait_mailbox(h, agent, loss_out);   // the kernel wrote it straight into mapped host memory
  // the last n losses are at ring positions [(ts-n) % cap, ts % cap): at most two contiguous pieces
  const size_t first = (size_t)((ts - n) % kLossCap);
  const size_t n1 = first + (size_t)n <= (size_t)kLossCap ? (size_t)n : (size_t)kLossCap - first;
  const float* src = h->loss_ring + (size_t)agent * kLossCap;
  float* dst = (size_t)n * 4 <= kBounceBytes ? (float*)h->bounce : loss_out;   // pinned bounce keeps small reads cheap
  CU(cudaMemcpyAsync(dst, src + first, n1 * 4, cudaMemcpyDeviceToHost, h->stream));
  if (n1 < (size_t)n) CU(cudaMemcpyAsync(dst + n1, src, ((size_t)n - n1) * 4, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (dst != loss_out) memcpy(loss_out, dst, (size_t)n * 4);
  return DQN_OK;
}

DQN_API int dqn_get_loss_lagged(dqn_handle* h, int32_t agent, int32_t lag, float* loss_out) {
  if (int rc = check_agent_raw(h, agent)) return rc;
  if (!loss_out || lag < 0 || lag > 1) return fail(DQN_E_INVALID, "dqn_get_loss_lagged: lag must be 0 or 1, loss_out not NULL");
  if (lag == 0) return dqn_get_losses(h, agent, 1, loss_out, nullptr);
  const long long T = h->hctl[agent].train_steps - 1;         // the step whose loss is wanted
  if (T < 1) return fail(DQN_E_INVALID, "dqn_get_loss_lagged: needs two train steps");
  if (h->session_active) {
    // the newest command stays in flight when it is the train step after T; every older answer is collected
    const int newest = (int)(h->session_seq & 1);
    const int keep = h->session_inflight > 0 && h->session_op[newest] == kOpStep && h->session_step_no[newest] == T + 1 ? 1 : 0;
    if (int rc = session_collect(h, nullptr, keep)) return rc;
    if (h->session_loss_step[T & 1] == T) { *loss_out = h->session_loss[T & 1]; return DQN_OK; }
  }
  if (int rc = check_agent(h, agent)) return rc;               // (ends a session) the step ran in an earlier launch: the loss ring has it
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaMemcpyAsync(h->bounce, h->loss_ring + (size_t)agent * kLossCap + (size_t)((T - 1) % kLossCap), 4, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  memcpy(loss_out, h->bounce, 4);
  return DQN_OK;
}

DQN_API int dqn_sync_target(dqn_handle* h, int32_t agent_begin, int32_t agent_end) {
  if (h && h->session_active && agent_begin == 0 && agent_end == 1) {
    const int srv = session_prepare(h);
    if (srv < 0) return srv;
    if (srv == DQN_OK) {
      session_publish(h, kOpSync, 0);
      return DQN_OK;
    }
  }
  if (int rc = check_range(h, agent_begin, agent_end)) return rc;
  CU(cudaSetDevice(h->cfg.device));
  CU(launch_sync_target(h->stream, h->params, h->dims, agent_begin, agent_end - agent_begin));
  return DQN_OK;
}

DQN_API int dqn_polyak_target(dqn_handle* h, int32_t agent_begin, int32_t agent_end, float tau) {
  if (int rc = check_range(h, agent_begin, agent_end)) return rc;      // (ends a session: the resident kernel serves the hard sync only)
  if (!(tau >= 0.f && tau <= 1.f)) return fail(DQN_E_INVALID, "dqn_polyak_target: tau must be in [0,1]");
  CU(cudaSetDevice(h->cfg.device));
  CU(launch_polyak_target(h->stream, h->params, h->dims, agent_begin, agent_end - agent_begin, tau));
  return DQN_OK;
}

DQN_API int dqn_set_loss_kind(dqn_handle* h, int32_t agent_begin, int32_t agent_end, int32_t kind) {
  if (int rc = check_range(h, agent_begin, agent_end)) return rc;
  if (kind != DQN_LOSS_HUBER && kind != DQN_LOSS_L2) return fail(DQN_E_INVALID, "dqn_set_loss_kind: kind must be DQN_LOSS_HUBER or DQN_LOSS_L2");
  CU(cudaSetDevice(h->cfg.device));
  for (int a = agent_begin; a < agent_end; ++a) {
    h->hctl[a].loss_kind = kind;
    CU(cudaMemcpyAsync(&h->ctl[a].loss_kind, &h->hctl[a].loss_kind, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  }
  CU(cudaStreamSynchronize(h->stream));
  return DQN_OK;
}

DQN_API int dqn_act_batch(dqn_handle* h, int32_t agent_begin, int32_t agent_end, const float* states, int32_t* actions_out) {
  if (int rc = check_range(h, agent_begin, agent_end)) return rc;
  if (!states || !actions_out) return fail(DQN_E_INVALID, "dqn_act: NULL argument");
  CU(cudaSetDevice(h->cfg.device));
  const int n = agent_end - agent_begin, D = h->dims.D;
  float* dstates = (float*)h->stage;
  int* dact = (int*)(h->stage + align_up((size_t)n * D * 4, 256));
  const bool small = (size_t)n * D * 4 <= kBounceBytes / 2 && (size_t)n * 4 <= kBounceBytes / 2;
  if (small) {   // bounce through pinned memory: truly asynchronous copies, one synchronisation
    memcpy(h->bounce, states, (size_t)n * D * 4);
    CU(cudaMemcpyAsync(dstates, h->bounce, (size_t)n * D * 4, cudaMemcpyHostToDevice, h->stream));
  } else {
    CU(cudaMemcpyAsync(dstates, states, (size_t)n * D * 4, cudaMemcpyHostToDevice, h->stream));
  }
  CU(launch_act(h->stream, h->params, h->dims, agent_begin, n, dstates, dact, nullptr));
  if (small) {
    int32_t* pa = (int32_t*)(h->bounce + kBounceBytes / 2);
    CU(cudaMemcpyAsync(pa, dact, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    memcpy(actions_out, pa, (size_t)n * 4);
  } else {
    CU(cudaMemcpyAsync(actions_out, dact, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  return DQN_OK;
}

DQN_API int dqn_act(dqn_handle* h, int32_t agent, const float* state, int32_t* action_out) {
  if (int rc = check_agent_raw(h, agent)) return rc;
  int srv = DQN_OK;
  if (h->session_enabled) {
    if (!state || !action_out) return fail(DQN_E_INVALID, "dqn_act: NULL argument");
    CU(cudaSetDevice(h->cfg.device));
    srv = session_prepare(h);
    if (srv < 0) return srv;
  }
  if (h->session_enabled && srv == DQN_OK) {
    const unsigned long long stamp = ((h->session_seq + 1) & 0xffffffffull) << 32;
    for (int k = 0; k < h->dims.D; ++k) {
      uint32_t bits;
      memcpy(&bits, state + k, 4);
      h->sess->stamped[(h->session_seq + 1) & 1][k] = stamp | bits;
    }
    session_publish(h, kOpAct, 0);
    uint32_t act = 0;
    if (int rc = session_collect(h, &act)) return rc;
    *action_out = (int32_t)act;
    return DQN_OK;
  }
  if (int rc = check_agent(h, agent)) return rc;
  return dqn_act_batch(h, agent, agent + 1, state, action_out);
}

}  // extern "C"
