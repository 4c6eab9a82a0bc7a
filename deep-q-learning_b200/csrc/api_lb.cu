// api_lb.cu -- C ABI of the large-batch data-parallel mode (dqn_lb_* in include/dqn_b200.h).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "../../include/dqn_b200.h"
#include "common.cuh"
#include "kernels.h"
#include "large.h"

using namespace dqn;

namespace dqn {
cudaError_t lb_gemm(cudaStream_t st, int gemm_mode, int kind, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                    float* C, int ldc, const float* aux, int ldaux, int splitk, const LbWorkspace& ws) {
  if (gemm_mode == kGemmModeTC3xTF32) return lb_gemm_tc(st, kind, M, N, K, A, lda, B, ldb, C, ldc, aux, ldaux, splitk, ws);
  return lb_gemm_ffma(st, kind, M, N, K, A, lda, B, ldb, C, ldc, aux, ldaux, splitk);
}
}  // namespace dqn

namespace {

int lbfail(int code, const std::string& m) { dqn::set_last_error(m.c_str()); return code; }

#define CU(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      char _b[512];                                                                                \
      snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return lbfail(DQN_E_CUDA, _b);                                                               \
    }                                                                                              \
  } while (0)

size_t up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
constexpr size_t kLbStage = 8u << 20;

struct LbCarve {
  size_t theta, theta_t, mu, nu, grads, ctl, ring, s, a, r, s2, done, idx, H1, H2, Q, dhd, dH2, dH1, partial, colpart, colred,
      gemmpart, tc, targets, maxa, stage, total;
};

LbCarve lb_carve(const LbDims& d) {
  LbCarve c;
  size_t o = 0;
  const size_t B = d.B;
  auto take = [&](size_t& field, size_t bytes) { field = o; o += up(bytes); };
  take(c.theta, (size_t)d.PF * 4); take(c.theta_t, (size_t)d.PF * 4); take(c.mu, (size_t)d.PF * 4); take(c.nu, (size_t)d.PF * 4);
  take(c.grads, (size_t)d.PF * 4);
  take(c.ctl, sizeof(AgentCtl));
  take(c.ring, (size_t)d.N * d.recw * 4);
  take(c.s, B * d.D * 4); take(c.a, B * 8); take(c.r, B * 4); take(c.s2, B * d.D * 4); take(c.done, B); take(c.idx, B * 8);
  take(c.H1, 3 * B * d.H1 * 4); take(c.H2, 3 * B * d.H2 * 4); take(c.Q, 3 * B * d.A * 4); take(c.dhd, B * 8 * 4);
  take(c.dH2, B * d.H2 * 4); take(c.dH1, B * d.H1 * 4);
  take(c.partial, (B / 256 + 1) * 16 * 4);
  const size_t colw = (size_t)(2 + kMaxA) * d.H2 > (size_t)(d.D + 1) * d.H1 ? (size_t)(2 + kMaxA) * d.H2 : (size_t)(d.D + 1) * d.H1;
  take(c.colpart, (B / lb_row_chunk((long long)B)) * colw * 4);
  take(c.colred, colw * 4);
  take(c.gemmpart, (size_t)16 * d.H1 * d.H2 * 4);
  take(c.tc, (size_t)d.H1 * d.H2 * 4);          // tcgen05 path: W2^T for the dh1 GEMM
  take(c.targets, B * d.A * 4); take(c.maxa, B * 4);
  take(c.stage, kLbStage);
  c.total = o;
  return c;
}

int lb_validate(const dqn_lb_config* cfg, LbDims* d) {
  if (!cfg) return lbfail(DQN_E_INVALID, "config is NULL");
  if (cfg->struct_size != (int32_t)sizeof(dqn_lb_config)) return lbfail(DQN_E_INVALID, "dqn_lb_config.struct_size mismatch");
  if (cfg->obs_dim < 1 || cfg->obs_dim > DQN_MAX_OBS_DIM) return lbfail(DQN_E_INVALID, "obs_dim must be in [1,16]");
  if (cfg->num_actions < 2 || cfg->num_actions > DQN_MAX_ACTIONS) return lbfail(DQN_E_INVALID, "num_actions must be in [2,7]");
  if (cfg->hidden1 < 256 || cfg->hidden1 % 256 || cfg->hidden2 < 256 || cfg->hidden2 % 256)
    return lbfail(DQN_E_INVALID, "large-batch mode needs hidden widths that are multiples of 256");
  if (cfg->batch_local < 128 || cfg->batch_local % 128) return lbfail(DQN_E_INVALID, "batch_local must be a positive multiple of 128");
  if (cfg->buffer_size < 1 || cfg->buffer_size >= (1ll << 31)) return lbfail(DQN_E_INVALID, "buffer_size must be in [1, 2^31)");
  if (cfg->world < 1 || cfg->rank < 0 || cfg->rank >= cfg->world) return lbfail(DQN_E_INVALID, "need 0 <= rank < world");
  if (cfg->gemm_mode != 0 && cfg->gemm_mode != 1) return lbfail(DQN_E_INVALID, "gemm_mode must be 0 (fp32 FFMA) or 1 (tcgen05 3xTF32)");
  if ((long long)cfg->batch_local * cfg->world >= (1ll << 31)) return lbfail(DQN_E_INVALID, "global batch too large");
  d->D = cfg->obs_dim; d->A = cfg->num_actions; d->H1 = cfg->hidden1; d->H2 = cfg->hidden2; d->B = cfg->batch_local;
  d->P = d->D * d->H1 + d->H1 + d->H1 * d->H2 + d->H2 + d->H2 + 1 + d->H2 * d->A + d->A;
  d->PF = (d->P + 1 + 3) & ~3;      // +1: the loss rides behind the gradients through the all-reduce
  d->recw = record_words_host(d->D);
  d->N = cfg->buffer_size;
  return DQN_OK;
}

}  // namespace

struct dqn_lb_handle {
  dqn_lb_config cfg;
  LbDims dims;
  LbCarve cv;
  LbWorkspace ws;
  cudaStream_t stream;
  uint8_t* arena;
  bool own_arena;
  AgentCtl* ctl;
  uint32_t* ring;
  uint8_t* stage;
  LbTaps taps;
  long long ring_counter;
  long long train_steps;
  int adam_count;
  float* pinned;
  // peer-memory all-reduce (dqn_lb_comm_*): the gradient lives in `window` instead of the arena once initialised
  float* window;
  CommPeers peers;
  void* opened[kMaxWorld];     // cudaIpcOpenMemHandle results to close
  bool connected;
  unsigned epoch[kCommChannels];
  cudaStream_t comm_stream;      // second stream: the W2 gradient is exchanged here while the main stream finishes backward
  cudaEvent_t ev_dw2, ev_comm;
  unsigned long long comm_timeout_ns;
  int loss_kind;
};

namespace {
// the all-reduce kernel mirrors its error flag into this word of the handle's pinned (mapped) host page
volatile unsigned* comm_host_error(const dqn_lb_handle* h) { return reinterpret_cast<volatile unsigned*>(h->pinned) + 64; }
int comm_failed(const dqn_lb_handle* h) {
  return *comm_host_error(h) ? lbfail(DQN_E_CUDA, "peer-memory all-reduce timed out waiting for a rank: the parameter update was "
                                                  "skipped and the replicas are out of step (rebuild the group)") : DQN_OK;
}
long long lb_size(const dqn_lb_handle* h) { return h->ring_counter < h->dims.N ? h->ring_counter : h->dims.N; }
Dims ring_dims(const dqn_lb_handle* h) {
  Dims d; d.D = h->dims.D; d.A = h->dims.A; d.P = h->dims.P; d.PK = 0; d.recw = h->dims.recw; d.N = h->dims.N;
  return d;
}
}  // namespace

extern "C" {

DQN_API int dqn_lb_arena_bytes(const dqn_lb_config* cfg, uint64_t* bytes_out) {
  LbDims d;
  if (int rc = lb_validate(cfg, &d)) return rc;
  if (!bytes_out) return lbfail(DQN_E_INVALID, "bytes_out is NULL");
  *bytes_out = lb_carve(d).total;
  return DQN_OK;
}

DQN_API int dqn_lb_create(const dqn_lb_config* cfg, dqn_lb_handle** out) {
  LbDims d;
  if (int rc = lb_validate(cfg, &d)) return rc;
  if (!out) return lbfail(DQN_E_INVALID, "out is NULL");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return lbfail(DQN_E_ARCH, "no CUDA device visible: libdqn_b200 has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return lbfail(DQN_E_INVALID, "device ordinal out of range");
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) return lbfail(DQN_E_ARCH, "libdqn_b200 is built for sm_100a (B200) only");
  CU(cudaSetDevice(cfg->device));
  dqn_lb_handle* h = new dqn_lb_handle();
  h->cfg = *cfg; h->dims = d; h->stream = (cudaStream_t)cfg->stream; h->cv = lb_carve(d);
  if (cfg->arena) {
    if (cfg->arena_bytes < h->cv.total || ((uintptr_t)cfg->arena & 255)) { delete h; return lbfail(DQN_E_INVALID, "arena too small or misaligned"); }
    h->arena = (uint8_t*)cfg->arena; h->own_arena = false;
  } else {
    if (cudaMalloc((void**)&h->arena, h->cv.total) != cudaSuccess) { delete h; return lbfail(DQN_E_NOMEM, "cudaMalloc(arena) failed"); }
    h->own_arena = true;
  }
  uint8_t* a = h->arena;
  const LbCarve& c = h->cv;
  LbWorkspace& w = h->ws;
  w.theta = (float*)(a + c.theta); w.theta_t = (float*)(a + c.theta_t); w.mu = (float*)(a + c.mu); w.nu = (float*)(a + c.nu);
  w.grads = (float*)(a + c.grads);
  h->ctl = (AgentCtl*)(a + c.ctl); h->ring = (uint32_t*)(a + c.ring);
  w.s = (float*)(a + c.s); w.a = (long long*)(a + c.a); w.r = (float*)(a + c.r); w.s2 = (float*)(a + c.s2); w.done = a + c.done;
  w.idx = (long long*)(a + c.idx);
  w.H1 = (float*)(a + c.H1); w.H2 = (float*)(a + c.H2); w.Q = (float*)(a + c.Q); w.dhd = (float*)(a + c.dhd);
  w.dH2 = (float*)(a + c.dH2); w.dH1 = (float*)(a + c.dH1); w.partial = (float*)(a + c.partial);
  w.colpart = (float*)(a + c.colpart); w.colred = (float*)(a + c.colred); w.gemmpart = (float*)(a + c.gemmpart);
  w.tc_scratch = (float*)(a + c.tc);
  h->taps.targets = (float*)(a + c.targets); h->taps.max_actions = (int*)(a + c.maxa); h->taps.enabled = 0;
  h->stage = a + c.stage;
  h->ring_counter = 0; h->train_steps = 0; h->adam_count = 0; h->pinned = nullptr;
  h->window = nullptr; h->connected = false; h->epoch[0] = h->epoch[1] = 0; h->loss_kind = kLossHuber;
  h->comm_stream = nullptr; h->ev_dw2 = h->ev_comm = nullptr;
  {
    const char* ms = getenv("DQN_B200_COMM_TIMEOUT_MS");
    const double v = ms ? atof(ms) : 20000.0;
    h->comm_timeout_ns = (unsigned long long)((v > 0.0 ? v : 20000.0) * 1e6);
  }
  memset(&h->peers, 0, sizeof h->peers); memset(h->opened, 0, sizeof h->opened);
  cudaError_t e = cudaMallocHost((void**)&h->pinned, 4096);
  if (e == cudaSuccess) memset(h->pinned, 0, 4096);
  if (e == cudaSuccess) e = cudaMemsetAsync(a, 0, c.s, h->stream);          // params, moments, grads, ctl, ring
  if (e == cudaSuccess) e = cudaMemsetAsync(a + c.dhd, 0, (size_t)d.B * 32, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) {
    std::string m = std::string("dqn_lb_create: device initialisation failed: ") + cudaGetErrorString(e);
    if (h->pinned) cudaFreeHost(h->pinned);
    if (h->own_arena) cudaFree(h->arena);
    delete h;
    return lbfail(DQN_E_CUDA, m);
  }
  *out = h;
  return DQN_OK;
}

DQN_API int dqn_lb_destroy(dqn_lb_handle* h) {
  if (!h) return DQN_OK;
  cudaSetDevice(h->cfg.device);
  cudaStreamSynchronize(h->stream);
  if (h->comm_stream) { cudaStreamSynchronize(h->comm_stream); cudaStreamDestroy(h->comm_stream); }
  if (h->ev_dw2) cudaEventDestroy(h->ev_dw2);
  if (h->ev_comm) cudaEventDestroy(h->ev_comm);
  for (int q = 0; q < kMaxWorld; ++q) if (h->opened[q]) cudaIpcCloseMemHandle(h->opened[q]);
  if (h->window) cudaFree(h->window);
  if (h->pinned) cudaFreeHost(h->pinned);
  if (h->own_arena) cudaFree(h->arena);
  delete h;
  return DQN_OK;
}

DQN_API int dqn_lb_param_count(const dqn_lb_handle* h, int32_t* p_out) {
  if (!h || !p_out) return lbfail(DQN_E_INVALID, "NULL argument");
  *p_out = h->dims.P;
  return DQN_OK;
}

DQN_API int dqn_lb_synchronize(dqn_lb_handle* h) {
  if (!h) return lbfail(DQN_E_INVALID, "handle is NULL");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaStreamSynchronize(h->stream));
  return DQN_OK;
}

DQN_API int dqn_lb_set_params(dqn_lb_handle* h, int32_t which, const float* host_flat, int32_t n) {
  if (!h || !host_flat || n != h->dims.P || (which != 0 && which != 1)) return lbfail(DQN_E_INVALID, "dqn_lb_set_params: bad argument");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaMemcpyAsync(which ? h->ws.theta_t : h->ws.theta, host_flat, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return DQN_OK;
}

DQN_API int dqn_lb_get_params(dqn_lb_handle* h, int32_t which, float* host_flat, int32_t n) {
  if (!h || !host_flat || n != h->dims.P || (which != 0 && which != 1)) return lbfail(DQN_E_INVALID, "dqn_lb_get_params: bad argument");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaMemcpyAsync(host_flat, which ? h->ws.theta_t : h->ws.theta, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return DQN_OK;
}

DQN_API int dqn_lb_set_opt_state(dqn_lb_handle* h, int32_t count, const float* mu, const float* nu, int32_t n) {
  if (!h || !mu || !nu || n != h->dims.P || count < 0) return lbfail(DQN_E_INVALID, "dqn_lb_set_opt_state: bad argument");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaMemcpyAsync(h->ws.mu, mu, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->ws.nu, nu, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  h->adam_count = count;
  return DQN_OK;
}

DQN_API int dqn_lb_get_opt_state(dqn_lb_handle* h, int32_t* count, float* mu, float* nu, int32_t n) {
  if (!h || n != h->dims.P) return lbfail(DQN_E_INVALID, "dqn_lb_get_opt_state: bad argument");
  CU(cudaSetDevice(h->cfg.device));
  if (mu) CU(cudaMemcpyAsync(mu, h->ws.mu, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
  if (nu) CU(cudaMemcpyAsync(nu, h->ws.nu, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (count) *count = h->adam_count;
  return DQN_OK;
}

DQN_API int dqn_lb_store_device(dqn_lb_handle* h, int64_t n, const float* s, const int64_t* a, const float* r, const float* s2, const uint8_t* done) {
  if (!h || n < 0 || (n > 0 && (!s || !a || !r || !s2 || !done))) return lbfail(DQN_E_INVALID, "dqn_lb_store: bad argument");
  if (n == 0) return DQN_OK;
  CU(cudaSetDevice(h->cfg.device));
  const long long N = h->dims.N, skip = n > N ? n - N : 0;
  const int D = h->dims.D;
  CU(launch_replay_store(h->stream, h->ring, ring_dims(h), h->ring_counter + skip, n - skip, s + skip * D, (const long long*)a + skip,
                         r + skip, s2 + skip * D, done + skip, h->ctl));
  h->ring_counter += n;
  return DQN_OK;
}

DQN_API int dqn_lb_store(dqn_lb_handle* h, int64_t n, const float* s, const int64_t* a, const float* r, const float* s2, const uint8_t* done) {
  if (!h || n < 0 || (n > 0 && (!s || !a || !r || !s2 || !done))) return lbfail(DQN_E_INVALID, "dqn_lb_store: bad argument");
  CU(cudaSetDevice(h->cfg.device));
  const int D = h->dims.D;
  const long long cap = (long long)((kLbStage - 256) / (8 * D + 13));
  for (long long off = 0; off < n; off += cap) {
    const long long m = n - off < cap ? n - off : cap;
    uint8_t* p = h->stage;
    long long* da = (long long*)p; p += up((size_t)m * 8, 16);
    float* ds = (float*)p; p += up((size_t)m * D * 4, 16);
    float* ds2 = (float*)p; p += up((size_t)m * D * 4, 16);
    float* dr = (float*)p; p += up((size_t)m * 4, 16);
    uint8_t* dd = p;
    CU(cudaMemcpyAsync(ds, s + off * D, (size_t)m * D * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(da, a + off, (size_t)m * 8, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(dr, r + off, (size_t)m * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(ds2, s2 + off * D, (size_t)m * D * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(dd, done + off, (size_t)m, cudaMemcpyHostToDevice, h->stream));
    if (int rc = dqn_lb_store_device(h, m, ds, (const int64_t*)da, dr, ds2, dd)) return rc;
  }
  return DQN_OK;
}

DQN_API int dqn_lb_buffer_state(dqn_lb_handle* h, int64_t* size_out, int64_t* counter_out) {
  if (!h) return lbfail(DQN_E_INVALID, "handle is NULL");
  if (size_out) *size_out = lb_size(h);
  if (counter_out) *counter_out = h->ring_counter;
  return DQN_OK;
}

DQN_API int dqn_lb_forward_backward(dqn_lb_handle* h, const int64_t* idx, int32_t debug) {
  if (!h) return lbfail(DQN_E_INVALID, "handle is NULL");
  const long long size = lb_size(h);
  if (size == 0) return lbfail(DQN_E_INVALID, "dqn_lb_forward_backward: the replay ring is empty");
  CU(cudaSetDevice(h->cfg.device));
  const int B = h->dims.B;
  if (idx) {
    for (int i = 0; i < B; ++i) if (idx[i] < 0 || idx[i] >= size) return lbfail(DQN_E_INVALID, "dqn_lb_forward_backward: index outside [0, size)");
    if ((size_t)B * 8 > kLbStage) return lbfail(DQN_E_INVALID, "index block exceeds the staging buffer");
    CU(cudaMemcpyAsync(h->ws.idx, idx, (size_t)B * 8, cudaMemcpyHostToDevice, h->stream));
  } else {   // this rank's slice of the global Philox draw of the step
    CU(launch_philox_indices(h->stream, h->ws.idx, B, h->cfg.seed, 0, h->train_steps, size, h->cfg.rank * B));
  }
  CU(launch_replay_gather(h->stream, h->ring, ring_dims(h), 0, h->ws.idx, 0, 0, 0, size, B, h->ws.s, h->ws.a, h->ws.r, h->ws.s2, h->ws.done));
  h->taps.enabled = debug ? 1 : 0;
  const float inv = 1.0f / ((float)B * (float)h->cfg.world);
  CU(lb_forward_backward(h->stream, h->dims, h->ws, h->cfg.gamma, inv, h->cfg.gemm_mode, h->loss_kind, h->taps));
  return DQN_OK;
}

DQN_API int dqn_lb_grads(dqn_lb_handle* h, void** dev_ptr_out, int64_t* count_out) {
  if (!h || !dev_ptr_out || !count_out) return lbfail(DQN_E_INVALID, "NULL argument");
  *dev_ptr_out = h->ws.grads;
  *count_out = h->dims.P + 1;
  return DQN_OK;
}

DQN_API int dqn_lb_comm_init(dqn_lb_handle* h, void* ipc_handle_out, void** window_out) {
  if (!h) return lbfail(DQN_E_INVALID, "handle is NULL");
  if (h->window) return lbfail(DQN_E_INVALID, "dqn_lb_comm_init: already initialised");
  const int W = h->cfg.world;
  if (W != 2 && W != 4 && W != 8) return lbfail(DQN_E_INVALID, "dqn_lb_comm_init: the peer-memory all-reduce supports world = 2, 4 or 8");
  CU(cudaSetDevice(h->cfg.device));
  const int n4 = h->dims.PF / 4;
  CU(cudaMalloc((void**)&h->window, comm_window_bytes(n4)));      // its own allocation: that is what cudaIpc shares
  CU(cudaMemsetAsync(h->window, 0, comm_window_bytes(n4), h->stream));
  CU(cudaStreamSynchronize(h->stream));
  h->ws.grads = h->window;                                        // backward writes / Adam reads the window from now on
  h->peers.win[h->cfg.rank] = h->window;
  if (ipc_handle_out) {
    cudaIpcMemHandle_t mh;
    CU(cudaIpcGetMemHandle(&mh, h->window));
    static_assert(sizeof(mh) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(ipc_handle_out, &mh, sizeof mh);
  }
  if (window_out) *window_out = h->window;
  return DQN_OK;
}

DQN_API int dqn_lb_comm_connect(dqn_lb_handle* h, const void* ipc_handles, void* const* peer_windows) {
  if (!h || !h->window) return lbfail(DQN_E_INVALID, "dqn_lb_comm_connect: call dqn_lb_comm_init first");
  if (!ipc_handles == !peer_windows) return lbfail(DQN_E_INVALID, "dqn_lb_comm_connect: pass either IPC handles or same-process pointers");
  CU(cudaSetDevice(h->cfg.device));
  for (int q = 0; q < h->cfg.world; ++q) {
    if (q == h->cfg.rank) continue;
    if (ipc_handles) {
      cudaIpcMemHandle_t mh;
      memcpy(&mh, (const uint8_t*)ipc_handles + (size_t)q * sizeof mh, sizeof mh);
      void* p = nullptr;
      CU(cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
      h->opened[q] = p;
      h->peers.win[q] = (float*)p;
    } else {
      if (!peer_windows[q]) return lbfail(DQN_E_INVALID, "dqn_lb_comm_connect: NULL peer window");
      cudaPointerAttributes at;
      CU(cudaPointerGetAttributes(&at, peer_windows[q]));
      if (at.device != h->cfg.device) {
        const cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e);
        cudaGetLastError();
      }
      h->peers.win[q] = (float*)peer_windows[q];
    }
  }
  h->connected = true;
  return DQN_OK;
}

DQN_API int dqn_lb_allreduce(dqn_lb_handle* h) {
  if (!h) return lbfail(DQN_E_INVALID, "handle is NULL");
  if (h->cfg.world == 1) return DQN_OK;
  if (!h->connected) return lbfail(DQN_E_INVALID, "dqn_lb_allreduce: dqn_lb_comm_connect has not been called");
  if (int rc = comm_failed(h)) return rc;
  CU(cudaSetDevice(h->cfg.device));
  CommRange whole = {{0, 0}, {h->dims.PF / 4, 0}};
  h->epoch[1] += 1;
  CU(lb_allreduce(h->stream, h->peers, h->cfg.world, h->cfg.rank, h->dims.PF / 4, whole, 1, h->epoch[1], h->comm_timeout_ns,
                  const_cast<unsigned*>(comm_host_error(h))));
  return DQN_OK;
}

// One whole train step: forward + backward, gradient exchange, Adam -- with the exchange of the W2 gradient overlapped
// with the rest of backward.  world = 1: no exchange.  Needs the peer-memory windows (dqn_lb_comm_*)
// when world > 1; with an external collective (NCCL) the caller uses forward_backward / its all-reduce / apply instead.
DQN_API int dqn_lb_train_step(dqn_lb_handle* h, const int64_t* idx, int32_t debug) {
  if (!h) return lbfail(DQN_E_INVALID, "handle is NULL");
  const int W = h->cfg.world;
  if (W == 1) {
    if (int rc = dqn_lb_forward_backward(h, idx, debug)) return rc;
    return dqn_lb_apply(h);
  }
  if (!h->connected) return lbfail(DQN_E_INVALID, "dqn_lb_train_step: dqn_lb_comm_connect has not been called");
  if (int rc = comm_failed(h)) return rc;
  CU(cudaSetDevice(h->cfg.device));
  if (!h->comm_stream) {
    CU(cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&h->ev_dw2, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_comm, cudaEventDisableTiming));
  }
  const long long size = lb_size(h);
  if (size == 0) return lbfail(DQN_E_INVALID, "dqn_lb_train_step: the replay ring is empty");
  const int B = h->dims.B;
  if (idx) {
    for (int i = 0; i < B; ++i) if (idx[i] < 0 || idx[i] >= size) return lbfail(DQN_E_INVALID, "dqn_lb_train_step: index outside [0, size)");
    if ((size_t)B * 8 > kLbStage) return lbfail(DQN_E_INVALID, "index block exceeds the staging buffer");
    CU(cudaMemcpyAsync(h->ws.idx, idx, (size_t)B * 8, cudaMemcpyHostToDevice, h->stream));
  } else {
    CU(launch_philox_indices(h->stream, h->ws.idx, B, h->cfg.seed, 0, h->train_steps, size, h->cfg.rank * B));
  }
  CU(launch_replay_gather(h->stream, h->ring, ring_dims(h), 0, h->ws.idx, 0, 0, 0, size, B, h->ws.s, h->ws.a, h->ws.r, h->ws.s2, h->ws.done));
  h->taps.enabled = debug ? 1 : 0;
  const float inv = 1.0f / ((float)B * (float)W);
  CU(lb_forward_backward(h->stream, h->dims, h->ws, h->cfg.gamma, inv, h->cfg.gemm_mode, h->loss_kind, h->taps, h->ev_dw2));
  const LbDims& d = h->dims;
  const int offW2 = d.D * d.H1 + d.H1, nW2 = d.H1 * d.H2;          // both multiples of 4 (hidden widths are multiples of 256)
  unsigned* herr = const_cast<unsigned*>(comm_host_error(h));
  // channel 0, second stream: the W2 gradient, as soon as it is final
  CU(cudaStreamWaitEvent(h->comm_stream, h->ev_dw2, 0));
  CommRange big = {{offW2 / 4, 0}, {nW2 / 4, 0}};
  h->epoch[0] += 1;
  CU(lb_allreduce(h->comm_stream, h->peers, W, h->cfg.rank, d.PF / 4, big, 0, h->epoch[0], h->comm_timeout_ns, herr));
  CU(cudaEventRecord(h->ev_comm, h->comm_stream));
  // channel 1, main stream: everything else (W1, b1 | b2, heads, loss), behind the dW1 pass
  CommRange rest = {{0, (offW2 + nW2) / 4}, {offW2 / 4, d.PF / 4 - (offW2 + nW2) / 4}};
  h->epoch[1] += 1;
  CU(lb_allreduce(h->stream, h->peers, W, h->cfg.rank, d.PF / 4, rest, 1, h->epoch[1], h->comm_timeout_ns, herr));
  CU(cudaStreamWaitEvent(h->stream, h->ev_comm, 0));
  return dqn_lb_apply(h);
}

DQN_API int dqn_lb_apply(dqn_lb_handle* h) {
  if (!h) return lbfail(DQN_E_INVALID, "handle is NULL");
  if (int rc = comm_failed(h)) return rc;
  CU(cudaSetDevice(h->cfg.device));
  const dqn_lb_config& c = h->cfg;
  const int t = h->adam_count == 0x7fffffff ? h->adam_count : h->adam_count + 1;
  // optax bias correction, decay**count correctly rounded to fp32 (oracle pow_f32)
  const float c1 = 1.0f - (float)pow((double)c.b1, (double)t);
  const float c2 = 1.0f - (float)pow((double)c.b2, (double)t);
  const float wd = c.opt_kind == DQN_OPT_ADAMW ? c.weight_decay : 0.f;
  CU(lb_adam(h->stream, h->dims, h->ws, c.b1, c.b2, c1, c2, c.eps, c.eps_root, c.lr, wd,
             h->window ? &comm_flags(h->window, h->dims.PF / 4)->error : nullptr));
  h->adam_count = t;
  h->train_steps += 1;
  return DQN_OK;
}

DQN_API int dqn_lb_sync_target(dqn_lb_handle* h) {
  if (!h) return lbfail(DQN_E_INVALID, "handle is NULL");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaMemcpyAsync(h->ws.theta_t, h->ws.theta, (size_t)h->dims.P * 4, cudaMemcpyDeviceToDevice, h->stream));
  return DQN_OK;
}

DQN_API int dqn_lb_polyak_target(dqn_lb_handle* h, float tau) {
  if (!h) return lbfail(DQN_E_INVALID, "handle is NULL");
  if (!(tau >= 0.f && tau <= 1.f)) return lbfail(DQN_E_INVALID, "dqn_lb_polyak_target: tau must be in [0,1]");
  CU(cudaSetDevice(h->cfg.device));
  CU(lb_polyak(h->stream, h->dims, h->ws, tau));
  return DQN_OK;
}

DQN_API int dqn_lb_set_loss_kind(dqn_lb_handle* h, int32_t kind) {
  if (!h) return lbfail(DQN_E_INVALID, "handle is NULL");
  if (kind != DQN_LOSS_HUBER && kind != DQN_LOSS_L2) return lbfail(DQN_E_INVALID, "dqn_lb_set_loss_kind: kind must be DQN_LOSS_HUBER or DQN_LOSS_L2");
  h->loss_kind = kind;
  return DQN_OK;
}

DQN_API int dqn_lb_get_loss(dqn_lb_handle* h, float* loss_out) {
  if (!h || !loss_out) return lbfail(DQN_E_INVALID, "NULL argument");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaMemcpyAsync(h->pinned, h->ws.grads + h->dims.P, 4, cudaMemcpyDeviceToHost, h->stream));
  if (h->window) CU(cudaMemcpyAsync(h->pinned + 1, &comm_flags(h->window, h->dims.PF / 4)->error, 4, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  *loss_out = h->pinned[0];
  if (h->window && ((const unsigned*)h->pinned)[1]) return lbfail(DQN_E_CUDA, "peer-memory all-reduce timed out waiting for a rank");
  return DQN_OK;
}

DQN_API int dqn_lb_debug_read(dqn_lb_handle* h, int32_t what, void* host_out, uint64_t nbytes) {
  if (!h || !host_out) return lbfail(DQN_E_INVALID, "NULL argument");
  CU(cudaSetDevice(h->cfg.device));
  const void* src = nullptr;
  size_t avail = 0;
  const size_t B = h->dims.B;
  switch (what) {
    case DQN_LB_READ_Q: src = h->ws.Q; avail = 3 * B * h->dims.A * 4; break;
    case DQN_LB_READ_TARGETS: src = h->taps.targets; avail = B * h->dims.A * 4; break;
    case DQN_LB_READ_MAX_ACTIONS: src = h->taps.max_actions; avail = B * 4; break;
    case DQN_LB_READ_GRADS: src = h->ws.grads; avail = (size_t)(h->dims.P + 1) * 4; break;
    case DQN_LB_READ_INDICES: src = h->ws.idx; avail = B * 8; break;
    case DQN_LB_READ_H1: src = h->ws.H1; avail = B * h->dims.H1 * 4; break;
    case DQN_LB_READ_H2: src = h->ws.H2; avail = B * h->dims.H2 * 4; break;
    default: return lbfail(DQN_E_INVALID, "dqn_lb_debug_read: unknown `what`");
  }
  if (nbytes > avail) return lbfail(DQN_E_INVALID, "dqn_lb_debug_read: nbytes exceeds the buffer");
  CU(cudaMemcpyAsync(host_out, src, nbytes, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return DQN_OK;
}

}  // extern "C"
