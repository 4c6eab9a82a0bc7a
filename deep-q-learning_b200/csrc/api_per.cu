// api_per.cu -- C ABI of the prioritized-replay sum tree (dqn_per_* in include/dqn_b200.h).
#include <stdio.h>
#include <string.h>

#include <string>

#include "../../include/dqn_b200.h"
#include "common.cuh"
#include "kernels.h"

using namespace dqn;

namespace {
int pfail(int code, const std::string& m) { dqn::set_last_error(m.c_str()); return code; }
#define CU(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      char _b[512];                                                                                \
      snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return pfail(DQN_E_CUDA, _b);                                                                \
    }                                                                                              \
  } while (0)
constexpr size_t kPerStage = 4u << 20;
long long pow2_at_least(long long n) { long long p = 1; while (p < n) p <<= 1; return p; }
}  // namespace

struct dqn_per_handle {
  int device;
  long long capacity, L;
  int levels;
  float alpha, eps;
  uint64_t seed;
  cudaStream_t stream;
  uint8_t* arena;
  bool own_arena;
  float* tree;
  uint8_t* stage;
  float* pinned;
};

extern "C" {

DQN_API int dqn_per_arena_bytes(int64_t capacity, uint64_t* bytes_out) {
  if (capacity < 2 || capacity > (1ll << 30) || !bytes_out) return pfail(DQN_E_INVALID, "dqn_per: capacity must be in [2, 2^30]");
  *bytes_out = (uint64_t)pow2_at_least(capacity) * 2 * 4 + kPerStage + 512;
  return DQN_OK;
}

DQN_API int dqn_per_create(int32_t device, int64_t capacity, float alpha, float eps, uint64_t seed, void* stream, void* arena,
                           uint64_t arena_bytes, dqn_per_handle** out) {
  uint64_t need = 0;
  if (int rc = dqn_per_arena_bytes(capacity, &need)) return rc;
  if (!out || !(alpha >= 0.f) || !(eps >= 0.f)) return pfail(DQN_E_INVALID, "dqn_per_create: bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return pfail(DQN_E_ARCH, "no CUDA device visible: libdqn_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) return pfail(DQN_E_INVALID, "device ordinal out of range");
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return pfail(DQN_E_ARCH, "libdqn_b200 is built for sm_100a (B200) only");
  CU(cudaSetDevice(device));
  dqn_per_handle* h = new dqn_per_handle();
  h->device = device; h->capacity = capacity; h->L = pow2_at_least(capacity);
  h->levels = 0;
  while ((1ll << h->levels) < h->L) ++h->levels;
  h->alpha = alpha; h->eps = eps; h->seed = seed; h->stream = (cudaStream_t)stream;
  if (arena) {
    if (arena_bytes < need || ((uintptr_t)arena & 255)) { delete h; return pfail(DQN_E_INVALID, "arena too small or misaligned"); }
    h->arena = (uint8_t*)arena; h->own_arena = false;
  } else {
    if (cudaMalloc((void**)&h->arena, need) != cudaSuccess) { delete h; return pfail(DQN_E_NOMEM, "cudaMalloc failed"); }
    h->own_arena = true;
  }
  h->tree = (float*)h->arena;
  h->stage = h->arena + (((size_t)h->L * 8 + 255) & ~(size_t)255);
  h->pinned = nullptr;
  cudaError_t e = cudaMallocHost((void**)&h->pinned, 256);
  if (e == cudaSuccess) e = cudaMemsetAsync(h->tree, 0, (size_t)h->L * 8, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) { if (h->pinned) cudaFreeHost(h->pinned); if (h->own_arena) cudaFree(h->arena); delete h; return pfail(DQN_E_CUDA, cudaGetErrorString(e)); }
  *out = h;
  return DQN_OK;
}

DQN_API int dqn_per_destroy(dqn_per_handle* h) {
  if (!h) return DQN_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  if (h->pinned) cudaFreeHost(h->pinned);
  if (h->own_arena) cudaFree(h->arena);
  delete h;
  return DQN_OK;
}

DQN_API int dqn_per_update(dqn_per_handle* h, const int64_t* idx_dev, const float* val_dev, int32_t n, int32_t is_td) {
  if (!h || n < 0 || (n > 0 && (!idx_dev || !val_dev))) return pfail(DQN_E_INVALID, "dqn_per_update: bad argument");
  CU(cudaSetDevice(h->device));
  CU(launch_per_update(h->stream, h->tree, h->L, h->levels, (const long long*)idx_dev, val_dev, n, is_td, h->alpha, h->eps));
  return DQN_OK;
}

DQN_API int dqn_per_update_host(dqn_per_handle* h, const int64_t* idx, const float* val, int32_t n, int32_t is_td) {
  if (!h || n < 0 || (n > 0 && (!idx || !val))) return pfail(DQN_E_INVALID, "dqn_per_update: bad argument");
  if ((size_t)n * 12 > kPerStage) return pfail(DQN_E_INVALID, "dqn_per_update_host: n too large for the staging buffer");
  for (int i = 0; i < n; ++i) if (idx[i] < 0 || idx[i] >= h->capacity) return pfail(DQN_E_INVALID, "dqn_per_update: index out of range");
  CU(cudaSetDevice(h->device));
  long long* di = (long long*)h->stage;
  float* dv = (float*)(h->stage + (((size_t)n * 8 + 255) & ~(size_t)255));
  CU(cudaMemcpyAsync(di, idx, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(dv, val, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
  return dqn_per_update(h, (const int64_t*)di, dv, n, is_td);
}

DQN_API int dqn_per_fill(dqn_per_handle* h, const float* prio_dev, int64_t n) {
  if (!h || !prio_dev || n < 0 || n > h->capacity) return pfail(DQN_E_INVALID, "dqn_per_fill: bad argument");
  CU(cudaSetDevice(h->device));
  CU(cudaMemsetAsync(h->tree + h->L, 0, (size_t)h->L * 4, h->stream));
  CU(cudaMemcpyAsync(h->tree + h->L, prio_dev, (size_t)n * 4, cudaMemcpyDeviceToDevice, h->stream));
  CU(launch_per_rebuild(h->stream, h->tree, h->L));
  return DQN_OK;
}

DQN_API int dqn_per_sample(dqn_per_handle* h, int64_t step, int32_t batch, int64_t* idx_dev, float* prio_dev) {
  if (!h || batch < 0 || (batch > 0 && (!idx_dev || !prio_dev))) return pfail(DQN_E_INVALID, "dqn_per_sample: bad argument");
  CU(cudaSetDevice(h->device));
  CU(launch_per_sample(h->stream, h->tree, h->L, h->levels, batch, h->seed, step, (long long*)idx_dev, prio_dev));
  return DQN_OK;
}

DQN_API int dqn_per_sample_host(dqn_per_handle* h, int64_t step, int32_t batch, int64_t* idx, float* prio) {
  if (!h || batch < 0 || (batch > 0 && (!idx || !prio))) return pfail(DQN_E_INVALID, "dqn_per_sample: bad argument");
  if ((size_t)batch * 12 > kPerStage) return pfail(DQN_E_INVALID, "dqn_per_sample_host: batch too large for the staging buffer");
  CU(cudaSetDevice(h->device));
  long long* di = (long long*)h->stage;
  float* dv = (float*)(h->stage + (((size_t)batch * 8 + 255) & ~(size_t)255));
  if (int rc = dqn_per_sample(h, step, batch, (int64_t*)di, dv)) return rc;
  CU(cudaMemcpyAsync(idx, di, (size_t)batch * 8, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(prio, dv, (size_t)batch * 4, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return DQN_OK;
}

DQN_API int dqn_per_total(dqn_per_handle* h, float* total_out) {
  if (!h || !total_out) return pfail(DQN_E_INVALID, "NULL argument");
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpyAsync(h->pinned, h->tree + 1, 4, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  *total_out = h->pinned[0];
  return DQN_OK;
}

DQN_API int dqn_per_read_nodes(dqn_per_handle* h, int64_t first, int64_t n, float* host_out) {
  if (!h || !host_out || first < 0 || n < 0 || first + n > 2 * h->L) return pfail(DQN_E_INVALID, "dqn_per_read_nodes: range out of bounds");
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpyAsync(host_out, h->tree + first, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return DQN_OK;
}

DQN_API int dqn_per_leaf_base(dqn_per_handle* h, int64_t* leaf_base_out) {
  if (!h || !leaf_base_out) return pfail(DQN_E_INVALID, "NULL argument");
  *leaf_base_out = h->L;
  return DQN_OK;
}

}  // extern "C"
