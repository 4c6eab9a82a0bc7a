// replay.cu -- HBM-resident replay ring: vectorised store, index draw, minibatch gather, export.
//
// Replaces ReplayBuffer.add (General/Base/replay_buffer.py:58-65), the index draw + five gathers of
// sample_batch (:77-84) and the array properties (:38-56).  All four kernels are HBM-bound byte
// movers: 16-byte accesses, consecutive lanes on consecutive 16-byte chunks of a record, grids
// sized to the work (they are far shorter than a wave on 148 SMs for reference batch sizes and
// several waves for the 32k..64k-sample stress configs).
#include "common.cuh"
#include "kernels.h"

namespace dqn {

enum { kIdxExplicit = 0, kIdxPhilox = 1, kIdxIdentity = 2 };

// One thread = one 16-byte chunk of kIlp records (strided by the grid, so a warp always covers consecutive chunks of
// consecutive records).  Reads the caller's SoA arrays (the reference's add() argument order), writes the AoS ring at
// slot (counter + i) % N.  All loads of the kIlp records are issued before the first store: the kernel is a byte mover
// whose rate is set by the bytes in flight per SM.  VEC = D % 4 == 0: a chunk is one float4 of s or s', or the meta
// words, or padding -- no per-word branches.
constexpr int kIlp = 4;

template <bool VEC>
__device__ __forceinline__ uint4 record_chunk(int c, long long rec, int D, const float* __restrict__ s, const long long* __restrict__ a,
                                              const float* __restrict__ r, const float* __restrict__ s2, const uint8_t* __restrict__ done) {
  if (VEC) {
    const int q = D >> 2;                       // float4 chunks per observation
    if (c < q) return __ldg(reinterpret_cast<const uint4*>(s + rec * D) + c);
    if (c < 2 * q) return __ldg(reinterpret_cast<const uint4*>(s2 + rec * D) + (c - q));
    if (c == 2 * q) {
      const unsigned long long av = (unsigned long long)__ldg(a + rec);
      return make_uint4((uint32_t)av, (uint32_t)(av >> 32), __float_as_uint(__ldg(r + rec)), __ldg(done + rec) ? 1u : 0u);
    }
    return make_uint4(0u, 0u, 0u, 0u);
  }
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int k = 4 * c + j;
    uint32_t v = 0u;
    if (k < D) v = __float_as_uint(s[rec * D + k]);
    else if (k < 2 * D) v = __float_as_uint(s2[rec * D + (k - D)]);
    else if (k == 2 * D) v = (uint32_t)((unsigned long long)a[rec]);
    else if (k == 2 * D + 1) v = (uint32_t)((unsigned long long)a[rec] >> 32);
    else if (k == 2 * D + 2) v = __float_as_uint(r[rec]);
    else if (k == 2 * D + 3) v = done[rec] ? 1u : 0u;
    w[j] = v;
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

template <bool VEC>
__global__ void __launch_bounds__(256)
replay_store_kernel(uint32_t* __restrict__ ring, long long N, int recw, int D, long long counter, long long n,
                    const float* __restrict__ s, const long long* __restrict__ a, const float* __restrict__ r,
                    const float* __restrict__ s2, const uint8_t* __restrict__ done, AgentCtl* ctl) {
  const int cpr = recw >> 2;
  const long long gid0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (gid0 == 0) ctl->ring_counter = counter + n;   // ReplayBuffer._counter += 1, n times
  const long long total = n * cpr;
  uint4 v[kIlp];
  long long rec[kIlp];
  int c[kIlp];
#pragma unroll
  for (int u = 0; u < kIlp; ++u) {
    const long long gid = gid0 + u * stride;
    rec[u] = gid / cpr;
    c[u] = (int)(gid - rec[u] * cpr);
    if (gid < total) v[u] = record_chunk<VEC>(c[u], rec[u], D, s, a, r, s2, done);
  }
#pragma unroll
  for (int u = 0; u < kIlp; ++u) {
    if (gid0 + u * stride < total) {
      const long long pos = (counter + rec[u]) % N;
      *reinterpret_cast<uint4*>(ring + pos * recw + 4 * c[u]) = v[u];
    }
  }
}

__global__ void __launch_bounds__(256)
philox_indices_kernel(long long* __restrict__ out, int batch, uint64_t seed, int agent, long long step, long long size, int first) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < batch) out[i] = philox_index(seed, agent, step, first + i, size);
}

// One thread = one 16-byte chunk of kIlp sampled records (strided by the grid): a warp reads whole 32-byte sectors of
// the records it touches and writes runs of consecutive words of the SoA outputs.  The kIlp random 16-byte loads are
// issued back to back (the gather is DRAM-latency bound: bytes in flight per SM set the rate).
constexpr int kGatherIlp = 4;

template <bool VEC>
__global__ void __launch_bounds__(256)
replay_gather_kernel(const uint32_t* __restrict__ ring, int recw, int D, int mode, const long long* __restrict__ idx,
                     uint64_t seed, int agent, long long step, long long size, long long batch,
                     float* __restrict__ s, long long* __restrict__ a, float* __restrict__ r,
                     float* __restrict__ s2, uint8_t* __restrict__ done) {
  const int cpr = recw >> 2;
  const long long gid0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long total = batch * cpr;
  uint4 v[kGatherIlp];
  long long row[kGatherIlp];
  int c[kGatherIlp];
#pragma unroll
  for (int u = 0; u < kGatherIlp; ++u) {
    const long long gid = gid0 + u * stride;
    row[u] = gid / cpr;
    c[u] = (int)(gid - row[u] * cpr);
    if (gid < total) {
      long long slot;
      if (mode == kIdxExplicit) slot = idx[row[u]];
      else if (mode == kIdxPhilox) slot = philox_index(seed, agent, step, (int)row[u], size);
      else slot = row[u];
      v[u] = __ldg(reinterpret_cast<const uint4*>(ring + slot * recw + 4 * c[u]));
    }
  }
#pragma unroll
  for (int u = 0; u < kGatherIlp; ++u) {
    if (gid0 + u * stride >= total) continue;
    const long long i = row[u];
    if (VEC) {
      const int q = D >> 2;
      if (c[u] < q) reinterpret_cast<uint4*>(s + i * D)[c[u]] = v[u];
      else if (c[u] < 2 * q) reinterpret_cast<uint4*>(s2 + i * D)[c[u] - q] = v[u];
      else if (c[u] == 2 * q) {
        a[i] = (long long)((unsigned long long)v[u].x | ((unsigned long long)v[u].y << 32));
        r[i] = __uint_as_float(v[u].z);
        done[i] = (uint8_t)(v[u].w != 0u);
      }
      continue;
    }
    const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = 4 * c[u] + j;
      if (k < D) s[i * D + k] = __uint_as_float(w[j]);
      else if (k < 2 * D) s2[i * D + (k - D)] = __uint_as_float(w[j]);
      else if (k == 2 * D)   // 2D is even, so the i64's two words never straddle a 16-byte chunk
        a[i] = (long long)((unsigned long long)w[j] | ((unsigned long long)w[(j + 1) & 3] << 32));
      else if (k == 2 * D + 2) r[i] = __uint_as_float(w[j]);
      else if (k == 2 * D + 3) done[i] = (uint8_t)(w[j] != 0u);
    }
  }
}

// ------------------------------------------------------------------------------------------------
cudaError_t launch_replay_store(cudaStream_t st, uint32_t* ring, const Dims& d, long long counter, long long n,
                                const float* s, const long long* a, const float* r, const float* s2,
                                const uint8_t* done, AgentCtl* ctl) {
  if (n <= 0) return cudaSuccess;
  const long long total = n * (d.recw / 4);
  const unsigned blocks = (unsigned)((total + 256 * kIlp - 1) / (256 * kIlp));
  // 16-byte source loads need D % 4 == 0 and 16-byte aligned observation arrays
  const bool vec = d.D % 4 == 0 && (((uintptr_t)s | (uintptr_t)s2) & 15) == 0;
  if (vec) replay_store_kernel<true><<<blocks, 256, 0, st>>>(ring, d.N, d.recw, d.D, counter, n, s, a, r, s2, done, ctl);
  else replay_store_kernel<false><<<blocks, 256, 0, st>>>(ring, d.N, d.recw, d.D, counter, n, s, a, r, s2, done, ctl);
  return cudaGetLastError();
}

cudaError_t launch_philox_indices(cudaStream_t st, long long* out, int batch, uint64_t seed, int agent,
                                  long long step, long long size, int first) {
  philox_indices_kernel<<<(batch + 255) / 256, 256, 0, st>>>(out, batch, seed, agent, step, size, first);
  return cudaGetLastError();
}

cudaError_t launch_replay_gather(cudaStream_t st, const uint32_t* ring, const Dims& d, int mode, const long long* idx,
                                 uint64_t seed, int agent, long long step, long long size, long long batch,
                                 float* s, long long* a, float* r, float* s2, uint8_t* done) {
  if (batch <= 0) return cudaSuccess;
  const long long total = batch * (d.recw / 4);
  const unsigned blocks = (unsigned)((total + 256 * kGatherIlp - 1) / (256 * kGatherIlp));
  const bool vec = d.D % 4 == 0 && (((uintptr_t)s | (uintptr_t)s2) & 15) == 0;
  if (vec) replay_gather_kernel<true><<<blocks, 256, 0, st>>>(ring, d.recw, d.D, mode, idx, seed, agent, step, size, batch, s, a, r, s2, done);
  else replay_gather_kernel<false><<<blocks, 256, 0, st>>>(ring, d.recw, d.D, mode, idx, seed, agent, step, size, batch, s, a, r, s2, done);
  return cudaGetLastError();
}

}  // namespace dqn
