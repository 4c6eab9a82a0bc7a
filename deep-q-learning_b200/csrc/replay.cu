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

// One thread = one 16-byte chunk of one record.  Reads the caller's SoA arrays (the reference's
// add() argument order), writes the AoS ring at slot (counter + i) % N.
__global__ void __launch_bounds__(256)
replay_store_kernel(uint32_t* __restrict__ ring, long long N, int recw, int D, long long counter, long long n,
                    const float* __restrict__ s, const long long* __restrict__ a, const float* __restrict__ r,
                    const float* __restrict__ s2, const uint8_t* __restrict__ done, AgentCtl* ctl) {
  const int cpr = recw >> 2;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid == 0) ctl->ring_counter = counter + n;   // ReplayBuffer._counter += 1, n times
  const long long total = n * cpr;
  if (gid >= total) return;
  const long long rec = gid / cpr;
  const int c = (int)(gid - rec * cpr);
  const long long pos = (counter + rec) % N;
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
  *reinterpret_cast<uint4*>(ring + pos * recw + 4 * c) = make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void __launch_bounds__(256)
philox_indices_kernel(long long* __restrict__ out, int batch, uint64_t seed, int agent, long long step, long long size, int first) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < batch) out[i] = philox_index(seed, agent, step, first + i, size);
}

// One thread = one 16-byte chunk of one sampled record: a warp reads whole 32-byte sectors of the
// records it touches and writes runs of consecutive words of the SoA outputs.
__global__ void __launch_bounds__(256)
replay_gather_kernel(const uint32_t* __restrict__ ring, int recw, int D, int mode, const long long* __restrict__ idx,
                     uint64_t seed, int agent, long long step, long long size, long long batch,
                     float* __restrict__ s, long long* __restrict__ a, float* __restrict__ r,
                     float* __restrict__ s2, uint8_t* __restrict__ done) {
  const int cpr = recw >> 2;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= batch * cpr) return;
  const long long i = gid / cpr;
  const int c = (int)(gid - i * cpr);
  long long slot;
  if (mode == kIdxExplicit) slot = idx[i];
  else if (mode == kIdxPhilox) slot = philox_index(seed, agent, step, (int)i, size);
  else slot = i;
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(ring + slot * recw + 4 * c));
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int k = 4 * c + j;
    if (k < D) s[i * D + k] = __uint_as_float(w[j]);
    else if (k < 2 * D) s2[i * D + (k - D)] = __uint_as_float(w[j]);
    else if (k == 2 * D)   // 2D is even, so the i64's two words never straddle a 16-byte chunk
      a[i] = (long long)((unsigned long long)w[j] | ((unsigned long long)w[(j + 1) & 3] << 32));
    else if (k == 2 * D + 2) r[i] = __uint_as_float(w[j]);
    else if (k == 2 * D + 3) done[i] = (uint8_t)(w[j] != 0u);
  }
}

// ------------------------------------------------------------------------------------------------
cudaError_t launch_replay_store(cudaStream_t st, uint32_t* ring, const Dims& d, long long counter, long long n,
                                const float* s, const long long* a, const float* r, const float* s2,
                                const uint8_t* done, AgentCtl* ctl) {
  if (n <= 0) return cudaSuccess;
  const long long total = n * (d.recw / 4);
  const unsigned blocks = (unsigned)((total + 255) / 256);
  replay_store_kernel<<<blocks, 256, 0, st>>>(ring, d.N, d.recw, d.D, counter, n, s, a, r, s2, done, ctl);
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
  const unsigned blocks = (unsigned)((total + 255) / 256);
  replay_gather_kernel<<<blocks, 256, 0, st>>>(ring, d.recw, d.D, mode, idx, seed, agent, step, size, batch,
                                               s, a, r, s2, done);
  return cudaGetLastError();
}

}  // namespace dqn
