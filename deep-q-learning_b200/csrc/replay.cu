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

// Both byte movers work on BLOCKS OF kRB = 256 CONSECUTIVE RECORDS: 384 threads, one thread = one 16-byte chunk of four
// records (consecutive threads on consecutive chunks, so a warp covers consecutive chunks of consecutive records), all
// four loads issued before the first store -- the kernels are DRAM-latency bound and bytes in flight per SM set the rate.
// Index arithmetic is 32-bit inside a block (the chunk -> (record, chunk) split is a division by the compile-time
// chunks-per-record CPR; the ring position wraps by one conditional subtract, not a 64-bit modulo per chunk).
//   VEC = D % 4 == 0: a chunk is one float4 of s or s', or the meta words, or padding -- no per-word branches.
constexpr int kRB = 256;          // records per block
constexpr int kRT = 384;          // threads per block

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

// ReplayBuffer.add x n (replay_buffer.py:58-65): the caller's SoA arrays (the reference's add() argument order) -> AoS
// ring slots (counter + i) % N.
template <int CPR, bool VEC>
__global__ void __launch_bounds__(kRT)
replay_store_kernel(uint32_t* __restrict__ ring, long long N, int D, long long counter, long long n,
                    const float* __restrict__ s, const long long* __restrict__ a, const float* __restrict__ r,
                    const float* __restrict__ s2, const uint8_t* __restrict__ done, AgentCtl* ctl) {
  constexpr int recw = 4 * CPR;
  constexpr int kIter = (kRB * CPR + kRT - 1) / kRT;
  const long long row0 = (long long)blockIdx.x * kRB;
  const int nrows = (int)(n - row0 < kRB ? n - row0 : kRB);
  if (blockIdx.x == 0 && threadIdx.x == 0) ctl->ring_counter = counter + n;   // ReplayBuffer._counter += 1, n times
  const long long base = (counter + row0) % N;      // block-uniform; rows of the block wrap at most once (n <= N)
  uint4 v[kIter];
  int rowl[kIter], c[kIter];
#pragma unroll
  for (int u = 0; u < kIter; ++u) {
    const int g = u * kRT + threadIdx.x;
    rowl[u] = g / CPR;
    c[u] = g - rowl[u] * CPR;
    if (rowl[u] < nrows) v[u] = record_chunk<VEC>(c[u], row0 + rowl[u], D, s, a, r, s2, done);
  }
#pragma unroll
  for (int u = 0; u < kIter; ++u) {
    if (rowl[u] < nrows) {
      long long pos = base + rowl[u];
      if (pos >= N) pos -= N;
      *reinterpret_cast<uint4*>(ring + pos * recw + 4 * c[u]) = v[u];
    }
  }
}

__global__ void __launch_bounds__(256)
philox_indices_kernel(long long* __restrict__ out, int batch, uint64_t seed, int agent, long long step, long long size, int first) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < batch) out[i] = philox_index(seed, agent, step, first + i, size);
}

// sample_batch (replay_buffer.py:77-84): ring -> the reference's five SoA arrays.  Per block of 256 samples:
//   1. the slot of every sample ONCE (explicit index / Philox / identity) into shared memory -- not once per chunk;
//   2. four random 16-byte loads per thread, back to back;
//   3. s / s' chunks go straight out (32-byte runs per record = whole sectors); the meta chunk (action, reward, done) is
//      parked in shared memory;
//   4. the block's 256 actions / rewards / dones leave as 16-byte stores of consecutive values (2 KB + 1 KB + 256 B
//      contiguous).  One thread per record storing 8 + 4 + 1 bytes left partially written sectors behind, and every
//      one of them cost a read-for-ownership: 186 B of DRAM reads per sample where the records account for 128 B
//      (profiles/r1_replay.md).
template <int CPR, bool VEC>
__global__ void __launch_bounds__(kRT)
replay_gather_kernel(const uint32_t* __restrict__ ring, int D, int mode, const long long* __restrict__ idx,
                     uint64_t seed, int agent, long long step, long long size, long long batch,
                     float* __restrict__ s, long long* __restrict__ a, float* __restrict__ r,
                     float* __restrict__ s2, uint8_t* __restrict__ done, int meta_vec) {
  constexpr int recw = 4 * CPR;
  constexpr int kIter = (kRB * CPR + kRT - 1) / kRT;
  __shared__ long long slot_s[kRB];
  __shared__ __align__(16) unsigned long long act_s[kRB];
  __shared__ __align__(16) uint32_t rew_s[kRB];
  __shared__ __align__(16) uint8_t done_s[kRB];
  const int t = threadIdx.x;
  const long long row0 = (long long)blockIdx.x * kRB;
  const int nrows = (int)(batch - row0 < kRB ? batch - row0 : kRB);
  if (t < nrows) {
    long long slot;
    if (mode == kIdxExplicit) slot = idx[row0 + t];
    else if (mode == kIdxPhilox) slot = philox_index(seed, agent, step, (int)(row0 + t), size);
    else slot = row0 + t;
    slot_s[t] = slot;
  }
  __syncthreads();
  uint4 v[kIter];
  int rowl[kIter], c[kIter];
#pragma unroll
  for (int u = 0; u < kIter; ++u) {
    const int g = u * kRT + t;
    rowl[u] = g / CPR;
    c[u] = g - rowl[u] * CPR;
    if (rowl[u] < nrows) v[u] = __ldg(reinterpret_cast<const uint4*>(ring + slot_s[rowl[u]] * recw + 4 * c[u]));
  }
#pragma unroll
  for (int u = 0; u < kIter; ++u) {
    if (rowl[u] >= nrows) continue;
    const long long i = row0 + rowl[u];
    if (VEC) {
      const int q = D >> 2;
      if (c[u] < q) reinterpret_cast<uint4*>(s + i * D)[c[u]] = v[u];
      else if (c[u] < 2 * q) reinterpret_cast<uint4*>(s2 + i * D)[c[u] - q] = v[u];
      else if (c[u] == 2 * q) {
        act_s[rowl[u]] = (unsigned long long)v[u].x | ((unsigned long long)v[u].y << 32);
        rew_s[rowl[u]] = v[u].z;
        done_s[rowl[u]] = (uint8_t)(v[u].w != 0u);
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
        act_s[rowl[u]] = (unsigned long long)w[j] | ((unsigned long long)w[(j + 1) & 3] << 32);
      else if (k == 2 * D + 2) rew_s[rowl[u]] = w[j];
      else if (k == 2 * D + 3) done_s[rowl[u]] = (uint8_t)(w[j] != 0u);
    }
  }
  __syncthreads();
  if (nrows == kRB && meta_vec) {          // full block, 16-byte aligned outputs: whole-sector stores
    if (t < kRB / 2) reinterpret_cast<uint4*>(a + row0)[t] = reinterpret_cast<const uint4*>(act_s)[t];
    else if (t < kRB / 2 + kRB / 4) reinterpret_cast<uint4*>(r + row0)[t - kRB / 2] = reinterpret_cast<const uint4*>(rew_s)[t - kRB / 2];
    else if (t < kRB / 2 + kRB / 4 + kRB / 16) reinterpret_cast<uint4*>(done + row0)[t - kRB / 2 - kRB / 4] = reinterpret_cast<const uint4*>(done_s)[t - kRB / 2 - kRB / 4];
  } else if (t < nrows) {
    a[row0 + t] = (long long)act_s[t];
    r[row0 + t] = __uint_as_float(rew_s[t]);
    done[row0 + t] = done_s[t];
  }
}

// ------------------------------------------------------------------------------------------------
cudaError_t launch_replay_store(cudaStream_t st, uint32_t* ring, const Dims& d, long long counter, long long n,
                                const float* s, const long long* a, const float* r, const float* s2,
                                const uint8_t* done, AgentCtl* ctl) {
  if (n <= 0) return cudaSuccess;
  if (n > d.N) return cudaErrorInvalidValue;        // callers drop the transitions a longer run would overwrite anyway
  const unsigned blocks = (unsigned)((n + kRB - 1) / kRB);
  // 16-byte source loads need D % 4 == 0 and 16-byte aligned observation arrays
  const bool vec = d.D % 4 == 0 && (((uintptr_t)s | (uintptr_t)s2) & 15) == 0;
  const int cpr = d.recw / 4;
#define DQN_STORE(CPR, VEC) replay_store_kernel<CPR, VEC><<<blocks, kRT, 0, st>>>(ring, d.N, d.D, counter, n, s, a, r, s2, done, ctl)
  switch (cpr) {                                    // record_words(): 32-byte multiples, 32 B (D <= 2) ... 160 B (D = 15, 16)
    case 2: if (vec) DQN_STORE(2, true); else DQN_STORE(2, false); break;
    case 4: if (vec) DQN_STORE(4, true); else DQN_STORE(4, false); break;
    case 6: if (vec) DQN_STORE(6, true); else DQN_STORE(6, false); break;
    case 8: if (vec) DQN_STORE(8, true); else DQN_STORE(8, false); break;
    case 10: if (vec) DQN_STORE(10, true); else DQN_STORE(10, false); break;
    default: return cudaErrorInvalidValue;
  }
#undef DQN_STORE
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
  const unsigned blocks = (unsigned)((batch + kRB - 1) / kRB);
  const bool vec = d.D % 4 == 0 && (((uintptr_t)s | (uintptr_t)s2) & 15) == 0;
  const int meta_vec = (((uintptr_t)a | (uintptr_t)r | (uintptr_t)done) & 15) == 0;
  const int cpr = d.recw / 4;
#define DQN_GATHER(CPR, VEC) replay_gather_kernel<CPR, VEC><<<blocks, kRT, 0, st>>>(ring, d.D, mode, idx, seed, agent, step, size, batch, s, a, r, s2, done, meta_vec)
  switch (cpr) {
    case 2: if (vec) DQN_GATHER(2, true); else DQN_GATHER(2, false); break;
    case 4: if (vec) DQN_GATHER(4, true); else DQN_GATHER(4, false); break;
    case 6: if (vec) DQN_GATHER(6, true); else DQN_GATHER(6, false); break;
    case 8: if (vec) DQN_GATHER(8, true); else DQN_GATHER(8, false); break;
    case 10: if (vec) DQN_GATHER(10, true); else DQN_GATHER(10, false); break;
    default: return cudaErrorInvalidValue;
  }
#undef DQN_GATHER
  return cudaGetLastError();
}

}  // namespace dqn
