// comm_p2p.cu -- gradient all-reduce of the large-batch data-parallel step as ONE kernel over NVLink peer memory.
//
// The loss of q_learning_functions.py:36 is a mean over the batch, so the gradients of the batch shards add up
// (SURVEY 8e).  Every rank keeps its gradient (P floats + the loss share) in a "window" that all ranks of the node
// can address (cudaIpc handles between processes; raw pointers inside one process).  The kernel is a two-shot
// all-reduce with the all-gather folded into the reduce:
//
//   phase 0   block 0 tells every rank "my gradient is complete" (release store into the peer's flag array);
//             every block waits until all ranks have said so (acquire loads of its OWN flags -- local polling);
//   phase 1   rank r owns slice r of the vector: it loads slice r from every rank over NVLink (16-byte loads),
//             sums them in rank order 0..W-1 and stores the sum into slice r of EVERY rank's window.  One reducer
//             per element and a fixed order: the result is deterministic and bit-identical on all replicas,
//             which is what keeps the replicated Adam states in lock step;
//   phase 2   the last block of the grid to finish tells every rank "slice r is in place" and waits until every
//             rank has said so; when the kernel retires, the whole window holds the global sum.
//
// One launch covers a RANGE of the vector on one of two flag CHANNELS, so that two exchanges can be in flight at once:
// dqn_lb_train_step sends the W2 gradient (4 MB of the 4.26 MB) on channel 0 from a second stream as soon as the dW2
// GEMM's split-K partials are reduced -- the transfer then runs under the dh1 GEMM and the dW1 pass -- and the small
// remainder (W1, biases, heads, loss) on channel 1 behind them.
//
// Flags carry the step number (monotonic), so nothing is ever reset.  Every spin is bounded (wall-clock time-out,
// DQN_B200_COMM_TIMEOUT_MS, default 20 s).  A rank that never shows up does not hang the GPU and does not corrupt
// training either: a block that times out in phase 0 raises the error flag (in the window and in a word of mapped host
// memory the C ABI polls) and LEAVES -- it neither sums stale peer gradients nor stores into anybody's window; a
// time-out in phase 2 raises the same flag.  The Adam kernel behind this one is predicated on the flag (mlp_large.cu),
// so a failed exchange never reaches the parameters, and the next dqn_lb_allreduce / dqn_lb_apply call returns
// DQN_E_CUDA.  The flag is sticky: the replicas are no longer in step and the caller must rebuild the group.
#include "common.cuh"
#include "large.h"

namespace dqn {

namespace {

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer4(const float4* p) {   // never served from a stale L1 line
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer4(float4* p, const float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// wait until flag >= epoch (wrap-safe); false on timeout
__device__ __forceinline__ bool wait_flag(const unsigned* flag, unsigned epoch, unsigned long long timeout_ns) {
  const unsigned long long t0 = globaltimer_ns();
  for (unsigned spin = 0;; ++spin) {
    if ((int)(ld_acquire_sys(flag) - epoch) >= 0) return true;
    if ((spin & 63u) == 63u && globaltimer_ns() - t0 > timeout_ns) return false;
    __nanosleep(100);
  }
}

template <int W>
__global__ void __launch_bounds__(256) lb_allreduce_kernel(const CommPeers peers, int rank, int n4, const CommRange range, int ch, unsigned epoch,
                                                           unsigned long long timeout_ns, unsigned* host_error) {
  CommFlags* const mine = comm_flags(peers.win[rank], n4);
  const int t = threadIdx.x;
  __shared__ int s_last, s_fail;
  if (t == 0) s_fail = 0;
  __syncthreads();

  // ---- phase 0 ----
  if (blockIdx.x == 0 && t < W) {
    __threadfence_system();
    st_release_sys(&comm_flags(peers.win[t], n4)->ready[ch][rank], epoch);
  }
  if (t < W && (mine->error || !wait_flag(&mine->ready[ch][t], epoch, timeout_ns))) {
    mine->error = 1; *host_error = 1u; s_fail = 1;
  }
  __syncthreads();
  if (s_fail) return;                    // nothing is summed and nothing is stored anywhere; Adam is predicated on the flag

  // ---- phase 1: reduce my slice, broadcast it ----
  const int total = range.n4[0] + range.n4[1];
  const int chunk = (total + W - 1) / W;
  const int lo = rank * chunk, hi = min(total, lo + chunk);
  for (int j = lo + blockIdx.x * blockDim.x + t; j < hi; j += gridDim.x * blockDim.x) {
    const int i = j < range.n4[0] ? range.lo4[0] + j : range.lo4[1] + (j - range.n4[0]);     // position in the gradient vector
    float4 acc = ld_peer4(reinterpret_cast<const float4*>(peers.win[0]) + i);
#pragma unroll
    for (int q = 1; q < W; ++q) {
      const float4 v = ld_peer4(reinterpret_cast<const float4*>(peers.win[q]) + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
#pragma unroll
    for (int q = 0; q < W; ++q) st_peer4(reinterpret_cast<float4*>(peers.win[q]) + i, acc);
  }

  // ---- phase 2 ----
  __threadfence_system();
  __syncthreads();
  if (t == 0) s_last = atomicAdd(&mine->blocks_done[ch], 1u) == gridDim.x - 1;
  __syncthreads();
  if (s_last) {
    if (t == 0) mine->blocks_done[ch] = 0;
    if (t < W) {
      st_release_sys(&comm_flags(peers.win[t], n4)->done[ch][rank], epoch);
      if (!wait_flag(&mine->done[ch][t], epoch, timeout_ns)) { mine->error = 1; *host_error = 1u; }
    }
  }
}

}  // namespace

cudaError_t lb_allreduce(cudaStream_t st, const CommPeers& peers, int world, int rank, int n4, const CommRange& range, int channel,
                         unsigned epoch, unsigned long long timeout_ns, unsigned* host_error) {
  if (channel < 0 || channel >= kCommChannels) return cudaErrorInvalidValue;
  const int chunk = (range.n4[0] + range.n4[1] + world - 1) / world;
  int grid = (chunk + 255) / 256;
  grid = grid < 1 ? 1 : (grid > 96 ? 96 : grid);     // all blocks co-resident on any B200 (they poll each other's progress)
  switch (world) {
    case 2: lb_allreduce_kernel<2><<<grid, 256, 0, st>>>(peers, rank, n4, range, channel, epoch, timeout_ns, host_error); break;
    case 4: lb_allreduce_kernel<4><<<grid, 256, 0, st>>>(peers, rank, n4, range, channel, epoch, timeout_ns, host_error); break;
    case 8: lb_allreduce_kernel<8><<<grid, 256, 0, st>>>(peers, rank, n4, range, channel, epoch, timeout_ns, host_error); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace dqn
