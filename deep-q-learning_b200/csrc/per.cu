// per.cu -- prioritized replay (sum-tree sampler + priority update), BASELINE configs[4].
//
// The reference has NO prioritized replay (General/Base/replay_buffer.py:68-85 is uniform), so this
// subsystem is self-specified (Schaul et al. 2016, proportional variant) and its oracle (oracle/per_oracle.py)
// restates THIS file, not the reference: parity unpinned.
//
//   tree   f32 [2 * L]   L = leaf count rounded up to a power of two; node 1 = root, leaves at [L, 2L)
//   parent = fl32(left + right)  -- always recomputed from its two children, never delta-updated, so the tree is a
//   deterministic function of the leaves (bit-exact against the numpy oracle) and never drifts.
//
// sample: stratified -- sample i draws u in [i, i+1) * total / B (Philox4x32-10, 24-bit mantissa uniform) and
//         walks log2(L) levels: go left if u < left, else u -= left and go right.  One 4-byte load per level.
// update: p = (|td| + eps)^alpha into the leaves, then one pass per level recomputes the touched parents
//         (duplicate indices recompute the same value: benign).
#include "common.cuh"
#include "kernels.h"

namespace dqn {

__global__ void __launch_bounds__(256)
per_set_leaves_kernel(float* __restrict__ tree, long long L, const long long* __restrict__ idx, const float* __restrict__ val,
                      int n, int is_td, float alpha, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float p = val[i];
  if (is_td) p = powf(fabsf(p) + eps, alpha);
  tree[L + idx[i]] = p;
}

__global__ void __launch_bounds__(256)
per_propagate_kernel(float* __restrict__ tree, long long L, const long long* __restrict__ idx, int n, int shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long node = (L + idx[i]) >> shift;
  tree[node] = tree[2 * node] + tree[2 * node + 1];
}

// full rebuild of one level (used after bulk initialisation): node in [first, 2*first)
__global__ void __launch_bounds__(256)
per_rebuild_level_kernel(float* __restrict__ tree, long long first) {
  const long long node = first + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (node < 2 * first) tree[node] = tree[2 * node] + tree[2 * node + 1];
}

__global__ void __launch_bounds__(256)
per_sample_kernel(const float* __restrict__ tree, long long L, int levels, int batch, uint64_t seed, long long step,
                  long long* __restrict__ idx_out, float* __restrict__ prio_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch) return;
  uint32_t o[4];
  philox4x32_10((uint32_t)i, (uint32_t)step, (uint32_t)((uint64_t)step >> 32), 0x50455221u /* 'PER!' stream */,
                (uint32_t)seed, (uint32_t)(seed >> 32), o);
  const float r = (float)(o[0] >> 8) * (1.0f / 16777216.0f);        // [0, 1), 24 bits
  const float total = tree[1];
  const float seg = total / (float)batch;
  float u = ((float)i + r) * seg;
  long long node = 1;
  for (int l = 0; l < levels; ++l) {
    const float left = __ldg(tree + 2 * node);
    if (u < left) node = 2 * node;
    else { u -= left; node = 2 * node + 1; }
  }
  idx_out[i] = node - L;
  prio_out[i] = __ldg(tree + node);
}

cudaError_t launch_per_update(cudaStream_t st, float* tree, long long L, int levels, const long long* idx, const float* val, int n,
                              int is_td, float alpha, float eps) {
  if (n <= 0) return cudaSuccess;
  const int blocks = (n + 255) / 256;
  per_set_leaves_kernel<<<blocks, 256, 0, st>>>(tree, L, idx, val, n, is_td, alpha, eps);
  for (int s = 1; s <= levels; ++s) per_propagate_kernel<<<blocks, 256, 0, st>>>(tree, L, idx, n, s);
  return cudaGetLastError();
}

cudaError_t launch_per_rebuild(cudaStream_t st, float* tree, long long L) {
  for (long long first = L / 2; first >= 1; first /= 2) {
    const unsigned blocks = (unsigned)((first + 255) / 256);
    per_rebuild_level_kernel<<<blocks, 256, 0, st>>>(tree, first);
  }
  return cudaGetLastError();
}

cudaError_t launch_per_sample(cudaStream_t st, const float* tree, long long L, int levels, int batch, uint64_t seed, long long step,
                              long long* idx_out, float* prio_out) {
  if (batch <= 0) return cudaSuccess;
  per_sample_kernel<<<(batch + 255) / 256, 256, 0, st>>>(tree, L, levels, batch, seed, step, idx_out, prio_out);
  return cudaGetLastError();
}

}  // namespace dqn
