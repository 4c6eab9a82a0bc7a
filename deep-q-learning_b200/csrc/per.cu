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
// update: p = (|td| + eps)^alpha into the leaves, then the parents of the touched leaves are recomputed, in THREE launches
//         (one launch per level -- 1 + 24 launches of ~4 us each -- was launch-bound: 102 us for 32 768 updates):
//           1. per_set_leaves_kernel      the new leaf values;
//           2. per_bottom_kernel          one WARP per update re-reduces the whole 256-leaf subtree around its leaf from
//                                         the (now final) leaves -- 1 KB, coalesced -- in the tree's own pairwise order
//                                         and stores the eight nodes on the leaf's path.  No thread ever reads another
//                                         thread's intermediate node, so updates that share ancestors cannot race: every
//                                         writer of a node writes the same bits;
//           3. per_top_kernel             the levels above (2^(levels-8) nodes and fewer: 65 536 at 16 M leaves) are
//                                         recomputed in FULL, 1024-input subtrees per CTA in shared memory, the last CTA
//                                         to finish (atomic ticket in the unused tree[0] slot) reduces the remaining top.
//         The result is the same deterministic function of the leaves as the level-by-level pass (IEEE addition is
//         commutative, the pairing is the tree's), which small trees (< 8 levels) still use.
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

// levels [levels - 8, levels): one warp per update.  Lane l holds leaves [8 l, 8 l + 8) of the 256-leaf subtree.
__global__ void __launch_bounds__(256)
per_bottom_kernel(float* __restrict__ tree, long long L, const long long* __restrict__ idx, int n) {
  const int w = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (w >= n) return;
  const long long leaf = idx[w];
  const long long node0 = L + (leaf & ~255ll);                 // first leaf node of the subtree
  const float4 a = __ldcg(reinterpret_cast<const float4*>(tree + node0) + 2 * lane);
  const float4 b = __ldcg(reinterpret_cast<const float4*>(tree + node0) + 2 * lane + 1);
  const float p0 = a.x + a.y, p1 = a.z + a.w, p2 = b.x + b.y, p3 = b.z + b.w;      // level - 1
  const float q0 = p0 + p1, q1 = p2 + p3;                                          // level - 2
  float v = q0 + q1;                                                               // level - 3
  const int o = (int)(leaf & 255), mine = o >> 3, j = o & 7;
  const long long path = L + leaf;
  if (lane == mine) {
    tree[path >> 1] = j < 2 ? p0 : (j < 4 ? p1 : (j < 6 ? p2 : p3));
    tree[path >> 2] = j < 4 ? q0 : q1;
    tree[path >> 3] = v;
  }
#pragma unroll
  for (int r = 0; r < 5; ++r) {
    v += __shfl_xor_sync(0xffffffffu, v, 1 << r);              // both partners now hold their common parent
    if (lane == mine) tree[path >> (4 + r)] = v;
  }
}

// levels [0, U): full recompute from the 2^U nodes of level U.  CTA b reduces the 2^h inputs [b 2^h, (b+1) 2^h), h = min(U, 10),
// in shared memory; if more than one CTA ran, the last one to finish reduces the 2^(U-h) <= 1024 nodes they produced.
__global__ void __launch_bounds__(256)
per_top_kernel(float* __restrict__ tree, int U) {
  __shared__ float buf[1024];
  __shared__ int s_last;
  const int t = threadIdx.x;
  int top = U;                                  // level of the current inputs
  int h = U < 10 ? U : 10;
  long long first = (1ll << top) + ((long long)blockIdx.x << h);
  for (int pass = 0; pass < 2; ++pass) {
    const int nin = 1 << h;
    for (int i = t; i < nin; i += blockDim.x) buf[i] = __ldcg(tree + first + i);
    __syncthreads();
    for (int d = 1; d <= h; ++d) {
      const int nn = nin >> d;
      float val[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) { const int i = t + q * 256; if (i < nn) val[q] = buf[2 * i] + buf[2 * i + 1]; }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int i = t + q * 256;
        if (i < nn) { buf[i] = val[q]; tree[(first >> d) + i] = val[q]; }
      }
      __syncthreads();
    }
    if (pass == 1 || top == h) return;          // reached the root
    // ticket: the last CTA continues with the nodes all CTAs produced (level top - h)
    __threadfence();
    if (t == 0) {
      unsigned* ticket = reinterpret_cast<unsigned*>(tree);      // node 0 is not part of the tree
      s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
      if (s_last) *ticket = 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    top -= h;
    h = top;                                    // <= 10 by construction of the grid
    first = 1ll << top;
  }
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
  const int U = levels - 8;                      // level whose nodes the warp-per-update pass leaves final
  if (levels < 8 || U > 20) {                    // tiny trees / > 2^28 leaves: one pass per level
    for (int s = 1; s <= levels; ++s) per_propagate_kernel<<<blocks, 256, 0, st>>>(tree, L, idx, n, s);
    return cudaGetLastError();
  }
  per_bottom_kernel<<<(unsigned)(((long long)n * 32 + 255) / 256), 256, 0, st>>>(tree, L, idx, n);
  if (U > 0) per_top_kernel<<<U > 10 ? 1u << (U - 10) : 1u, 256, 0, st>>>(tree, U);
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
