// large.h -- internal declarations of the large-batch data-parallel path (mlp_large.cu, gemm_tc.cu, api_lb.cu).
#pragma once
#include "common.cuh"

namespace dqn {

enum { kEpiSplitK = 0, kEpiBiasRelu = 1, kEpiReluMask = 2 };
enum { kGemmNN_BiasRelu = 0, kGemmNT_ReluMask = 1, kGemmTN_SplitK = 2, kGemmNN_ReluMask = 3 };
enum { kGemmModeFFMA = 0, kGemmModeTC3xTF32 = 1 };

struct LbDims {
  int D, A, H1, H2;
  int B;          // local batch rows (multiple of 128)
  int P, PF;      // flat parameter count (+1 slot for the loss lives at grads[P]); PF = padded stride
  int recw;
  long long N;    // ring slots
};

struct LbTaps {
  float* targets;     // [B][A]
  int* max_actions;   // [B]
  int enabled;
};

// rows per block of the two column-sum passes of the backward (dh2, dW1): 128, or 32 for small local batches -- at 8192 rows
// (the 8-GPU shard of configs[3]) 128-row chunks are only 256 blocks, too few loads in flight to reach HBM speed
inline int lb_row_chunk(long long B) { return B <= 16384 ? 32 : 128; }

struct LbWorkspace {
  float *theta, *theta_t, *mu, *nu, *grads;       // flat layout, PF floats each (grads has the loss at [P])
  float *s, *r, *s2; long long* a; uint8_t* done; long long* idx;   // gathered local batch (SoA)
  float *H1, *H2;      // [3B][H1], [3B][H2]   rows: (theta,s) | (theta,s') | (theta^-,s')
  float *Q;            // [3B][A]
  float *dhd;          // [B][8]
  float *dH2, *dH1;    // [B][H2], [B][H1]
  float *partial;      // [nblk][16]  targets-kernel block partials
  float *colpart;      // column-sum partials  [B / lb_row_chunk(B)][max((2+kMaxA)*H2, (D+1)*H1)]
  float *colred;       // [(2+kMaxA)*H2]
  float *gemmpart;     // split-K partials [16][H1*H2]
  float *tc_scratch;   // tcgen05 path: W2^T [H2][H1] for the dh1 GEMM (so that it runs in the faster NN form)
};

// ---- peer-memory gradient all-reduce (comm_p2p.cu) ----
constexpr int kMaxWorld = 8;
struct CommPeers { float* win[kMaxWorld]; };          // every rank's window, addressable from this GPU
constexpr int kCommChannels = 2;                       // independent exchanges that may be in flight at once (one per stream)
struct CommFlags {                                     // lives behind the n4 * 4 floats of a window
  unsigned ready[kCommChannels][kMaxWorld];            // ready[c][q] = epoch once rank q's part of channel c is complete
  unsigned done[kCommChannels][kMaxWorld];             // done[c][q]  = epoch once rank q has broadcast its slice
  unsigned blocks_done[kCommChannels], error;
};
// The part of the gradient vector one exchange covers: up to two ranges of float4 indices, treated as one index space.
struct CommRange { int lo4[2], n4[2]; };
__host__ __device__ inline CommFlags* comm_flags(float* win, int n4) { return reinterpret_cast<CommFlags*>(win + 4 * (size_t)n4); }
inline size_t comm_window_bytes(int n4) { return (((size_t)n4 * 16 + sizeof(CommFlags)) + 255) / 256 * 256; }
// host_error: a word of mapped pinned host memory that mirrors CommFlags.error (polled by the C ABI without a sync)
cudaError_t lb_allreduce(cudaStream_t st, const CommPeers& peers, int world, int rank, int n4, const CommRange& range, int channel,
                         unsigned epoch, unsigned long long timeout_ns, unsigned* host_error);

cudaError_t lb_gemm_ffma(cudaStream_t st, int kind, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                         float* C, int ldc, const float* aux, int ldaux, int splitk);
cudaError_t lb_gemm_tc(cudaStream_t st, int kind, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                       float* C, int ldc, const float* aux, int ldaux, int splitk, const LbWorkspace& ws);
// h2 = relu(A . B + bias) with the dueling head's partial sums evaluated in the epilogue (gemm_tc.cu: gemm_tc3_kernel<.., HEAD>):
// hp[(column tile * M + row) * 8 + c]; only rows < store_rows of C are written.  cudaErrorNotSupported when the shape / alignment
// does not fit the SM-pair TMA kernel (the caller falls back to the separate head pass).
cudaError_t lb_gemm_tc_nn_head(cudaStream_t st, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                               const float* bias, const float* head_w, int head_a, float* hp, int store_rows);
// dispatcher (gemm_mode: kGemmModeFFMA / kGemmModeTC3xTF32)
cudaError_t lb_gemm(cudaStream_t st, int gemm_mode, int kind, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                    float* C, int ldc, const float* aux, int ldaux, int splitk, const LbWorkspace& ws);
// after_dw2: nullptr, or an event recorded on `st` as soon as the W2 gradient (the 4 MB bulk of the vector) is final -- the
// caller starts exchanging it on another stream while dh1 / dW1 are still being computed
cudaError_t lb_forward_backward(cudaStream_t st, const LbDims& d, const LbWorkspace& ws, float gamma, float inv_global_batch,
                                int gemm_mode, int loss_kind, const LbTaps& taps, cudaEvent_t after_dw2 = nullptr);
cudaError_t lb_polyak(cudaStream_t st, const LbDims& d, const LbWorkspace& ws, float tau);
// comm_error: nullptr, or the window's CommFlags.error -- a non-zero value turns the update into a no-op
cudaError_t lb_adam(cudaStream_t st, const LbDims& d, const LbWorkspace& ws, float b1, float b2, float c1, float c2,
                    float eps, float eps_root, float lr, float wd, const unsigned* comm_error);

}  // namespace dqn
