// gemm_tc.cu -- tcgen05 (5th-gen tensor core) GEMMs for the large-batch path, 3xTF32 error-compensated.
#include "common.cuh"
#include "kernels.h"
#include "large.h"

namespace dqn {

cudaError_t lb_gemm_tc(cudaStream_t st, int kind, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                       float* C, int ldc, const float* aux, int ldaux, int splitk, const LbWorkspace& ws) {
  return cudaErrorNotSupported;   // filled in below once validated against the FFMA path
}

}  // namespace dqn
