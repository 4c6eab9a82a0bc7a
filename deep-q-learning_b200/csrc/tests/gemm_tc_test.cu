// Stand-alone check of the tcgen05 3xTF32 GEMMs against the fp32 FFMA2 GEMMs and an fp64 host reference.
// build: make -C deep-q-learning_b200/csrc gemm_tc_test ; run on a B200: ./gemm_tc_test
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../large.h"

using namespace dqn;

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

static int run(int kind, int M, int N, int K, int splitk, bool timing) {
  // operand shapes as they lie in memory
  const bool ta = kind == kGemmTN_SplitK, tb = kind == kGemmNT_ReluMask;
  const int lda = ta ? M : K, ldb = tb ? K : N;
  std::vector<float> hA((size_t)M * K), hB((size_t)K * N), hAux((size_t)M * N), hBias(N);
  for (auto& v : hA) v = frand();
  for (auto& v : hB) v = frand() * 0.1f;
  for (auto& v : hAux) v = frand();
  for (auto& v : hBias) v = frand();
  float *dA, *dB, *dC0, *dC1, *dAux, *dBias;
  cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dB, hB.size() * 4);
  cudaMalloc(&dC0, (size_t)M * N * 4 * splitk); cudaMalloc(&dC1, (size_t)M * N * 4 * splitk);
  cudaMalloc(&dAux, hAux.size() * 4); cudaMalloc(&dBias, N * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dAux, hAux.data(), hAux.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dBias, hBias.data(), N * 4, cudaMemcpyHostToDevice);
  const float* aux = kind == kGemmNN_BiasRelu ? dBias : dAux;
  LbWorkspace ws{};
  const int kind_ffma = kind == kGemmNN_ReluMask ? -1 : kind;
  cudaError_t e0 = kind_ffma < 0 ? cudaMemset(dC0, 0, (size_t)M * N * 4) : lb_gemm_ffma(0, kind, M, N, K, dA, lda, dB, ldb, dC0, N, aux, N, splitk);
  cudaError_t e1 = lb_gemm_tc(0, kind, M, N, K, dA, lda, dB, ldb, dC1, N, aux, N, splitk, ws);
  cudaError_t e2 = cudaDeviceSynchronize();
  if (e0 || e1 || e2) { printf("kind %d: CUDA error %s / %s / %s\n", kind, cudaGetErrorString(e0), cudaGetErrorString(e1), cudaGetErrorString(e2)); return 1; }
  std::vector<float> c0((size_t)M * N * splitk), c1((size_t)M * N * splitk);
  cudaMemcpy(c0.data(), dC0, c0.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(c1.data(), dC1, c1.size() * 4, cudaMemcpyDeviceToHost);
  // sum split-K partials on the host
  std::vector<double> s0((size_t)M * N, 0.0), s1((size_t)M * N, 0.0);
  for (int z = 0; z < splitk; ++z) for (size_t i = 0; i < (size_t)M * N; ++i) { s0[i] += c0[z * (size_t)M * N + i]; s1[i] += c1[z * (size_t)M * N + i]; }
  double maxref = 0, err_tc_ffma = 0, err_tc_64 = 0, err_ffma_64 = 0;
  for (size_t i = 0; i < (size_t)M * N; ++i) maxref = fmax(maxref, fabs(s0[i]));
  for (size_t i = 0; i < (size_t)M * N; ++i) err_tc_ffma = fmax(err_tc_ffma, fabs(s0[i] - s1[i]));
  // fp64 reference on a sample of entries
  for (int t = 0; t < 2000; ++t) {
    const int m = rand() % M, n = rand() % N;
    double acc = 0;
    for (int k = 0; k < K; ++k) {
      const double a = ta ? hA[(size_t)k * M + m] : hA[(size_t)m * K + k];
      const double b = tb ? hB[(size_t)n * K + k] : hB[(size_t)k * N + n];
      acc += a * b;
    }
    if (kind == kGemmNN_BiasRelu) acc = fmax(acc + hBias[n], 0.0);
    if (kind == kGemmNT_ReluMask || kind == kGemmNN_ReluMask) acc = hAux[(size_t)m * N + n] > 0 ? acc : 0.0;
    err_tc_64 = fmax(err_tc_64, fabs(s1[(size_t)m * N + n] - acc));
    err_ffma_64 = fmax(err_ffma_64, fabs(s0[(size_t)m * N + n] - acc));
  }
  printf("kind %d M %d N %d K %d splitk %d: max|C| %.3f  max|tc-ffma| %.3e  max|tc-f64| %.3e  max|ffma-f64| %.3e  -> %s\n", kind, M, N, K, splitk,
         maxref, err_tc_ffma, err_tc_64, err_ffma_64, err_tc_64 <= 4 * err_ffma_64 + 2e-6 * maxref ? "OK" : "MISMATCH");
  if (kind_ffma < 0) {      // no FFMA counterpart of this epilogue: judge the tensor-core result against fp64 alone
    double m1 = 0; for (size_t i = 0; i < (size_t)M * N; ++i) m1 = fmax(m1, fabs(s1[i]));
    printf("   (kind 3) max|C| %.3f  max|tc-f64| %.3e -> %s\n", m1, err_tc_64, err_tc_64 <= 5e-6 * m1 ? "OK" : "MISMATCH");
    cudaFree(dA); cudaFree(dB); cudaFree(dC0); cudaFree(dC1); cudaFree(dAux); cudaFree(dBias);
    return err_tc_64 <= 5e-6 * m1 ? 0 : 1;
  }
  if (timing) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int mode = 0; mode < 2; ++mode) {
      for (int w = 0; w < 2; ++w) mode ? lb_gemm_tc(0, kind, M, N, K, dA, lda, dB, ldb, dC1, N, aux, N, splitk, ws) : lb_gemm_ffma(0, kind, M, N, K, dA, lda, dB, ldb, dC0, N, aux, N, splitk);
      cudaEventRecord(a);
      const int reps = 5;
      for (int r = 0; r < reps; ++r) mode ? lb_gemm_tc(0, kind, M, N, K, dA, lda, dB, ldb, dC1, N, aux, N, splitk, ws) : lb_gemm_ffma(0, kind, M, N, K, dA, lda, dB, ldb, dC0, N, aux, N, splitk);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b); ms /= reps;
      printf("   %s: %.3f ms  %.1f TFLOP/s (fp32-equivalent)\n", mode ? "tcgen05 3xTF32" : "FFMA2 fp32    ", ms, 2.0 * M * N * K / ms / 1e9);
    }
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dC0); cudaFree(dC1); cudaFree(dAux); cudaFree(dBias);
  return err_tc_64 <= 4 * err_ffma_64 + 2e-6 * maxref ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += run(kGemmNN_BiasRelu, 256, 256, 256, 1, false);
  bad += run(kGemmNN_BiasRelu, 512, 384, 512, 1, false);
  bad += run(kGemmNN_ReluMask, 256, 256, 256, 1, false);
  bad += run(kGemmNN_ReluMask, 1024, 512, 1024, 1, false);
  bad += run(kGemmNT_ReluMask, 256, 256, 512, 1, false);
  bad += run(kGemmTN_SplitK, 256, 256, 1024, 2, false);
  bad += run(kGemmNN_BiasRelu, 16384, 1024, 1024, 1, true);
  bad += run(kGemmNT_ReluMask, 16384, 1024, 1024, 1, true);
  bad += run(kGemmTN_SplitK, 1024, 1024, 16384, 4, true);
  bad += run(kGemmTN_SplitK, 1024, 1024, 65536, 16, true);
  bad += run(kGemmTN_SplitK, 1024, 1024, 8192, 16, true);
  bad += run(kGemmNN_BiasRelu, 131072, 1024, 1024, 1, true);
  printf(bad ? "FAILED\n" : "ALL OK\n");
  return bad;
}
