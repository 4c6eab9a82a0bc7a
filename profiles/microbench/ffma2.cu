// Microbenchmark (B200): issue rate of FFMA vs packed FFMA2 (fma.rn.f32x2) on ONE SM, as a function of
// resident warps.  Decides whether the fused DQN kernel's register tiles should use f32x2.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu ; run: ./ffma2
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}

template <int MODE>
__global__ void k(float* out, int iters, float x, float y) {
  float acc[16];
  unsigned long long acc2[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc2[i] = ((unsigned long long)__float_as_uint(acc[2 * i + 1]) << 32) | __float_as_uint(acc[2 * i]);
  unsigned long long a2 = ((unsigned long long)__float_as_uint(x) << 32) | __float_as_uint(x);
  unsigned long long b2 = ((unsigned long long)__float_as_uint(y) << 32) | __float_as_uint(y);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], x, y);      // 64 FFMA, 16 independent chains
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) ffma2(acc2[i], a2, b2);            // 32 FFMA2 = 64 FMA lanes-ops, 8 chains
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += __uint_as_float((unsigned)acc2[i]) + __uint_as_float((unsigned)(acc2[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) out[4096 + blockIdx.x] = (float)(t1 - t0);
}

int main() {
  float* d;
  cudaMalloc(&d, 1 << 20);
  const int iters = 20000;
  for (int warps : {1, 2, 4, 8, 16, 32}) {
    for (int mode = 0; mode < 2; ++mode) {
      float cyc = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<1, warps * 32>>>(d, iters, 1.0001f, 1e-7f);
        else k<1><<<1, warps * 32>>>(d, iters, 1.0001f, 1e-7f);
        cudaDeviceSynchronize();
        cudaMemcpy(&cyc, d + 4096, 4, cudaMemcpyDeviceToHost);
      }
      const double fma_per_cycle = (double)iters * 64 * 32 * warps / cyc;
      printf("warps=%2d %-5s cycles=%.0f  FMA/cycle/SM=%.1f  warp-instr/cycle/SM=%.2f\n", warps, mode ? "FFMA2" : "FFMA", cyc,
             fma_per_cycle, fma_per_cycle / 32 / (mode ? 2 : 1));
    }
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
