// Microbenchmark: issue rate of packed FFMA2 (fma.rn.f32x2) against scalar FFMA on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int PACKED>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
    if (PACKED) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        unsigned long long xx, aa, bb;
        asm("mov.b64 %0, {%1,%2};" : "=l"(xx) : "f"(x[i]), "f"(x[i + 1]));
        asm("mov.b64 %0, {%1,%2};" : "=l"(aa) : "f"(a), "f"(a));
        asm("mov.b64 %0, {%1,%2};" : "=l"(bb) : "f"(b), "f"(b));
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(xx) : "l"(xx), "l"(aa), "l"(bb));
        asm("mov.b64 {%0,%1}, %2;" : "=f"(x[i]), "=f"(x[i + 1]) : "l"(xx));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
  const int iters = 20000;
  for (int warps_per_smsp = 1; warps_per_smsp <= 4; warps_per_smsp *= 2) {
    const int blocks = sms * warps_per_smsp / 2;   // 256 threads = 8 warps = 2 per SMSP
    for (int packed = 0; packed < 2; ++packed) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (packed) k<1><<<blocks < sms ? sms : blocks, 256>>>(out, iters, 1.0001f, 1e-4f);
        else k<0><<<blocks < sms ? sms : blocks, 256>>>(out, iters, 1.0001f, 1e-4f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
      }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const int nb = blocks < sms ? sms : blocks;
      const double fma = (double)nb * 256 * 16 * iters;
      printf("blocks/SM %.1f  %s  %.3f ms  %.2f TFMA/s  (%.1f fma/clk/SM at 1.965 GHz)\n", (double)nb / sms,
             packed ? "FFMA2" : "FFMA ", ms, fma / ms / 1e9, fma / (ms * 1e-3) / sms / 1.965e9);
    }
  }
  return 0;
}
