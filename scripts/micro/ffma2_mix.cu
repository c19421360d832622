// Microbenchmark: does a packed FFMA2 (fma.rn.f32x2) take ONE issue slot or two?  Each iteration runs 16 fp32 FMAs
// (16 FFMA or 8 FFMA2) interleaved with 16 independent LOP3 (ALU pipe).  If the packed form frees issue slots the
// packed loop needs 24 issue cycles per iteration against 32.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_mix ffma2_mix.cu && ./ffma2_mix
#include <cstdio>
#include <cuda_runtime.h>

template <int PACKED, int NALU>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b, unsigned m) {
  float x[16];
  unsigned y[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { x[i] = threadIdx.x * 1e-3f + i; y[i] = threadIdx.x * 77u + i; }
  for (int it = 0; it < iters; ++it) {
    if (PACKED) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        unsigned long long xx, aa, bb;
        asm volatile("mov.b64 %0, {%1,%2};" : "=l"(xx) : "f"(x[i]), "f"(x[i + 1]));
        asm volatile("mov.b64 %0, {%1,%2};" : "=l"(aa) : "f"(a), "f"(a));
        asm volatile("mov.b64 %0, {%1,%2};" : "=l"(bb) : "f"(b), "f"(b));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(xx) : "l"(xx), "l"(aa), "l"(bb));
        asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(x[i]), "=f"(x[i + 1]) : "l"(xx));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
    }
#pragma unroll
    for (int i = 0; i < NALU; ++i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(m), "r"(it));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i] + (float)y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int PACKED, int NALU>
void run(float* out, int sms, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int nb = sms * 2;   // 16 warps per SM = 4 per scheduler
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    k<PACKED, NALU><<<nb, 256>>>(out, iters, 1.0001f, 1e-4f, 0x5a5a5a5au);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
  }
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  // cycles per iteration per scheduler: 4 warps share one scheduler
  const double cyc = ms * 1e-3 * 1.965e9 / iters / 4.0;
  printf("%s + %2d LOP3: %.3f ms, %.1f scheduler cycles per warp-iteration (at 1.965 GHz)\n", PACKED ? " 8 FFMA2" : "16 FFMA ", NALU, ms, cyc);
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
  const int iters = 20000;
  run<0, 0>(out, sms, iters);
  run<1, 0>(out, sms, iters);
  run<0, 8>(out, sms, iters);
  run<1, 8>(out, sms, iters);
  run<0, 16>(out, sms, iters);
  run<1, 16>(out, sms, iters);
  return 0;
}
