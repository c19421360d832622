// Micro-benchmark: how fast can 148 CTAs x 384 threads write a tile-major workspace
// ([tile][row][128 particles] fp32) with (a) one 4-byte store per lane per row, (b) 16-byte stores per lane
// (4 particles of a row per lane), (c) TMA bulk stores from shared memory.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kRows = 576;   // rows per tile (gphi 384 + acts 192)

__global__ void __launch_bounds__(384) st32(float* out, int64_t ntiles) {
  const int wg = threadIdx.x >> 7, t = threadIdx.x & 127;
  for (int64_t tile = (int64_t)blockIdx.x * 3 + wg; tile < ntiles; tile += (int64_t)gridDim.x * 3) {
    float* base = out + (size_t)tile * kRows * 128 + t;
    const float v = (float)tile;
#pragma unroll 16
    for (int r = 0; r < kRows; ++r) base[(size_t)r * 128] = v + r;
  }
}

__global__ void __launch_bounds__(384) st128(float* out, int64_t ntiles) {
  const int wg = threadIdx.x >> 7, t = threadIdx.x & 127;
  const int warp = t >> 5, lane = t & 31;
  for (int64_t tile = (int64_t)blockIdx.x * 3 + wg; tile < ntiles; tile += (int64_t)gridDim.x * 3) {
    // warp w writes rows w, w+4, ...: lane covers 4 particles
    float* base = out + (size_t)tile * kRows * 128 + 4 * lane;
    const float v = (float)tile;
#pragma unroll 8
    for (int r = warp; r < kRows; r += 4) *reinterpret_cast<float4*>(base + (size_t)r * 128) = make_float4(v, v + 1, v + 2, v + r);
  }
}

__global__ void __launch_bounds__(384) st_tma(float* out, int64_t ntiles) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int wg = threadIdx.x >> 7, t = threadIdx.x & 127;
  float* buf = reinterpret_cast<float*>(smem) + (size_t)wg * 64 * 128;   // 64 rows x 128 = 32 KB per WG
  for (int64_t tile = (int64_t)blockIdx.x * 3 + wg; tile < ntiles; tile += (int64_t)gridDim.x * 3) {
    for (int c = 0; c < kRows / 64; ++c) {
      // wait until the previous bulk store has read the buffer
      if (t == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      asm volatile("bar.sync %0, 128;" ::"r"(1 + wg));
#pragma unroll 16
      for (int r = 0; r < 64; ++r) buf[r * 128 + t] = (float)tile + r;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync %0, 128;" ::"r"(1 + wg));
      if (t == 0) {
        float* dst = out + ((size_t)tile * kRows + c * 64) * 128;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                     "r"((uint32_t)__cvta_generic_to_shared(buf)), "r"(64 * 128 * 4) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
  }
  if (t == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  const int64_t n = 1000000, ntiles = (n + 127) / 128;
  float* out;
  const size_t bytes = (size_t)ntiles * kRows * 128 * 4;
  cudaMalloc(&out, bytes);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaFuncSetAttribute(st_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 32768);
  for (int which = 0; which < 3; ++which) {
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(a);
      if (which == 0) st32<<<148, 384>>>(out, ntiles);
      else if (which == 1) st128<<<148, 384>>>(out, ntiles);
      else st_tma<<<148, 384, 3 * 32768>>>(out, ntiles);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms;
      cudaEventElapsedTime(&ms, a, b);
      if (ms < best) best = ms;
    }
    printf("%s: %.3f ms  %.2f TB/s  (%s)\n", which == 0 ? "st32 " : which == 1 ? "st128" : "tma  ", best, bytes / best * 1e-9,
           cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
