// Base noise of the flow on the device: z ~ N(0, I), drawn with the Philox4x32-10 stream and the
// element -> (subsequence, counter, lane) assignment of torch.randn(..., device="cuda").
//
// Replaces generate/flows/zuko.py:15-16,24-26 (`flow.base.rsample`, the DiagNormal draw of
// zuko's NormalizingFlow.rsample_and_log_prob).  With the same (seed, offset) as torch's CUDA
// generator the tensor is the one torch.randn would return, so "identical seeds" parity with the
// reference holds for the whole step and not only for given z: ATen's normal_ kernel
// (distribution_elementwise_grid_stride_kernel, unroll 4, 256-thread blocks, grid =
// min(SMs * maxThreadsPerSM / 256, ceil(numel / 256))) gives element li to thread li % T of the launch
// (T = 256 * grid) as output (li / T) % 4 of that thread's (li / T / 4)-th curand_normal4 call.  The same
// loop is written here on cuRAND's own device API (curand_init / curand_normal4), which fixes the
// Box-Muller arithmetic (logf, sqrtf, __sincosf) to the one torch uses.
//
// The (seed, offset) pair can be immediate or live in device memory; in the second form
// mfb_philox_advance moves the offset on after the draw, so a captured CUDA graph draws fresh noise at
// every replay without any host involvement.
#include <curand_kernel.h>

#include "common.cuh"

namespace mfb {

constexpr int kRandThreads = 256;

__global__ void __launch_bounds__(kRandThreads)
randn_philox_kernel(float* __restrict__ out, int64_t numel, uint64_t seed, uint64_t offset,
                    const uint64_t* __restrict__ state) {
  pdl_enter();
  if (state) {
    seed = state[0];
    offset = state[1];
  }
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)blockDim.x * gridDim.x;
  curandStatePhilox4_32_10_t st;
  curand_init(seed, idx, offset, &st);
  const int64_t rounded = ((numel - 1) / (total * 4) + 1) * total * 4;
  for (int64_t li = idx; li < rounded; li += total * 4) {
    const float4 r = curand_normal4(&st);
    const float rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const int64_t e = li + total * ii;
      if (e < numel) out[e] = rr[ii] * 1.0f + 0.0f;   // at::transformation::normal(rand, mean = 0, std = 1)
    }
  }
}

__global__ void philox_advance_kernel(uint64_t* state, uint64_t inc) {
  pdl_enter();
  state[1] += inc;
}

// launch shape of ATen's calc_execution_policy for `numel` elements on the current device
static int randn_grid(int64_t numel, uint64_t* counter_offset) {
  int dev = 0, sms = 148, tps = 2048;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&tps, cudaDevAttrMaxThreadsPerMultiProcessor, dev);
  const uint64_t blocks = ((uint64_t)numel + kRandThreads - 1) / kRandThreads;
  uint64_t grid = (uint64_t)sms * (uint64_t)(tps / kRandThreads);
  if (blocks < grid) grid = blocks;
  if (counter_offset) *counter_offset = (((uint64_t)numel - 1) / ((uint64_t)kRandThreads * grid * 4) + 1) * 4;
  return (int)grid;
}

}  // namespace mfb

using namespace mfb;

extern "C" {

int64_t mfb_randn_offset_increment(int64_t numel) {
  if (numel <= 0) return 0;
  uint64_t inc = 0;
  randn_grid(numel, &inc);
  return (int64_t)inc;
}

int mfb_randn_philox(float* out, int64_t numel, uint64_t seed, uint64_t offset, void* stream) {
  MFB_CHECK_ARG(numel >= 0 && (out || numel == 0));
  if (numel == 0) return 0;
  const int grid = randn_grid(numel, nullptr);
  MFB_CUDA(launch_pdl(randn_philox_kernel, dim3(grid), dim3(kRandThreads), 0, (cudaStream_t)stream, out, numel, seed, offset,
                      (const uint64_t*)nullptr));
  return launch_status();
}

int mfb_randn_philox_state(float* out, int64_t numel, uint64_t* state, int advance, void* stream) {
  MFB_CHECK_ARG(numel >= 0 && (out || numel == 0) && state);
  if (numel == 0) return 0;
  uint64_t inc = 0;
  const int grid = randn_grid(numel, &inc);
  MFB_CUDA(launch_pdl(randn_philox_kernel, dim3(grid), dim3(kRandThreads), 0, (cudaStream_t)stream, out, numel, (uint64_t)0,
                      (uint64_t)0, (const uint64_t*)state));
  int rc = launch_status();
  if (rc) return rc;
  if (advance) {
    MFB_CUDA(launch_pdl(philox_advance_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, state, inc));
    rc = launch_status();
  }
  return rc;
}

}  // extern "C"
