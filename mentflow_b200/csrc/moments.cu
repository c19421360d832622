// Particle moments for the entropy estimators.
//
// Replaces entropy.py:58-62 (torch.mean(log_prob), torch.mean(prior.log_prob(x))) with
// prior.py:25-26 (MultivariateNormal(0, s^2 I).log_prob = -|x|^2/(2 s^2) - D log s - D/2 log 2pi)
// and entropy.py:35-38 (torch.cov): one pass producing, in double precision,
//   out[0] = sum logq, out[1] = sum |x|^2, out[2+i] = sum x_i, out[2+d+i*d+j] = sum x_i x_j.
// Two-stage reduction with a fixed order => deterministic.
#include "common.cuh"

namespace mfb {

constexpr int kMomThreads = 256;

template <int D, bool COV>
__global__ void __launch_bounds__(kMomThreads)
moments_kernel(const float* __restrict__ x, const float* __restrict__ logq, int64_t n, double* __restrict__ partial) {
  pdl_enter();
  constexpr int M = COV ? 2 + D + D * D : 2;
  double acc[M];
#pragma unroll
  for (int i = 0; i < M; ++i) acc[i] = 0.0;
  for (int64_t p = (int64_t)blockIdx.x * kMomThreads + threadIdx.x; p < n; p += (int64_t)gridDim.x * kMomThreads) {
    float xr[D];
#pragma unroll
    for (int i = 0; i < D; ++i) xr[i] = x[p * D + i];
    if (logq) acc[0] += (double)logq[p];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < D; ++i) ss = fmaf(xr[i], xr[i], ss);
    acc[1] += (double)ss;
    if (COV) {
#pragma unroll
      for (int i = 0; i < D; ++i) {
        acc[2 + i] += (double)xr[i];
#pragma unroll
        for (int j = 0; j < D; ++j) acc[2 + D + i * D + j] += (double)xr[i] * (double)xr[j];
      }
    }
  }
  __shared__ double red[kMomThreads / 32][M];
#pragma unroll
  for (int i = 0; i < M; ++i) {
    const double s = warp_sum(acc[i]);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][i] = s;
  }
  __syncthreads();
  if (threadIdx.x < M) {
    double s = 0.0;
    for (int w = 0; w < kMomThreads / 32; ++w) s += red[w][threadIdx.x];
    partial[(size_t)blockIdx.x * M + threadIdx.x] = s;
  }
}

__global__ void moments_finish_kernel(const double* __restrict__ partial, int nparts, int m, double* __restrict__ out) {
  pdl_enter();
  // one warp per output: lanes stride over the partials, fixed shuffle tree => deterministic
  const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x >> 5; i < m; i += nwarps) {
    double s = 0.0;
    for (int c = lane; c < nparts; c += 32) s += partial[(size_t)c * m + i];
    s = warp_sum(s);
    if (lane == 0) out[i] = s;
  }
}

static int moments_grid(int64_t n) {
  int64_t g = (n + kMomThreads - 1) / kMomThreads;
  int64_t cap = (int64_t)sm_count() * 4;
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

template <int D>
static int launch_moments(const float* x, const float* logq, int64_t n, int cov, double* out, double* partial,
                          cudaStream_t st) {
  const int grid = moments_grid(n);
  const int m = cov ? 2 + D + D * D : 2;
  if (cov) MFB_CUDA(launch_pdl(moments_kernel<D, true>, dim3(grid), dim3(kMomThreads), 0, st, x, logq, n, partial));
  else MFB_CUDA(launch_pdl(moments_kernel<D, false>, dim3(grid), dim3(kMomThreads), 0, st, x, logq, n, partial));
  int rc = launch_status();
  if (rc) return rc;
  MFB_CUDA(launch_pdl(moments_finish_kernel, dim3(1), dim3(256), 0, st, (const double*)partial, grid, m, out));
  return launch_status();
}

// double <-> (hi, lo) float pairs: lets a few double partial sums ride at the tail of a float32 all-reduce
// (out[i] = hi, out[n + i] = lo; hi + lo = in to 2^-48, and sums of hi and of lo over ranks are formed separately)
__global__ void f64_split_kernel(const double* __restrict__ in, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float hi = (float)in[i];
    out[i] = hi;
    out[n + i] = (float)(in[i] - (double)hi);
  }
}
__global__ void f64_join_kernel(const float* __restrict__ in, int n, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (double)in[i] + (double)in[n + i];
}

// H = a * sums[0] + b * sums[1] - c in double, rounded once to float (replaces four elementwise launches)
__global__ void mc_entropy_kernel(const double* __restrict__ sums, double a, double b, double c, float* __restrict__ h) {
  pdl_enter();
  if (threadIdx.x == 0) h[0] = (float)(fma(a, sums[0], b * sums[1]) - c);
}

// L = H + mu * mean_k D_k: one warp, lanes stride over k, fixed shuffle tree (deterministic)
__global__ void loss_tail_kernel(const float* __restrict__ d, int k, const float* __restrict__ h, float mu,
                                 float* __restrict__ out) {
  pdl_enter();
  float s = 0.f;
  for (int i = threadIdx.x; i < k; i += 32) s += d[i];
  s = warp_sum(s);
  if (threadIdx.x == 0) {
    const float mean = s / (float)k;
    out[1] = mean;
    out[0] = fmaf(mu, mean, h ? h[0] : 0.f);
  }
}

}  // namespace mfb

using namespace mfb;

extern "C" {

int mfb_mc_entropy(const double* sums, double a, double b, double c, float* h, void* stream) {
  MFB_CHECK_ARG(sums && h);
  MFB_CUDA(launch_pdl(mc_entropy_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, sums, a, b, c, h));
  return launch_status();
}

int mfb_loss_tail(const float* d, int k, const float* h, float mu, float* out, void* stream) {
  MFB_CHECK_ARG(d && out && k >= 1);
  MFB_CUDA(launch_pdl(loss_tail_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, d, k, h, mu, out));
  return launch_status();
}

int mfb_f64_split(const double* in, int n, float* out_hi_lo, void* stream) {
  MFB_CHECK_ARG(in && out_hi_lo && n >= 1);
  f64_split_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(in, n, out_hi_lo);
  return launch_status();
}

int mfb_f64_join(const float* in_hi_lo, int n, double* out, void* stream) {
  MFB_CHECK_ARG(in_hi_lo && out && n >= 1);
  f64_join_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(in_hi_lo, n, out);
  return launch_status();
}

int64_t mfb_moments_workspace_bytes(int64_t n, int d) {
  if (d < 1 || d > kMaxDim) return 0;
  return (int64_t)moments_grid(n > 0 ? n : 1) * (2 + d + d * d) * 8;
}

int mfb_moments(const float* x, const float* logq, int64_t n, int d, int with_cov, double* out, void* workspace,
                int64_t workspace_bytes, void* stream) {
  MFB_CHECK_ARG(x && out && workspace && n >= 1 && d >= 1 && d <= kMaxDim);
  if (workspace_bytes < mfb_moments_workspace_bytes(n, d)) return MFB_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  double* partial = (double*)workspace;
  switch (d) {
    case 1: return launch_moments<1>(x, logq, n, with_cov, out, partial, st);
    case 2: return launch_moments<2>(x, logq, n, with_cov, out, partial, st);
    case 3: return launch_moments<3>(x, logq, n, with_cov, out, partial, st);
    case 4: return launch_moments<4>(x, logq, n, with_cov, out, partial, st);
    case 5: return launch_moments<5>(x, logq, n, with_cov, out, partial, st);
    case 6: return launch_moments<6>(x, logq, n, with_cov, out, partial, st);
    case 7: return launch_moments<7>(x, logq, n, with_cov, out, partial, st);
    default: return launch_moments<8>(x, logq, n, with_cov, out, partial, st);
  }
}

}  // extern "C"
