// Neural spline flow layer: masked-MLP conditioner fused with the monotonic rational-quadratic
// spline and its log|det J|.
//
// Replaces, per autoregressive layer (reference call sites generate/flows/zuko.py:24-29 ->
// zuko 1.3.1 MaskedAutoregressiveTransform.meta / MonotonicRQSTransform.call_and_ladj, restated
// in SURVEY.md App. A and oracle/zuko_nsf.py):
//   phi = MaskedMLP(v)            4 masked SGEMMs + 3 ReLU            (zuko/nn.py)
//   soft-clip, 2 softmax, pad, 2 cumsum, exp, searchsorted, 6 gathers, RQ formula, log, where, sum
//
// This file is the fp32 CUDA-core version ("stage A"): one thread per particle, the layer's
// (pre-masked, transposed) weights resident in shared memory for the lifetime of a persistent
// CTA, activations of the current layer in registers, outputs of a layer staged through a
// per-thread column of shared memory (conflict free).  Spline parameters never touch HBM.
#include <math.h>

#include "common.cuh"

namespace mfb {

constexpr int kH = 64;          // hidden units (reference config gen/flow.yaml: hidden_units 64)
constexpr int kPP = 64;         // per-feature parameter block, 3*bins-1 padded to 64
constexpr int kNsfThreads = 256;
constexpr float kBound = 5.0f;  // zuko MonotonicRQSTransform(bound=5.0, slope=1e-3)
constexpr float kClipW = 0.28952965460216789f;  // 2 / |log(1e-3)|
constexpr float kClipD = 0.14476482730108395f;  // 1 / |log(1e-3)|
constexpr float kHalfLog2Pi = 0.91893853320467274f;

struct FeatureOrder {
  int v[kMaxDim];
};

// packed layout of one layer (floats); all weights are pre-masked and stored [in][out]
//   W1t [D][64] | b1 [64] | (Wt_l [64][64] | b_l [64]) x (L-1) | Wout_t [D][64][64] | bout [D][64]
__host__ __device__ inline int64_t nsf_param_floats(int d, int hidden_layers) {
  return (int64_t)d * kH + kH + (int64_t)(hidden_layers - 1) * (kH * kH + kH) + (int64_t)d * kH * kPP +
         (int64_t)d * kPP;
}

// dense 64 -> 8 block: acc[q] = bias[q] + sum_i h[i] * Wt[i][q]   (Wt row stride = 64 floats)
__device__ __forceinline__ void dense8(const float (&h)[kH], const float* __restrict__ wt,
                                       const float* __restrict__ bias, float (&acc)[8]) {
  const float4 b0 = *reinterpret_cast<const float4*>(bias);
  const float4 b1 = *reinterpret_cast<const float4*>(bias + 4);
  acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
  acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
#pragma unroll
  for (int i = 0; i < kH; ++i) {
    const float4 w0 = *reinterpret_cast<const float4*>(wt + i * kH);
    const float4 w1 = *reinterpret_cast<const float4*>(wt + i * kH + 4);
    acc[0] = fmaf(h[i], w0.x, acc[0]); acc[1] = fmaf(h[i], w0.y, acc[1]);
    acc[2] = fmaf(h[i], w0.z, acc[2]); acc[3] = fmaf(h[i], w0.w, acc[3]);
    acc[4] = fmaf(h[i], w1.x, acc[4]); acc[5] = fmaf(h[i], w1.y, acc[5]);
    acc[6] = fmaf(h[i], w1.z, acc[6]); acc[7] = fmaf(h[i], w1.w, acc[7]);
  }
}

// softmax over nb raw parameters held in a strided shared-memory column; leaves
// exp(w - max) in place and returns their sum.
__device__ __forceinline__ float softmax_inplace(float* col, int stride, int nb) {
  float m = -INFINITY;
  for (int j = 0; j < nb; ++j) {
    float w = col[j * stride];
    w = w / (1.0f + kClipW * fabsf(w));
    col[j * stride] = w;
    m = fmaxf(m, w);
  }
  float sum = 0.f;
  for (int j = 0; j < nb; ++j) {
    const float e = expf(col[j * stride] - m);
    col[j * stride] = e;
    sum += e;
  }
  return sum;
}

// One univariate spline: parameters (3*nb-1 raw conditioner outputs) in col[j*stride].
// Returns y and adds log dy/dv to ladj.
//
// Conditioning: zuko differences the cumulative knot arrays (x1 - x0, y1 - y0), which cancels
// catastrophically in narrow bins (widths go down to 1e-3 of the mean).  Here the bin width and
// height are taken directly from the softmax values (dx = 2B W_k, dy = 2B H_k), which is the same
// number in exact arithmetic but accurate to an ulp; only the bin origin comes from the
// cumulative sum.  This puts the kernel closer to the float64 truth than a plain fp32 evaluation.
__device__ __forceinline__ float rq_spline_forward(float* col, int stride, int nb, float v, float& ladj) {
  // horizontal knots + bin search: k = #(knots < v) - 1
  const float sum_w = softmax_inplace(col, stride, nb);
  float cum = 0.f, xl = -kBound, x0 = 0.f, wk = 0.f;
  int kbin = -1;
  for (int j = 0; j < nb; ++j) {
    const float wj = col[j * stride] / sum_w;
    cum += wj;
    const float xr = fmaf(2.0f * kBound, cum, -kBound);
    if (kbin < 0 && xl < v && v <= xr) {
      kbin = j;
      x0 = xl;
      wk = wj;
    }
    xl = xr;
  }
  if (kbin < 0) return v;  // outside [-bound, bound]: identity, ladj += 0
  float* colh = col + nb * stride;
  const float sum_h = softmax_inplace(colh, stride, nb);
  cum = 0.f;
  for (int j = 0; j < kbin; ++j) cum += colh[j * stride] / sum_h;
  const float y0 = fmaf(2.0f * kBound, cum, -kBound);
  const float hk = colh[kbin * stride] / sum_h;
  const float* cold = col + 2 * nb * stride;
  float d0 = 1.0f, d1 = 1.0f;
  if (kbin > 0) {
    const float r = cold[(kbin - 1) * stride];
    d0 = expf(r / (1.0f + kClipD * fabsf(r)));
  }
  if (kbin < nb - 1) {
    const float r = cold[kbin * stride];
    d1 = expf(r / (1.0f + kClipD * fabsf(r)));
  }
  const float dx = 2.0f * kBound * wk, dy = 2.0f * kBound * hk;
  const float s = hk / wk;
  float t = (v - x0) / dx;
  t = fminf(fmaxf(t, 0.0f), 1.0f);
  const float omt = 1.0f - t;
  const float tomt = t * omt;
  const float den = fmaf(d0 + d1 - 2.0f * s, tomt, s);
  const float y = y0 + dy * (s * t * t + d0 * tomt) / den;
  const float jac = s * s * (2.0f * s * tomt + d0 * omt * omt + d1 * t * t) / (den * den);
  ladj += logf(jac);
  return y;
}

template <int D>
__global__ void __launch_bounds__(kNsfThreads, 1)
nsf_layer_fwd_kernel(const float* __restrict__ v, int64_t n, int hidden_layers, int nb,
                     const float* __restrict__ params, int64_t nparams, FeatureOrder order,
                     const float* __restrict__ logq_in, int first_layer, float* __restrict__ y,
                     float* __restrict__ logq_out) {
  extern __shared__ __align__(16) float smem[];
  float* s_par = smem;
  float* s_scr = smem + ((nparams + 3) & ~(int64_t)3);  // [64][kNsfThreads]
  const int tid = threadIdx.x;
  {
    const float4* src = reinterpret_cast<const float4*>(params);
    float4* dst = reinterpret_cast<float4*>(s_par);
    for (int i = tid; i < (int)(nparams >> 2); i += kNsfThreads) dst[i] = src[i];
    for (int i = (int)(nparams & ~(int64_t)3) + tid; i < (int)nparams; i += kNsfThreads) s_par[i] = params[i];
  }
  __syncthreads();
  const float* W1t = s_par;
  const float* b1 = W1t + D * kH;
  const float* hid = b1 + kH;  // (L-1) blocks of [64][64] + [64]
  const float* Wout = hid + (size_t)(hidden_layers - 1) * (kH * kH + kH);
  const float* bout = Wout + (size_t)D * kH * kPP;
  float* col = s_scr + tid;

  const int64_t ntiles = (n + kNsfThreads - 1) / kNsfThreads;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t p = tile * kNsfThreads + tid;
    const bool valid = p < n;
    float vin[D];
#pragma unroll
    for (int i = 0; i < D; ++i) vin[i] = valid ? v[p * D + i] : 0.f;

    float h[kH];
    // first masked layer (D -> 64)
#pragma unroll
    for (int j4 = 0; j4 < kH / 4; ++j4) {
      float4 acc = *reinterpret_cast<const float4*>(b1 + 4 * j4);
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const float4 w = *reinterpret_cast<const float4*>(W1t + i * kH + 4 * j4);
        acc.x = fmaf(vin[i], w.x, acc.x); acc.y = fmaf(vin[i], w.y, acc.y);
        acc.z = fmaf(vin[i], w.z, acc.z); acc.w = fmaf(vin[i], w.w, acc.w);
      }
      h[4 * j4 + 0] = fmaxf(acc.x, 0.f); h[4 * j4 + 1] = fmaxf(acc.y, 0.f);
      h[4 * j4 + 2] = fmaxf(acc.z, 0.f); h[4 * j4 + 3] = fmaxf(acc.w, 0.f);
    }
    // hidden -> hidden layers
    for (int l = 0; l < hidden_layers - 1; ++l) {
      const float* wt = hid + (size_t)l * (kH * kH + kH);
      const float* bias = wt + kH * kH;
#pragma unroll 1
      for (int jc = 0; jc < kH / 8; ++jc) {
        float acc[8];
        dense8(h, wt + jc * 8, bias + jc * 8, acc);
#pragma unroll
        for (int q = 0; q < 8; ++q) col[(jc * 8 + q) * kNsfThreads] = fmaxf(acc[q], 0.f);
      }
#pragma unroll
      for (int j = 0; j < kH; ++j) h[j] = col[j * kNsfThreads];
    }
    // output layer, one feature at a time, fused with the spline
    float ladj = 0.f;
    float yout[D];
    const int ptotal = 3 * nb - 1;
#pragma unroll 1
    for (int f = 0; f < D; ++f) {
      const float* bf = bout + f * kPP;
      if (order.v[f] == 0) {
        // first feature in the layer's order: its spline is unconditional (bias only)
        for (int j = 0; j < ptotal; ++j) col[j * kNsfThreads] = bf[j];
      } else {
        const float* wf = Wout + (size_t)f * kH * kPP;
        const int nchunk = (ptotal + 7) >> 3;
#pragma unroll 1
        for (int jc = 0; jc < nchunk; ++jc) {
          float acc[8];
          dense8(h, wf + jc * 8, bf + jc * 8, acc);
#pragma unroll
          for (int q = 0; q < 8; ++q) col[(jc * 8 + q) * kNsfThreads] = acc[q];
        }
      }
      float vf = vin[0];
#pragma unroll
      for (int i = 1; i < D; ++i) vf = (f == i) ? vin[i] : vf;
      const float yf = rq_spline_forward(col, kNsfThreads, nb, vf, ladj);
#pragma unroll
      for (int i = 0; i < D; ++i)
        if (f == i) yout[i] = yf;
    }
    if (valid) {
#pragma unroll
      for (int i = 0; i < D; ++i) y[p * D + i] = yout[i];
      if (logq_out) {
        float base;
        if (first_layer) {
          float ss = 0.f;
#pragma unroll
          for (int i = 0; i < D; ++i) ss = fmaf(vin[i], vin[i], ss);
          base = -0.5f * ss - (float)D * kHalfLog2Pi;
        } else {
          base = logq_in[p];
        }
        logq_out[p] = base - ladj;
      }
    }
  }
}

template <int D>
static int launch_nsf_fwd(const float* v, int64_t n, int hidden_layers, int nb, const float* params,
                          const FeatureOrder& order, const float* logq_in, int first, float* y, float* logq_out,
                          cudaStream_t st) {
  const int64_t np = nsf_param_floats(D, hidden_layers);
  const size_t smem = (size_t)((np + 3) & ~(int64_t)3) * 4 + (size_t)kPP * kNsfThreads * 4;
  if (smem > 227 * 1024) return MFB_E_UNSUPPORTED;
  MFB_CUDA(cudaFuncSetAttribute(nsf_layer_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t tiles = (n + kNsfThreads - 1) / kNsfThreads;
  int64_t grid = sm_count();
  if (grid > tiles) grid = tiles;
  nsf_layer_fwd_kernel<D><<<(int)grid, kNsfThreads, smem, st>>>(v, n, hidden_layers, nb, params, np, order,
                                                                  logq_in, first, y, logq_out);
  return launch_status();
}

}  // namespace mfb

using namespace mfb;

extern "C" {

int64_t mfb_nsf_layer_param_floats(int d, int hidden_units, int hidden_layers, int bins) {
  if (d < 2 || d > 6 || hidden_units != kH || hidden_layers < 1 || bins < 2 || 3 * bins - 1 > kPP) return 0;
  return nsf_param_floats(d, hidden_layers);
}

int mfb_nsf_layer_fwd(const float* v, int64_t n, int d, int hidden_units, int hidden_layers, int bins,
                      const float* params, const int32_t* order_host, const float* logq_in, int first_layer,
                      float* y, float* logq_out, void* stream) {
  MFB_CHECK_ARG(v && params && y && n >= 0);
  if (hidden_units != kH || hidden_layers < 1 || bins < 2 || 3 * bins - 1 > kPP) return MFB_E_UNSUPPORTED;
  MFB_CHECK_ARG(first_layer || !logq_out || logq_in);
  if (n == 0) return 0;
  FeatureOrder ord;
  for (int i = 0; i < kMaxDim; ++i) ord.v[i] = (order_host && i < d) ? order_host[i] : 1;
  cudaStream_t st = (cudaStream_t)stream;
  switch (d) {
    case 2: return launch_nsf_fwd<2>(v, n, hidden_layers, bins, params, ord, logq_in, first_layer, y, logq_out, st);
    case 3: return launch_nsf_fwd<3>(v, n, hidden_layers, bins, params, ord, logq_in, first_layer, y, logq_out, st);
    case 4: return launch_nsf_fwd<4>(v, n, hidden_layers, bins, params, ord, logq_in, first_layer, y, logq_out, st);
    case 5: return launch_nsf_fwd<5>(v, n, hidden_layers, bins, params, ord, logq_in, first_layer, y, logq_out, st);
    case 6: return launch_nsf_fwd<6>(v, n, hidden_layers, bins, params, ord, logq_in, first_layer, y, logq_out, st);
    default: return MFB_E_UNSUPPORTED;
  }
}

}  // extern "C"
