// Device code shared by the neural-spline-flow kernels (forward, backward, inverse).
#pragma once
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

// Position of v in the horizontal knot array, from the un-normalised softmax terms e_j (col[j*stride],
// as left by softmax_inplace).  The running sums are carried in double: in fp32 they are the largest
// rounding error of the whole spline (scripts/emul_spline.py; every fp32 add at the magnitude of the
// sum moves the knot by an ulp of the BOX, not of the bin).  These CUDA-core kernels are the fallback /
// cross-check path, so the few double adds per feature are not a cost that matters.
// Returns the bin k, the un-normalised bin width e_k, its left knot sum and the total, all exact to
// double rounding; v must be inside (-B, B].
struct KnotPos {
  int k;
  double cum, sum, target;
};
__device__ __forceinline__ KnotPos knot_search(const float* col, int stride, int nb, float v) {
  KnotPos kp;
  double sum = 0.0;
  for (int j = 0; j < nb; ++j) sum += (double)col[j * stride];
  kp.sum = sum;
  kp.target = ((double)v + (double)kBound) * (0.5 / (double)kBound) * sum;
  double cum = 0.0;
  int k = nb - 1;
  for (int j = 0; j < nb - 1; ++j) {   // knot_k < v <= knot_{k+1}
    const double nxt = cum + (double)col[j * stride];
    if (kp.target <= nxt) {
      k = j;
      break;
    }
    cum = nxt;
  }
  kp.k = k;
  kp.cum = cum;
  return kp;
}

// One univariate spline: parameters (3*nb-1 raw conditioner outputs) in col[j*stride].
// Returns y and adds log dy/dv to ladj.
//
// Conditioning: zuko differences the cumulative knot arrays (x1 - x0, y1 - y0), which cancels
// catastrophically in narrow bins (widths go down to 1e-3 of the mean).  Here the bin width and
// height are taken directly from the softmax values (dx = 2B W_k, dy = 2B H_k), which is the same
// number in exact arithmetic but accurate to an ulp, and the position inside the bin comes from
// knot sums carried in double (knot_search).
__device__ __forceinline__ float rq_spline_forward(float* col, int stride, int nb, float v, float& ladj) {
  if (!(v > -kBound && v <= kBound)) return v;  // outside (-bound, bound]: identity, ladj += 0
  softmax_inplace(col, stride, nb);
  const KnotPos kp = knot_search(col, stride, nb, v);
  const int kbin = kp.k;
  const float ek = col[kbin * stride];
  const float wk = (float)((double)ek / kp.sum);
  float* colh = col + nb * stride;
  softmax_inplace(colh, stride, nb);
  double sum_h = 0.0, cumh = 0.0;
  for (int j = 0; j < nb; ++j) {
    const double e = (double)colh[j * stride];
    if (j < kbin) cumh += e;
    sum_h += e;
  }
  const float y0 = (float)(2.0 * (double)kBound * (cumh / sum_h) - (double)kBound);
  const float hk = (float)((double)colh[kbin * stride] / sum_h);
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
  const float dy = 2.0f * kBound * hk;
  const float s = hk / wk;
  float t = (float)((kp.target - kp.cum) / (double)ek);
  t = fminf(fmaxf(t, 0.0f), 1.0f);
  const float omt = 1.0f - t;
  const float tomt = t * omt;
  const float den = fmaf(d0 + d1 - 2.0f * s, tomt, s);
  const float y = y0 + dy * (s * t * t + d0 * tomt) / den;
  const float jac = s * s * (2.0f * s * tomt + d0 * omt * omt + d1 * t * t) / (den * den);
  ladj += logf(jac);
  return y;
}

}  // namespace mfb
