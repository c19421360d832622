// zuko-layout parameters of the whole flow  <->  the kernels' packed per-layer blocks, one launch each way.
//
// Replaces ~30 eager torch kernels per optimisation step (mask multiply, transpose, pad, cat for the
// forward layout of nsf_common.cuh and the out-major layout of the backward kernels, and autograd's
// un-packing of the gradient): generate/flows/zuko.py keeps its parameters in zuko's MaskedLinear layout
// (weight [out][in], mask applied on every call, zuko/nn.py), the kernels want them pre-masked,
// transposed to [in][out] and the 59 spline parameters of a feature padded to 64.
//
//   forward layout of one layer (nsf_param_floats):  W1t [D][64] | b1 [64] | (Wt_l [64][64] | b_l [64]) x (L-1) |
//                                                     Wout_t [D][64 in][64] | bout [D][64]
//   out-major layout (mfb_nsf_layer_param_om_floats): W1 [64][D] | Wl [64 out][64 in] x (L-1) | Wout [D*64][64 in]
#include "nsf_common.cuh"

namespace mfb {

struct PackDims {
  int T, D, L, P;   // layers, features, hidden layers, parameters per feature (3 * bins - 1)
};

struct ZukoTensors {       // [T][...] contiguous, fp32; masks are 0/1 floats of the weights' shapes
  const float *w_in, *b_in, *w_hid, *b_hid, *w_out, *b_out;
  const float *m_in, *m_hid, *m_out;
};
struct ZukoGrads {
  float *w_in, *b_in, *w_hid, *b_hid, *w_out, *b_out;
};

__global__ void __launch_bounds__(256)
nsf_pack_kernel(const __grid_constant__ ZukoTensors z, PackDims dm, int64_t np, int64_t nom, float* __restrict__ packed,
                float* __restrict__ packed_om) {
  const int D = dm.D, L = dm.L, P = dm.P;
  const int64_t total = (int64_t)dm.T * np;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(g / np);
    int64_t e = g - (int64_t)t * np;
    float v = 0.f;
    int64_t om = -1;          // position of the same (masked) weight in the out-major block, if it has one
    if (e < (int64_t)D * kH) {
      const int i = (int)(e / kH), o = (int)(e % kH);
      const int64_t src = ((int64_t)t * kH + o) * D + i;
      v = z.w_in[src] * z.m_in[src];
      om = (int64_t)o * D + i;
    } else if ((e -= (int64_t)D * kH) < kH) {
      v = z.b_in[(int64_t)t * kH + e];
    } else if ((e -= kH) < (int64_t)(L - 1) * (kH * kH + kH)) {
      const int l = (int)(e / (kH * kH + kH));
      const int64_t r = e - (int64_t)l * (kH * kH + kH);
      if (r < kH * kH) {
        const int i = (int)(r / kH), o = (int)(r % kH);
        const int64_t src = (((int64_t)t * (L - 1) + l) * kH + o) * kH + i;
        v = z.w_hid[src] * z.m_hid[src];
        om = (int64_t)kH * D + (int64_t)l * kH * kH + (int64_t)o * kH + i;
      } else {
        v = z.b_hid[((int64_t)t * (L - 1) + l) * kH + (r - kH * kH)];
      }
    } else if ((e -= (int64_t)(L - 1) * (kH * kH + kH)) < (int64_t)D * kH * kPP) {
      const int f = (int)(e / (kH * kPP)), i = (int)((e / kPP) % kH), q = (int)(e % kPP);
      if (q < P) {
        const int64_t src = ((int64_t)t * D * P + (int64_t)f * P + q) * kH + i;
        v = z.w_out[src] * z.m_out[src];
      }
      om = (int64_t)kH * D + (int64_t)(L - 1) * kH * kH + ((int64_t)f * kPP + q) * kH + i;
    } else {
      e -= (int64_t)D * kH * kPP;
      const int f = (int)(e / kPP), q = (int)(e % kPP);
      if (q < P) v = z.b_out[(int64_t)t * D * P + (int64_t)f * P + q];
    }
    packed[g] = v;
    if (packed_om && om >= 0) packed_om[(int64_t)t * nom + om] = v;
  }
}

// gradient of the packed blocks -> gradients of the zuko-layout tensors (masked entries get exactly 0)
__global__ void __launch_bounds__(256)
nsf_unpack_grad_kernel(const float* __restrict__ gpacked, int64_t np, const __grid_constant__ ZukoTensors z, PackDims dm,
                       const __grid_constant__ ZukoGrads g) {
  const int D = dm.D, L = dm.L, P = dm.P, T = dm.T;
  const int64_t n_win = (int64_t)T * kH * D, n_bin = (int64_t)T * kH, n_whid = (int64_t)T * (L - 1) * kH * kH,
                n_bhid = (int64_t)T * (L - 1) * kH, n_wout = (int64_t)T * D * P * kH, n_bout = (int64_t)T * D * P;
  const int64_t total = n_win + n_bin + n_whid + n_bhid + n_wout + n_bout;
  const int64_t off_b1 = (int64_t)D * kH, off_hid = off_b1 + kH, off_wout = off_hid + (int64_t)(L - 1) * (kH * kH + kH),
                off_bout = off_wout + (int64_t)D * kH * kPP;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = e;
    if (r < n_win) {
      const int t = (int)(r / ((int64_t)kH * D)), o = (int)((r / D) % kH), i = (int)(r % D);
      g.w_in[r] = gpacked[(int64_t)t * np + (int64_t)i * kH + o] * z.m_in[r];
    } else if ((r -= n_win) < n_bin) {
      const int t = (int)(r / kH), o = (int)(r % kH);
      g.b_in[r] = gpacked[(int64_t)t * np + off_b1 + o];
    } else if ((r -= n_bin) < n_whid) {
      const int t = (int)(r / ((int64_t)(L - 1) * kH * kH)), l = (int)((r / (kH * kH)) % (L - 1));
      const int o = (int)((r / kH) % kH), i = (int)(r % kH);
      g.w_hid[r] = gpacked[(int64_t)t * np + off_hid + (int64_t)l * (kH * kH + kH) + (int64_t)i * kH + o] * z.m_hid[r];
    } else if ((r -= n_whid) < n_bhid) {
      const int t = (int)(r / ((int64_t)(L - 1) * kH)), l = (int)((r / kH) % (L - 1)), o = (int)(r % kH);
      g.b_hid[r] = gpacked[(int64_t)t * np + off_hid + (int64_t)l * (kH * kH + kH) + kH * kH + o];
    } else if ((r -= n_bhid) < n_wout) {
      const int t = (int)(r / ((int64_t)D * P * kH));
      const int row = (int)((r / kH) % ((int64_t)D * P)), i = (int)(r % kH);
      const int f = row / P, q = row % P;
      g.w_out[r] = gpacked[(int64_t)t * np + off_wout + ((int64_t)f * kH + i) * kPP + q] * z.m_out[r];
    } else {
      r -= n_wout;
      const int t = (int)(r / ((int64_t)D * P)), row = (int)(r % ((int64_t)D * P));
      const int f = row / P, q = row % P;
      g.b_out[r] = gpacked[(int64_t)t * np + off_bout + (int64_t)f * kPP + q];
    }
  }
}

}  // namespace mfb

using namespace mfb;

extern "C" {

int mfb_nsf_pack_params(const float* w_in, const float* b_in, const float* w_hid, const float* b_hid, const float* w_out,
                        const float* b_out, const float* m_in, const float* m_hid, const float* m_out, int transforms,
                        int d, int hidden_units, int hidden_layers, int bins, float* packed, float* packed_om,
                        void* stream) {
  MFB_CHECK_ARG(w_in && b_in && w_out && b_out && m_in && m_out && packed && transforms >= 1);
  MFB_CHECK_ARG(hidden_layers == 1 || (w_hid && b_hid && m_hid));
  if (hidden_units != kH || hidden_layers < 1 || bins < 2 || 3 * bins - 1 > kPP || d < 1 || d > kMaxDim) return MFB_E_UNSUPPORTED;
  const ZukoTensors z{w_in, b_in, w_hid, b_hid, w_out, b_out, m_in, m_hid, m_out};
  const PackDims dm{transforms, d, hidden_layers, 3 * bins - 1};
  const int64_t np = nsf_param_floats(d, hidden_layers);
  const int64_t nom = (int64_t)kH * d + (int64_t)(hidden_layers - 1) * kH * kH + (int64_t)d * kPP * kH;
  const int64_t total = (int64_t)transforms * np;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 4096) blocks = 4096;
  nsf_pack_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(z, dm, np, nom, packed, packed_om);
  return launch_status();
}

int mfb_nsf_unpack_grads(const float* gpacked, const float* m_in, const float* m_hid, const float* m_out, int transforms,
                         int d, int hidden_units, int hidden_layers, int bins, float* g_w_in, float* g_b_in,
                         float* g_w_hid, float* g_b_hid, float* g_w_out, float* g_b_out, void* stream) {
  MFB_CHECK_ARG(gpacked && m_in && m_out && g_w_in && g_b_in && g_w_out && g_b_out && transforms >= 1);
  MFB_CHECK_ARG(hidden_layers == 1 || (m_hid && g_w_hid && g_b_hid));
  if (hidden_units != kH || hidden_layers < 1 || bins < 2 || 3 * bins - 1 > kPP || d < 1 || d > kMaxDim) return MFB_E_UNSUPPORTED;
  const ZukoTensors z{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, m_in, m_hid, m_out};
  const ZukoGrads g{g_w_in, g_b_in, g_w_hid, g_b_hid, g_w_out, g_b_out};
  const PackDims dm{transforms, d, hidden_layers, 3 * bins - 1};
  const int64_t np = nsf_param_floats(d, hidden_layers);
  const int P = 3 * bins - 1;
  const int64_t total = (int64_t)transforms * ((int64_t)kH * d + kH + (int64_t)(hidden_layers - 1) * (kH * kH + kH) +
                                               (int64_t)d * P * kH + (int64_t)d * P);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 4096) blocks = 4096;
  nsf_unpack_grad_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(gpacked, np, z, dm, g);
  return launch_status();
}

}  // extern "C"
