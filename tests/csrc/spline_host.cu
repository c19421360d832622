// TEST INFRASTRUCTURE: compiles the register-resident spline of the tcgen05 flow kernels
// (mentflow_b200/csrc/nsf_spline_regs.cuh) for the HOST, so that its fp32 arithmetic can be checked
// against the float64 oracle on a machine without a GPU.  ex2.approx / rcp.approx become exp2f / a
// divide here; everything else (summation order, centred differences, fused multiply-adds, the
// two-level bin search) is the code the GPU runs.
#include <stdint.h>

#include "../../mentflow_b200/csrc/nsf_spline_regs.cuh"

using namespace mfb;

extern "C" {

// phi: [n][64] raw conditioner outputs in NATURAL units (the kernel sees them times log2 e, which the
// operand image folds into the weights); v: [n].  y, jac: [n].
void spline_host_fwd(const float* phi, const float* v, int64_t n, float* y, float* jac) {
  for (int64_t p = 0; p < n; ++p) {
    float a[64];
    for (int j = 0; j < 64; ++j) a[j] = phi[p * 64 + j] * kLog2e;
    float jc = 1.0f;
    y[p] = tc::rq_spline_regs<20>(a, v[p], jc);
    jac[p] = jc;
  }
}

// forward + backward: gphi [n][64] = dL/d(raw natural parameter), gv [n] = direct dL/dv
void spline_host_bwd(const float* phi, const float* v, const float* gy, const float* gl, int64_t n, float* gphi,
                     float* gv) {
  for (int64_t p = 0; p < n; ++p) {
    float a[64];
    for (int j = 0; j < 64; ++j) a[j] = phi[p * 64 + j] * kLog2e;
    gv[p] = tc::rq_spline_regs_bwd<20>(a, v[p], gy[p], gl[p]);
    for (int j = 0; j < 64; ++j) gphi[p * 64 + j] = j < 59 ? a[j] : 0.f;
  }
}

// inverse: v with RQS(v) = y, jac = dy/dv at v
void spline_host_inv(const float* phi, const float* y, int64_t n, float* v, float* jac) {
  for (int64_t p = 0; p < n; ++p) {
    float a[64];
    for (int j = 0; j < 64; ++j) a[j] = phi[p * 64 + j] * kLog2e;
    float jc = 1.0f;
    v[p] = tc::rq_spline_regs_inv<20>(a, y[p], jc);
    jac[p] = jc;
  }
}

int spline_host_comp() { return MFB_SPLINE_COMP; }

}  // extern "C"
