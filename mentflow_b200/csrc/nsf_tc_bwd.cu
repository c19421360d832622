// Data-gradient chain of one neural-spline-flow layer on the tensor cores (tcgen05 + TMEM), sm_100a.
//
// Part of the hand-written backward that replaces torch autograd through zuko's MaskedMLP (see
// nsf_bwd.cu for the pipeline).  Given dL/dphi (gradient w.r.t. the raw conditioner outputs, from the
// spline backward) this kernel walks the conditioner backwards,
//   g3 = relu'(h3) * sum_f dL/dphi_f Wout_f        (five output-feature GEMMs into one accumulator)
//   g2 = relu'(h2) * g3 W3,   g1 = relu'(h1) * g2 W2,   dL/dv = direct + g1 W1,
// with the same machinery as the forward kernel (nsf_tc.cu): 128-particle tiles, fp32 operands split
// into fp16 (hi, lo) pairs, three MMAs per K step, accumulators in TMEM, 3 compute warpgroups + MMA
// issuer warps.  It replaces three nsf_dgrad_kernel launches and nsf_bwd_input_kernel (4.7 ms per
// 1e6 particles and layer on the CUDA cores).
//
// Gradients have no natural scale (1/N of a mean loss is 1e-6 and less, far below fp16's range), so
// every A-operand row (= particle) is multiplied by a power of two that brings its largest entry to
// [1, 2) before the split, and the accumulator row is multiplied back afterwards: exact, and the
// relative precision of the split (2^-22) no longer depends on the magnitude.  For the five
// output-feature tiles of a particle the scale comes from max |dL/dphi| over all of them, which the
// spline-backward kernel records (gmax).
#include "nsf_tc_common.cuh"

namespace mfb {
namespace tc {

struct DgradMeta {
  int perm[kH];               // perm[c] = original hidden unit at sorted position c (global row of column c)
  int slot_feature[kMaxDim];  // feature of output slot s (order s+1)
  int slot_nk[kMaxDim];       // 16 * slot_nk[s] hidden units (sorted) are read by slot s
  int nslots;
};

// image: [slot s: hi 8K | lo 8K] x S   B = Wout_s^T  rows = hidden unit (sorted), K = parameter j
//        [hidden l: hi 8K | lo 8K] x 2 B = W_(l+2)^T rows = input unit (sorted), K = output unit (sorted)
//        [first: hi 2K | lo 2K]        B = W1^T      rows = feature i (16 rows), K = hidden unit (sorted)
__host__ __device__ constexpr int dgrad_off_slot(int s) { return s * 2 * kTileBytes; }
__host__ __device__ constexpr int dgrad_off_hid(int S, int l) { return (S + l) * 2 * kTileBytes; }
__host__ __device__ constexpr int dgrad_off_first(int S, int L) { return (S + L - 1) * 2 * kTileBytes; }
__host__ __device__ constexpr int dgrad_image_bytes(int D, int L) {
  return (dgrad_off_first(D - 1, L) + 2 * 2048 + 1023) & ~1023;
}

// one block per layer; params = packed fp32 block of nsf_common.cuh (pre-masked, [in][out])
__global__ void __launch_bounds__(256)
nsf_tc_dgrad_prepare_kernel(const float* __restrict__ params, int D, int L, const __grid_constant__ DgradMeta meta,
                            unsigned char* __restrict__ img, int img_bytes) {
  const int S = meta.nslots;
  const float* W1t = params;
  const float* hid = W1t + D * kH + kH;
  const float* Wout = hid + (size_t)(L - 1) * (kH * kH + kH);
  for (int i = threadIdx.x; i < img_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(img)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  const int n_slot = S * kH * 8, n_hid = (L - 1) * kH * 8, n_first = 16 * 8;
  for (int task = threadIdx.x; task < n_slot + n_hid + n_first; task += blockDim.x) {
    int t2 = task;
    float x[8];
    if (t2 < n_slot) {
      const int s = t2 / (kH * 8), row = (t2 / 8) % kH, c = t2 % 8;   // row = hidden unit (sorted), 8 parameters
      const float* wf = Wout + ((size_t)meta.slot_feature[s] * kH + meta.perm[row]) * kPP;
      for (int e = 0; e < 8; ++e) x[e] = wf[c * 8 + e];
      store_split8(img + dgrad_off_slot(s), img + dgrad_off_slot(s) + kTileBytes, row, c, x);
      continue;
    }
    t2 -= n_slot;
    if (t2 < n_hid) {
      const int l = t2 / (kH * 8), row = (t2 / 8) % kH, c = t2 % 8;   // row = input unit, 8 output units
      const float* wt = hid + (size_t)l * (kH * kH + kH) + (size_t)meta.perm[row] * kH;
      for (int e = 0; e < 8; ++e) x[e] = wt[meta.perm[c * 8 + e]];
      store_split8(img + dgrad_off_hid(S, l), img + dgrad_off_hid(S, l) + kTileBytes, row, c, x);
      continue;
    }
    t2 -= n_hid;
    {
      const int row = t2 / 8, c = t2 % 8;                              // row = input feature, 8 hidden units
      for (int e = 0; e < 8; ++e) x[e] = row < D ? W1t[row * kH + meta.perm[c * 8 + e]] : 0.f;
      store_split8(img + dgrad_off_first(S, L), img + dgrad_off_first(S, L) + 2048, row, c, x);
    }
  }
}

// power of two s with amax * s in [1, 2) (1 for amax = 0 / denormal / non-finite)
__device__ __forceinline__ float pow2_scale(float amax, float& inv) {
  const int e = (__float_as_int(amax) >> 23) & 0xFF;
  const bool ok = e > 0 && e < 0xFF;
  inv = ok ? __int_as_float(e << 23) : 1.0f;            // 2^(e-127)
  return ok ? __int_as_float((254 - e) << 23) : 1.0f;   // 2^(127-e)
}

// (hi, lo) fp16 split of a scaled row into the A tile (signed values, round to nearest)
__device__ __forceinline__ void store_row_split(const float (&x)[64], float scale, unsigned char* a_hi, unsigned char* a_lo,
                                                int row) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float x0 = x[8 * c + 2 * e] * scale, x1 = x[8 * c + 2 * e + 1] * scale;
      const __half2 h = __floats2half2_rn(x0, x1);
      const float2 hf = __half22float2(h);
      const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
      hi[e] = *reinterpret_cast<const uint32_t*>(&h);
      lo[e] = *reinterpret_cast<const uint32_t*>(&l);
    }
    const uint32_t off = umma::sw128_offset(row, c);
    *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

template <int D, int L>
__global__ void __launch_bounds__(kThreads, 1)
nsf_tc_dgrad_kernel(const float* __restrict__ gphi /* [D*64][n] */, const float* __restrict__ gmax /* [n] */,
                    const float* __restrict__ acts /* [L][64][n] */, const float* __restrict__ gvd /* [n][D] */,
                    int64_t n, const unsigned char* __restrict__ image, const __grid_constant__ DgradMeta meta,
                    float* __restrict__ gz /* [L][64][n]: dL/d(pre-activation) of hidden layer l */,
                    float* __restrict__ gv /* [n][D] */) {
  static_assert(L == 3, "compiled for three hidden layers");
  constexpr int S = D - 1;
  constexpr int kImg = dgrad_image_bytes(D, L);
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* img = smem;
  unsigned char* a_all = smem + kImg;
  // mbarriers: [0] image; [1 + wg] MMA group of wg complete (count 1); [1 + kWG + wg] request of wg (count 4)
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_all + kWG * kABytes);
  uint64_t* done = bars + 1;
  uint64_t* reqs = done + kWG;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(reqs + kWG);

  const int tid = threadIdx.x;
  const int wg = __shfl_sync(0xffffffffu, tid >> 7, 0);
  const int t = tid & 127;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    for (int i = 0; i < kWG; ++i) {
      mbar_init(&done[i], 1);
      mbar_init(&reqs[i], 4);
    }
    fence_mbar_init();
  }
  if (tid < 32) umma::tmem_alloc(tmem_slot, 256);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  if (tid == 0) {
    mbar_expect_tx(&bars[0], (uint32_t)kImg);
    tma_load_1d(img, image, (uint32_t)kImg, &bars[0]);
  }
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  mbar_wait_bounded(&bars[0], 0);

  const int64_t ntiles = (n + 127) / 128;
  const int64_t tstride = (int64_t)gridDim.x * kWG;
  const uint32_t idesc64 = umma::make_idesc_f16(128, 64);

  if (wg == kWG) {
    // ===== MMA issuers: warp i serves compute warpgroup i; per tile S slot requests, then three GEMMs
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsIssuer));
    const int w = __shfl_sync(0xffffffffu, (tid - kWG * 128) >> 5, 0);
    if (w < kWG) {
      const int64_t first = (int64_t)blockIdx.x * kWG + w;
      const int cnt = first < ntiles ? (int)((ntiles - first + tstride - 1) / tstride) : 0;
      uint32_t rp = 0;
      unsigned char* wa_hi = a_all + w * kABytes;
      const uint64_t dA_hi = umma::make_desc_sw128(smem_u32(wa_hi));
      const uint64_t dA_lo = umma::make_desc_sw128(smem_u32(wa_hi + kABytes / 2));
      const uint32_t col0 = tmem_base + (uint32_t)(w * 64);
#pragma unroll 1
      for (int i = 0; i < cnt; ++i) {
#pragma unroll 1
        for (int s = S - 1; s >= 0; --s) {
          mbar_wait_polite(&reqs[w], rp);
          rp ^= 1;
          umma::fence_after_sync();
          const uint32_t idesc = umma::make_idesc_f16(128, 16 * meta.slot_nk[s]);
          const uint64_t dBh = umma::make_desc_sw128(smem_u32(img + dgrad_off_slot(s)));
          const uint64_t dBl = umma::make_desc_sw128(smem_u32(img + dgrad_off_slot(s) + kTileBytes));
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) mma_cross(col0, dA_hi, dA_lo, dBh, dBl, ks, idesc, (s < S - 1 || ks > 0) ? 1u : 0u);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) mma_main(col0, dA_hi, dBh, ks, idesc);
            umma::commit(&done[w]);
          }
          __syncwarp();
        }
#pragma unroll 1
        for (int l = L - 2; l >= -1; --l) {
          mbar_wait_polite(&reqs[w], rp);
          rp ^= 1;
          umma::fence_after_sync();
          const int boff = l >= 0 ? dgrad_off_hid(S, l) : dgrad_off_first(S, L);
          const int lo_off = l >= 0 ? kTileBytes : 2048;
          const uint32_t idesc = l >= 0 ? idesc64 : umma::make_idesc_f16(128, 16);
          const uint64_t dBh = umma::make_desc_sw128(smem_u32(img + boff));
          const uint64_t dBl = umma::make_desc_sw128(smem_u32(img + boff + lo_off));
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) mma_cross(col0, dA_hi, dA_lo, dBh, dBl, ks, idesc, ks > 0 ? 1u : 0u);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) mma_main(col0, dA_hi, dBh, ks, idesc);
            umma::commit(&done[w]);
          }
          __syncwarp();
        }
      }
    }
    __syncwarp();
  } else {
    // ===== compute warpgroups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsCompute));
    unsigned char* a_hi = a_all + wg * kABytes;
    unsigned char* a_lo = a_hi + kABytes / 2;
    uint32_t ph = 0;
    const uint32_t taddr = tmem_base + (uint32_t)(wg * 64) + ((uint32_t)((t >> 5) * 32) << 16);
    auto publish = [&]() {   // A rows written: hand the tile to the issuer
      fence_proxy_async();
      umma::fence_before_sync();
      request_arrive(&reqs[wg]);
    };
    auto wait_done = [&]() {
      mbar_wait_bounded(&done[wg], ph);
      ph ^= 1;
      umma::fence_after_sync();
    };
    for (int64_t tile = (int64_t)blockIdx.x * kWG + wg; tile < ntiles; tile += tstride) {
      const int64_t p = tile * 128 + t;
      const bool valid = p < n;
      // ---- output layer: slots in descending order (the last slot reads every hidden unit: it
      //      initialises all 64 accumulator columns)
      float inv0;
      const float sc0 = pow2_scale(valid ? gmax[p] : 0.f, inv0);
      float g[64];
      {
        const float* gp = gphi + (size_t)meta.slot_feature[S - 1] * kPP * n + p;
#pragma unroll
        for (int j = 0; j < 64; ++j) g[j] = (valid && j < 59) ? gp[(size_t)j * n] : 0.f;
      }
      // global loads are always issued BEFORE waiting for the tensor core and global stores AFTER the
      // hand-off to the issuer, so HBM latency and the release-fence of the arrive overlap the MMAs
      auto load_mask = [&](int l, float (&hv)[64]) {
        const float* hl = acts + (size_t)l * kH * n + p;
#pragma unroll
        for (int c = 0; c < 64; ++c) hv[c] = valid ? hl[(size_t)meta.perm[c] * n] : 0.f;
      };
#pragma unroll 1
      for (int s = S - 1; s >= 0; --s) {
        store_row_split(g, sc0, a_hi, a_lo, t);
        publish();
        if (s > 0) {   // next slot's gradient rows travel while the tensor core works
          const float* gp = gphi + (size_t)meta.slot_feature[s - 1] * kPP * n + p;
#pragma unroll
          for (int j = 0; j < 64; ++j) g[j] = (valid && j < 59) ? gp[(size_t)j * n] : 0.f;
        } else {
          load_mask(L - 1, g);   // g now holds h3 (sorted unit order): the ReLU mask of the first chain step
        }
        wait_done();
      }
      // ---- hidden layers, last to first; then the first layer
      float unscale = inv0;
#pragma unroll 1
      for (int l = L - 1; l >= 0; --l) {
        float acc[64];
        tmem_ld64(taddr, acc);
        float amax = 0.f;
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          acc[c] = g[c] > 0.f ? acc[c] * unscale : 0.f;
          amax = fmaxf(amax, fabsf(acc[c]));
        }
        float inv;
        const float sc = pow2_scale(amax, inv);
        store_row_split(acc, sc, a_hi, a_lo, t);
        unscale = inv;
        publish();
        if (valid) {
          float* gl = gz + (size_t)l * kH * n + p;
#pragma unroll
          for (int c = 0; c < 64; ++c) gl[(size_t)meta.perm[c] * n] = acc[c];
        }
        if (l > 0) load_mask(l - 1, g);
        wait_done();
      }
      // ---- dL/dv = direct (through the spline) + g1 W1
      {
        float acc16[32];
        umma::tmem_ld32(taddr, acc16);
        if (valid) {
#pragma unroll
          for (int i = 0; i < D; ++i) gv[p * D + i] = fmaf(acc16[i], unscale, gvd[p * D + i]);
        }
        umma::fence_before_sync();
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (tid < 32) umma::tmem_dealloc(tmem_base, 256);
}

template <int D>
static int launch_dgrad(const float* gphi, const float* gmax, const float* acts, const float* gvd, int64_t n,
                        const float* params, const int32_t* order, float* gz, float* gv, unsigned char* image,
                        cudaStream_t st) {
  constexpr int L = 3;
  DgradMeta meta = {};
  int cls[kH];
  hidden_classes(D, cls, meta.perm);
  int feat_of_order[kMaxDim];
  for (int i = 0; i < D; ++i) feat_of_order[order[i]] = i;
  meta.nslots = D - 1;
  for (int s = 0; s < D - 1; ++s) {
    meta.slot_feature[s] = feat_of_order[s + 1];
    const int o = s + 1;
    const int cnt = o >= D - 1 ? kH : o * (kH / (D - 1)) + (o < kH % (D - 1) ? o : kH % (D - 1));
    meta.slot_nk[s] = (cnt + 15) / 16;
  }
  const int img_bytes = dgrad_image_bytes(D, L);
  nsf_tc_dgrad_prepare_kernel<<<1, 256, 0, st>>>(params, D, L, meta, image, img_bytes);
  int rc = launch_status();
  if (rc) return rc;
  const size_t smem = (size_t)img_bytes + kWG * kABytes + 128 + 1024;
  auto kern = nsf_tc_dgrad_kernel<D, L>;
  MFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t ntiles = (n + 127) / 128;
  int64_t grid = sm_count();
  if (grid * kWG > ntiles) grid = (ntiles + kWG - 1) / kWG;
  kern<<<(int)grid, kThreads, smem, st>>>(gphi, gmax, acts, gvd, n, image, meta, gz, gv);
  return launch_status();
}

}  // namespace tc

// Called by the backward orchestration in nsf_bwd.cu.  gz receives dL/d(pre-activation) of the three
// hidden layers ([3][64][n], feature-major like acts), gv the gradient w.r.t. the layer input.
// image: scratch of nsf_tc_dgrad_image_bytes(d) bytes.  Returns MFB_E_UNSUPPORTED for shapes that are
// not compiled (the caller then uses the CUDA-core kernels).
int64_t nsf_tc_dgrad_image_bytes(int d) { return (d >= 2 && d <= 6) ? tc::dgrad_image_bytes(d, 3) : 0; }

int nsf_tc_dgrad(const float* gphi, const float* gmax, const float* acts, const float* gvd, int64_t n, int d,
                 int hidden_layers, const float* params, const int32_t* order, float* gz, float* gv, void* image,
                 cudaStream_t st) {
  if (hidden_layers != 3 || d < 2 || d > 6) return MFB_E_UNSUPPORTED;
  unsigned char* img = reinterpret_cast<unsigned char*>(image);
  switch (d) {
    case 2: return tc::launch_dgrad<2>(gphi, gmax, acts, gvd, n, params, order, gz, gv, img, st);
    case 3: return tc::launch_dgrad<3>(gphi, gmax, acts, gvd, n, params, order, gz, gv, img, st);
    case 4: return tc::launch_dgrad<4>(gphi, gmax, acts, gvd, n, params, order, gz, gv, img, st);
    case 5: return tc::launch_dgrad<5>(gphi, gmax, acts, gvd, n, params, order, gz, gv, img, st);
    case 6: return tc::launch_dgrad<6>(gphi, gmax, acts, gvd, n, params, order, gz, gv, img, st);
    default: return MFB_E_UNSUPPORTED;
  }
}

}  // namespace mfb
