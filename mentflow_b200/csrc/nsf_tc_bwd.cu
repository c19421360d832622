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
  // the image was zeroed by a memset node before the launch; the rows are dealt out over gridDim.x blocks
  const int n_slot = S * kH * 8, n_hid = (L - 1) * kH * 8, n_first = 16 * 8;
  for (int task = blockIdx.x * blockDim.x + threadIdx.x; task < n_slot + n_hid + n_first; task += blockDim.x * gridDim.x) {
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
nsf_tc_dgrad_kernel(const float* __restrict__ gphi /* D*kGRows rows, tile-major, compact (nsf_tc_common.cuh) */,
                    const float* __restrict__ gmax /* [n] */,
                    const uint32_t* __restrict__ masks /* L*2 rows, tile-major: ReLU masks */, const float* __restrict__ gvd /* [n][D] */,
                    int64_t n, const unsigned char* __restrict__ image, const __grid_constant__ DgradMeta meta,
                    float* __restrict__ gz /* L*64 rows, tile-major: dL/d(pre-activation) of hidden layer l */,
                    float* __restrict__ gv /* [n][D] */,
                    int* __restrict__ gmaxes /* [1 + L] float bits: batch maxima of |gphi|, |gz[l]| (atomicMax) */) {
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
      const int64_t first = first_tile_of(w);
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
    float g[64];
    float gmax_next = 0.f;
    // dL/dphi of one feature of a tile from its compact rows: 2 NB dense, then (left, right, bin) of the derivative block
    auto load_gphi = [&](int64_t tl, int f) {
      constexpr int NB2 = kGRows - 4;
      const bool ok = tl * 128 + t < n;
      const float* gp = gphi + ((size_t)tl * (D * kGRows) + f * kGRows) * 128 + t;
#pragma unroll
      for (int j = 0; j < NB2; ++j) g[j] = ok ? gp[j * 128] : 0.f;
      const float left = ok ? gp[NB2 * 128] : 0.f, right = ok ? gp[(NB2 + 1) * 128] : 0.f;
      const int k = ok ? (int)gp[(NB2 + 2) * 128] : 0;
#pragma unroll
      for (int j = 0; j < 64 - NB2; ++j) g[NB2 + j] = (j == k - 1) ? left : ((j == k) ? right : 0.f);   // j >= NB - 1 never matches a stored value: left / right are zero there
    };
    // first rows of a tile (slot S-1) and its row scale: requested during the hidden chain of the tile before, so
    // that no tile starts with an exposed trip to HBM
    auto prefetch_tile = [&](int64_t tl) {
      if (tl >= ntiles) return;
      gmax_next = (tl * 128 + t < n) ? gmax[tl * 128 + t] : 0.f;
      load_gphi(tl, meta.slot_feature[S - 1]);
    };
    prefetch_tile(first_tile_of(wg));
    for (int64_t tile = first_tile_of(wg); tile < ntiles; tile += tstride) {
      const int64_t p = tile * 128 + t;
      const bool valid = p < n;
      // ---- output layer: slots in descending order (the last slot reads every hidden unit: it
      //      initialises all 64 accumulator columns)
      float inv0;
      const float sc0 = pow2_scale(gmax_next, inv0);
      // global loads are always issued BEFORE waiting for the tensor core and global stores AFTER the
      // hand-off to the issuer, so HBM latency and the release-fence of the arrive overlap the MMAs
      uint32_t m0 = 0, m1 = 0;
      auto load_mask = [&](int l) {
        const uint32_t* mk = masks + ((size_t)tile * (L * 2) + l * 2) * 128 + t;
        m0 = valid ? mk[0] : 0u;
        m1 = valid ? mk[128] : 0u;
      };
#pragma unroll 1
      for (int s = S - 1; s >= 0; --s) {
        store_row_split(g, sc0, a_hi, a_lo, t);
        publish();
        if (s > 0) load_gphi(tile, meta.slot_feature[s - 1]);   // next slot's gradient rows travel while the tensor core works
        else load_mask(L - 1);                            // ReLU mask of the first chain step
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
          const bool on = ((c < 32 ? m0 >> c : m1 >> (c - 32)) & 1u) != 0u;
          acc[c] = on ? acc[c] * unscale : 0.f;
          amax = fmaxf(amax, fabsf(acc[c]));
        }
        float inv;
        const float sc = pow2_scale(amax, inv);
        store_row_split(acc, sc, a_hi, a_lo, t);
        unscale = inv;
        publish();
        {   // batch maximum of |gz[l]|: one atomic per warp (non-negative floats order like ints)
          float wm = amax;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
          if ((t & 31) == 0) atomicMax(gmaxes + 1 + l, __float_as_int(wm));
        }
        if (valid) {
          float* gl = gz + ((size_t)tile * (L * kH) + l * kH) * 128 + t;
#pragma unroll
          for (int c = 0; c < 64; ++c) gl[c * 128] = acc[c];   // sorted unit order, like acts
        }
        if (l > 0) load_mask(l - 1);
        if (l == L - 1) prefetch_tile(tile + tstride);   // g[] is free from here on
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
static int launch_dgrad(const float* gphi, const float* gmax, const uint32_t* masks, const float* gvd, int64_t n,
                        const float* params, const int32_t* order, float* gz, float* gv, unsigned char* image,
                        int* gmaxes, cudaStream_t st) {
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
  MFB_CUDA(cudaMemsetAsync(image, 0, (size_t)img_bytes, st));
  nsf_tc_dgrad_prepare_kernel<<<8, 256, 0, st>>>(params, D, L, meta, image, img_bytes);
  int rc = launch_status();
  if (rc) return rc;
  const size_t smem = (size_t)img_bytes + kWG * kABytes + 128 + 1024;
  auto kern = nsf_tc_dgrad_kernel<D, L>;
  MFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t ntiles = (n + 127) / 128;
  int64_t grid = sm_count();
  if (grid * kWG > ntiles) grid = (ntiles + kWG - 1) / kWG;
  kern<<<(int)grid, kThreads, smem, st>>>(gphi, gmax, masks, gvd, n, image, meta, gz, gv, gmaxes);
  return launch_status();
}

// =============================================================================================
// weight + bias gradients of one layer: dW[i][j] = sum_p H[i][p] G[j][p] for the eight (G, H) pairs
//   (dL/dphi_f, h3) x S,  (g3, h2),  (g2, h1),  (g1, v^T)
// as tcgen05 GEMMs whose K dimension is the particle axis.  The workspace matrices are tile-major (see
// BwdIO in nsf_tc.cu): the 128 particles of a row of a tile are contiguous, exactly a K-major operand row.
// A producer thread streams the 64-row blocks (32 KB, contiguous) into a shared-memory staging ring
// with cp.async.bulk, two blocks ahead of their use, so HBM latency is never on the critical path;
// loader warps turn the staged rows into (hi, lo) fp16 SWIZZLE_128B tiles (64 rows x 128 K, two 64-wide
// halves), the issuer warp multiplies them into accumulators that stay in TMEM for the whole
// launch (464 columns: 5 x 64 + 64 + 64 + 16), bias gradients are row sums taken on the way.  G is
// scaled by a power of two from the batch maximum (fp16 has no range for 1/N-sized gradients).
// =============================================================================================
constexpr int kWgTile = 32768;                 // one operand tile: hi (2 halves x 8 KB) | lo (2 x 8 KB)
#ifndef MFB_WG_NOSLACK
#define MFB_WG_NOSLACK 1   // three staging blocks fit only without the 1 KB alignment slack
#endif
#ifndef MFB_WG_H
#define MFB_WG_H 2
#endif
#ifndef MFB_WG_S
#define MFB_WG_S 3   // bytes in flight bound this kernel: 0.677 -> 0.604 ms per layer and 1e6 particles with a third block
#endif
#ifndef MFB_WG_LOADERS
#define MFB_WG_LOADERS 12
#endif
constexpr int kWgG = 2, kWgH = MFB_WG_H;       // ring depths of the fp16 operand tiles
constexpr int kWgS = MFB_WG_S;                        // ring depth of the fp32 staging blocks (64 rows x 512 B, filled by TMA)
constexpr int kWgStage = 32768;
constexpr int kWgLoaders = MFB_WG_LOADERS;     // loader warps
constexpr int kWgThreads = (kWgLoaders + 4) * 32;   // + issuer, producer, two spare warps (epilogue uses warps 0..3)
constexpr int kWgCols = 464;                   // accumulator columns

struct WgradMeta {
  int perm[kH];               // acts / gz rows are in sorted unit order: row c = hidden unit perm[c]
  int slot_feature[kMaxDim];
  int const_feature;
  int nslots;
};

__host__ __device__ constexpr int wgrad_rows(int D) { return kWgCols + D + 3; }   // + bias rows: D features, 3 hidden

template <int D>
__global__ void __launch_bounds__(kWgThreads, 1)
nsf_tc_wgrad_kernel(const float* __restrict__ gphi /* D*kGRows rows, compact */, const float* __restrict__ gz /* 192 rows */,
                    const float* __restrict__ acts /* per tile: 3 x 32 KB fp16 (hi, lo) operand tiles (nsf_tc.cu store_hidden) */, const float* __restrict__ v /* [n][D] */,
                    int64_t n, const int* __restrict__ gmaxes, const __grid_constant__ WgradMeta meta,
                    float* __restrict__ partial /* [grid][wgrad_rows][64] */) {
  constexpr int S = D - 1;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
#if MFB_WG_NOSLACK
  unsigned char* smem = smem_raw;   // no static shared memory in this kernel: the dynamic window starts 1 KB aligned
  if (smem_u32(smem_raw) & 1023u) __trap();
#else
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
#endif
  unsigned char* g_ring = smem;
  unsigned char* h_ring = smem + kWgG * kWgTile;
  // an M = 128 MMA reads 64 rows past a 64-row tile half: the staging ring that follows is the slack
  float* stage_ring = reinterpret_cast<float*>(smem + (kWgG + kWgH) * kWgTile);
  float* bias = reinterpret_cast<float*>(smem + (kWgG + kWgH) * kWgTile + kWgS * kWgStage);   // [D + 3][64]
  uint64_t* g_full = reinterpret_cast<uint64_t*>(bias + (D + 3) * 64);
  uint64_t* g_empty = g_full + kWgG;
  uint64_t* h_full = g_empty + kWgG;
  uint64_t* h_empty = h_full + kWgH;
  uint64_t* s_full = h_empty + kWgH;
  uint64_t* s_empty = s_full + kWgS;
  uint64_t* all_done = s_empty + kWgS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(all_done + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (D + 3) * 64; i += kWgThreads) bias[i] = 0.f;
  if (tid == 0) {
    for (int i = 0; i < kWgG; ++i) {
      mbar_init(&g_full[i], kWgLoaders);
      mbar_init(&g_empty[i], 1);
    }
    for (int i = 0; i < kWgH; ++i) {
      mbar_init(&h_full[i], kWgLoaders);
      mbar_init(&h_empty[i], 1);
    }
    for (int i = 0; i < kWgS; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], kWgLoaders);
    }
    mbar_init(all_done, 1);
    fence_mbar_init();
  }
  if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  const int64_t ntiles = (n + 127) / 128;
  // scales of the G operands from the batch maxima: [0] dL/dphi, [1 + l] gz[l]
  float gsc[4], ginv[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) gsc[i] = pow2_scale(__int_as_float(gmaxes[i]), ginv[i]);

  if (warp < kWgLoaders) {
    // ===== loaders: rows w, w + 12, ... of every tile =====
    uint32_t gi = 0, hi_ = 0, sc_ = 0;   // running use counters of the rings
    constexpr int kRows = (64 + kWgLoaders - 1) / kWgLoaders;
    // rows w, w + 12, ... of the next staged block -> registers (masked beyond particle n), then the
    // staging slot is handed back to the producer
    auto take = [&](float (&x)[kRows][4], int nrows, int64_t p0, bool strided_v, bool compact = false) {
      const uint32_t slot = sc_ % kWgS, par = (sc_ / kWgS) & 1;
      mbar_wait_bounded(&s_full[slot], par);
      const float* st = stage_ring + (size_t)slot * (kWgStage / 4);
#pragma unroll
      for (int i = 0; i < kRows; ++i) {
        const int r = warp + i * kWgLoaders;
        x[i][0] = x[i][1] = x[i][2] = x[i][3] = 0.f;
        if (r < nrows) {
          constexpr int NB2 = kGRows - 4;
          if (compact && r >= NB2) {   // derivative block of a compact dL/dphi block: rows (left, right, bin) -> row r
            const float4 lf = *reinterpret_cast<const float4*>(st + (size_t)NB2 * 128 + 4 * lane);
            const float4 rt = *reinterpret_cast<const float4*>(st + (size_t)(NB2 + 1) * 128 + 4 * lane);
            const float4 kb = *reinterpret_cast<const float4*>(st + (size_t)(NB2 + 2) * 128 + 4 * lane);
            const float j = (float)(r - NB2);
            x[i][0] = (j == kb.x - 1.f) ? lf.x : ((j == kb.x) ? rt.x : 0.f);
            x[i][1] = (j == kb.y - 1.f) ? lf.y : ((j == kb.y) ? rt.y : 0.f);
            x[i][2] = (j == kb.z - 1.f) ? lf.z : ((j == kb.z) ? rt.z : 0.f);
            x[i][3] = (j == kb.w - 1.f) ? lf.w : ((j == kb.w) ? rt.w : 0.f);
          } else if (!strided_v) {
            const float4 q = *reinterpret_cast<const float4*>(st + (size_t)r * 128 + 4 * lane);
            x[i][0] = q.x; x[i][1] = q.y; x[i][2] = q.z; x[i][3] = q.w;
          } else {   // the block is v[p0 .. p0+127][D], particle-major: row r of v^T is feature r
#pragma unroll
            for (int e = 0; e < 4; ++e) x[i][e] = st[(4 * lane + e) * D + r];
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) x[i][e] = (p0 + 4 * lane + e < n) ? x[i][e] : 0.f;
        }
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_empty[slot])) : "memory");
      ++sc_;
    };
    auto convert = [&](unsigned char* tile, const float (&x)[kRows][4], int nrows, float scale, float* bsum) {
#pragma unroll
      for (int i = 0; i < kRows; ++i) {
        const int r = warp + i * kWgLoaders;
        if (r >= nrows) break;
        if (bsum) {
          float sacc = (x[i][0] + x[i][1]) + (x[i][2] + x[i][3]);
          sacc = warp_sum(sacc);
          if (lane == 0) bsum[r] += sacc;
        }
        if (tile) {
          const float x0 = x[i][0] * scale, x1 = x[i][1] * scale, x2 = x[i][2] * scale, x3 = x[i][3] * scale;
          const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
          const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
          const __half2 l01 = __floats2half2_rn(x0 - f01.x, x1 - f01.y), l23 = __floats2half2_rn(x2 - f23.x, x3 - f23.y);
          // K index 4*lane: half = lane / 16, 16-byte chunk (lane % 16) / 2, upper or lower 8 bytes of it
          unsigned char* dst = tile + (lane >> 4) * 8192 + umma::sw128_offset(r, (lane & 15) >> 1) + (lane & 1) * 8;
          *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
          *reinterpret_cast<uint2*>(dst + 16384) =
              make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
        }
      }
    };
    auto fill_g = [&](int64_t p0, float scale, float* bsum, bool compact = false) {
      float x[kRows][4];
      take(x, 64, p0, false, compact);
      const uint32_t slot = gi % kWgG, par = (gi / kWgG) & 1;
      mbar_wait_bounded(&g_empty[slot], par ^ 1);
      convert(g_ring + slot * kWgTile, x, 64, scale, bsum);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&g_full[slot])) : "memory");
      ++gi;
    };
    auto fill_h = [&](int nrows, int64_t p0, bool strided_v) {
      float x[kRows][4];
      take(x, nrows, p0, strided_v);
      const uint32_t slot = hi_ % kWgH, par = (hi_ / kWgH) & 1;
      mbar_wait_bounded(&h_empty[slot], par ^ 1);
      convert(h_ring + slot * kWgTile, x, nrows, 1.0f, nullptr);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&h_full[slot])) : "memory");
      ++hi_;
    };
    // the block order below is the producer's (next branch) and the issuer's
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t p0 = tile * 128;
      // h3, h2, h1 come straight from HBM into the H ring (activation producer below): uses 0..2 of this tile
      for (int s = 0; s < S; ++s) fill_g(p0, gsc[0], bias + meta.slot_feature[s] * 64, true);   // dL/dphi of the slots
      {                                                                                   // bias-only feature: row sums
        float x[kRows][4];
        take(x, 64, p0, false, true);
        convert(nullptr, x, 64, 1.0f, bias + meta.const_feature * 64);
      }
      fill_g(p0, gsc[3], bias + (D + 0) * 64);                                            // g3
      fill_g(p0, gsc[2], bias + (D + 1) * 64);                                            // g2
      hi_ += 3;                                                                           // H ring uses of h3, h2, h1
      fill_h(D, p0, true);                                                                // v^T
      fill_g(p0, gsc[1], bias + (D + 2) * 64);                                            // g1
    }
  } else if (warp == kWgLoaders + 1) {
    // ===== producer: one thread streams the blocks into the staging ring =====
    if (lane == 0) {
      uint32_t sc_ = 0;
      auto push = [&](const float* src, uint32_t bytes) {
        const uint32_t slot = sc_ % kWgS, par = (sc_ / kWgS) & 1;
        mbar_wait_bounded(&s_empty[slot], par ^ 1);
        float* dst = stage_ring + (size_t)slot * (kWgStage / 4);
        const uint32_t bulk = bytes & ~15u;
        for (uint32_t i = bulk / 4; i < bytes / 4; ++i) dst[i] = src[i];   // ragged tail (< 16 B) of the last v block
        mbar_expect_tx(&s_full[slot], bulk);
        if (bulk) tma_load_1d(dst, src, bulk, &s_full[slot]);
        ++sc_;
      };
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t p0 = tile * 128;
        const float* zt = gz + (size_t)tile * (3 * kH) * 128;
        const float* pt = gphi + (size_t)tile * (D * kGRows) * 128;
        constexpr uint32_t kGBlock = kGRows * 128 * 4;                                      // compact dL/dphi block
        for (int s = 0; s < S; ++s) push(pt + meta.slot_feature[s] * kGRows * 128, kGBlock);
        push(pt + meta.const_feature * kGRows * 128, kGBlock);
        push(zt + 2 * kH * 128, kWgStage);                                                  // g3
        push(zt + 1 * kH * 128, kWgStage);                                                  // g2
        const int64_t rows = (n - p0 < 128) ? (n - p0) : 128;
        push(v + p0 * D, (uint32_t)(rows * D * 4));                                         // v rows of the tile
        push(zt, kWgStage);                                                                 // g1
      }
    }
  } else if (warp == kWgLoaders + 2) {
    // ===== activation producer: the (hi, lo) fp16 operand tiles the first backward kernel mirrored to HBM go
    //       straight into the H ring (uses 0, 1, 2 of every tile = h3, h2, h1; use 3 = v^T is built by the loaders).
    //       The barrier expects kWgLoaders arrivals per phase: this thread supplies all of them.
    if (lane == 0) {
      uint32_t use = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, use += 4) {
        const unsigned char* img = reinterpret_cast<const unsigned char*>(acts) + (size_t)tile * 3 * kWgTile;
        for (int u = 0; u < 3; ++u) {
          const uint32_t idx = use + u, slot = idx % kWgH, par = (idx / kWgH) & 1;
          mbar_wait_bounded(&h_empty[slot], par ^ 1);
          mbar_expect_tx(&h_full[slot], (uint32_t)kWgTile);
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&h_full[slot])), "r"(kWgLoaders - 1) : "memory");
          tma_load_1d(h_ring + slot * kWgTile, img + (size_t)(2 - u) * kWgTile, (uint32_t)kWgTile, &h_full[slot]);
        }
      }
    }
  } else if (warp == kWgLoaders) {
    // ===== issuer =====
    uint32_t gi = 0, hi_ = 0;
    const uint32_t idesc64 = umma::make_idesc_f16(128, 64), idesc16 = umma::make_idesc_f16(128, 16);
    bool first = true;
    // one (G, H) pair: K = 128 particles.  h_mn: H is an activation tile as the forward kernel wrote it -- [128
    // particles][64 units], SWIZZLE_128B -- read as an MN-major operand (units contiguous, 8 particles per swizzle
    // atom, 16 particles = one K step = 2 KB); otherwise a K-major tile built by the loaders (two 64-particle halves)
    auto job = [&](uint32_t dcol, uint32_t hslot, uint32_t idesc, bool h_mn) {
      const uint32_t slot = gi % kWgG, par = (gi / kWgG) & 1;
      mbar_wait_polite(&g_full[slot], par);
      umma::fence_after_sync();
      const uint32_t ga = smem_u32(g_ring + slot * kWgTile), ha = smem_u32(h_ring + hslot * kWgTile);
      if (elect_one()) {
        if (h_mn) {
          const uint32_t id = idesc | (1u << 16);   // B operand MN-major
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint64_t aH = umma::desc_advance_k(umma::make_desc_sw128(ga + (kk >> 2) * 8192), kk & 3);
            const uint64_t aL = umma::desc_advance_k(umma::make_desc_sw128(ga + 16384 + (kk >> 2) * 8192), kk & 3);
            const uint64_t bH = umma::make_desc_sw128(ha + kk * 2048), bL = umma::make_desc_sw128(ha + 16384 + kk * 2048);
            umma::mma_f16_ss(dcol, aH, bL, id, (first && kk == 0) ? 0u : 1u);
            umma::mma_f16_ss(dcol, aL, bH, id, 1);
          }
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint64_t aH = umma::desc_advance_k(umma::make_desc_sw128(ga + (kk >> 2) * 8192), kk & 3);
            umma::mma_f16_ss(dcol, aH, umma::make_desc_sw128(ha + kk * 2048), id, 1);
          }
        } else {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint64_t aH = umma::make_desc_sw128(ga + half * 8192), aL = umma::make_desc_sw128(ga + 16384 + half * 8192);
            const uint64_t bH = umma::make_desc_sw128(ha + half * 8192), bL = umma::make_desc_sw128(ha + 16384 + half * 8192);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) mma_cross(dcol, aH, aL, bH, bL, ks, idesc, (first && half == 0 && ks == 0) ? 0u : 1u);
          }
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint64_t aH = umma::make_desc_sw128(ga + half * 8192), bH = umma::make_desc_sw128(ha + half * 8192);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) mma_main(dcol, aH, bH, ks, idesc);
          }
        }
        umma::commit(&g_empty[slot]);
      }
      __syncwarp();
      ++gi;
    };
    auto take_h = [&]() {
      const uint32_t slot = hi_ % kWgH, par = (hi_ / kWgH) & 1;
      mbar_wait_polite(&h_full[slot], par);
      ++hi_;
      return slot;
    };
    auto release_h = [&](uint32_t slot) {
      if (elect_one()) umma::commit(&h_empty[slot]);
      __syncwarp();
    };
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      uint32_t hs = take_h();                                   // h3
      for (int s = 0; s < S; ++s) job(tmem_base + 64 * s, hs, idesc64, true);
      release_h(hs);
      hs = take_h();                                            // h2
      job(tmem_base + 320, hs, idesc64, true);
      release_h(hs);
      hs = take_h();                                            // h1
      job(tmem_base + 384, hs, idesc64, true);
      release_h(hs);
      hs = take_h();                                            // v^T
      job(tmem_base + 448, hs, idesc16, false);
      release_h(hs);
      first = false;
    }
    if (elect_one()) umma::commit(all_done);
    __syncwarp();
  }
  // ===== epilogue: accumulators (lanes 0..63 = G row j) and bias sums -> this CTA's partial block
  __syncthreads();   // every loader has added its last bias sums
  float* out = partial + (size_t)blockIdx.x * wgrad_rows(D) * 64;
  if (warp < 2) {
    mbar_wait_bounded(all_done, 0);
    umma::fence_after_sync();
    const bool any = (int64_t)blockIdx.x < ntiles;
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int j = warp * 32 + lane;
    for (int c0 = 0; c0 < kWgCols; c0 += 32) {
      float acc[32];
      umma::tmem_ld32(taddr + c0, acc);
      const float unscale = c0 < 320 ? ginv[0] : (c0 < 384 ? ginv[3] : (c0 < 448 ? ginv[2] : ginv[1]));
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (c0 + i < kWgCols) out[(size_t)(c0 + i) * 64 + j] = any ? acc[i] * unscale : 0.f;
    }
    umma::fence_before_sync();
  } else if (warp >= 2 && warp < 4) {
    for (int i = (warp - 2) * 32 + lane; i < (D + 3) * 64; i += 64) out[(size_t)kWgCols * 64 + i] = bias[i];
  }
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base, 512);
}

// gparams (packed forward layout) (+)= sum over CTAs of the partial blocks
__global__ void nsf_tc_wgrad_reduce_kernel(const float* __restrict__ partial, int nparts, int D, int L,
                                           const __grid_constant__ WgradMeta meta, float* __restrict__ gparams,
                                           int accumulate) {
  const int rows = wgrad_rows(D);
  // eight lanes per entry: lane q adds partials q, q + 8, ...; fixed shuffle tree => deterministic
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int idx = gid >> 3, q = gid & 7;
  const bool live = idx < rows * 64;
  float sacc = 0.f;
  if (live)
    for (int k = q; k < nparts; k += 8) sacc += partial[(size_t)k * rows * 64 + idx];
  sacc += __shfl_xor_sync(0xffffffffu, sacc, 4);
  sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
  sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
  if (!live || q != 0) return;
  const int c = idx / 64, j = idx % 64;
  const int64_t off_w1 = 0, off_b1 = (int64_t)D * kH, off_hid = off_b1 + kH;
  const int64_t off_wout = off_hid + (int64_t)(L - 1) * (kH * kH + kH);
  const int64_t off_bout = off_wout + (int64_t)D * kH * kPP;
  int64_t dst = -1;
  // hidden-unit indices arrive in the sorted order of the workspace rows: unit = perm[sorted index]
  const int uj = meta.perm[j];
  if (c < 320) {                                  // dWout_t[f][i][q]: column block s, i = perm[c % 64], q = j
    const int sl = c / 64;
    if (sl < meta.nslots) dst = off_wout + ((int64_t)meta.slot_feature[sl] * kH + meta.perm[c % 64]) * kPP + j;
  } else if (c < 384) {                           // W3: hidden block l = 1, Wt[in i][out j]
    dst = off_hid + 1 * (kH * kH + kH) + (int64_t)meta.perm[c - 320] * kH + uj;
  } else if (c < 448) {                           // W2: hidden block l = 0
    dst = off_hid + (int64_t)meta.perm[c - 384] * kH + uj;
  } else if (c < kWgCols) {                       // W1t[i][j]: i = input feature
    if (c - 448 < D) dst = off_w1 + (int64_t)(c - 448) * kH + uj;
  } else {
    const int kind = c - kWgCols;                 // bias rows: features 0..D-1, then g3, g2, g1
    if (kind < D) dst = off_bout + (int64_t)kind * kPP + j;
    else if (kind == D) dst = off_hid + 1 * (kH * kH + kH) + kH * kH + uj;
    else if (kind == D + 1) dst = off_hid + kH * kH + uj;
    else dst = off_b1 + uj;
  }
  if (dst >= 0) gparams[dst] = accumulate ? gparams[dst] + sacc : sacc;
}

// the output-layer weights of the bias-only feature are masked: their gradient block is defined as zero
__global__ void nsf_tc_wgrad_zero_const_kernel(float* __restrict__ gparams, int64_t off, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) gparams[off + i] = 0.f;
}

template <int D>
static int launch_wgrad(const float* gphi, const float* gz, const float* acts, const float* v, int64_t n,
                        const int32_t* order, const int* gmaxes, float* partial, float* gparams, int accumulate,
                        cudaStream_t st) {
  constexpr int L = 3;
  WgradMeta meta = {};
  int cls[kH];
  hidden_classes(D, cls, meta.perm);
  int feat_of_order[kMaxDim];
  for (int i = 0; i < D; ++i) feat_of_order[order[i]] = i;
  meta.nslots = D - 1;
  meta.const_feature = feat_of_order[0];
  for (int s = 0; s < D - 1; ++s) meta.slot_feature[s] = feat_of_order[s + 1];
  const size_t smem = (size_t)(kWgG + kWgH) * kWgTile + (size_t)kWgS * kWgStage + (size_t)(D + 3) * 64 * 4 + 256 + (MFB_WG_NOSLACK ? 0 : 1024);
  auto kern = nsf_tc_wgrad_kernel<D>;
  MFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t ntiles = (n + 127) / 128;
  int grid = sm_count();
  if (grid > ntiles) grid = (int)ntiles;
  kern<<<grid, kWgThreads, smem, st>>>(gphi, gz, acts, v, n, gmaxes, meta, partial);
  int rc = launch_status();
  if (rc) return rc;
  const int entries = wgrad_rows(D) * 64;
  nsf_tc_wgrad_reduce_kernel<<<(entries * 8 + 255) / 256, 256, 0, st>>>(partial, grid, D, L, meta, gparams, accumulate);
  if (!accumulate) {
    const int64_t off_wout = (int64_t)D * kH + kH + (int64_t)(L - 1) * (kH * kH + kH);
    nsf_tc_wgrad_zero_const_kernel<<<(kH * kPP + 255) / 256, 256, 0, st>>>(
        gparams, off_wout + (int64_t)meta.const_feature * kH * kPP, kH * kPP);
  }
  return launch_status();
}

}  // namespace tc

// Called by the backward orchestration in nsf_bwd.cu.  gz receives dL/d(pre-activation) of the three
// hidden layers ([3][64][n], feature-major like acts), gv the gradient w.r.t. the layer input.
// image: scratch of nsf_tc_dgrad_image_bytes(d) bytes.  Returns MFB_E_UNSUPPORTED for shapes that are
// not compiled (the caller then uses the CUDA-core kernels).
int64_t nsf_tc_dgrad_image_bytes(int d) { return (d >= 2 && d <= 6) ? tc::dgrad_image_bytes(d, 3) : 0; }

int nsf_tc_dgrad(const float* gphi, const float* gmax, const uint32_t* masks, const float* gvd, int64_t n, int d,
                 int hidden_layers, const float* params, const int32_t* order, float* gz, float* gv, void* image,
                 int* gmaxes, cudaStream_t st) {
  if (hidden_layers != 3 || d < 2 || d > 6) return MFB_E_UNSUPPORTED;
  unsigned char* img = reinterpret_cast<unsigned char*>(image);
  switch (d) {
    case 2: return tc::launch_dgrad<2>(gphi, gmax, masks, gvd, n, params, order, gz, gv, img, gmaxes, st);
    case 3: return tc::launch_dgrad<3>(gphi, gmax, masks, gvd, n, params, order, gz, gv, img, gmaxes, st);
    case 4: return tc::launch_dgrad<4>(gphi, gmax, masks, gvd, n, params, order, gz, gv, img, gmaxes, st);
    case 5: return tc::launch_dgrad<5>(gphi, gmax, masks, gvd, n, params, order, gz, gv, img, gmaxes, st);
    case 6: return tc::launch_dgrad<6>(gphi, gmax, masks, gvd, n, params, order, gz, gv, img, gmaxes, st);
    default: return MFB_E_UNSUPPORTED;
  }
}

int64_t nsf_tc_wgrad_partial_floats(int d) {
  return (d >= 2 && d <= 6) ? (int64_t)sm_count() * tc::wgrad_rows(d) * 64 : 0;
}

int nsf_tc_wgrad(const float* gphi, const float* gz, const float* acts, const float* v, int64_t n, int d,
                 int hidden_layers, const int32_t* order, const int* gmaxes, float* partial, float* gparams,
                 int accumulate, cudaStream_t st) {
  if (hidden_layers != 3 || d < 2 || d > 6) return MFB_E_UNSUPPORTED;
  switch (d) {
    case 2: return tc::launch_wgrad<2>(gphi, gz, acts, v, n, order, gmaxes, partial, gparams, accumulate, st);
    case 3: return tc::launch_wgrad<3>(gphi, gz, acts, v, n, order, gmaxes, partial, gparams, accumulate, st);
    case 4: return tc::launch_wgrad<4>(gphi, gz, acts, v, n, order, gmaxes, partial, gparams, accumulate, st);
    case 5: return tc::launch_wgrad<5>(gphi, gz, acts, v, n, order, gmaxes, partial, gparams, accumulate, st);
    case 6: return tc::launch_wgrad<6>(gphi, gz, acts, v, n, order, gmaxes, partial, gparams, accumulate, st);
    default: return MFB_E_UNSUPPORTED;
  }
}

}  // namespace mfb
