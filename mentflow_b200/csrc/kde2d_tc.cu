// Two-dimensional KDE screens on the tensor cores.
//
// The reference evaluates a screen as the GEMM  P = Kx^T Ky  over the particle axis
// (mentflow/diagnostics/histogram.py:47-74: two dense (N, B) Gaussian kernel matrices and a matmul).
// This kernel does the same contraction with tcgen05: for a tile of 64 particles the loader warps write
// the dense kernel rows  Kx[bx][p] = exp(-(ux_p - cx_bx)^2 / 2 sx^2)  and  Ky[by][p]  straight into
// K-major SWIZZLE_128B operand tiles (the particle axis is the MMA K dimension), one elected thread
// multiplies them into an accumulator P[bx][by] that stays in TMEM while the CTA walks its particles,
// five screens at a time (5 x 96 of the 512 TMEM columns).  Nothing is windowed: the dense rows keep
// every tail the reference keeps.
//
// Precision: operands are split bf16 pairs (hi, mid), x = hi + mid to 2^-17; bf16 and not fp16 because
// the far tails (1e-12 of the peak and below, which the KL's log(p + 1e-12) can see) need fp32's
// exponent range.  Per 16 particles three MMAs, cross terms first: mid*hi + hi*mid + hi*hi (products of
// bf16 are exact in the fp32 accumulator); the dropped mid*mid term is 2^-18.  All terms are positive,
// so every bin is accurate to ~2e-5 relative.
//
// Determinism: a CTA owns a contiguous particle range and accumulates in a fixed order; per-CTA partial
// screens are merged in CTA order by kde2d_tc_reduce_kernel, which also writes the 44-bit fixed-point
// planes that ranks all-reduce (integer adds: independent of the reduction order across ranks).
#include <cuda_bf16.h>

#include "nsf_tc_common.cuh"

#ifndef MFB_KDE2D_PACKED
#define MFB_KDE2D_PACKED 1
#endif

namespace mfb {
namespace k2tc {

using tc::elect_one;
using tc::mbar_wait_bounded;
using tc::mbar_wait_polite;

constexpr int kThreads = 512;
constexpr int kLoaderWarps = 12;                 // warps 0..11 build operands, warp 12 issues the MMAs
constexpr int kLoaders = kLoaderWarps * 32;
constexpr int kKT = 64;                          // particles per operand tile (one 128-byte K row)
constexpr int kCRow = kKT + 4;                   // coordinate row: the second half shifted by 16 B, so that the eight
                                                 // 32-byte chunks a quarter warp reads fall into distinct banks
__host__ __device__ constexpr int crow_index(int p) { return p + ((p >> 5) << 2); }
constexpr int kGroup = 5;                        // screens resident in TMEM
constexpr int kCoordPer = (kGroup * 2 * kKT + kLoaders - 1) / kLoaders;   // coordinates per loader thread and tile
constexpr int kCols = 96;                        // accumulator columns per screen (By <= 96)
constexpr int kABytes = 128 * 128;               // 128 rows x 128 B (Bx <= 128; M = 128 reads all of them)
constexpr int kBBytes = 96 * 128;
constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;   // A hi | A mid | B hi | B mid = 56 KB
constexpr int kStages = 3;
constexpr float kHalfScale = 4194304.0f;         // 2^22 (fixed-point planes of kde2d.cu)

__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return umma::make_idesc_f16(M, N) | (1u << 7) | (1u << 10);   // a_format = b_format = BF16
}

// two fp32 lanes for one issue slot (FADD2 / FMUL2; IEEE round-to-nearest like the scalar forms)
__device__ __forceinline__ void sub2(float a0, float a1, float b, float& c0, float& c1) {
  unsigned long long a, bb, c;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(-b));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(bb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(c0), "=f"(c1) : "l"(c));
}
__device__ __forceinline__ void mul2(float a0, float a1, float b0, float b1, float& c0, float& c1) {
  unsigned long long a, b, c;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(c0), "=f"(c1) : "l"(c));
}

struct ScreenAxis {
  float c0, inv_delta, scale;   // scale = sqrt(0.5 log2 e) * delta / sigma: kernel value = 2^-((a - r) scale)^2
};

__global__ void __launch_bounds__(kThreads, 1)
kde2d_tc_kernel(const float* __restrict__ x, int64_t n, int d, const float* __restrict__ proj /* [K][2][d] */,
                const float* __restrict__ geom, int K, int BX, int BY, int64_t tiles_per_cta,
                float* __restrict__ partial /* [gridDim.x][K][BX][BY] */) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* stages = smem;
  float* coord = reinterpret_cast<float*>(smem + kStages * kStageBytes);   // [2 buffers][kGroup][2][kCRow]
  float* s_w = coord + 2 * kGroup * 2 * kCRow;                               // [kGroup][2][kMaxDim]
  ScreenAxis* s_ax = reinterpret_cast<ScreenAxis*>(s_w + kGroup * 2 * kMaxDim);   // [kGroup][2]
  uint64_t* full = reinterpret_cast<uint64_t*>(s_ax + kGroup * 2);
  uint64_t* empty = full + kStages;
  uint64_t* acc_full = empty + kStages;
  uint64_t* acc_free = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], kLoaderWarps);
      mbar_init(&empty[i], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_free, 4);
    fence_mbar_init();
  }
  if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  const int64_t ntiles = (n + kKT - 1) / kKT;
  const int64_t t0 = (int64_t)blockIdx.x * tiles_per_cta;
  int64_t t1 = t0 + tiles_per_cta;
  if (t1 > ntiles) t1 = ntiles;
  const int ngroups = (K + kGroup - 1) / kGroup;
  const int nmma = (BY + 15) & ~15;                 // MMA N
  const uint32_t idesc = make_idesc_bf16(128, nmma);

  uint32_t it = 0;   // running stage use counter (same sequence in loaders and issuer)
  for (int g = 0; g < ngroups; ++g) {
    const int s0 = g * kGroup;
    const int ns = (K - s0 < kGroup) ? (K - s0) : kGroup;
    // ---- per-group constants (all threads take part; previous group's readers are past the barrier below)
    __syncthreads();
    for (int i = tid; i < ns * 2 * kMaxDim; i += kThreads) {
      const int s = i / (2 * kMaxDim), a = (i / kMaxDim) % 2, c = i % kMaxDim;
      s_w[i] = c < d ? proj[((size_t)(s0 + s) * 2 + a) * d + c] : 0.f;
    }
    for (int i = tid; i < ns * 2; i += kThreads) {
      const float* gp = geom + (size_t)(2 * s0 + i) * MFB_GEOM_STRIDE;
      const float r = gp[1] / gp[2];
      s_ax[i].c0 = gp[0];
      s_ax[i].inv_delta = 1.0f / gp[1];
      s_ax[i].scale = sqrtf(0.5f * kLog2e) * r;
    }
    __syncthreads();

    if (warp < kLoaderWarps) {
      // ===== loaders =====
      int cb = 0;   // coordinate buffer of this K tile
      // Scaled bin coordinates a = (u - c0) / delta of a tile's particles on both axes of every screen.
      // The particle rows of tile t+1 are loaded into registers before the operands of tile t are generated
      // and turned into coordinates afterwards, so their L2 / HBM latency is never waited for.
      float xr[kCoordPer][kMaxDim];
      auto load_rows = [&](int64_t t) {
#pragma unroll
        for (int j = 0; j < kCoordPer; ++j) {
          const int i = tid + j * kLoaders;
          const int64_t p = t * kKT + (i % kKT);
          const bool live = i < ns * 2 * kKT && t < t1 && p < n;
#pragma unroll
          for (int c = 0; c < kMaxDim; ++c) xr[j][c] = (live && c < d) ? x[p * d + c] : 0.f;
        }
      };
      auto store_coords = [&](int64_t t, float* cbuf) {
#pragma unroll
        for (int j = 0; j < kCoordPer; ++j) {
          const int i = tid + j * kLoaders;
          if (i >= ns * 2 * kKT) continue;
          const int sa = i / kKT, p = i % kKT;       // sa = screen * 2 + axis
          float a = -1.0e4f;                         // no particle: every kernel value is 0
          if (t * kKT + p < n) {
            const float* w = s_w + sa * kMaxDim;
            float u = 0.f;
#pragma unroll
            for (int c = 0; c < kMaxDim; ++c) u = fmaf(w[c], xr[j][c], u);   // w is zero beyond d
            a = (u - s_ax[sa].c0) * s_ax[sa].inv_delta;
            a = fminf(fmaxf(a, -1.0e4f), 1.0e4f);
          }
          cbuf[sa * kCRow + crow_index(p)] = a;
        }
      };
      load_rows(t0);
      store_coords(t0, coord);
      asm volatile("bar.sync 1, %0;" ::"n"(kLoaders) : "memory");
      for (int64_t t = t0; t < t1; ++t, cb ^= 1) {
        float* cbuf = coord + cb * (kGroup * 2 * kCRow);
        load_rows(t + 1);
        for (int s = 0; s < ns; ++s, ++it) {
          const uint32_t slot = it % kStages, par = (it / kStages) & 1;
          mbar_wait_bounded(&empty[slot], par ^ 1);
          unsigned char* st = stages + slot * kStageBytes;
          const float ax_scale = s_ax[2 * s].scale, ay_scale = s_ax[2 * s + 1].scale;
          const float* cx = cbuf + (2 * s) * kCRow;
          const float* cy = cx + kCRow;
          // A thread owns one 8-particle chunk of the tile and every (kLoaders / 8)-th operand row: the chunk's
          // coordinates on both axes are read once per stage and stay in registers (re-reading them per row
          // cost as much shared-memory bandwidth as writing the operands).  Per (row, chunk): dense kernel
          // values 2^-(cs - r s)^2, bf16 (hi, mid), 16 B each.  hi is the value truncated to its upper 16 bits (one
          // AND gives it back as a float, one PRMT packs a pair), mid the rounded remainder: 16 significant bits
          // like a rounded hi, for 9 instead of 15 instructions per value (profiles/r2_kde2d_tc_metrics.txt).
          const int chunk = tid & 7;
          auto operand_rows = [&](const float* crow, float sc, int nrows, unsigned char* hi_tile, unsigned char* mid_tile) {
            float cs[8];
            {
              const float4 a0 = *reinterpret_cast<const float4*>(crow + crow_index(chunk * 8));
              const float4 a1 = *reinterpret_cast<const float4*>(crow + crow_index(chunk * 8) + 4);
              cs[0] = a0.x; cs[1] = a0.y; cs[2] = a0.z; cs[3] = a0.w; cs[4] = a1.x; cs[5] = a1.y; cs[6] = a1.z; cs[7] = a1.w;
            }
            for (int r = tid >> 3; r < nrows; r += kLoaders / 8) {
              const float rs = (float)r;
              uint32_t hi[4], mid[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                // (a - r) first: r is an integer and a is within a few bins of it wherever the value matters, so the
                // difference is exact; scaling the coordinates beforehand costs 2x the rounding error in the exponent
#if MFB_KDE2D_PACKED
                float da, db, ta, tb, qa, qb;
                sub2(cs[2 * e], cs[2 * e + 1], rs, da, db);
                mul2(da, db, sc, sc, ta, tb);
                mul2(ta, tb, ta, tb, qa, qb);
                const float va = fast_exp2(-qa), vb = fast_exp2(-qb);
#else
                const float ta = (cs[2 * e] - rs) * sc, tb = (cs[2 * e + 1] - rs) * sc;
                const float va = fast_exp2(-(ta * ta)), vb = fast_exp2(-(tb * tb));
#endif
                const uint32_t ua = __float_as_uint(va), ub = __float_as_uint(vb);
                hi[e] = __byte_perm(ua, ub, 0x7632);                         // (bf16 of vb) << 16 | bf16 of va, truncated
#if MFB_KDE2D_PACKED
                float ra, rb;
                {
                  unsigned long long pv, ph, pr;
                  asm("mov.b64 %0, {%1, %2};" : "=l"(pv) : "f"(va), "f"(vb));
                  asm("mov.b64 %0, {%1, %2};" : "=l"(ph) : "f"(-__uint_as_float(ua & 0xFFFF0000u)), "f"(-__uint_as_float(ub & 0xFFFF0000u)));
                  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(pr) : "l"(pv), "l"(ph));
                  asm("mov.b64 {%0, %1}, %2;" : "=f"(ra), "=f"(rb) : "l"(pr));
                }
                const __nv_bfloat162 m = __floats2bfloat162_rn(ra, rb);
#else
                const __nv_bfloat162 m = __floats2bfloat162_rn(va - __uint_as_float(ua & 0xFFFF0000u),
                                                               vb - __uint_as_float(ub & 0xFFFF0000u));
#endif
                mid[e] = *reinterpret_cast<const uint32_t*>(&m);
              }
              const uint32_t off = umma::sw128_offset(r, chunk);
              *reinterpret_cast<uint4*>(hi_tile + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(mid_tile + off) = make_uint4(mid[0], mid[1], mid[2], mid[3]);
            }
          };
          operand_rows(cx, ax_scale, BX, st, st + kABytes);
          operand_rows(cy, ay_scale, BY, st + 2 * kABytes, st + 2 * kABytes + kBBytes);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0)
            asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[slot])) : "memory");
        }
        // coordinates of the next tile into the other buffer (its last readers passed the previous barrier)
        if (t + 1 < t1) store_coords(t + 1, coord + (cb ^ 1) * (kGroup * 2 * kCRow));
        asm volatile("bar.sync 1, %0;" ::"n"(kLoaders) : "memory");
      }
      // ===== epilogue of the group (warps 0..3: TMEM lanes 32 w .. 32 w + 31 = screen rows bx) =====
      if (warp < 4) {
        mbar_wait_bounded(acc_full, g & 1);
        umma::fence_after_sync();
        const int bx = warp * 32 + lane;
        for (int s = 0; s < ns; ++s) {
          float* out = partial + (((size_t)blockIdx.x * K + (s0 + s)) * BX + bx) * BY;
          for (int c0 = 0; c0 < nmma; c0 += 32) {
            float acc[32];
            umma::tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(s * kCols + c0), acc);
            if (bx < BX) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (c0 + i < BY) out[c0 + i] = acc[i];
            }
          }
        }
        umma::fence_before_sync();
        __syncwarp();
        if (lane == 0)
          asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(acc_free)) : "memory");
      }
    } else if (warp == kLoaderWarps) {
      // ===== issuer =====
      if (g > 0) {   // the previous group's accumulators have been read out
        mbar_wait_polite(acc_free, (g - 1) & 1);
        umma::fence_after_sync();
      }
      for (int64_t t = t0; t < t1; ++t) {
        for (int s = 0; s < ns; ++s, ++it) {
          const uint32_t slot = it % kStages, par = (it / kStages) & 1;
          mbar_wait_polite(&full[slot], par);
          umma::fence_after_sync();
          const uint32_t sa = smem_u32(stages + slot * kStageBytes);
          const uint64_t aH = umma::make_desc_sw128(sa), aM = umma::make_desc_sw128(sa + kABytes);
          const uint64_t bH = umma::make_desc_sw128(sa + 2 * kABytes), bM = umma::make_desc_sw128(sa + 2 * kABytes + kBBytes);
          const uint32_t dcol = tmem_base + (uint32_t)(s * kCols);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              umma::mma_f16_ss(dcol, umma::desc_advance_k(aM, ks), umma::desc_advance_k(bH, ks), idesc,
                               (t == t0 && ks == 0) ? 0u : 1u);
              umma::mma_f16_ss(dcol, umma::desc_advance_k(aH, ks), umma::desc_advance_k(bM, ks), idesc, 1u);
            }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma::mma_f16_ss(dcol, umma::desc_advance_k(aH, ks), umma::desc_advance_k(bH, ks), idesc, 1u);
            umma::commit(&empty[slot]);
          }
          __syncwarp();
        }
      }
      if (elect_one()) umma::commit(acc_full);
      __syncwarp();
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base, 512);
}

// sums[i] = sum over CTAs (fixed order) of the partial screens; also the two fixed-point planes
// (units 2^-22 and 2^-44) that kde2d.cu's integer path produces, for the all-reduce across ranks
__global__ void kde2d_tc_reduce_kernel(const float* __restrict__ partial, int nparts, int64_t len,
                                       float* __restrict__ sums, unsigned long long* __restrict__ acc) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
    double v = 0.0;
    for (int c = 0; c < nparts; ++c) v += (double)partial[(size_t)c * len + i];
    const double v22 = v * (double)kHalfScale;
    const long long hi = __double2ll_rn(v22);
    const long long lo = __double2ll_rn((v22 - (double)hi) * (double)kHalfScale);
    acc[i] = (unsigned long long)hi;
    acc[len + i] = (unsigned long long)lo;
    sums[i] = (float)(((double)hi + (double)lo / (double)kHalfScale) / (double)kHalfScale);
  }
}

}  // namespace k2tc

// host side, called from kde2d.cu -----------------------------------------------------------------------
bool kde2d_tc_supported(int64_t n, int d, int bx, int by) {
  return n >= 4096 && d >= 1 && d <= kMaxDim && bx >= 2 && bx <= 128 && by >= 2 && by <= k2tc::kCols;
}

int64_t kde2d_tc_partial_bytes(int k, int bx, int by) { return (int64_t)sm_count() * k * bx * by * 4; }

int kde2d_tc_forward(const float* x, int64_t n, int d, const float* proj, const float* geom, int k, int bx, int by,
                     float* sums, unsigned long long* acc, float* partial, cudaStream_t st) {
  using namespace k2tc;
  const int64_t ntiles = (n + kKT - 1) / kKT;
  int64_t grid = sm_count();
  if (grid > ntiles) grid = ntiles;
  const int64_t per = (ntiles + grid - 1) / grid;
  grid = (ntiles + per - 1) / per;     // every CTA has at least one tile
  const size_t smem = (size_t)kStages * kStageBytes + (size_t)2 * kGroup * 2 * kCRow * 4 + (size_t)kGroup * 2 * kMaxDim * 4 +
                      (size_t)kGroup * 2 * sizeof(ScreenAxis) + 256 + 1024;
  MFB_CUDA(cudaFuncSetAttribute(kde2d_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kde2d_tc_kernel<<<(int)grid, kThreads, smem, st>>>(x, n, d, proj, geom, k, bx, by, per, partial);
  int rc = launch_status();
  if (rc) return rc;
  const int64_t len = (int64_t)k * bx * by;
  int rgrid = (int)((len + 255) / 256);
  if (rgrid > 4096) rgrid = 4096;
  kde2d_tc_reduce_kernel<<<rgrid, 256, 0, st>>>(partial, (int)grid, len, sums, acc);
  return launch_status();
}

}  // namespace mfb
