/*
 * mentflow_b200 -- C ABI of the B200-native MENT-Flow hot path.
 *
 * The reference (austin-hoover/ment-flow) is pure Python/PyTorch and has no FFI layer of
 * its own; the functions below are what a reference-side binding (ctypes, see
 * INTEGRATION.md) calls in place of the reference Python functions cited on each entry
 * (paths relative to the reference's `mentflow/` package, per-file line numbers).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 / int32 / int64 data owned by the caller
 *     (PyTorch tensors in the shipped host layer); row-major, contiguous, 16-byte aligned;
 *   - `stream` is a cudaStream_t passed as void*; nothing here allocates, synchronises or
 *     keeps global mutable state, so calls are safe from autograd worker threads;
 *   - return value: 0 on success, a cudaError_t code (>0) for CUDA failures,
 *     negative MFB_E_* for argument errors; mfb_error_string() decodes both;
 *   - workspaces are caller-allocated; query the size with the matching *_workspace_bytes.
 *   - there is no CPU implementation behind any entry point.
 */
#ifndef MENTFLOW_B200_H
#define MENTFLOW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFB_ABI_VERSION 2

#define MFB_E_BADARG (-1)      /* null pointer / non-positive size / unsupported shape   */
#define MFB_E_UNSUPPORTED (-2) /* combination not compiled (e.g. D > 8, hidden != 64)    */
#define MFB_E_WORKSPACE (-3)   /* workspace too small                                    */

/* per-call `flags` of the entry points that have a tensor-core and a CUDA-core implementation of the same
 * result (A/B tests, parity cross-checks).  There is no process-wide switch: the library keeps no mutable state. */
#define MFB_FLAG_NO_TENSOR_CORES 1

int mfb_abi_version(void);
const char* mfb_error_string(int code);
/* number of SMs of the current device (grid sizing is done inside the library) */
int mfb_sm_count(void);

/* ------------------------------------------------------------------------------------
 * Screen geometry of one 1-D projection, 8 floats per projection (device array [K][8]):
 *   [0] c0        first bin centre            (diagnostics/diagnostics.py:111 `coords`)
 *   [1] spacing   centre spacing, accurate: (c[B-1]-c[0])/(B-1) evaluated in float64
 *   [2] sigma     absolute kernel width       (diagnostics/diagnostics.py:113-114)
 *   [3] delta     the reference's fp32 c[1]-c[0] used in the normalisation
 *                 (diagnostics/histogram.py:40); differs from [1] by up to 1e-5 relative
 *   [4..7]        reserved (0)
 * 2-D screens use two consecutive records (x axis, then y axis).
 * ------------------------------------------------------------------------------------ */
#define MFB_GEOM_STRIDE 8

/* ---- fused linear projection + Gaussian KDE, 1-D screens ------------------------------
 * Replaces, for all K (transform, Histogram1D) pairs at once:
 *   simulate/simulate.py:29-33        for transform ...: u = transform(x.clone())
 *   simulate/transform.py:67-68       u = x @ M.T            (only the measured row)
 *   diagnostics/diagnostics.py:116-127  project -> kde_histogram_1d
 *   diagnostics/histogram.py:37-39    K_nb = exp(-0.5((u-c_b)/sigma)^2), sum over n
 * x[n][d]; proj[K][d] (row `axis` of M_k, or M_k^T d_hat); out partial sums are reduced
 * deterministically into sums[K][B] = S_kb = sum_n K_nb (unnormalised; this is what ranks
 * all-reduce).  max_sigma_over_delta = largest sigma/delta of the K screens (host value; sets
 * the deposit window: bins further than ~8.9 sigma from a particle are skipped).         */
int64_t mfb_kde1d_workspace_bytes(int64_t n, int d, int k, int b);
int mfb_project_kde1d_fwd(const float* x, int64_t n, int d, const float* proj, const float* geom,
                          int k, int b, float max_sigma_over_delta, float* sums, void* workspace,
                          int64_t workspace_bytes, void* stream);
/* diagnostics/histogram.py:39-43: p = (S/N) / (sum_b (S_b/N) * delta + 1e-10)            */
int mfb_kde1d_normalize(const float* sums, double n_total, const float* geom, int k, int b,
                        float* profiles, void* stream);
/* backward of the normalisation: gsums = dL/dS given gprof = dL/dp    */
int mfb_kde1d_normalize_bwd(const float* sums, double n_total, const float* geom, int k, int b,
                            const float* gprof, float* gsums, void* stream);
/* Fused forward tail for the training loss (core.py:89-117 with loss.py:15-17): deposit, merge,
 * normalise and -- when meas[K][B] is given -- kl[k] = sum_b (xlogy(t,t) - t log(p + pad)) / B in
 * two launches.  sums[K][B] (unnormalised, kept for the backward), profiles[K][B], kl[K] (NULL iff
 * meas is NULL).  Workspace as for mfb_project_kde1d_fwd.                                  */
int mfb_project_kde1d_loss_fwd(const float* x, int64_t n, int d, const float* proj, const float* geom,
                               int k, int b, float max_sigma_over_delta, double n_total,
                               const float* meas, float pad, float* sums, float* profiles, float* kl,
                               void* workspace, int64_t workspace_bytes, void* stream);
/* The same tail from already merged (e.g. all-reduced across ranks) sums.                   */
int mfb_kde1d_finish(const float* sums, double n_total, const float* geom, int k, int b,
                     const float* meas, float pad, float* profiles, float* kl, void* stream);
/* gsums = d/dS of <gprof, profiles> + <gkl, kl>; gprof or gkl may be NULL (not both).       */
int mfb_kde1d_finish_bwd(const float* sums, double n_total, const float* geom, int k, int b,
                         const float* meas, float pad, const float* gprof, const float* gkl,
                         float* gsums, void* stream);
/* The same tail at N > 1 GPUs with the cross-rank sum INSIDE the kernel (replaces the NCCL all-reduce of
 * S[K][B] + the entropy sums that sat between deposit and tail): every rank owns a peer-mapped block of
 * mfb_kde1d_p2p_block_floats(k, b, tail_doubles) floats, zero-initialised, whose addresses AS MAPPED IN THIS PROCESS
 * are passed in peer_blocks_host[world] (host array).  The kernel copies local_sums[K][B] (+ local_tail doubles)
 * into its own block, exchanges an epoch over NVLink, adds all ranks' rows in rank order (bit-identical on every
 * rank) and finishes like mfb_kde1d_finish; sums[K][B] receives the reduced sums, tail_out the reduced doubles.
 * state: 3 zero-initialised uint32 in local device memory (epoch, two arrival counters), advanced by the kernel
 * itself so that a captured graph replays correctly.  Every rank must launch it once per step; k*b even,
 * k <= 1024.                                                                                                 */
int64_t mfb_kde1d_p2p_block_floats(int k, int b, int tail_doubles);
int mfb_kde1d_finish_p2p(const uint64_t* peer_blocks_host, int rank, int world, uint32_t* state,
                         const float* local_sums, const double* local_tail, int tail_doubles, double n_total,
                         const float* geom, int k, int b, const float* meas, float pad, float* sums,
                         float* profiles, float* kl, double* tail_out, void* stream);
/* deposit + (merge of the deposit's partials, cross-rank sum, normalise, KL) in two launches: the sharded counterpart
 * of mfb_project_kde1d_loss_fwd; workspace as for mfb_project_kde1d_fwd.                                       */
int mfb_project_kde1d_loss_fwd_p2p(const float* x, int64_t n, int d, const float* proj, const float* geom, int k,
                                   int b, float max_sigma_over_delta, double n_total, const float* meas, float pad,
                                   const uint64_t* peer_blocks_host, int rank, int world, uint32_t* state,
                                   const double* local_tail, int tail_doubles, float* sums, float* profiles,
                                   float* kl, double* tail_out, void* workspace, int64_t workspace_bytes,
                                   void* stream);
/* dL/dx[n][d] (+)= sum_k proj_k * sum_b gsums[k][b] K_nb (-(u-c_b)/sigma^2); accumulate!=0
 * adds into gx instead of overwriting it.                                               */
int mfb_project_kde1d_bwd(const float* x, int64_t n, int d, const float* proj, const float* geom,
                          int k, int b, float max_sigma_over_delta, const float* gsums, float* gx,
                          int accumulate, void* stream);

/* ---- the same three kernels behind a thin multipole kick -------------------------------
 * Replaces CompositeTransform(linear.., MultipoleTransform, linear..) followed by Histogram1D
 * (simulate/transform.py:35-55,78-146; experiments/rec_2d/nonlinear/setup.py:24-44):
 *   u = proj_k . x + a_k Re(z^m) + b_k Im(z^m),  z = wa_k . x + i wb_k . x,  m = order - 1.
 * mp[K][2 d + 4] = [wa (d) | wb (d) | a | b | order | 0]; everything else as in the functions
 * without _mp.  The host side folds the matrices, the kick strength / (order-1)! and the
 * reference's U[:,3] = X[:,1] + .. quirk into proj, a and b (simulate.multipole_terms).     */
int mfb_project_kde1d_mp_fwd(const float* x, int64_t n, int d, const float* proj, const float* mp,
                             const float* geom, int k, int b, float max_sigma_over_delta, float* sums,
                             void* workspace, int64_t workspace_bytes, void* stream);
int mfb_project_kde1d_mp_bwd(const float* x, int64_t n, int d, const float* proj, const float* mp,
                             const float* geom, int k, int b, float max_sigma_over_delta,
                             const float* gsums, float* gx, int accumulate, void* stream);
int mfb_project_hist1d_mp(const float* x, int64_t n, int d, const float* proj, const float* mp,
                          const float* edges, int k, int b, int64_t* counts, void* stream);

/* ---- fused projection + exact histogram, 1-D screens -----------------------------------
 * Replaces diagnostics/diagnostics.py:128-131 (torch.histogram(x_proj, edges)): bin i holds
 * edges[i] <= u < edges[i+1], last bin closed, everything else dropped.  edges[K][B+1];
 * counts[K][B] int64, ADDED to (zero them first; integer adds => bit-reproducible).     */
int mfb_project_hist1d(const float* x, int64_t n, int d, const float* proj, const float* edges,
                       int k, int b, int64_t* counts, void* stream);

/* ---- 2-D screens ------------------------------------------------------------------------
 * diagnostics/diagnostics.py:179-191 -> histogram.py:89-101, 47-74: P = Kx^T Ky.
 * proj[K][2][d], geom[K][2][8]; sums[K][bx][by] unnormalised.  Deposits are accumulated in
 * fixed point (44 fractional bits; 64-bit integers globally) so the result is independent
 * of the atomics' order, the CTA decomposition and the rank count.
 * workspace = int64 accumulators [2][K][bx][by] (plane 0 in units of 2^-22, plane 1 in units
 * of 2^-44; this is what ranks all-reduce exactly).                                      */
#define MFB_KDE2D_FRAC_BITS 44
int64_t mfb_kde2d_workspace_bytes(int64_t n, int d, int k, int bx, int by);
/* Screens of up to 128 x 96 bins and batches of at least 4096 particles run as tcgen05 GEMMs over the
 * particle axis (the reference's own formulation, diagnostics/histogram.py:47-74: P = Kx^T Ky), dense
 * kernel rows in split bf16; otherwise (or with MFB_FLAG_NO_TENSOR_CORES in `flags`) windowed fixed-point
 * deposits.                                                                                          */
int mfb_project_kde2d_fwd(const float* x, int64_t n, int d, const float* proj, const float* geom,
                          int k, int bx, int by, float max_sigma_over_delta, float* sums,
                          void* workspace, int64_t workspace_bytes, int flags, void* stream);
/* P <- P / (sum(P) dx dy + 1e-10)  (histogram.py:70-73) and its backward                */
int mfb_kde2d_normalize(const float* sums, const float* geom, int k, int bx, int by,
                        float* profiles, void* stream);
int mfb_kde2d_normalize_bwd(const float* sums, const float* geom, int k, int bx, int by,
                            const float* gprof, float* gsums, void* stream);
int mfb_project_kde2d_bwd(const float* x, int64_t n, int d, const float* proj, const float* geom,
                          int k, int bx, int by, float max_sigma_over_delta, const float* gsums,
                          float* gx, int accumulate, void* stream);
/* diagnostics/diagnostics.py:192-201 (np.histogramdd): edges_x[K][bx+1], edges_y[K][by+1];
 * counts[K][bx][by] int64, added to                                                      */
int mfb_project_hist2d(const float* x, int64_t n, int d, const float* proj, const float* edges_x,
                       const float* edges_y, int k, int bx, int by, int64_t* counts, void* stream);

/* ---- neural spline flow (zuko 1.3.1 NSF as built by generate/build.py:36-46) ------------
 * One autoregressive layer per call: y = RQS(MaskedMLP(v))(v), logq_out = logq_in - ladj.
 * Replaces generate/flows/zuko.py:24-29 (rsample_and_log_prob / transform.inv) layer by layer.
 * params: packed fp32 block of one layer (layout: mfb_nsf_layer_param_floats / nsf.cu).
 * first_layer != 0: logq_in is ignored and replaced by log N(v; 0, I).
 * logq_in/logq_out may be NULL (sample only).  order_host: HOST array of d ints, the
 * layer's autoregressive order (feature with order 0 has a bias-only spline); may be NULL.*/
int64_t mfb_nsf_layer_param_floats(int d, int hidden_units, int hidden_layers, int bins);
int mfb_nsf_layer_fwd(const float* v, int64_t n, int d, int hidden_units, int hidden_layers,
                      int bins, const float* params, const int32_t* order_host, const float* logq_in,
                      int first_layer, float* y, float* logq_out, void* stream);

/* Tensor-core version of mfb_nsf_layer_fwd (same reference lines): the conditioner's masked GEMMs
 * run as tcgen05.mma tiles over fp16 (hi, lo) splits of the fp32 operands with fp32 accumulators
 * in TMEM, the spline is the epilogue.  Compiled for hidden_units = 64, hidden_layers = 3,
 * bins = 20, d = 2..6 (mfb_nsf_tc_supported); other shapes use mfb_nsf_layer_fwd.
 * mfb_nsf_tc_prepare turns the packed fp32 parameters of n_layers layers (layer l at
 * params + l * layer_stride_floats; they MUST be pre-masked: all-zero blocks are skipped) into
 * n_layers operand images of mfb_nsf_tc_image_bytes bytes each; orders_host = HOST array
 * [n_layers][d].  mfb_nsf_tc_layer_fwd then runs one layer from its image.                  */
int mfb_nsf_tc_supported(int d, int hidden_units, int hidden_layers, int bins);
int64_t mfb_nsf_tc_image_bytes(int d, int hidden_layers);
int64_t mfb_nsf_tc_prepare_workspace_bytes(int n_layers);
int mfb_nsf_tc_prepare(const float* params, int64_t layer_stride_floats, int n_layers, int d,
                       int hidden_units, int hidden_layers, int bins, const int32_t* orders_host,
                       void* images, void* workspace, int64_t workspace_bytes, void* stream);
int mfb_nsf_tc_layer_fwd(const float* v, int64_t n, int d, int hidden_units, int hidden_layers,
                         int bins, const void* image, const int32_t* order_host,
                         const float* logq_in, int first_layer, float* y, float* logq_out,
                         void* stream);

/* Density direction of one layer: v = A^-1(y) (d conditioner sweeps per particle) and
 * ladj_out = ladj_in + log|det dA/dv|(v).  Replaces generate/flows/zuko.py:21-22,31-32,43-50
 * (log_prob / inverse / inverse_steps).  Call the layers in REVERSE order; with last_layer != 0
 * (the flow's first layer) ladj_out receives log q(x) = log N(v;0,I) - sum ladj instead.
 * ladj_in / ladj_out may be NULL.  order_host is required (HOST array of d ints).          */
int mfb_nsf_layer_inv(const float* y, int64_t n, int d, int hidden_units, int hidden_layers,
                      int bins, const float* params, const int32_t* order_host, const float* ladj_in,
                      int last_layer, float* v, float* ladj_out, void* stream);
/* The same on the tensor cores, from one layer's operand image of mfb_nsf_tc_prepare (shapes of
 * mfb_nsf_tc_supported): the feature first in the order inverts its bias-only spline, every further
 * feature takes one pass of the conditioner (first layer, two hidden GEMMs, its own output tile) and
 * the inverse spline in registers -- S x 4 GEMM round trips per 128-particle tile instead of D sweeps. */
int mfb_nsf_tc_layer_inv(const float* y, int64_t n, int d, int hidden_units, int hidden_layers, int bins,
                         const void* image, const int32_t* order_host, const float* ladj_in, int last_layer,
                         float* v, float* ladj_out, void* stream);

/* Backward of one layer (replaces torch autograd through the zuko graph).  Activations are
 * recomputed from the layer input v.  gy = dL/dy [n][d], glogq = dL/dlogq_out [n] (may be NULL);
 * outputs gv = dL/dv [n][d] and gparams = dL/dparams in the packed forward layout (added to
 * when accumulate != 0).  params_om: the same masked weights in out-major layout
 * W1 [64][d] | Wl [64 out][64 in] x (L-1) | Wout [d*64 (59->64 padded rows)][64 in], no biases
 * (mfb_nsf_layer_param_om_floats floats).  dL/dlogq_in = glogq (pass-through).           */
int64_t mfb_nsf_layer_param_om_floats(int d, int hidden_units, int hidden_layers);
/* The backward of a layer runs as three tcgen05 kernels for hidden_layers = 3, bins = 20 (nsf_tc.cu kBwd,
 * nsf_tc_bwd.cu) and on CUDA-core kernels otherwise, or when `flags` has MFB_FLAG_NO_TENSOR_CORES.   */
int64_t mfb_nsf_layer_bwd_workspace_bytes(int64_t n, int d, int hidden_layers);
int mfb_nsf_layer_bwd(const float* v, const float* gy, const float* glogq, int64_t n, int d,
                      int hidden_units, int hidden_layers, int bins, const float* params,
                      const float* params_om, const int32_t* order_host, int first_layer, float* gv,
                      float* gparams, int accumulate, void* workspace, int64_t workspace_bytes,
                      int flags, void* stream);

/* The same with the layer's tcgen05 operand image that mfb_nsf_tc_prepare built for the forward pass of
 * this step (tc_image, mfb_nsf_tc_image_bytes bytes, 16-byte aligned; NULL = build it here): a
 * training step then builds each image once instead of twice.                              */
int mfb_nsf_layer_bwd_img(const float* v, const float* gy, const float* glogq, int64_t n, int d,
                          int hidden_units, int hidden_layers, int bins, const float* params,
                          const float* params_om, const int32_t* order_host, int first_layer,
                          const void* tc_image, float* gv, float* gparams, int accumulate,
                          void* workspace, int64_t workspace_bytes, int flags, void* stream);

/* ---- parameter layouts --------------------------------------------------------------------
 * zuko keeps every MaskedLinear as weight [out][in] + a 0/1 mask applied on each call (zuko/nn.py, reached from
 * generate/build.py:36-46); the kernels take pre-masked, transposed blocks.  One launch converts the whole flow:
 * w_in [T][64][d], b_in [T][64], w_hid [T][L-1][64][64], b_hid [T][L-1][64], w_out [T][d*P][64], b_out [T][d*P]
 * (P = 3*bins-1) and masks of the weights' shapes -> packed [T][mfb_nsf_layer_param_floats] and, if not NULL,
 * packed_om [T][mfb_nsf_layer_param_om_floats].  mfb_nsf_unpack_grads is its backward: dL/dpacked -> dL/d(each
 * tensor), masked entries exactly 0.                                                              */
int mfb_nsf_pack_params(const float* w_in, const float* b_in, const float* w_hid, const float* b_hid,
                        const float* w_out, const float* b_out, const float* m_in, const float* m_hid,
                        const float* m_out, int transforms, int d, int hidden_units, int hidden_layers, int bins,
                        float* packed, float* packed_om, void* stream);
int mfb_nsf_unpack_grads(const float* gpacked, const float* m_in, const float* m_hid, const float* m_out,
                         int transforms, int d, int hidden_units, int hidden_layers, int bins, float* g_w_in,
                         float* g_b_in, float* g_w_hid, float* g_b_hid, float* g_w_out, float* g_b_out,
                         void* stream);

/* ---- Monte-Carlo entropy pieces (entropy.py:58-62, prior.py:25-26) ----------------------
 * out[0] = sum logq, out[1] = sum |x|^2, out[2+i] = sum x_i, out[2+d+i*d+j] = sum x_i x_j
 * (double precision, deterministic two-stage reduction; the x_i / x_i x_j block only when
 * with_cov != 0; logq may be NULL).  entropy.py:35-38 (torch.cov) uses the second block. */
int64_t mfb_moments_workspace_bytes(int64_t n, int d);
int mfb_moments(const float* x, const float* logq, int64_t n, int d, int with_cov, double* out,
                void* workspace, int64_t workspace_bytes, void* stream);
/* The scalar ends of MENTFlow.loss as one launch each (they sit on the critical path of every step):
 * h[0] = (float)(a * sums[0] + b * sums[1] - c) in double -- entropy.py:58-62 with a = 1/N,
 * b = 1/(2 s^2 N), c = log prior normalisation (prior.py:25-26);
 * out[0] = h[0] + mu * mean(d[0..k)), out[1] = mean(d[0..k)) -- core.py:111-113 (h may be NULL: 0).   */
int mfb_mc_entropy(const double* sums, double a, double b, double c, float* h, void* stream);
int mfb_loss_tail(const float* d, int k, const float* h, float mu, float* out, void* stream);
/* Multi-GPU plumbing of the entropy sums (ONE packed all-reduce per forward step): n doubles
 * <-> 2n floats (hi[0..n), lo[0..n)), hi + lo = value to 2^-48, so that they can ride at the tail of the
 * float32 all-reduce of the profile sums; the two halves are summed over ranks separately.            */
int mfb_f64_split(const double* in, int n, float* out_hi_lo, void* stream);
int mfb_f64_join(const float* in_hi_lo, int n, double* out, void* stream);

/* ---- classical MENT (ment.py, sample.py) --------------------------------------------------
 * rho(x) = exp(log prior(x)) * prod_k clamp(h_k(proj_k . x), 0, 1e10); h_k = linear interpolation
 * of tables[k][0..b) on the bin centres coords[k][0..b), 0 outside the first/last centre,
 * evaluated in double like scipy's RegularGridInterpolator (ment.py:45-52, 227-249).
 * prior: log p(x) = prior_neg_half_inv_s2 * |x|^2 + prior_log_norm (Gaussian prior.py:25-26;
 * a flat prior passes 0 and its log constant).                                            */
int mfb_ment_prob(const float* x, int64_t g, int d, const float* proj, const float* coords,
                  const float* tables, int k, int b, float prior_neg_half_inv_s2,
                  float prior_log_norm, float* out, void* stream);
/* same on the cell centres of a regular grid generated from the index ('ij' order, last axis
 * fastest) instead of res^D x D materialised points (sample.py:99-104 GridSampler).  shape /
 * first_centre / step are HOST arrays of d entries.                                       */
int mfb_ment_prob_grid(int d, const int32_t* shape_host, const float* first_centre_host,
                       const float* step_host, const float* proj, const float* coords,
                       const float* tables, int k, int b, float prior_neg_half_inv_s2,
                       float prior_log_norm, float* out, void* stream);
/* integrate mode for a 1-D screen (ment.py:267-317): pred[i] = sum over the integration grid of
 * rho(Minv [meas_coords[i] on meas_axis ; grid point on the other axes]).                 */
int mfb_ment_integrate(int d, const float* meas_coords, int nb_meas, int meas_axis, int n_int_axes,
                       const int32_t* int_shape_host, const float* int_first_host,
                       const float* int_step_host, const float* minv, const float* proj,
                       const float* coords, const float* tables, int k, int b,
                       float prior_neg_half_inv_s2, float prior_log_norm, float* pred, void* stream);
/* The same three entry points with two-dimensional screens as well (ment.py:36-49: LagrangeFunction over an N-D
 * RegularGridInterpolator; experiments/config/rec_nd_2d_ment.yaml): k2 tables tables2[k2][bx][by] on the bin-centre
 * grids cx2[k2][bx] x cy2[k2][by], bilinear, zero outside the box of the centres, evaluated in double like scipy;
 * proj2[k2][2][d] are the two measured rows of each transfer matrix.  rho = prior * prod(1-D tables) * prod(2-D
 * tables); either family may be empty (k = 0 / k2 = 0).  _integrate_nd: meas_axis2 >= 0 selects a 2-D screen
 * (pred[nb_meas][nb_meas2], integration over the other d-2 axes), meas_axis2 = -1 a 1-D one.              */
int mfb_ment_prob_nd(const float* x, int64_t g, int d, const float* proj, const float* coords, const float* tables,
                     int k, int b, const float* proj2, const float* cx2, const float* cy2, const float* tables2,
                     int k2, int bx, int by, float prior_neg_half_inv_s2, float prior_log_norm, float* out,
                     void* stream);
int mfb_ment_prob_grid_nd(int d, const int32_t* shape_host, const float* first_centre_host, const float* step_host,
                          const float* proj, const float* coords, const float* tables, int k, int b,
                          const float* proj2, const float* cx2, const float* cy2, const float* tables2, int k2,
                          int bx, int by, float prior_neg_half_inv_s2, float prior_log_norm, float* out,
                          void* stream);
int mfb_ment_integrate_nd(int d, const float* meas_coords, int nb_meas, int meas_axis, const float* meas_coords2,
                          int nb_meas2, int meas_axis2, int n_int_axes, const int32_t* int_shape_host,
                          const float* int_first_host, const float* int_step_host, const float* minv,
                          const float* proj, const float* coords, const float* tables, int k, int b,
                          const float* proj2, const float* cx2, const float* cy2, const float* tables2, int k2,
                          int bx, int by, float prior_neg_half_inv_s2, float prior_log_norm, float* pred,
                          void* stream);
/* sample.py:27-57: cdf of (rho + pad) over the cells in double, then `size` draws: cell by
 * inverse-CDF search with a Philox4x32-10 stream (seed, offset), uniform position inside the
 * cell (+ 0.5*U(-delta,delta) when jitter != 0).  workspace[0] (device double) = total mass. */
int64_t mfb_cdf_workspace_bytes(int64_t g);
int mfb_cdf_build(const float* rho, int64_t g, double pad, double* cdf, void* workspace,
                  int64_t workspace_bytes, void* stream);
int mfb_cdf_sample(const double* cdf, int64_t g, const void* workspace, int d,
                   const int32_t* shape_host, const float* first_edge_host, const float* cell_host,
                   int jitter, uint64_t seed, uint64_t offset, int64_t size, float* out, void* stream);
/* ment.py:360-367: pred[pred<thresh]=0; where meas!=0 and pred!=0: h *= 1 + lr*(meas/pred-1) */
int mfb_gs_update(float* table, const float* meas, const float* pred, int n, float lr, float thresh,
                  void* stream);

/* ---- base noise of the flow ---------------------------------------------------------------
 * generate/flows/zuko.py:15-16,24-26 (`flow.base.rsample`: zuko DiagNormal -> torch.randn).
 * out[numel] ~ N(0,1) from Philox4x32-10 with torch's CUDA element assignment: for the (seed,
 * offset) of torch's CUDA generator the result is bitwise the tensor torch.randn(numel,
 * device="cuda") returns, and mfb_randn_offset_increment(numel) is what torch adds to the offset.
 * _state form: state[0] = seed, state[1] = offset live in DEVICE memory; with advance != 0 the
 * offset is moved on after the draw (stream-ordered), so a captured graph draws fresh noise at
 * every replay.                                                                              */
int64_t mfb_randn_offset_increment(int64_t numel);
int mfb_randn_philox(float* out, int64_t numel, uint64_t seed, uint64_t offset, void* stream);
int mfb_randn_philox_state(float* out, int64_t numel, uint64_t* state, int advance, void* stream);

/* ---- diagnostics --------------------------------------------------------------------------
 * Self-test of the tcgen05/TMEM building blocks: d[128][n] = a[128][64] * b[n][64]^T through
 * fp16 (hi,lo)-split kind::f16 MMAs with TMEM accumulators (n = 64, 128 or 256).  *err != 0
 * if the MMA never completed.                                                              */
int mfb_selftest_umma(const float* a, const float* b, int n, float* d, int* err, void* stream);
/* the same GEMM with the A operand in tensor memory (tcgen05.st + tcgen05.mma with a TMEM A address), n <= 128 */
int mfb_selftest_umma_ts(const float* a, const float* b, int n, float* d, int* err, void* stream);
/* One K step (16) with 32-byte-row SWIZZLE_32B operand tiles: d[128][n] = a[128][16] * b[n][16]^T
 * (values rounded to fp16).  mode 0: both operands SWIZZLE_32B; mode 1: a sits in K step `kstep`
 * of a 64-wide SWIZZLE_128B tile (the layouts the flow kernel mixes).  n = 64 or 128.        */
int mfb_selftest_umma_sw32(const float* a, const float* b, int n, int mode, int kstep, float* d, int* err,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MENTFLOW_B200_H */
