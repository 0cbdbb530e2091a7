/*
 * gptb200 -- C ABI of the B200-native exact-GP transport posterior engine (libgptb200.so).
 *
 * This is the drop-in boundary for ONE hot path of vyasakash231/gaussian_process_transportation: the exact-GP
 * posterior behind `GaussianProcess.fit / predict(return_std) / derivative` and the
 * `PolicyTransportation` fit/apply flow.  Every entry point names the reference interface it replaces
 * (paths relative to the reference root; `sklearn:` = scikit-learn's gaussian_process package, the un-vendored
 * dependency that does the arithmetic there).
 *
 * Conventions
 *   - plain pointers and sizes only; all arrays are C-contiguous IEEE float64 unless the name ends in `_dev`.
 *   - host pointers may be pageable or pinned; `_dev` pointers are CUDA device pointers on the handle's device.
 *   - every function returns 0 on success, >0 for a numerical failure (leading minor of that order is not positive
 *     definite, LAPACK `info` convention, cf. sklearn:_gpr.py:351-361 / 590-593), <0 for CUDA / argument errors.
 *     `gptb_last_error()` returns a human-readable message for the last non-zero status on that handle.
 *   - a handle owns one device, one stream set and all device state; it is NOT thread-safe, distinct handles are
 *     independent.  There is no CPU fallback anywhere behind this ABI.
 *   - kernel family: k(x,y) = c * exp(-0.5 * sum_a ((x_a - y_a)/ell_a)^2)  [+ s2 on the training diagonal],
 *     i.e. sklearn's ConstantKernel * RBF + WhiteKernel (sklearn:kernels.py:1273-1296,1403-1419,1558-1587).
 *     `ell` always has d entries (isotropic kernels repeat the scalar).
 */
#ifndef GPTB200_H
#define GPTB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gptb_handle gptb_handle;

/* query flags (OR-able) */
#define GPTB_MEAN      0x001u  /* posterior mean  k* alpha                  gaussian_process.py:46-55 / sklearn:_gpr.py:446-447 */
#define GPTB_STD       0x002u  /* sqrt(max(c+s2-|L^-1 k*|^2,0)) - sqrt(s2)   gaussian_process.py:49    / sklearn:_gpr.py:460-500 */
#define GPTB_JAC       0x004u  /* d mean / d x, layout (M,p,d)               gaussian_process.py:72-90 */
#define GPTB_JACVAR    0x008u  /* c/ell_a^2 - |L^-1 dk/dx_a|^2, (M,p,d)       gaussian_process.py:92-101 */
#define GPTB_AFFINE_IN 0x010u  /* evaluate at gamma(x) = s R (x - Sbar) + Tbar affine_trasformation.py:51-53 */
#define GPTB_TRANSPORT 0x020u  /* xhat = gamma(x) + mean                     policy_transportation.py:26-35 */
#define GPTB_VELOCITY  0x040u  /* vhat = (R + J R) v, vvar = Jvar (R v)^2    policy_transportation.py:37-59 */
#define GPTB_JPHI      0x080u  /* Jphi = R + J R, layout (M,d,d)             policy_transportation.py:61-64 */
#define GPTB_DVAR      0x100u  /* -2 (dk/dx_a)^T K^-1 kstar, layout (d,M)      gaussian_process.py:104-126 */

/* ---- lifetime ------------------------------------------------------------------------------------------------ */
int  gptb_create(int device, gptb_handle** out);
void gptb_destroy(gptb_handle* h);
const char* gptb_last_error(gptb_handle* h);
int  gptb_version(void);

/* ---- training data: replaces the X/Y copies of GaussianProcess.fit (gaussian_process.py:25-35; sklearn:_gpr.py:296-297).
 * X is (N,d), Y is (N,p).  Limits: 1 <= d <= 4, 1 <= p <= 4. */
int gptb_set_train(gptb_handle* h, const double* X, const double* Y, int64_t N, int d, int p);

/* ---- radial profile of the stationary factor: 0 = RBF (default), 1 = Matern nu=1.5, 2 = Matern nu=2.5, 3 = Matern nu=0.5
 * (exponential; example/comparisons/multi_reference_frames/multiple_source_same_target_with_gpt.py:46)
 * (sklearn:kernels.py Matern.__call__; the dynamics GPs of the reference's demos use C*Matern(nu=2.5)+White,
 * example/2D/surface_generalization.py:49).  Applies to every later factorize / lml / query call.
 * Note: `derivative` keeps the reference's closed form dk = k (X - x)/ell^2 for every profile (gaussian_process.py:82-87). */
int gptb_set_kernel_kind(gptb_handle* h, int kind);

/* ---- fit at fixed hyper-parameters: Gram build, L = chol(c R + (s2+jitter) I), alpha = L^-T L^-1 Y.
 * Replaces sklearn:_gpr.py:347-367.  On return *lml (may be NULL) holds the log marginal likelihood
 * (sklearn:_gpr.py:613-617).  Non-PD => returns the failing order (>0). */
int gptb_factorize(gptb_handle* h, double c, const double* ell, double s2, double jitter, double* lml);

/* ---- append ONE training point (x (d), y (p)) to a fitted model at its current hyper-parameters: the re-fit of the greedy
 * active-learning loop (gaussian_process_al.py:41-55 adds the pool point of largest predictive std and calls fit again) when the
 * hyper-parameters are fixed.  Rank-1 update of L, L^-1, alpha in O(N^2) instead of the O(N^3) re-factorisation; every 128th point
 * (a new row tile) re-factorises.  *lml (may be NULL) receives the log marginal likelihood of the extended model.  Returns the
 * order of the non-positive leading minor (> 0) when the extended kernel matrix is not positive definite. */
int gptb_append_point(gptb_handle* h, const double* x, const double* y, double* lml);

/* ---- one log-marginal-likelihood evaluation with gradient w.r.t. log-hyper-parameters, the objective that
 * scipy's L-BFGS-B drives in sklearn:_gpr.py:302-309,541-656.  grad has 2+d entries:
 * [d/dlog c, d/dlog ell_0 .. ell_{d-1}, d/dlog s2]; an isotropic kernel's gradient is the sum of the ell entries.
 * Non-PD => returns >0 and the caller substitutes (-inf, 0) as sklearn does (sklearn:_gpr.py:590-593).
 * Leaves L/alpha for these hyper-parameters in the handle (a later gptb_factorize with the same values is free to
 * recompute). */
int gptb_lml(gptb_handle* h, double c, const double* ell, double s2, double jitter, int want_grad,
             double* lml, double* grad);

/* ---- how the variance products |L^-1 k*|^2 are evaluated: mode 0 (default) = FP64 DMMA tile engine; mode 1 = INT8-sliced
 * ("Ozaki") evaluation on tcgen05 tensor cores with `slices` in {5,6,7} 7-bit digit planes per operand (exact int32
 * accumulation of the digit products, FP64 recombination).  6 planes reproduce the FP64 std to ~1e-9 of sqrt(c+s2)
 * (tolerance 1e-7), 7 to ~1e-11; the posterior mean and Jacobian are unaffected.  mode 2 = the same evaluation with
 * `slices` in {4,5,6} 8-bit digit planes (the full int8 range): 5 planes (15 plane products instead of 21) give ~5e-9,
 * 6 give ~2e-11; limited to N <= 26112 by the worst-case exactness bound of the int32 accumulators (S * N * 2^14 < 2^31); beyond it the slicer's data-dependent
 * bound (128 * largest row sum of |digit| of the inverse factor < 2^31) decides, and the call fails if that does not hold either.
 * mode 3 = mode 2 with slices = 5 plus the first DROPPED diagonal of plane products (a + b = S: 19 products instead of 15).  With
 * 8-bit planes the error of the 15-product scheme is that dropped diagonal, not the 40-bit operands (tools/plane_error_study.py: std
 * error 9.0e-9 -> 5.5e-10 at N = 4096), so this buys the accuracy of a sixth plane for 27 % instead of 40 % more tensor work and no
 * extra operand traffic.
 * Whatever the mode, a query call with at most eight right-hand-side rows in all (queries x (1 | 1+d | 1+2d) for STD | JACVAR | DVAR:
 * a control-loop query, a rollout step) is evaluated by the exact FP64 matrix-vector product -- it streams L^-1 once and is faster
 * than any tiled form at that size; FP64-mode batches too small to fill the GPU with tiles run the tile product split along k. */
int gptb_set_variance_mode(gptb_handle* h, int mode, int slices);

/* ---- run-time accuracy guard of the INT8-sliced path (on by default).  Before the first variance query of a model (and in
 * gptb_prepare_variance) 2048 probe queries -- half of them next to training points, where the std is most sensitive -- are
 * evaluated on the INT8 path and on the FP64 path of the same handle.  If max |std_int8 - std_fp64| / sqrt(c + s2) exceeds
 * `threshold` (default 2e-8, a fifth of the 1e-7 std tolerance of gaussian_process.py:46-49 parity) one digit plane is added and
 * the probe repeated (five 8-bit planes first gain the extra diagonal of mode 3, which needs no new planes); with the plane count
 * exhausted the model is served by the FP64 path.  threshold = 0 switches the guard off.
 * The report returns the requested and the effective plane count (0 = FP64 path), the probe error of the effective mode and the
 * one measured with the requested plane count. */
int gptb_set_variance_guard(gptb_handle* h, double threshold);
int gptb_variance_guard_report(gptb_handle* h, int* requested_slices, int* used_slices, int* used_extra_diagonal, double* probe_err,
                               double* first_err, double* threshold);

/* ---- spatial mode (call before gptb_set_train; off by default).  The training points are kept in Morton (Z-curve) order
 * inside the handle, every query batch of the INT8-sliced path with 8-bit planes is processed in Morton order too (radix sort
 * of the batch by key, results scattered back to the caller's order), and the generator / the slicer record per
 * (128-query tile | 64-row tile of L^-1, 64-point chunk) which digit planes hold a non-zero digit.  The product kernel skips
 * leading all-zero planes: for a compact kernel most (tile, chunk) pairs are far apart, k(x*, X) and the entries of L^-1 there
 * are tiny and their top digit planes vanish.  Exact -- the skipped products are sums of zeros; only the summation ORDER of
 * the training points changes with respect to the natural order (1e-13-level differences).  gptb_export_alpha/_Kinv return
 * the caller's order; gptb_export_L is refused (the factor belongs to the permuted system).
 * on = 2 keeps the ordering and the sorted batches but issues every plane product (bit-identical results: A/B of the skipping). */
int gptb_set_spatial(gptb_handle* h, int on);

/* ---- explicit inverse factor for the variance queries (built lazily by gptb_query when needed). */
int gptb_prepare_variance(gptb_handle* h);

/* ---- affine pre-alignment gamma(x) = s R (x - Sbar) + Tbar (affine_trasformation.py:15-57).  R is (d,d) row-major.
 * The d x d SVD stays on the host (numpy) for bit-parity; only the apply and the Jacobian algebra are fused here. */
int gptb_set_affine(gptb_handle* h, const double* R, double s, const double* Sbar, const double* Tbar);

/* ---- posterior / transport query on M points.  Any output pointer may be NULL when its flag is absent.
 *   x (M,d) in; vel (M,d) in (GPTB_VELOCITY);
 *   mean (M,p); std (M,p) [columns identical, sklearn:_gpr.py:494]; jac (M,p,d); jacvar (M,p,d) [identical over p];
 *   xhat (M,d); vhat (M,d); vvar (M,p); jphi (M,d,d) [needs d==p]; dvar (d,M).
 * Host-pointer version: copies in/out inside the call (this is the end-to-end path the benchmark's `e2e` times). */
int gptb_query(gptb_handle* h, const double* x, int64_t M, uint32_t flags, const double* vel,
               double* mean, double* std, double* jac, double* jacvar,
               double* xhat, double* vhat, double* vvar, double* jphi, double* dvar);

/* Device-pointer version: inputs/outputs already resident in HBM (benchmark `value`; sharded multi-GPU queries). */
int gptb_query_dev(gptb_handle* h, const double* x_dev, int64_t M, uint32_t flags, const double* vel_dev,
                   double* mean_dev, double* std_dev, double* jac_dev, double* jacvar_dev,
                   double* xhat_dev, double* vhat_dev, double* vvar_dev, double* jphi_dev, double* dvar_dev);

/* ---- minimum-variance stabilised rollouts (plot_utils.py:298-310, and :283-296 for the one-step grid form): K start points are
 * advanced `steps` times by  pos <- pos + mean(pos) - gain * std(pos) * g / |g|,  g = derivative_of_variance(pos)
 * (gaussian_process.py:46-49, 104-126), all K points per step in ONE batched query; the loop is device-resident (one step is
 * captured as a CUDA graph and replayed -- for small K a step is ~20 launch-bound kernels), only the finished trajectories travel.
 * start (K,d) in; traj (steps, K, d) out: the positions AFTER each step.  Needs d == p in {2, 3} (a dynamics model).  The
 * reference's loop is the K = 1, gain = 1 case. */
int gptb_rollout_min_variance(gptb_handle* h, const double* start, int64_t K, int steps, double gain, double* traj);

/* ---- dense transport grid (BASELINE config 5; the query shape of plot_utils.py:10-15, 353-358: a regular lattice over the workspace).
 * The lattice points are GENERATED ON THE DEVICE: point g of the row-major lattice `dims` (last dimension fastest) is
 * x_a = origin[a] + step[a] * i_a; this call evaluates points [first, first + count) -- a rank's shard of the lattice -- with
 * flags from {GPTB_MEAN, GPTB_STD, GPTB_JAC, GPTB_AFFINE_IN} and reduces the outputs on the device, so a 2^29-point grid needs no host
 * buffers.  Packed columns per point: [mean (p) | std (1) | jac (p*d)] (those requested, ncol in total).
 *   stats (ncol, 4) out: per column {sum, sum of squares, min, max} over the evaluated points (fixed-order reductions: deterministic;
 *          shard sums add up to the whole-lattice sums -- the size-independent parity property at full size);
 *   sample_stride > 0: lattice points j * sample_stride that fall into [first, first + count) are also returned, packed rows in
 *          sample_out (ceil-counted: (ceil((first+count)/stride) - ceil(first/stride), ncol)), for checks against the CPU oracle. */
int gptb_query_grid(gptb_handle* h, const double* origin, const double* step, const int64_t* dims, int64_t first, int64_t count,
                    uint32_t flags, double* stats, int64_t sample_stride, double* sample_out);

/* ---- joint posterior on M points: mean (M,p) and covariance (M,M) = k(x,x) + s2 I - K* K^-1 K*^T, the quantity behind
 * GaussianProcess.predict(return_cov=True) and samples() (gaussian_process.py:50-60; sklearn:_gpr.py:470-475, 502-539).
 * The covariance is output-independent (the host replicates it over p as sklearn does). */
int gptb_query_cov(gptb_handle* h, const double* x, int64_t M, double* mean, double* cov);

/* ---- orientation transport (policy_transportation.py:61-77): Jphi = R + Jpsi(pos) R evaluated at the UN-rotated positions
 * (the reference's behaviour), q_hat = quat(Jphi) (x) q with the Bar-Itzhack quaternion of the non-orthogonal 3x3 Jphi.
 * pos (M,3), ori (M,4) as (w,x,y,z) in; ori_out (M,4); jphi (M,3,3) optional (may be NULL).  Needs d == p == 3. */
int gptb_transport_orientation(gptb_handle* h, const double* pos, const double* ori, int64_t M, double* ori_out, double* jphi);
/* the composition of the older diffeomorphic variant (gaussian_process_transportation_diffeomorphic.py:94-101): the Jacobian of the
 * residual map at the ROTATED positions, q_hat = quat(I + Jpsi(gamma(pos))) (x) (quat(R) (x) q). */
int gptb_transport_orientation_diffeo(gptb_handle* h, const double* pos, const double* ori, int64_t M, double* ori_out);

/* ---- stiffness transport: K_hat = Jphi K Jphi^T per point, Jphi(x) = (I + Jpsi(gamma(x))) R the Jacobian of the whole map
 * at x (the linearisation the velocity transport uses, policy_transportation.py:37-46).  NOT in the reference code: its
 * README (README.md:6) and the paper announce stiffness transport, the repository has no implementation; for an orthogonal
 * Jphi every candidate form reduces to this congruence.  pos (M,d), stiff (M,d,d) in;
 * stiff_out (M,d,d); jphi (M,d,d) optional (may be NULL).  Needs d == p. */
int gptb_transport_stiffness(gptb_handle* h, const double* pos, const double* stiff, int64_t M, double* stiff_out, double* jphi);

/* ---- read-back of fitted state (GaussianProcess attributes `gp.L_`, `gp.alpha_`, `K_inv`; gaussian_process.py:42-43).
 * L is (N,N) lower (upper part zero), alpha is (N,p), Kinv is (N,N) symmetric. */
int gptb_export_L(gptb_handle* h, double* L);
int gptb_export_alpha(gptb_handle* h, double* alpha);
int gptb_export_Kinv(gptb_handle* h, double* Kinv);

/* ---- model-state exchange for query sharding: after a fit on rank 0 the immutable state is
 * broadcast (NCCL, by the host layer) into identically-shaped buffers on every rank.
 *   gptb_state_alloc : non-fitting ranks allocate state for (N,d,p) without training data.
 *   gptb_state_buffer: device pointer + byte size of state buffer `which` (0: packed header/params, 1: X, 2: alpha,
 *                      3: inverse factor) so the host can wrap it (cuda array interface) and broadcast in place.
 *   gptb_state_commit: called on every rank after the broadcast to (re)derive cached quantities. */
int gptb_state_alloc(gptb_handle* h, int64_t N, int d, int p, int with_variance);
int gptb_state_buffer(gptb_handle* h, int which, void** dev_ptr, int64_t* bytes);
int gptb_state_commit(gptb_handle* h);

/* ---- instrumentation */
/* number of kernels this library has launched on the handle since creation (benchmark `gpu_launches`). */
int64_t gptb_launch_count(gptb_handle* h);
/* CUDA-event time (ms) of the dominant kernel class accumulated since the last reset: which = 0 triangular-multiply
 * (variance; FP64 DMMA or INT8-sliced, whichever mode is active), 1 mean/Jacobian generator, 2 factorisation trailing
 * update, 3 digit split of the right-hand sides (INT8-sliced mode).  Returns launches through *n. */
int gptb_kernel_time(gptb_handle* h, int which, double* ms, int64_t* n);
int gptb_timing_enable(gptb_handle* h, int on);
int gptb_timing_reset(gptb_handle* h);
/* stream the handle launches on (cudaStream_t as void*), so callers can record events on it. */
void* gptb_stream(gptb_handle* h);
/* schedule of the Cholesky trailing update: 0 = 128x128 tiles, one persistent CTA per SM; 1 (default) = 128x64 half tiles,
 * two CTAs per SM with a dynamic job queue (tuning / A-B measurement knob; results are bit-identical). */
int gptb_set_trailing_variant(gptb_handle* h, int variant);
/* workspace cap for query batches in bytes (default 16 GiB). */
int gptb_set_workspace_limit(gptb_handle* h, int64_t bytes);
/* INT8-sliced path only: overlap the k* generator of batch i+1 (FP64 pipe, low-priority stream) with the digit-plane
 * products of batch i (tensor pipe) through double-buffered batches.  Off by default: on this pool's B200 the product
 * kernel runs at the 1 kW power cap (SM clock ~1.72 GHz), the two kernels then share one energy budget and the overlapped
 * step is no faster than the serialised one (profiles/r01_pipeline_ab.log; re-measured in spatial mode after every kernel change of round 2:
 * profiles/r02_pipeline_ab_spatial_*.log).  Results are bit-identical either way. */
int gptb_set_query_pipeline(gptb_handle* h, int on);

/* number of (digit-plane pair, 64-byte k-chunk, 128 x 64 tile) products the INT8-sliced product kernel has issued since the last
 * reset (each is 2 * 128 * 64 * 64 int8 operations): the EXECUTED work of the roofline -- in spatial mode the kernel skips
 * products of all-zero planes, so this is below the algorithmic S(S+1)/2 per chunk.  Synchronises the handle's stream. */
int gptb_executed_products(gptb_handle* h, int64_t* pairs, int reset);
/* developer switches, none of them on a hot path (A/B measurements only): "spatial_shuffle" (1 default; 0 keeps the plain Z-order,
 * before gptb_set_train), "oz_force_skip_variant" (1: the dense case runs through the skipping loops), "oz_whatif" (bit mask, acts
 * only in a -DGPTB_OZ_WHATIF build of the library: tools/whatif.py), "batch_cap" (queries per device batch), "back_substitution_variant"
 * (1 default: one flag-chained launch; 0: one launch per 128-row block), "spine_variant" (1 default: diag -> spine kernel -> diag on a
 * stream of its own; 0: the round-1 factorisation schedule), "variance_splitk" (1 default: small batches take the matrix-vector /
 * split-k forms of the FP64 variance product; 0: always one CTA per output tile). */
int gptb_set_debug_option(gptb_handle* h, const char* name, int value);
/* 64 cycle counters of CTA 0's roles in the last skipping product launch (all zero unless the library was built with
 * -DGPTB_OZ_WHATIF); the first call allocates the buffer, later launches fill it.  Layout: tools/whatif.py. */
int gptb_debug_read_profile(gptb_handle* h, int64_t* out64);

/* ---- unit-test hooks: exercise the DMMA tile engine and the small factor kernels in isolation.
 * C (128*mt,128*nt) = A (128*mt,K) * B(128*nt,K)^T, all row-major host arrays, K multiple of 128. */
int gptb_test_gemm_nt(gptb_handle* h, const double* A, const double* B, double* C, int mt, int nt, int K,
                      int maskA, int maskB);
int gptb_test_potrf_tile(gptb_handle* h, const double* A128, double* L128, double* Linv128, int* info);

#ifdef __cplusplus
}
#endif
#endif /* GPTB200_H */
