/*
 * batman_b200 -- C ABI of the B200-native Kriging-over-Saltelli -> Sobol' hot path.
 *
 * The reference (tupui/batman 1.9) is pure Python and has no FFI of its own;
 * its boundary for this path is two Python classes.  Each entry point below
 * names the reference call site it replaces (paths relative to the reference
 * root; "sklearn:" = site-packages/sklearn/gaussian_process/ of 1.9.0).
 * INTEGRATION.md shows the ctypes binding a batman maintainer would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch types.
 *   - every function returns 0 on success, non-zero on error; the message is
 *     in bk_last_error() (thread-local).  Nothing throws.
 *   - "_dev" pointers are device memory owned by the caller (torch tensors in
 *     the Python host); "_host" pointers are host memory read during the call.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - all arithmetic is IEEE binary64.
 */
#ifndef BATMAN_B200_H
#define BATMAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bk_gp_s* bk_gp_t;

/* stationary factor of the covariance, sklearn:kernels.py:1723-1730, 1557-1563 */
enum bk_kind {
    BK_MATERN12 = 0,  /* exp(-r)                                   Matern(nu=0.5) */
    BK_MATERN32 = 1,  /* (1 + sqrt3 r) exp(-sqrt3 r)               Matern(nu=1.5), batman default kriging.py:85-88 */
    BK_MATERN52 = 2,  /* (1 + sqrt5 r + 5r^2/3) exp(-sqrt5 r)      Matern(nu=2.5) */
    BK_RBF = 3,       /* exp(-0.5 r^2) on the squared distance     RBF */
    BK_MATERN_INF = 4, /* exp(-(r*r)/2) on the rooted distance     Matern(nu=inf) */
    BK_GENERAL = 5     /* expanded Sum/Product tree with several stationary leaves (bk_gp_set_output_general) */
};

/* Marginals of the UQ inputs: ot.<Name>(...) strings of uq.py:151-155, OpenTURNS 1.15 parametrisations.  Every marginal
 * carries BK_MARGINAL_NPAR = 4 doubles (unused ones 0); a sample is the inverse CDF of a Philox uniform in (0, 1). */
enum bk_marginal_kind {
    BK_UNIFORM = 0,      /* (a, b)                      a + (b - a) u */
    BK_NORMAL = 1,       /* (mu, sigma)                 mu + sigma ndtri(u) */
    BK_LOGNORMAL = 2,    /* (muLog, sigmaLog, gamma)    gamma + exp(muLog + sigmaLog ndtri(u)) */
    BK_TRIANGULAR = 3,   /* (a, m, b) */
    BK_TRUNCNORMAL = 4,  /* (mu, sigma, Phi((a-mu)/sigma), Phi((b-mu)/sigma)): the host evaluates the two CDF values */
    BK_BETA = 5,         /* (alpha, beta, a, b)         a + (b - a) betaincinv(alpha, beta, u); also BetaMuSigma */
    BK_GUMBEL = 6,       /* (beta, gamma)               gamma - beta log(-log u) */
    BK_EXPONENTIAL = 7,  /* (lambda, gamma)             gamma - log(1 - u) / lambda */
    BK_LOGUNIFORM = 8,   /* (aLog, bLog)                exp(aLog + (bLog - aLog) u) */
    BK_WEIBULLMIN = 9,   /* (beta, alpha, gamma)        gamma + beta (-log(1 - u))^(1/alpha) */
    BK_LOGISTIC = 10,    /* (mu, beta)                  mu + beta log(u / (1 - u)) */
    BK_RAYLEIGH = 11,    /* (beta, gamma)               gamma + beta sqrt(-2 log(1 - u)) */
    BK_MARGINAL_COUNT = 12
};
#define BK_MARGINAL_NPAR 4

int bk_version(void);
const char* bk_last_error(void);
int bk_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* Optional per-kernel timing with CUDA events on the launching stream (bench.py roofline numbers).
 * Categories: 0 = gp_kstar_mean (K1), 1 = gp_var_trmm (K2), 2 = fused Saltelli (K1+K3+K6),
 * 3 = Saltelli reducer over supplied outputs, 4 = everything else, 5 = LML / Cholesky / inverse (K5), 6 = POD kernels (K4).  bk_profile_read synchronises,
 * returns the summed milliseconds and launch counts since the last read (arrays of 8) and clears. */
int bk_profile_enable(int on);
int bk_profile_read(double* ms_by_cat, long* launches_by_cat);

/* Diagnostics: the kernels' own branch-free sqrt(v) and exp(-v) (csrc/fastmath.cuh) on n values v >= 0. */
int bk_debug_fastmath(const double* in_dev, long n, double* sqrt_dev, double* exp_neg_dev, void* stream);

/* ---- model state: replaces the fitted attributes of the m GaussianProcessRegressor
 *      objects held in Kriging.gp_models (kriging.py:148; sklearn:_gpr.py:349-367) ---- */
int bk_gp_create(bk_gp_t* out, int n_train, int p, int m);
int bk_gp_destroy(bk_gp_t h);
/* padded input dimension the kernels are instantiated for (xs_dev row pitch) */
int bk_gp_padded_dim(int p);
/* padded training size (multiple of 128) used by the packed L^-1 tiles */
int bk_gp_padded_n(int n_train);
/* doubles in the packed tile image of one n_train x n_train lower-triangular W = L^-1 */
size_t bk_gp_wtiles_len(int n_train);
/* output q:  k(x, x') = c0 + c1 * stat_kind(||(x - x')/ls||),  kdiag = kernel_.diag(x)
 * (WhiteKernel noise included).  xs_dev = X_train / ls, row-major n x P, zero padded.
 * w_tiles_dev may be NULL if sigma is never requested. */
int bk_gp_set_output(bk_gp_t h, int q, int kind, double c0, double c1, double kdiag,
                     const double* ls_host, const double* xs_dev, const double* alpha_dev,
                     const double* w_tiles_dev, double y_mean, double y_std);
/* General Sum/Product trees over ConstantKernel / WhiteKernel / RBF / Matern leaves (sklearn:kernels.py:838-871,
 * 936-990; the reference accepts any such expression, kriging.py:80-96, driver.py:181-190), expanded by the host to
 *     k(x, x') = sum_{t < nt} tcoef[t] * prod_{f < nf} stat_{fkind[f]}(||(x - x')/ls_f||) ^ texp[t*nf + f]
 * with nf <= 3 stationary leaves, nt <= 6 terms and nf * bk_gp_padded_dim(p) <= 48.  fls_host is nf x p,
 * xs_dev is row-major n_pad x (nf * P): row j = [X_j/ls_0 | X_j/ls_1 | ...], each part zero padded to P.
 * Trees with ONE leaf and an affine expansion go through bk_gp_set_output (the fast kernels). */
int bk_gp_set_output_general(bk_gp_t h, int q, int nf, const int* fkind_host, const double* fls_host, int nt,
                             const double* tcoef_host, const int* texp_host, double kdiag, const double* xs_dev,
                             const double* alpha_dev, const double* w_tiles_dev, double y_mean, double y_std);
/* Optional int8 image of L^-1 for the tcgen05 variance kernel (error-free 8 x 7-bit slicing, csrc/var_ozaki.cu):
 * wq_dev bk_gp_wq_bytes(n) bytes of slice tiles, wsig_dev n_pad power-of-two row scales; both produced by
 * bk_tri_inverse_pack / bk_gp_factorise with wq_scheme 0 (8 signed 7-bit slices, n_train <= 16384) or 1 (7 8-bit
 * slices, W signed / K* unsigned bytes, n_train <= 8192, c0, c1 >= 0: 22 % fewer MMAs).  bk_gp_set_var_mode selects
 * the sigma kernel: 0 = FP64 DMMA, 1 = int8 tensor cores + TMEM with scheme 0, 2 = with scheme 1 (the slices set with
 * bk_gp_set_output_int8 must have been produced with the matching wq_scheme).  Results agree with mode 0 to ~1e-14 of
 * the prior variance. */
size_t bk_gp_wq_bytes(int n_train);
int bk_gp_set_output_int8(bk_gp_t h, int q, const int8_t* wq_dev, const double* wsig_dev);
int bk_gp_set_var_mode(bk_gp_t h, int mode);
/* CTAs that share one tile of 128 test points in the int8 sigma kernel (csrc/var_ozaki.cu): chosen from n_train so that
 * the K* tiles in flight (sm_count / G of them, 896 * n_pad bytes each) stay in the L2.  Pure host arithmetic. */
int bk_sigma_gang_size(int n_train, int sm_count);
/* on = 1: a sigma prediction of several units (chunks of points x outputs) runs on two internal streams, K1 of unit
 * u+1 (FP64 pipe) beside the sigma kernel of unit u (tensor pipe), on two K* buffers inside the workspace; the
 * caller's stream forks into them and joins again before bk_gp_predict returns.  on = 0 (default): every kernel on the
 * caller's stream, one K* buffer.  Same results either way.  On a B200 both kernels run into the board's power cap on
 * their own, so the overlap trades clock for concurrency and the step time does not move (profiles/r02_overlap_probe.jsonl);
 * the switch exists for boards with power headroom. */
int bk_gp_set_overlap(bk_gp_t h, int on);
/* MinMaxScaler of SurrogateModel (surrogate_model.py:108-109,181): x' = x*scale + min.
 * NULL, NULL = identity (Kriging.evaluate receives already-scaled points). */
int bk_gp_set_scaler(bk_gp_t h, const double* scale_host, const double* min_host);

/* row-major lower-triangular W (n x n) -> packed DMMA-fragment tiles (bk_gp_wtiles_len doubles) */
int bk_pack_w(const double* w_dev, int n_train, double* w_tiles_dev, void* stream);

/* workspace (bytes) bk_gp_predict needs for k points when sigma is requested */
size_t bk_gp_predict_workspace(bk_gp_t h, long k, int want_std);
/* Kriging.evaluate (kriging.py:192-214) == gp.predict(return_std=True) per output
 * (sklearn:_gpr.py:444-500), batched: pts_dev k x p row-major -> mean_dev, std_dev k x m
 * row-major.  std_dev may be NULL (mean only). */
int bk_gp_predict(bk_gp_t h, const double* pts_dev, long k, double* mean_dev, double* std_dev,
                  void* workspace_dev, size_t workspace_bytes, void* stream);
/* The int8 sigma kernel bounds every mbarrier wait so that a pipeline fault cannot hang the device; a wait that
 * timed out is recorded in a device-side flag that lives in the LAST 256 bytes of the caller's workspace (zeroed by
 * bk_gp_predict on `stream` before the kernels run -- callers must not use that tail).  This call waits for `stream`
 * and returns the flag in *fault_host: 0 = every launch since the last bk_gp_predict completed normally, 1 = sigma
 * of that call is invalid (the return status of bk_gp_predict cannot carry it: the kernels are asynchronous). */
int bk_gp_predict_status(const void* workspace_dev, size_t workspace_bytes, int* fault_host, void* stream);

/* ---- hyper-parameter fit: replaces log_marginal_likelihood(theta) (sklearn:_gpr.py:584-617) as driven by
 *      Kriging._optim_evolution / the L-BFGS-B restarts (kriging.py:99-106, 151-190), for R candidates at once,
 *      and the final factorisation of GaussianProcessRegressor.fit (sklearn:_gpr.py:349-367). ---- */
/* bytes of workspace for R candidates (with_inverse = 1 adds room for the triangular inverse) */
size_t bk_lml_workspace(int n_train, int n_candidates, int with_inverse);
/* lml_dev[r] = LML of k_r(x,x') = c0[r] + c1[r]*stat_kind[r](||(x-x')/ls[r]||), K_ii = kdiag[r] + 1e-10;
 * -inf where the Cholesky fails.  x_dev n x p (scaled design).  y_dev is n_targets x n row-major: the normalised
 * targets of n_targets outputs; candidate r is scored on row target_host[r] (NULL = all on row 0), so that the
 * populations of ALL restarts (kriging.py:176-181) of ALL outputs (kriging.py:117-143) share one batched call. */
int bk_lml_batched(const double* x_dev, const double* y_dev, int n_targets, const int* target_host, int n_train, int p,
                   int n_candidates, const int* kind_host, const double* c0_host, const double* c1_host, const double* kdiag_host,
                   const double* ls_host /* R x p */, double* lml_dev, void* workspace_dev, size_t workspace_bytes,
                   void* stream);
/* The same for general trees (bk_gp_set_output_general): the tree shape (nf, fkind, nt, texp) is shared by the
 * candidates; tcoef_host is R x nt, fls_host R x nf x p, kdiag_host R. */
int bk_lml_batched_general(const double* x_dev, const double* y_dev, int n_targets, const int* target_host, int n_train,
                           int p, int n_candidates, int nf,
                           const int* fkind_host, int nt, const int* texp_host, const double* tcoef_host,
                           const double* fls_host, const double* kdiag_host, double* lml_dev, void* workspace_dev,
                           size_t workspace_bytes, void* stream);
/* Final state at one theta: l_dev (n_pad x n_pad row-major, lower triangle = L_), alpha_dev (n_pad) = alpha_,
 * w_tiles_dev (bk_gp_wtiles_len doubles or NULL) = packed L^-1, lml_dev (1), info_dev (1; 0 = positive definite). */
int bk_gp_factorise(const double* x_dev, const double* y_dev, int n_train, int p, int kind, double c0, double c1,
                    double kdiag, const double* ls_host, double* l_dev, double* alpha_dev, double* w_tiles_dev,
                    int8_t* wq_dev /* or NULL */, double* wsig_dev /* or NULL */, int wq_scheme, double* lml_dev,
                    int* info_dev, void* workspace_dev, size_t workspace_bytes, void* stream);
int bk_gp_factorise_general(const double* x_dev, const double* y_dev, int n_train, int p, int nf,
                            const int* fkind_host, int nt, const int* texp_host, const double* tcoef_host,
                            const double* fls_host, double kdiag, double* l_dev, double* alpha_dev,
                            double* w_tiles_dev, int8_t* wq_dev /* or NULL */, double* wsig_dev /* or NULL */,
                            int wq_scheme, double* lml_dev, int* info_dev, void* workspace_dev,
                            size_t workspace_bytes, void* stream);
/* Packed L^-1 tiles from a given Cholesky factor l_dev (n x n row-major lower), e.g. scikit-learn's L_;
 * workspace: bk_lml_workspace(n, 1, 1). */
int bk_tri_inverse_pack(const double* l_dev, int n_train, double* w_tiles_dev, int8_t* wq_dev /* or NULL */,
                        double* wsig_dev /* or NULL */, int wq_scheme, void* workspace_dev, size_t workspace_bytes,
                        void* stream);

/* ---- Saltelli experiment + estimator: replaces ot.SobolIndicesExperiment(...).generate(),
 *      UQ.func over the design, and ot.SaltelliSensitivityAlgorithm (uq/uq.py:331-349,367-375) ---- */
/* number of accumulated sums per output for input dimension p (the last one is the row count of the slot) */
int bk_sobol_nsums(int p, int second_order);
/* number of per-CTA partial slots the accumulate kernel writes */
int bk_sobol_nslots(void);
/* Counter-based base samples (Philox4x32-10; oracle/philox.py is the CPU statement):
 * rows [row_lo, row_lo+count) of stream 0 (A) or 1 (B) -> out_dev count x p. */
int bk_design_rows(uint64_t seed, long row_lo, long count, int p, const int* marg_kind_host,
                   const double* marg_par_host, int stream_id, double* out_dev, void* stream);
/* Predict the surrogate mean on the Saltelli blocks [A;B;E^i;(C^i)] of base rows
 * [row_lo, row_lo+count) and ADD the compensated sums into partials_dev
 * ([nslots][m][nsums][2] doubles, zero it before the first call).
 * a_dev/b_dev: explicit base samples (count x p) or NULL to generate in-kernel.
 * shift_host: m pilot shifts subtracted before products (any value; the GP's y_mean is free). */
int bk_sobol_accumulate(bk_gp_t h, const double* a_dev, const double* b_dev, uint64_t seed,
                        const int* marg_kind_host, const double* marg_par_host, long row_lo,
                        long count, int second_order, const double* shift_host,
                        double* partials_dev, void* stream);
/* Same design and prediction, but the per-block outputs are STORED instead of reduced:
 * ymodes_dev[(b*count + k)*m + q] = output q of block b at base row row_lo + k  (input of bk_pod_sobol_accumulate). */
int bk_sobol_modes(bk_gp_t h, const double* a_dev, const double* b_dev, uint64_t seed,
                   const int* marg_kind_host, const double* marg_par_host, long row_lo, long count,
                   int second_order, double* ymodes_dev, void* stream);
/* Pod.inverse_transform (pod/pod.py:219-228): out_dev (k x F) = mean_dev (F) + modes_dev (k x m) . u_dev^T (F x m),
 * FP64 tensor cores.  SurrogateModel.__call__ applies it to the mean and to sigma (surrogate_model.py:191-196). */
int bk_pod_inverse(const double* modes_dev, long k, int m, const double* u_dev, const double* mean_dev, int F,
                   double* out_dev, void* stream);
/* POD field + Saltelli sums fused: lifts the stored mode outputs to the F-feature field on the tensor cores and
 * ADDs the sums into partials_dev ([nslots][F][nsums][2]) without materialising the field (uq.py:333-343 with a POD
 * surrogate).  'block' indices: pass F = 1 with u_dev = trapz weights applied to U, mean_dev = trapz(mean). */
int bk_pod_sobol_accumulate(const double* ymodes_dev, long count, int p, int m, int second_order,
                            const double* u_dev, const double* mean_dev, int F, const double* shift_dev,
                            double* partials_dev, void* stream);
/* Same sums from a user-supplied output sample (ensemble mode, uq.py:336-339):
 * y_dev is the block-major (nblocks*N) x F output design restricted to base rows
 * [row_lo,row_lo+count): element (block b, row k, feature f) at y_dev[(b*ld_block + k)*F + f]. */
int bk_sobol_accumulate_y(const double* y_dev, long ld_block, long count, int p, int F,
                          int second_order, const double* shift_dev, double* partials_dev,
                          void* stream);
/* fixed-order reduction of the slots: totals_dev [F][nsums][2] (this is what is all-reduced) */
int bk_sobol_reduce_slots(const double* partials_dev, int F, int nsums, double* totals_dev, void* stream);
/* Confidence intervals of the aggregated indices: replaces sobol.getFirstOrderIndicesInterval() /
 * getTotalOrderIndicesInterval() (uq/uq.py:341, 425-426; OpenTURNS' asymptotic intervals) by a delete-a-group jackknife
 * over the partial slots, which are disjoint groups of base rows.  totals_dev: the GLOBAL totals (after the all-reduce)
 * of N base rows; partials_dev: this process's slots.  out_dev [nslots][2 + 2p] =
 * {active (1/0), rows in the group, S1agg without the group [p], STagg without the group [p]}. */
int bk_sobol_jackknife(const double* partials_dev, const double* totals_dev, long N, int p, int F, int second_order,
                       double* out_dev, void* stream);
/* Asymptotic (delta-method) intervals, the method OpenTURNS itself uses for uq.py:341, 425-426: from the STORED outputs of
 * the A, B and E^i blocks (z_dev[(b * ld_block + k) * r + c], b = 0: A, 1: B, 2+i: E^i; r outputs per row; what
 * bk_sobol_modes / bk_sobol_accumulate_keep wrote) and the per-output mean over all rows of the experiment (center_dev,
 * r values, known after bk_sobol_finalize), ADD into sums_dev ((2 + 6p) x 2 doubles (hi, lo), zero before the first
 * chunk) the sums over base rows of  b, b^2, and per input i: a1, a1^2, a1 b, a2, a2^2, a2 b  with
 * b = <zA, zA>, a1 = <zB, zE^i>, a2 = <zA, zE^i> (dot products over the r centred outputs).  The host forms
 * Var(S_i) = grad(psi)^T Cov(a1, b) grad(psi) / N for psi(x, y) = x / y (total order: 1 - x / y). */
int bk_sobol_asymptotic(const double* z_dev, long ld_block, long count, int p, int r, const double* center_dev,
                        double* sums_dev, void* stream);
/* bk_sobol_accumulate that ALSO keeps the outputs: ykeep_dev[(b * count + k) * m + q] for every block b (the layout of
 * bk_sobol_modes), so that the interval pass does not predict again. */
int bk_sobol_accumulate_keep(bk_gp_t h, const double* a_dev, const double* b_dev, uint64_t seed,
                             const int* marg_kind_host, const double* marg_par_host, long row_lo, long count,
                             int second_order, const double* shift_host, double* partials_dev, double* ykeep_dev,
                             void* stream);
/* The N-point randomised LHS of UQ.__init__ (ot.LHSExperiment(dist, N, True, True), uq.py:159-160): row i, column c is
 * the inverse CDF of (perm_dev[i * p + c] + shift) / count, shift = the Philox uniform of (row i, column c, stream 2);
 * perm_dev (count x p int64): every column a permutation of 0 .. count-1.  out_dev count x p. */
int bk_lhs_rows(uint64_t seed, long count, int p, const int* marg_kind_host, const double* marg_par_host,
                const long* perm_dev, double* out_dev, void* stream);
/* Moments of a stored output sample y_dev (rows x F row-major), what UQ.error_propagation reports (uq.py:501-529):
 * out_dev F x 6 = {sum hi, sum lo, sum of squares hi, lo (both of y - shift_dev[f]; shift_dev may be NULL), min, max}. */
int bk_moments(const double* y_dev, long rows, int F, const double* shift_dev, double* out_dev, void* stream);
/* doubles in the packed result of bk_sobol_finalize */
size_t bk_sobol_result_len(int p, int F);
/* Estimators from the totals.  out_dev layout (row-major, all double):
 *   S1[F][p] ST[F][p] S2[F][p][p] V1[F][p] VT[F][p] var[F] outvar[F] J1[F][p] JT[F][p]
 *   S1agg[p] STagg[p] S2agg[p][p]
 *   M1[F][p] MT[F][p] (Martinez) K1[F][p] KT[F][p] (Mauntz-Kucherenko) -- the other OpenTURNS estimators uq.py:340 names
 *   J1agg[p] JTagg[p] M1agg[p] MTagg[p] K1agg[p] KTagg[p]  mean[F] (of y - shift over all rows) */
int bk_sobol_finalize(const double* totals_dev, long N, int p, int F, int second_order,
                      double* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BATMAN_B200_H */
