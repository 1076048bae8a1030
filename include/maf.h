/*
 * maf.h -- C ABI of the B200-native acquisition-maximisation hot path.
 *
 * The reference (j-wilson/MaximizingAcquisitionFunctions) is pure Python on
 * TensorFlow 1.x and has no FFI of its own; its "operator interface" for this
 * path is the set of Python call sites listed next to each entry point below
 * (file:line relative to the reference root).  A maintainer binds these with
 * ctypes; the in-tree binding is
 * maximizingacquisitionfunctions_b200/_native.py.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure;
 *     maf_last_error(ctx) returns a human-readable reason;
 *   - arrays are C-contiguous float64, caller-owned; "host" entry points take
 *     host pointers and do their own H2D/D2H copies on the context's stream and
 *     synchronise before returning; "_dev" entry points take device pointers,
 *     enqueue on the context's stream and do NOT synchronise;
 *   - one CUDA stream per context, no global state; a context is not
 *     thread-safe;
 *   - candidates ("pools") are X[S][q][d]: S restarts of q points; rows
 *     0..p-1 of each pool are pending points (constant, no gradient), rows
 *     p..q-1 are trainable (reference: bayesian_optimization.prepare,
 *     src/agents/bayesian_optimization.py:1051-1125).
 *   - limits: 1 <= q <= MAF_MAX_Q, 1 <= d <= MAF_MAX_D.
 *
 * Environment (read once in maf_create; tuning / diagnostics, the defaults are the measured best):
 *   MAF_GEMM_I8=0|1      contraction engine (1: exact INT8 digit planes on tcgen05, default; 0: FP64 DMMA)
 *   MAF_OZ_VARIANT=1|3|4|5  INT8 kernel: generic / static schedule / CTA pairs / CTA pairs + register-stash epilogue (default)
 *   MAF_OZ_NS_BWD=3..7   digit planes of the reverse contraction (default 7; 6 is 25 % faster but leaves the 1e-9 bar
 *                        for candidates next to training points)
 *   MAF_OZ_NS_FWD=3..7   digit planes of K10 in the forward contraction (default 7)
 *   MAF_MU_ALPHA=0|1     posterior mean as m + K10 . alpha in FP64 inside the cross-covariance kernel and its gradient
 *                        inside the reverse covariance kernel (default 1) instead of m + V^T gamma through the contractions
 *   MAF_LEVEL_SWITCH=0|1 error-bounded level switch of the two contractions (default 1, see maf_set_level_switch);
 *                        MAF_LEVEL_SWITCH_TOL=<rel> its tolerance (default 2.5e-10), MAF_LEVEL_SWITCH_MIN_N=<n> smallest padded
 *                        training-set size it applies to (default 1024)
 *   MAF_CENTRE_K10=0|1   digit planes of K10 - a^2/2 with the rank-one term added back by the forward epilogue (default 1: one more
 *                        bit of fixed-point range, half the truncation error) or of K10 itself
 *   MAF_G_ROUTE=0|1      the forward covariance kernel stores d kappa / d r2 of every (candidate, training point) pair and the
 *                        reverse one streams it (default 1) instead of recomputing square root and exponential per pair
 *   MAF_PEND_DEDUPE=0|1  pending points shared by all restarts are contracted once per evaluation (default 1)
 *   MAF_VBAR_TIGHT=0|1   row scales of the Vbar digit planes from the true row maxima (one more sweep; default 0)
 *   MAF_OZ_HINTS=<a><t>  L2 policy of the A / T plane loads, 0 normal, 1 evict-first, 2 evict-last (default 22)
 *   MAF_OZ_SJ=<k>        row blocks per super-row of the persistent tile order (default 2)
 *   MAF_OZ_SCHED=0|1     balanced static tile schedule of the CTA pairs (default 1; 0: round-robin walk)
 *   MAF_OZ_GROUPS=<g>    column groups of the balanced schedule: the T planes of a group stay in L2 while all row
 *                        blocks stream past (default 2)
 *   MAF_OZ_FUSE=<bits>   bit 0: cross-covariance straight to digit planes, bit 1: Vbar straight to digit planes (default 3)
 *   MAF_OZ_ROWMAX=0|1    row maxima of V^T from the forward contraction's epilogue (default 1) or from a sweep of the Vbar producer
 *   MAF_GRAPH=0|1        maf_adam_steps with >= 4 steps replays one captured step (CUDA graph; default 1), bit-identical to
 *                        the host-driven loop
 *   MAF_CHUNK_ROWS=<r>   candidate rows per chunk (default 65536: 4.3 GB of workspace)
 */
#ifndef MAF_H_
#define MAF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAF_MAX_Q 16
#define MAF_MAX_D 16

typedef struct maf_ctx maf_ctx;

/* MAF_F64: the reference's default dtype (src/core/module.py:29), 1e-9 bar.  MAF_F32: its dtype='float32' mode
 * (jitter 1e-4, src/models/gaussian_process.py:42-44; 1e-4 bar): the two contractions keep 5 digit planes per
 * operand (39-bit fixed point, 15 exact INT8 products instead of 28); all other arithmetic and every buffer of
 * this ABI stay double. */
enum maf_dtype { MAF_F64 = 0, MAF_F32 = 1 };
/* src/models/kernels.py:32-49 */
enum maf_kernel { MAF_KERNEL_SE = 0, MAF_KERNEL_MATERN52 = 1, MAF_KERNEL_MATERN32 = 2 };
/* src/losses/{negative_ei,negative_pi,simple_regret,confidence_bound}.py */
enum maf_acq { MAF_ACQ_EI = 0, MAF_ACQ_PI = 1, MAF_ACQ_SR = 2, MAF_ACQ_CB = 3 };

typedef struct maf_acq_params {
  int32_t acq;        /* enum maf_acq */
  int32_t lower;      /* CB: 1 = lower bound (default of the reference), 0 = upper */
  double level;       /* EI/PI improvement level; NaN => min(y0) (negative_ei.py:31-34) */
  double temperature; /* PI: NaN => hard indicator, else soft_less temperature (math_ops.py:56-69) */
  double beta;        /* CB: confidence parameter (confidence_bound.py:27-33) */
} maf_acq_params;

/* ---- life cycle ------------------------------------------------------- */
int maf_create(int device, int dtype, maf_ctx** out);
int maf_destroy(maf_ctx* ctx);
const char* maf_last_error(const maf_ctx* ctx);
const char* maf_version(void);
/* the context's CUDA stream (cudaStream_t as void*) and a full sync */
void* maf_stream(maf_ctx* ctx);
int maf_sync(maf_ctx* ctx);

/* ---- model + training set --------------------------------------------- */
/* Replaces gaussian_process.__init__/state (src/models/gaussian_process.py:37-75,
 * 377-411): ARD lengthscales[d], amplitude a (cov = a^2 * kappa), noise variance
 * (NaN => noiseless), constant prior mean (NaN => empirical mean of y0, :229-231),
 * jitter (1e-6 f64 / 1e-4 f32, :42-44). */
int maf_set_model(maf_ctx* ctx, int kernel_id, int d, const double* lenscales,
                  double amplitude, double noise, double mean, double jitter);

/* Replaces model.compute_cholesky(X0, noisy=True) + the cached-factor feed in
 * bayesian_optimization.appraise_inputs (src/agents/bayesian_optimization.py:188-198;
 * src/models/gaussian_process.py:352-360): builds K00 + (noise + jitter) I, its
 * Cholesky factor L0, the explicit inverse factor L0^-1 and gamma = L0^-1 (y0 - m)
 * on the device.  Non-zero return if K00 is not positive definite. */
int maf_set_train(maf_ctx* ctx, int n, const double* X0 /*[n][d]*/, const double* y0 /*[n]*/);
/* copies of the cached factors, for tests: L0[n][n] lower, Linv[n][n] lower */
int maf_get_factors(maf_ctx* ctx, double* L0, double* Linv, double* gamma);

/* Data term of the negative log marginal likelihood of the loaded training set under the loaded
 * model and its gradient w.r.t. (log lengthscales[d], log amplitude, log noise, mean): replaces
 * gaussian_process.log_likelihood(as_negative=True) + tf.gradients inside gaussian_process.fit
 * (src/models/gaussian_process.py:78-106,283-325); the hyper-priors (priors.py) are scalar
 * functions the host adds.  grad has d + 3 entries. */
int maf_gp_nll(maf_ctx* ctx, double* nll, double* grad);

/* ---- GP posterior ------------------------------------------------------ */
/* Replaces gaussian_process.predict/_predict_nd (src/models/gaussian_process.py:
 * 109-117,172-223) for rank-3 pools against a rank-2 training set.
 * mu[S][q]; full_cov != 0: cov[S][q][q], else cov[S][q] (marginal variances). */
int maf_posterior(maf_ctx* ctx, int S, int q, const double* X, int full_cov,
                  double* mu, double* cov);

/* ---- acquisition value + gradient -------------------------------------- */
/* Replaces loss_fn(pools, ...) -> myopic_loss.evaluate/monte_carlo/closed_form
 * (src/losses/myopic_loss.py:63-244) plus the tf.gradients pass that
 * optimizer.minimize builds (src/agents/bayesian_optimization.py:420-426).
 * q == 1 with z == NULL evaluates the closed form (myopic_loss.py:84-90),
 * otherwise the MC estimator with base samples z[M][q] (shared by all restarts),
 * optional weights[M] (sum_m w_m u_m instead of the mean) and incumbents[M].
 * out_loss[S]; out_grad[S][q-p][d] = d(sum_s loss_s)/dX (NULL => value only). */
int maf_acq_value_grad(maf_ctx* ctx, const maf_acq_params* prm, int S, int q, int p_pending,
                       const double* X, int M, const double* z, const double* weights,
                       const double* incumbents, double* out_loss, double* out_grad);
/* same, device pointers, asynchronous on maf_stream(ctx) */
int maf_acq_value_grad_dev(maf_ctx* ctx, const maf_acq_params* prm, int S, int q, int p_pending,
                           const double* dX, int M, const double* dz, const double* dweights,
                           const double* dincumbents, double* dout_loss, double* dout_grad);
/* Monte-Carlo estimator from GIVEN posterior moments: replaces the loss classes'
 * _monte_carlo(means, cov, base_rvs=...) bodies (negative_ei.py:52-74, negative_pi.py:52-80,
 * simple_regret.py:32-49, confidence_bound.py:48-116) incl. jitter_cholesky of cov with the
 * model's jitter.  mu[S][q], cov[S][q][q] (lower triangle read), z[M][q]; outputs (NULL to skip)
 * out_loss[S], out_mubar[S][q], out_covbar[S][q][q] (d sum(loss) / d mu, / d cov), out_chol. */
int maf_mc_from_moments(maf_ctx* ctx, const maf_acq_params* prm, int S, int q, const double* mu,
                        const double* cov, int M, const double* z, const double* weights,
                        const double* incumbents, double* out_loss, double* out_mubar,
                        double* out_covbar, double* out_chol);
/* ---- fantasy-conditioned closed form (incremental greedy) ----------------- */
/* Replaces the rank-4 ("padded") observation-set path: scripts/run_bayesopt_incremental.py:138-218
 * builds inputs_old[F,1,n,d] (identical copies) and outputs_old[F,1,n,1]; gaussian_process._predict_nd
 * (:195-206) solves once with chol[0,0] and appraise_inputs (:227-229) averages the [F,S,1,1] closed-form
 * losses over fantasies.  maf_set_fantasies keeps the factor of the loaded training inputs and loads
 * F output vectors Y[F][n] (per-fantasy prior mean if the model mean is empirical, per-fantasy level
 * min(Y_f) unless prm->level is given).  maf_fantasies_value_grad: loss[S] = mean_f loss(mu_sf, var_s,
 * level_f) for single-point candidates X[S][d], out_grad[S][d] (NULL => value only).
 * maf_fantasies_posterior: mu[F][S], var[S]. */
int maf_set_fantasies(maf_ctx* ctx, int F, const double* Y);
int maf_fantasies_value_grad(maf_ctx* ctx, const maf_acq_params* prm, int S, const double* X,
                             double* out_loss, double* out_grad);
int maf_fantasies_posterior(maf_ctx* ctx, int S, const double* X, double* mu, double* var);

/* extras of the last evaluation (myopic_loss.py:179 returns (means, cov, chol)):
 * mu[S][q], cov[S][q][q], chol[S][q][q] (NULL to skip) */
int maf_last_extras(maf_ctx* ctx, double* mu, double* cov, double* chol);

/* ---- multi-start Adam loop --------------------------------------------- */
/* Replaces bayesian_optimization._optimize_inputs_sgd (src/agents/
 * bayesian_optimization.py:377-482): TF-1 Adam on sum_s loss_s, then
 * clip_by_value(x, 0, 1); a fresh z[M][q] per step.  maf_adam_begin uploads
 * the starts X[S][q][d] (pending rows included) and zeroes the slots;
 * maf_adam_steps runs `steps` updates with z_steps[steps][M][q] and may be called
 * repeatedly (the host owns the time limit); out_loss_trace[steps][S] may be NULL;
 * maf_adam_end downloads the iterate. */
int maf_adam_begin(maf_ctx* ctx, const maf_acq_params* prm, int S, int q, int p_pending,
                   const double* X, double lr, double beta1, double beta2, double eps);
/* Optimizer of the loop, between maf_adam_begin and the first step (default Adam).  MAF_OPT_GD / MAF_OPT_MOMENTUM are
 * the reference's tf.train.GradientDescentOptimizer / MomentumOptimizer(., momentum) branches (:405-412): learning
 * rate lr * (global_step + 1)^-0.7 (lr of maf_adam_begin, global_step from 0), accum = momentum*accum + g,
 * x -= lr_t*accum, then the same clip. */
enum maf_optimizer { MAF_OPT_ADAM = 0, MAF_OPT_GD = 1, MAF_OPT_MOMENTUM = 2 };
int maf_adam_set_method(maf_ctx* ctx, int method, double momentum);
int maf_adam_steps(maf_ctx* ctx, int steps, int M, const double* z_steps, double* out_loss_trace);
int maf_adam_peek(maf_ctx* ctx, double* X_out);   /* current iterate, loop stays open */
int maf_adam_end(maf_ctx* ctx, double* X_out);

/* ---- multi-start L-BFGS loop ------------------------------------------- */
/* Replaces bayesian_optimization._optimize_inputs_scipy (src/agents/bayesian_optimization.py:485-555:
 * ScipyOptimizerInterface, method 'L-BFGS-B', bounds (0, 1), fixed base samples), the default subroutine for pools of
 * one point (:311-313), of suggest_answers (:938-1048) and of the incremental greedy construction
 * (scripts/run_bayesopt_incremental.py:138-218).  The reference minimises mean_s loss_s jointly; the objective is
 * separable over restarts, so every restart runs its own box-constrained L-BFGS here (history 10, projected Armijo
 * backtracking; csrc/lbfgs.cuh), all restarts in lock step, one value+gradient evaluation per iteration, and nothing
 * leaves the device between iterations (the host looks at a counter every 8 iterations: all converged? out of time?).
 * X[S][q][d] host, in: starts (pending rows first), out: the best (= last accepted) iterate of every restart;
 * z[M][q] fixed base samples (NULL with q == 1: closed form); fantasies != 0: the fantasy-conditioned closed form of
 * maf_set_fantasies (q == 1).  Stops per restart on |projected gradient|_inf <= pgtol or on an accepted step improving
 * the loss by <= ftol * max(|f|, |f'|, 1) (SciPy's factr test).  out_loss[S]; out_info[2] = {evaluations per restart,
 * restarts not converged}; out_trace[maxiter][S] (NULL to skip): the loss of every evaluated trial point, rows
 * [0, out_info[0]) valid.  time_limit_s: NaN / inf for none. */
int maf_lbfgs_run(maf_ctx* ctx, const maf_acq_params* prm, int S, int q, int p_pending, double* X, int M,
                  const double* z, int fantasies, int maxiter, double ftol, double pgtol, double time_limit_s,
                  double* out_loss, int32_t* out_info, double* out_trace);

/* ---- selection --------------------------------------------------------- */
/* np.argmin semantics (first minimum; src/agents/bayesian_optimization.py:152) over
 * a host loss vector is the caller's; this is the device-side variant over the losses
 * of the last evaluation: index and value of the first minimum (NaNs never win). */
int maf_argmin_last(maf_ctx* ctx, int64_t* idx, double* value);
/* Multi-GPU selection (restarts sharded over ranks, src/misc/sharded_fn.py:106-132 maps over independent restarts):
 * writes this rank's record [loss*, index_offset + local index, X*[q][d]] of the last evaluation to DEVICE memory
 * d_rec[2 + q*d], asynchronously on maf_stream(ctx) (index -1 / loss +inf when no loss is finite).  One all-gather
 * of these records over NCCL and an arg-min with the lowest index winning ties reproduces np.argmin over all
 * restarts (src/agents/bayesian_optimization.py:152). */
int maf_best_record_dev(maf_ctx* ctx, int64_t index_offset, double* d_rec);

/* ---- profiling (CUDA events on the context's stream) -------------------- */
enum maf_prof_slot {
  MAF_PROF_CROSS_FWD = 0, MAF_PROF_GEMM_FWD = 1, MAF_PROF_POSTERIOR = 2, MAF_PROF_MC = 3,
  MAF_PROF_VBAR = 4, MAF_PROF_GEMM_BWD = 5, MAF_PROF_CROSS_BWD = 6, MAF_PROF_ADAM = 7,
  MAF_PROF_NSLOTS = 8
};
int maf_profile_enable(maf_ctx* ctx, int on);
/* accumulated device milliseconds and launch counts per slot since the last reset */
int maf_profile_read(maf_ctx* ctx, double* ms /*[MAF_PROF_NSLOTS]*/, int64_t* launches /*[MAF_PROF_NSLOTS]*/);
int maf_profile_reset(maf_ctx* ctx);
/* total kernels launched by this context since creation */
int64_t maf_launch_count(const maf_ctx* ctx);
/* a pair of CUDA events on the context's stream: start is recorded now; stop records the
 * second event, waits for it and returns the device time between the two */
int maf_timer_start(maf_ctx* ctx);
int maf_timer_stop(maf_ctx* ctx, double* ms);

/* ---- raw building blocks (exported for tests / micro-benchmarks) -------- */
/* C[J][N] = A[J][K] * T[N][K]^T, row-major device pointers, J,N,K multiples of 128;
 * tri: 0 full, 1 T lower-triangular (K==N), 2 T upper-triangular (K==N). */
int maf_gemm_nt_dev(maf_ctx* ctx, int J, int N, int K, const double* dA, int lda,
                    const double* dT, int ldt, double* dC, int ldc, int tri);
/* Same contraction on the INT8 tensor pipe by exact base-256 digit slicing (csrc/ozaki_i8.cuh):
 * ns digit planes per operand (<= 7), levels up to lmax kept (ns=7, lmax=8: FP64-grade). */
int maf_gemm_nt_i8_dev(maf_ctx* ctx, int J, int N, int K, const double* dA, int lda, const double* dT,
                       int ldt, double* dC, int ldc, int tri, int ns, int lmax);
/* 1 (default): FP64-grade contractions by exact INT8 digit slicing on tcgen05 (n <= 8192; larger n uses DMMA);
 * 0: FP64 DMMA contractions.  Takes effect at the next maf_set_train.  Environment override: MAF_GEMM_I8. */
int maf_set_gemm_mode(maf_ctx* ctx, int mode);
/* Digit planes of the two INT8 contractions (levels <= planes + 1 are kept): forward V^T = K10 L0^-T and reverse
 * Kbar10 = Vbar^T L0^-1.  7 / 7 (28 + 28 exact products) is the FP64-grade default, 3 <= ns <= 7 (5 / 5 in MAF_F32 mode).
 * Accuracy studies and the level switch of the host layer use it; takes effect at the next evaluation. */
int maf_set_digits(maf_ctx* ctx, int ns_fwd, int ns_bwd);
/* Level switch (default on for n >= 1024; MAF_LEVEL_SWITCH=0 turns it off): both contractions first keep
 * one digit plane less (21 instead of 28 exact products; MAF_F32 mode: 10 instead of 15, tol 2.5e-5).  maf_set_train measures the error that level leaves on the
 * posterior covariance and on the gradient against the full level on 512 probe candidates; every restart whose own result
 * (max |Sigma_s|, max |gradient_s|) does not tolerate that error within `tol` (relative; default 2.5e-10, a quarter of
 * the 1e-9 parity bar; NaN keeps the current value) is recomputed with all planes, on the device, inside the same call.
 * A direction whose cheaper level fails for more than 40 % of the restarts runs on all planes directly for the next 64
 * evaluations (a host decision between evaluations; results never depend on it).
 * maf_level_switch_stats: counts[4] = restarts tested / recomputed forward, tested / recomputed reverse since the last
 * reset, errs[3] = calibrated covariance error, calibrated gradient error per unit row scale, tolerance. */
int maf_set_level_switch(maf_ctx* ctx, int on, double tol);
int maf_level_switch_stats(maf_ctx* ctx, int64_t* counts, double* errs, int reset);
void* maf_dev_alloc(maf_ctx* ctx, size_t bytes);
int maf_dev_free(maf_ctx* ctx, void* p);
int maf_h2d(maf_ctx* ctx, void* dst, const void* src, size_t bytes);
int maf_d2h(maf_ctx* ctx, void* dst, const void* src, size_t bytes);
/* page-locked host memory (so that the host entry points copy at full PCIe rate) */
void* maf_host_alloc(maf_ctx* ctx, size_t bytes);
int maf_host_free(maf_ctx* ctx, void* p);

#ifdef __cplusplus
}
#endif
#endif /* MAF_H_ */
