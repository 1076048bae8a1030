// Monte-Carlo acquisition: per-restart q x q Cholesky, reparameterised samples
// y = mu + L z, utility, sample mean -- and the matching reverse pass to
// (mubar, Sigmabar), fused in one warp-per-restart kernel.
//
// Reference op chain (one TF graph op each, [S,M,q] intermediates materialised):
//   utils.jitter_cholesky (src/utils/linalg_ops.py:131-150) at myopic_loss.py:174,
//   myopic_loss.draw_samples :247-272, negative_ei._monte_carlo (negative_ei.py:52-74),
//   negative_pi._monte_carlo (negative_pi.py:52-80), simple_regret._monte_carlo
//   (simple_regret.py:32-49), confidence_bound._monte_carlo/draw_samples/
//   transform_residuals (confidence_bound.py:48-116), reduce_samples :275-284, and the
//   tf.gradients chain back through tensordot / tf.cholesky (symmetrised gradient).
// Closed forms (q == 1): negative_ei.py:37-49, negative_pi.py:38-49, simple_regret.py:24-29,
//   confidence_bound.py:36-45 with utils.normal_pdf/cdf (src/utils/stats_ops.py:39-51).
#pragma once
#include "common.cuh"

#define MC_WARPS 4
#define MC_CHUNK 128

template <int Q>
struct McSmem {
  static constexpr int ZLD = (Q % 2 == 0) ? Q + 1 : Q;   // odd row stride: conflict-free lane-per-sample reads
  double zs[MC_CHUNK][ZLD];
  double wz[MC_CHUNK];
  double incs[MC_CHUNK];
  double L[MC_WARPS][Q][Q];
  double Lb[MC_WARPS][Q][Q];
  double P[MC_WARPS][Q][Q];
  double gw[MC_WARPS][MC_CHUNK];
  unsigned msk[MC_WARPS][MC_CHUNK];
  unsigned neg[MC_WARPS][MC_CHUNK];
  unsigned zer[MC_WARPS][MC_CHUNK];
};

template <int Q>
__global__ void __launch_bounds__(MC_WARPS * 32)
mc_acq_kernel(int S, int M, const double* __restrict__ mu, const double* __restrict__ Sigma,
              const double* __restrict__ z, const double* __restrict__ weights,
              const double* __restrict__ incumbents, AcqDev ap, double jitter, int need_grad,
              double* __restrict__ loss, double* __restrict__ mubar, double* __restrict__ Sigbar,
              double* __restrict__ chol_out) {
  extern __shared__ __align__(16) unsigned char mc_raw[];
  McSmem<Q>& sm = *reinterpret_cast<McSmem<Q>*>(mc_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * MC_WARPS + warp;
  const bool active = s < S;
  double(*L)[Q] = sm.L[warp];
  double(*Lb)[Q] = sm.Lb[warp];
  double(*P)[Q] = sm.P[warp];

  // ---- phase 0: L = chol(Sigma + jitter I), lower triangle of Sigma is read ------------
  if (active) {
    for (int t = lane; t < Q * Q; t += 32) {
      int i = t / Q, j = t - i * Q;
      double v = Sigma[(size_t)s * Q * Q + t];
      if (i == j) v += jitter;
      L[i][j] = (j <= i) ? v : 0.0;
    }
    __syncwarp();
    for (int k = 0; k < Q; ++k) {
      double lkk = sqrt(L[k][k]);
      __syncwarp();
      if (lane == k) L[k][k] = lkk;
      if (lane > k && lane < Q) L[lane][k] = L[lane][k] / lkk;
      __syncwarp();
      for (int t = lane; t < Q * Q; t += 32) {
        int i = t / Q, j = t - i * Q;
        if (j > k && i >= j) L[i][j] -= L[i][k] * L[j][k];
      }
      __syncwarp();
    }
    if (chol_out != nullptr)
      for (int t = lane; t < Q * Q; t += 32) chol_out[(size_t)s * Q * Q + t] = L[t / Q][t % Q];
  }

  constexpr int NPAIR = Q * (Q + 1) / 2;
  constexpr int NP = (NPAIR + 31) / 32;
  int pi_[NP], pj_[NP];
  double pacc[NP];
#pragma unroll
  for (int u = 0; u < NP; ++u) {
    int p = lane + 32 * u;
    pacc[u] = 0.0;
    if (p < NPAIR) {
      int i = 0;
      while ((i + 1) * (i + 2) / 2 <= p) ++i;
      pi_[u] = i;
      pj_[u] = p - i * (i + 1) / 2;
    } else {
      pi_[u] = -1;
      pj_[u] = 0;
    }
  }
  double mub = 0.0, loss_acc = 0.0;
  double mur[Q];
#pragma unroll
  for (int i = 0; i < Q; ++i) mur[i] = active ? mu[(size_t)s * Q + i] : 0.0;

  const bool is_cb = ap.acq == MAF_ACQ_CB;
  const bool use_max = is_cb && !ap.lower;

  for (int m0 = 0; m0 < M; m0 += MC_CHUNK) {
    __syncthreads();
    for (int t = threadIdx.x; t < MC_CHUNK * Q; t += MC_WARPS * 32) {
      int mm = t / Q, j = t - mm * Q;
      int m = m0 + mm;
      sm.zs[mm][j] = (m < M) ? z[(size_t)m * Q + j] : 0.0;
    }
    for (int t = threadIdx.x; t < MC_CHUNK; t += MC_WARPS * 32) {
      int m = m0 + t;
      sm.wz[t] = (m < M) ? (weights != nullptr ? weights[m] : 1.0 / (double)M) : 0.0;
      sm.incs[t] = (m < M && incumbents != nullptr) ? incumbents[m] : nan("");
    }
    __syncthreads();
    if (active) {
      // ---- phase 1: lanes own samples --------------------------------------------------
      for (int mm = lane; mm < MC_CHUNK; mm += 32) {
        if (m0 + mm >= M) {
          sm.gw[warp][mm] = 0.0;
          sm.msk[warp][mm] = 0u;
          sm.neg[warp][mm] = 0u;
          sm.zer[warp][mm] = 0u;
          continue;
        }
        double zr[Q], y[Q];
#pragma unroll
        for (int j = 0; j < Q; ++j) zr[j] = sm.zs[mm][j];
        unsigned negm = 0u, zerm = 0u;
#pragma unroll
        for (int i = 0; i < Q; ++i) {
          double e = 0.0;
#pragma unroll
          for (int j = 0; j <= i; ++j) e += L[i][j] * zr[j];
          if (is_cb) {
            double t = ap.cb_scale * e;
            double a = fabs(t);
            y[i] = ap.lower ? mur[i] - a : mur[i] + a;
            if (t < 0.0) negm |= 1u << i;
            if (t == 0.0) zerm |= 1u << i;
          } else {
            y[i] = mur[i] + e;
          }
        }
        double ext = y[0];
#pragma unroll
        for (int i = 1; i < Q; ++i) ext = use_max ? fmax(ext, y[i]) : fmin(ext, y[i]);
        unsigned mk = 0u;
#pragma unroll
        for (int i = 0; i < Q; ++i)
          if (y[i] == ext) mk |= 1u << i;
        const int cnt = __popc(mk);       // reduce_min/max gradient splits equally among ties
        const double inc = sm.incs[mm];
        bool pass = true;
        double ext2 = ext;
        if (inc == inc) {                 // tf.minimum/maximum: gradient to x where x <= y (>=)
          if (use_max) { pass = ext >= inc; ext2 = fmax(ext, inc); }
          else         { pass = ext <= inc; ext2 = fmin(ext, inc); }
        }
        const double w = sm.wz[mm];
        double contrib, dl;
        if (ap.acq == MAF_ACQ_EI) {
          double t = ap.level - ext2;
          contrib = -w * fmax(t, 0.0);
          dl = (t > 0.0) ? w : 0.0;
        } else if (ap.acq == MAF_ACQ_PI) {
          if (!ap.soft) {
            contrib = (ext2 < ap.level) ? -w : 0.0;
            dl = 0.0;
          } else {
            double a = ap.inv_temp * ((ap.level - ext2) - MAF_EPS_F64);
            double sg = 1.0 / (1.0 + exp(-a));
            contrib = -w * sg;
            dl = w * sg * (1.0 - sg) * ap.inv_temp;
          }
        } else {
          contrib = w * ext2;
          dl = w;
        }
        loss_acc += contrib;
        sm.gw[warp][mm] = (pass ? dl : 0.0) / (double)cnt;
        sm.msk[warp][mm] = mk;
        sm.neg[warp][mm] = negm;
        sm.zer[warp][mm] = zerm;
      }
    }
    __syncwarp();
    if (active && need_grad) {
      // ---- phase 2: lanes own (i,j) entries of Lbar = sum_m ebar_m z_m^T -------------------
      const double cbf = ap.lower ? -ap.cb_scale : ap.cb_scale;
#pragma unroll
      for (int u = 0; u < NP; ++u) {
        const int i = pi_[u], j = pj_[u];
        if (i >= 0) {
          double a = 0.0;
          for (int mm = 0; mm < MC_CHUNK; ++mm) {
            if ((sm.msk[warp][mm] >> i) & 1u) {
              double f = sm.gw[warp][mm];
              if (is_cb) {
                double sg = ((sm.zer[warp][mm] >> i) & 1u) ? 0.0 : (((sm.neg[warp][mm] >> i) & 1u) ? -1.0 : 1.0);
                f *= cbf * sg;
              }
              a += f * sm.zs[mm][j];
            }
          }
          pacc[u] += a;
        }
      }
      if (lane < Q) {
        double a = 0.0;
        for (int mm = 0; mm < MC_CHUNK; ++mm)
          if ((sm.msk[warp][mm] >> lane) & 1u) a += sm.gw[warp][mm];
        mub += a;
      }
    }
  }

  if (!active) return;
  loss_acc = warp_sum(loss_acc);
  if (lane == 0) loss[s] = loss_acc;
  if (!need_grad) return;

  // ---- phase 3: Cholesky reverse:  Sbar = sym(L^-T Phi(L^T Lbar) L^-1) ---------------------
  for (int t = lane; t < Q * Q; t += 32) Lb[t / Q][t % Q] = 0.0;
  __syncwarp();
#pragma unroll
  for (int u = 0; u < NP; ++u)
    if (pi_[u] >= 0) Lb[pi_[u]][pj_[u]] = pacc[u];
  if (lane < Q) mubar[(size_t)s * Q + lane] = mub;
  __syncwarp();
  for (int t = lane; t < Q * Q; t += 32) {
    int a = t / Q, b = t - a * Q;
    double v = 0.0;
    if (a >= b) {
      for (int i = a; i < Q; ++i) v += L[i][a] * Lb[i][b];
      if (a == b) v *= 0.5;
    }
    P[a][b] = v;
  }
  __syncwarp();
  if (lane < Q) {               // X = L^-T P, one column per lane
    const int b = lane;
    for (int a = Q - 1; a >= 0; --a) {
      double v = P[a][b];
      for (int i = a + 1; i < Q; ++i) v -= L[i][a] * P[i][b];
      P[a][b] = v / L[a][a];
    }
  }
  __syncwarp();
  if (lane < Q) {               // S = X L^-1, one row per lane
    const int a = lane;
    for (int j = Q - 1; j >= 0; --j) {
      double v = P[a][j];
      for (int i = j + 1; i < Q; ++i) v -= P[a][i] * L[i][j];
      P[a][j] = v / L[j][j];
    }
  }
  __syncwarp();
  for (int t = lane; t < Q * Q; t += 32) {
    int i = t / Q, j = t - i * Q;
    Sigbar[(size_t)s * Q * Q + t] = 0.5 * (P[i][j] + P[j][i]);
  }
}

// ---- closed forms (q == 1): one thread per restart -------------------------------------
__device__ __forceinline__ double norm_pdf(double x) { return 0.3989422804014327 * exp(-0.5 * x * x); }
__device__ __forceinline__ double norm_cdf(double x) { return 0.5 * erfc(-0.7071067811865475 * x); }

// loss and its derivatives w.r.t. (mean, variance) for one (candidate, level) pair
__device__ __forceinline__ void closed_form_eval(const AcqDev& ap, double level, double m, double v,
                                                 double& l, double& dm, double& dv) {
  if (ap.acq == MAF_ACQ_SR) {
    l = m; dm = 1.0; dv = 0.0;
  } else if (ap.acq == MAF_ACQ_CB) {
    double sb = sqrt(ap.beta * v);
    double sgn = ap.lower ? -1.0 : 1.0;
    l = m + sgn * sb;
    dm = 1.0;
    dv = sgn * ap.beta / (2.0 * sb);
  } else {
    double r = level - m;
    double sd = sqrt(v);
    double zz = (r == 0.0 && sd == 0.0) ? 0.0 : r / sd;     // utils.safe_div
    double pdf = norm_pdf(zz), cdf = norm_cdf(zz);
    if (ap.acq == MAF_ACQ_EI) {
      l = -(sd * pdf + r * cdf);
      dm = cdf;
      dv = -pdf / (2.0 * sd);
    } else {
      l = -cdf;
      dm = pdf / sd;
      dv = pdf * zz / (2.0 * v);
    }
  }
}

__global__ void closed_form_kernel(int S, const double* __restrict__ mu, const double* __restrict__ var,
                                   AcqDev ap, int need_grad, double* __restrict__ loss,
                                   double* __restrict__ mubar, double* __restrict__ varbar) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  double l, dm, dv;
  closed_form_eval(ap, ap.level, mu[s], var[s], l, dm, dv);
  loss[s] = l;
  if (need_grad) { mubar[s] = dm; varbar[s] = dv; }
}

// Fantasy-conditioned closed form (rank-4 observation sets of the incremental greedy driver,
// scripts/run_bayesopt_incremental.py:138-218; gaussian_process._predict_nd :195-206;
// bayesian_optimization.appraise_inputs :227-229): all F fantasies share X0 (one factor, one variance),
// differ in outputs, hence in gamma_f, prior mean m_f and level_f;  loss_s = mean_f loss(mu_sf, var_s, level_f).
// One warp per candidate; VG[s][f] = v_s . gamma_f.
__global__ void closed_form_fant_kernel(int S, int F, int Fpad, const double* __restrict__ VG,
                                        const double* __restrict__ fmeans, const double* __restrict__ flevels,
                                        const double* __restrict__ var, AcqDev ap, int use_flevels, int need_grad,
                                        double* __restrict__ loss, double* __restrict__ mubar /*[S][Fpad]*/,
                                        double* __restrict__ varbar, double* __restrict__ mu_out /*[S][Fpad] or null*/) {
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (s >= S) return;
  const double v = var[s];
  const double invF = 1.0 / (double)F;
  double lsum = 0.0, dvsum = 0.0;
  for (int f = lane; f < Fpad; f += 32) {
    double dm_out = 0.0;
    if (f < F) {
      const double m = fmeans[f] + VG[(size_t)s * Fpad + f];
      double l, dm, dv;
      closed_form_eval(ap, use_flevels ? flevels[f] : ap.level, m, v, l, dm, dv);
      lsum += l;
      dvsum += dv;
      dm_out = dm * invF;
      if (mu_out) mu_out[(size_t)s * Fpad + f] = m;
    }
    if (need_grad) mubar[(size_t)s * Fpad + f] = dm_out;
  }
  lsum = warp_sum(lsum);
  dvsum = warp_sum(dvsum);
  if (lane == 0) {
    loss[s] = lsum * invF;
    if (need_grad) varbar[s] = dvsum * invF;
  }
}

// Vbar[s][c] = G[s][c] - 2 varbar_s V[s][c]   (G = mubar Gamma), in place into V.
__global__ void fant_vbar_kernel(double* __restrict__ Vt, const double* __restrict__ G, const double* __restrict__ varbar,
                                 int S, int npad) {
  const int s = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S || c >= npad) return;
  const size_t o = (size_t)s * npad + c;
  Vt[o] = G[o] - 2.0 * varbar[s] * Vt[o];
}

// ---- Adam + clip (TF-1 ApplyAdam kernel form, then clip_by_value(x, 0, 1)) ---------------
// X is the full pool tensor [S][q][d]; only rows >= p are trainable; g is [S][q-p][d].
__global__ void adam_clip_kernel(double* __restrict__ X, const double* __restrict__ g,
                                 double* __restrict__ m, double* __restrict__ v, size_t ntrain,
                                 int q, int p, int d, double alpha, double beta1, double beta2, double eps,
                                 const double* __restrict__ coef = nullptr, const long long* __restrict__ kcur = nullptr) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntrain) return;
  if (coef) alpha = coef[*kcur];                            // replayed step (CUDA graph): this step's lr_t from the table
  const int qn = q - p;
  size_t s = t / ((size_t)qn * d);
  size_t rem = t - s * (size_t)qn * d;
  size_t xi = s * (size_t)q * d + (size_t)p * d + rem;
  double gi = g[t];
  double mi = m[t], vi = v[t];
  mi += (gi - mi) * (1.0 - beta1);
  vi += (gi * gi - vi) * (1.0 - beta2);
  double x = X[xi] - (mi * alpha) / (sqrt(vi) + eps);
  m[t] = mi;
  v[t] = vi;
  X[xi] = fmin(fmax(x, 0.0), 1.0);
}

// ---- replayed optimizer steps (one CUDA graph per session, maf_adam_steps) ------------------------------------
// The step index lives on the device: ctr[0] = index of the step being executed, ctr[1] = index of the next one.
// step_in (one block, first kernel of a step) advances it and stages the step's base samples at a fixed address;
// step_out files the step's losses in the trace.
__global__ void adam_step_in_kernel(long long* __restrict__ ctr, const double* __restrict__ zbase, size_t zsz,
                                    double* __restrict__ zcur) {
  const long long k = ctr[1];
  __syncthreads();                                          // every thread has read the index before it moves on
  for (size_t i = threadIdx.x; i < zsz; i += blockDim.x) zcur[i] = zbase[(size_t)k * zsz + i];
  if (threadIdx.x == 0) {
    ctr[0] = k;
    ctr[1] = k + 1;
  }
}

__global__ void adam_step_out_kernel(const long long* __restrict__ ctr, const double* __restrict__ loss, int S,
                                     double* __restrict__ trace) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < S) trace[(size_t)ctr[0] * S + i] = loss[i];
}

// ---- GradientDescent / Momentum + clip (bayesian_optimization.py:405-412; TF ApplyMomentum without Nesterov:
// accum = momentum * accum + g, x -= lr_t * accum; momentum == 0 is plain gradient descent) ----
__global__ void sgd_clip_kernel(double* __restrict__ X, const double* __restrict__ g, double* __restrict__ accum,
                                size_t ntrain, int q, int p, int d, double lr_t, double momentum,
                                const double* __restrict__ coef = nullptr, const long long* __restrict__ kcur = nullptr) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntrain) return;
  if (coef) lr_t = coef[*kcur];
  const int qn = q - p;
  size_t s = t / ((size_t)qn * d);
  size_t rem = t - s * (size_t)qn * d;
  size_t xi = s * (size_t)q * d + (size_t)p * d + rem;
  const double a = momentum * accum[t] + g[t];
  accum[t] = a;
  X[xi] = fmin(fmax(X[xi] - lr_t * a, 0.0), 1.0);
}

// ---- first-minimum argmin over the loss vector (np.argmin; NaNs never win) ----------------
__global__ void argmin_kernel(const double* __restrict__ loss, int S, double* out_val, long long* out_idx) {
  __shared__ double sv[32];
  __shared__ long long si[32];
  double bv = INFINITY;
  long long bi = -1;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    double v = loss[s];
    if (v == v && (bi < 0 || v < bv)) { bv = v; bi = s; }   // strided scan keeps the first index per thread
  }
  auto better = [](double v1, long long i1, double v2, long long i2) {
    if (i1 < 0) return false;
    if (i2 < 0) return true;
    return v1 < v2 || (v1 == v2 && i1 < i2);
  };
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, bv, o);
    long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    int nw = blockDim.x >> 5;
    bv = (lane < nw) ? sv[lane] : INFINITY;
    bi = (lane < nw) ? si[lane] : -1;
    for (int o = 16; o > 0; o >>= 1) {
      double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { *out_val = bv; *out_idx = bi; }
  }
}

// ---- selection record of one rank: rec = [loss*, global index, X*[q][d]] (index -1, loss +inf if nothing finite) ------
// Follows argmin_kernel on the same stream; what the single all-gather of the multi-GPU selection carries
// (np.argmin over all restarts, src/agents/bayesian_optimization.py:152, lowest index wins ties).
__global__ void best_record_kernel(const double* __restrict__ val, const long long* __restrict__ idx,
                                   const double* __restrict__ X, int qd, long long offset, double* __restrict__ rec) {
  const long long i = *idx;
  if (threadIdx.x == 0) {
    rec[0] = i >= 0 ? *val : INFINITY;
    rec[1] = i >= 0 ? (double)(offset + i) : -1.0;
  }
  for (int t = threadIdx.x; t < qd; t += blockDim.x) rec[2 + t] = i >= 0 ? X[(size_t)i * qd + t] : 0.0;
}

