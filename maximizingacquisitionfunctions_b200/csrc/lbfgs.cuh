// Batched, box-constrained L-BFGS on the device: one independent optimiser per restart, all restarts advanced in lock
// step (one value+gradient evaluation of every restart's trial point per iteration, then this kernel).
//
// Replaces the SciPy loop behind bayesian_optimization._optimize_inputs_scipy (src/agents/bayesian_optimization.py:
// 485-555: ScipyOptimizerInterface(method='L-BFGS-B', var_to_bounds={inputs_new: (0, 1)}) on mean_s loss_s), the default
// subroutine for pools of one point (:311-313), of suggest_answers (:938-1048) and so of the whole incremental greedy
// construction.  The reference optimises the mean of the per-restart losses jointly; the restarts only share a line
// search there (the objective is separable), so each restart gets its own L-BFGS here (value / gradient
// parity is the contract, the trajectory is not) and nothing leaves the GPU between iterations.
//
// Per restart: limited-memory BFGS two-loop recursion (history of LB_M pairs) on the free variables, projected
// backtracking line search on [0, 1]^D with the Armijo test along the projected path, curvature pairs skipped unless
// s.y > 1e-10 |s| |y|.  A restart stops when its projected gradient is below pgtol, when an accepted step improves the
// loss by less than ftol * max(|f|, |f'|, 1) (SciPy's factr test), or when the step has shrunk to nothing.
// Where the loss has no positive curvature along the step (the far tail of -EI is concave: s.y <= 0, no pair can be
// stored) the restart is in steepest-descent mode and the trial DISTANCE grows eightfold after every accepted step
// (backtracking still halves it on failure).  Starts in the flat tail of an improvement-type loss (|loss| ~ 1e-40 and a
// gradient to match) stop at once under these absolute tests -- as they do, in effect, inside the reference's joint
// problem, whose `best S (iterate, restart) pairs over the whole trajectory` rule (:541-550) then fills their places with
// further iterates of the restarts that did move; sample_starts draws starts in proportion to the marginal loss, so such
// starts are rare.
#pragma once
#include "common.cuh"

#define LB_M 10          // history pairs (SciPy's default maxcor)
#define LB_MAX_D 256     // (q - p) * d <= MAF_MAX_Q * MAF_MAX_D

struct LbfgsState {
  double* x;        // [S][D]   accepted iterate (trainable rows only)
  double* f;        // [S]
  double* g;        // [S][D]
  double* dir;      // [S][D]   search direction of the running line search
  double* step;     // [S]      current step length
  double* gd;       // [S]      steepest-descent mode: length of the last accepted move
  double* hs;       // [S][LB_M][D]
  double* hy;       // [S][LB_M][D]
  double* rho;      // [S][LB_M]
  int* hcount;      // [S]      pairs stored (<= LB_M)
  int* hpos;        // [S]      next slot (ring)
  int* status;      // [S]      0 running, 1 converged (gradient), 2 converged (loss), 3 line search exhausted
  int* iters;       // [S]      accepted steps
  int* active;      // [1]      restarts still running (refreshed by every step)
};

// Trial pools Xt[S][q][d] (pending rows untouched); ft[S], gt[S][D] = loss and gradient at the trial point.
// first != 0: the trial point IS the start (no line search yet).
__global__ void lbfgs_step_kernel(LbfgsState st, int S, int q, int p, int d, double* __restrict__ Xt,
                                  const double* __restrict__ ft, const double* __restrict__ gt, int first, double ftol,
                                  double pgtol) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const int D = (q - p) * d;
  double* x = st.x + (size_t)s * D;
  double* g = st.g + (size_t)s * D;
  double* dir = st.dir + (size_t)s * D;
  double* xt = Xt + ((size_t)s * q + p) * d;               // trainable rows are contiguous
  const double* gn = gt + (size_t)s * D;
  double* HS = st.hs + (size_t)s * LB_M * D;
  double* HY = st.hy + (size_t)s * LB_M * D;
  double* rho = st.rho + (size_t)s * LB_M;
  if (first) {
    st.hcount[s] = 0;
    st.hpos[s] = 0;
    st.status[s] = 0;
    st.iters[s] = 0;
  } else if (st.status[s] != 0) {
    return;                                                // finished: the trial pool already holds the final iterate
  }
  const double fn = ft[s];
  bool accepted = first != 0;
  if (!first) {
    // Armijo along the projected path: f(x_t) <= f(x) + c1 g . (x_t - x)
    double gdx = 0.0;
    for (int t = 0; t < D; ++t) gdx += g[t] * (xt[t] - x[t]);
    accepted = (fn == fn) && fn <= st.f[s] + 1e-4 * gdx;
    if (!accepted) {
      const double a = st.step[s] * 0.5;
      st.step[s] = a;
      bool moved = false;
      for (int t = 0; t < D; ++t) {
        const double v = fmin(fmax(x[t] + a * dir[t], 0.0), 1.0);
        moved = moved || v != x[t];
        xt[t] = v;
      }
      if (a < 1e-12 || !moved) {                           // nothing left to try: stay at the accepted iterate
        for (int t = 0; t < D; ++t) xt[t] = x[t];
        st.status[s] = 3;
        atomicSub(st.active, 1);
      }
      return;
    }
    // accepted: curvature pair, convergence on the loss
    double sy = 0.0, ss = 0.0, yy = 0.0;
    const int slot = st.hpos[s];
    for (int t = 0; t < D; ++t) {
      const double sv = xt[t] - x[t], yv = gn[t] - g[t];
      HS[(size_t)slot * D + t] = sv;
      HY[(size_t)slot * D + t] = yv;
      sy += sv * yv;
      ss += sv * sv;
      yy += yv * yv;
    }
    if (sy > 1e-10 * sqrt(ss * yy) && sy > 0.0) {
      rho[slot] = 1.0 / sy;
      st.hpos[s] = (slot + 1) % LB_M;
      st.hcount[s] = min(st.hcount[s] + 1, LB_M);
    }
    st.gd[s] = sqrt(ss);
    st.iters[s] += 1;
    const double fo = st.f[s];
    if (fo - fn <= ftol * fmax(fmax(fabs(fo), fabs(fn)), 1.0)) st.status[s] = 2;
  }
  // new accepted iterate
  st.f[s] = fn;
  double pgmax = 0.0;
  for (int t = 0; t < D; ++t) {
    x[t] = xt[t];
    g[t] = gn[t];
    const double pg = x[t] - fmin(fmax(x[t] - g[t], 0.0), 1.0);     // projected gradient
    pgmax = fmax(pgmax, fabs(pg));
  }
  if (st.status[s] == 0 && pgmax <= pgtol) st.status[s] = 1;
  if (st.status[s] != 0) {
    atomicSub(st.active, 1);
    return;
  }
  // direction: two-loop recursion on the free variables (bound-active components with the gradient pushing outward
  // are held: their direction is zero)
  double al[LB_M];
  const int hc = st.hcount[s], hp = st.hpos[s];
  for (int t = 0; t < D; ++t) {
    const bool held = (x[t] <= 0.0 && g[t] > 0.0) || (x[t] >= 1.0 && g[t] < 0.0);
    dir[t] = held ? 0.0 : g[t];
  }
  for (int u = 0; u < hc; ++u) {
    const int k = (hp - 1 - u + 2 * LB_M) % LB_M;
    double a = 0.0;
    for (int t = 0; t < D; ++t) a += HS[(size_t)k * D + t] * dir[t];
    a *= rho[k];
    al[u] = a;
    for (int t = 0; t < D; ++t) dir[t] -= a * HY[(size_t)k * D + t];
  }
  if (hc > 0) {
    const int k = (hp - 1 + LB_M) % LB_M;
    double yy = 0.0;
    for (int t = 0; t < D; ++t) yy += HY[(size_t)k * D + t] * HY[(size_t)k * D + t];
    const double gam = 1.0 / (rho[k] * yy);               // s.y / y.y
    for (int t = 0; t < D; ++t) dir[t] *= gam;
  }
  for (int u = hc - 1; u >= 0; --u) {
    const int k = (hp - 1 - u + 2 * LB_M) % LB_M;
    double b = 0.0;
    for (int t = 0; t < D; ++t) b += HY[(size_t)k * D + t] * dir[t];
    b *= rho[k];
    for (int t = 0; t < D; ++t) dir[t] += (al[u] - b) * HS[(size_t)k * D + t];
  }
  double gd = 0.0, gnorm = 0.0;
  for (int t = 0; t < D; ++t) {
    const bool held = (x[t] <= 0.0 && g[t] > 0.0) || (x[t] >= 1.0 && g[t] < 0.0);
    dir[t] = held ? 0.0 : -dir[t];
    gd += g[t] * dir[t];
    gnorm += held ? 0.0 : g[t] * g[t];
  }
  if (!(gd < 0.0)) {                                       // not a descent direction (stale curvature): steepest descent
    st.hcount[s] = 0;
    gd = 0.0;
    for (int t = 0; t < D; ++t) {
      const bool held = (x[t] <= 0.0 && g[t] > 0.0) || (x[t] >= 1.0 && g[t] < 0.0);
      dir[t] = held ? 0.0 : -g[t];
      gd += g[t] * dir[t];
    }
  }
  // no curvature model (first step, or negative curvature so far): steepest descent over a trial distance L -- first
  // min(1, |g|) (SciPy's first step), then 8 x the last accepted move, at most the diagonal of the box; with a model:
  // the quasi-Newton unit step
  double a = 1.0;
  if (st.hcount[s] == 0) {
    const double gn2 = sqrt(fmax(gnorm, 1e-300));
    const double L = first ? fmin(1.0, gn2) : fmin(8.0 * st.gd[s], sqrt((double)D));
    a = L / gn2;
  }
  st.step[s] = a;
  bool moved = false;
  for (int t = 0; t < D; ++t) {
    const double v = fmin(fmax(x[t] + a * dir[t], 0.0), 1.0);
    moved = moved || v != x[t];
    xt[t] = v;
  }
  if (!moved) {
    st.status[s] = 1;
    atomicSub(st.active, 1);
  }
}
