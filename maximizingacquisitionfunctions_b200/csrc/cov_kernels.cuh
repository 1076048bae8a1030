// Covariance kernels: K00 build, fused ARD cross-covariance K(Xcand, Xtrain) forward,
// and its reverse-mode reduction to d loss / d Xcand.
// Reference: gaussian_process.covar_fn (src/models/gaussian_process.py:244-269),
// kernels.py:32-49, utils.pdist2 (src/utils/linalg_ops.py:102-109).
#pragma once
#include "common.cuh"

// X0 [n][d] (AoS, raw) -> X0s [d][npad] (SoA, multiplied by 1/l; pad = 0).
__global__ void scale_train_kernel(const double* __restrict__ X0, int n, int npad, ModelParams mp,
                                   double* __restrict__ X0s) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= npad) return;
  for (int c = 0; c < mp.d; ++c)
    X0s[(size_t)c * npad + k] = (k < n) ? X0[(size_t)k * mp.d + c] * mp.inv_ls[c] : 0.0;
}

// K00[i][j] = (a^2 kappa(r2_ij) + noise*delta_ij) + jitter*delta_ij ; identity on the pad.
// (noise_fn :272-280 then jitter_diagonal linalg_ops.py:112-128, same order of additions.)
__global__ void k00_kernel(const double* __restrict__ X0s, int n, int npad, ModelParams mp,
                           double* __restrict__ K00) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y;
  if (j >= npad) return;
  double v;
  if (i < n && j < n) {
    double r2 = 0.0;
    for (int c = 0; c < mp.d; ++c) {
      double df = X0s[(size_t)c * npad + i] - X0s[(size_t)c * npad + j];
      r2 += df * df;
    }
    v = mp.amp2 * kern_val(mp.kernel_id, r2);
    if (i == j) v = (v + mp.noise) + mp.jitter;
  } else {
    v = (i == j) ? 1.0 : 0.0;
  }
  K00[(size_t)i * npad + j] = v;
}

// ---------------------------------------------------------------------------
// Forward: K10[j][k] = a^2 kappa(|xs_j - x0s_k|^2), j < J candidate rows, k < n.
// One thread per training point k (its scaled coordinates live in registers for
// the whole tile), CROSS_TJ candidate rows per CTA broadcast from shared memory;
// stores are coalesced along k.  Rows j >= J and columns k >= n are written as 0 so
// the GEMM can run on whole 128x128 tiles.
// D > 0: compile-time dimension; D == 0: runtime d <= MAF_MAX_D.
// ---------------------------------------------------------------------------
#define CROSS_TJ 32
#define CROSS_TK 256

template <int D>
__global__ void __launch_bounds__(CROSS_TK)
cross_fwd_kernel(const double* __restrict__ X, int J, const double* __restrict__ X0s, int n, int npad,
                 ModelParams mp, double* __restrict__ K10) {
  constexpr int DD = D > 0 ? D : MAF_MAX_D;
  const int d = D > 0 ? D : mp.d;
  __shared__ double xs[CROSS_TJ][DD];
  const int k = blockIdx.x * CROSS_TK + threadIdx.x;
  const int j0 = blockIdx.y * CROSS_TJ;
  for (int t = threadIdx.x; t < CROSS_TJ * d; t += CROSS_TK) {
    int jj = t / d, c = t - jj * d;
    int j = j0 + jj;
    xs[jj][c] = (j < J) ? X[(size_t)j * d + c] * mp.inv_ls[c] : 0.0;
  }
  __syncthreads();
  if (k >= npad) return;          // npad is a multiple of 128, the k-tile is 256 wide
  double x0[DD];
#pragma unroll
  for (int c = 0; c < DD; ++c) x0[c] = (c < d) ? X0s[(size_t)c * npad + k] : 0.0;
  const bool kvalid = k < n;
#pragma unroll 4
  for (int jj = 0; jj < CROSS_TJ; ++jj) {
    double r2 = 0.0;
#pragma unroll
    for (int c = 0; c < DD; ++c) {
      if (c < d) {
        double df = xs[jj][c] - x0[c];
        r2 += df * df;
      }
    }
    double v = (kvalid && (j0 + jj) < J) ? mp.amp2 * kern_val(mp.kernel_id, r2) : 0.0;
    K10[(size_t)(j0 + jj) * npad + k] = v;
  }
}

// ---------------------------------------------------------------------------
// Backward: grad[s][i-p][c] = 2 a^2 / l_c * ( sum_k Kbar10[j][k] kappa'(r2_jk) (xs_jc - x0s_kc)
//                       + sum_{i' != i} (Sbar_ii' + Sbar_i'i) kappa'(r2_ii') (xs_ic - xs_i'c) )
// One warp per candidate row (block = one restart, q warps); lanes stride over k with
// coalesced reads of Kbar10 and of the SoA training set; warp-shuffle reduction.
// ---------------------------------------------------------------------------
#ifndef CROSS_BWD_MINB
#define CROSS_BWD_MINB 2
#endif
#ifndef CROSS_BWD_UNROLL
#define CROSS_BWD_UNROLL 2
#endif
template <int D, int KID, bool SEL = false>
__global__ void __launch_bounds__(512, (D > 0 && D <= 8) ? CROSS_BWD_MINB : 1)      // MINB = 2: <= 64 registers, 32 warps per SM
cross_bwd_kernel(const double* __restrict__ X, int S, int q, int p,
                                 const double* __restrict__ X0s, int n, int npad, ModelParams mp,
                                 const double* __restrict__ Kbar /*[S*q][npad]*/,
                                 const double* __restrict__ Sigbar /*[S][q][q]*/,
                                 double* __restrict__ grad /*[S][q-p][d]*/,
                                 const int* __restrict__ sel = nullptr, const int* __restrict__ sel_count = nullptr,
                                 int kq = 0) {
  // One warp per trainable row (block = 32 (q - p) threads).  Kbar holds kq rows per pool: kq = q (all rows, pending
  // ones included and ignored; kq = 0 means q) or kq = q - p (pending points contracted once elsewhere: trainable rows
  // only); pool row i lives at row blockIdx.x kq + i - (q - kq).
  // SEL (second pass of the reverse level switch): CTA r serves restart sel[r] from compact Kbar rows.
  // (With the mean computed as m + K10 . alpha the mean part of d loss / d K10, mubar_j alpha_k, is added to Kbar by the
  // reverse contraction's epilogue in FP64: this kernel always sees the complete Kbar.)
  constexpr int DD = D > 0 ? D : MAF_MAX_D;
  const int d = D > 0 ? D : mp.d;
  __shared__ double etab_sh[MAF_ETAB_WORDS];
  maf_fill_etab(etab_sh, threadIdx.x, blockDim.x);
  const double* etab = etab_sh + MAF_ETAB_LANE(threadIdx.x);
  __syncthreads();
  int s = blockIdx.x;
  if (kq == 0) kq = q;
  size_t krow0 = (size_t)s * kq;
  if (SEL) {
    if ((int)blockIdx.x >= __ldg(sel_count)) return;
    s = __ldg(sel + blockIdx.x);
  }
  const int i = p + (threadIdx.x >> 5);                      // warp w <-> trainable row p + w
  const int lane = threadIdx.x & 31;
  if (i >= q) return;
  const size_t j = (size_t)s * q + i;
  double xi[DD], acc[DD];
#pragma unroll
  for (int c = 0; c < DD; ++c) {
    xi[c] = (c < d) ? X[j * d + c] * mp.inv_ls[c] : 0.0;
    acc[c] = 0.0;
  }
  const double* kb = Kbar + (krow0 + i - (q - kq)) * npad;
  constexpr int UNR = CROSS_BWD_UNROLL;
#pragma unroll UNR
  for (int k = lane; k < n; k += 32) {
    double x0[DD];
    double r2 = 0.0;
#pragma unroll
    for (int c = 0; c < DD; ++c) {
      if (c < d) {
        x0[c] = X0s[(size_t)c * npad + k];
        double df = xi[c] - x0[c];
        r2 += df * df;
      }
    }
    double g = __ldcs(kb + k) * kern_dr2_fast<KID>(r2, etab);   // streamed once: keep the training set in L1 instead
#pragma unroll
    for (int c = 0; c < DD; ++c)
      if (c < d) acc[c] += g * (xi[c] - x0[c]);
  }
  // self-covariance term (K11 = a^2 kappa(r2(X_s, X_s)); diagonal has zero gradient)
  if (lane < q && lane != i) {
    const double* xo = X + ((size_t)s * q + lane) * d;
    double xj[DD];
    double r2 = 0.0;
#pragma unroll
    for (int c = 0; c < DD; ++c) {
      if (c < d) {
        xj[c] = xo[c] * mp.inv_ls[c];
        double df = xi[c] - xj[c];
        r2 += df * df;
      }
    }
    const double* sb = Sigbar + (size_t)s * q * q;
    double g = (sb[i * q + lane] + sb[lane * q + i]) * kern_dr2_fast<KID>(r2, etab);
#pragma unroll
    for (int c = 0; c < DD; ++c)
      if (c < d) acc[c] += g * (xi[c] - xj[c]);
  }
#pragma unroll
  for (int c = 0; c < DD; ++c) {
    if (c < d) {
      double v = warp_sum(acc[c]);
      if (lane == 0)
        grad[((size_t)s * (q - p) + (i - p)) * d + c] = 2.0 * mp.amp2 * mp.inv_ls[c] * v;
    }
  }
}

// ---------------------------------------------------------------------------
// Backward on stored kernel derivatives: the forward covariance kernel left G[j][k] = kappa'(r2_jk) behind
// (cross_fwd_planes_kernel, gout), so the element costs two streamed loads and 13 FP64 instructions instead of a
// square root and an exponential (40): the kernel moves from the FP64 pipe to the HBM stream (Kbar + G, 16 B per pair).
// One CTA per restart, R trainable rows per warp: a training point's coordinates are loaded once per R rows.
// Row layout of Kbar as in cross_bwd_kernel (kq rows per pool, compact positions in the second pass of the level
// switch); G always lives at the restart's own rows.
// ---------------------------------------------------------------------------
#ifndef CROSS_BWD_G_R
#define CROSS_BWD_G_R 2
#endif
#ifndef CROSS_BWD_G_MINB
#define CROSS_BWD_G_MINB 2
#endif
#ifndef CROSS_BWD_G_UNROLL
#define CROSS_BWD_G_UNROLL 1
#endif
#ifndef CROSS_BWD_G_VEC
#define CROSS_BWD_G_VEC 2
#endif
template <int D, int KID, int R, bool SEL = false>
__global__ void __launch_bounds__(32 * ((MAF_MAX_Q + R - 1) / R), CROSS_BWD_G_MINB)
cross_bwd_g_kernel(const double* __restrict__ X, int S, int q, int p, const double* __restrict__ X0s, int n, int npad,
                   ModelParams mp, const double* __restrict__ Kbar, const double* __restrict__ G,
                   const double* __restrict__ Sigbar, double* __restrict__ grad, const int* __restrict__ sel,
                   const int* __restrict__ sel_count, int kq) {
  static_assert(D > 0, "compile-time dimension only");
  __shared__ double etab_sh[MAF_ETAB_WORDS];
  maf_fill_etab(etab_sh, threadIdx.x, blockDim.x);
  const double* etab = etab_sh + MAF_ETAB_LANE(threadIdx.x);
  __syncthreads();
  int s = blockIdx.x;
  if (kq == 0) kq = q;
  const size_t krow0 = (size_t)s * kq;                         // rows of Kbar: position of this CTA
  if (SEL) {
    if ((int)blockIdx.x >= __ldg(sel_count)) return;
    s = __ldg(sel + blockIdx.x);
  }
  const size_t grow0 = (size_t)s * kq;                         // rows of G: the restart itself
  const int lane = threadIdx.x & 31;
  const int i0 = p + (threadIdx.x >> 5) * R;                   // first row of this warp
  if (i0 >= q) return;
  double xi[R][D], acc[R][D];
  const double* kb[R];
  const double* gb[R];
#pragma unroll
  for (int rr = 0; rr < R; ++rr) {
    const int i = min(i0 + rr, q - 1);                         // a missing last row repeats the previous one (never stored)
    const size_t j = (size_t)s * q + i;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      xi[rr][c] = X[j * D + c] * mp.inv_ls[c];
      acc[rr][c] = 0.0;
    }
    kb[rr] = Kbar + (krow0 + i - (q - kq)) * npad;
    gb[rr] = G + (grow0 + i - (q - kq)) * npad;
  }
  constexpr int UNRG = CROSS_BWD_G_UNROLL;
#if CROSS_BWD_G_VEC == 2
  // two training points per lane and 16-byte loads (rows are 1 KB aligned: npad is a multiple of 128); the odd last
  // column of an odd n is masked, pad columns are never read past n rounded up to 2
  const int n2 = n & ~1;
#pragma unroll UNRG
  for (int k = 2 * lane; k < n2; k += 64) {
    double2 x0[D];
#pragma unroll
    for (int c = 0; c < D; ++c) x0[c] = *reinterpret_cast<const double2*>(X0s + (size_t)c * npad + k);
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      const double2 kv = __ldcs(reinterpret_cast<const double2*>(kb[rr] + k));
      const double2 gv = __ldcs(reinterpret_cast<const double2*>(gb[rr] + k));
      const double g0 = kv.x * gv.x, g1 = kv.y * gv.y;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        acc[rr][c] = fma(g0, xi[rr][c] - x0[c].x, acc[rr][c]);
        acc[rr][c] = fma(g1, xi[rr][c] - x0[c].y, acc[rr][c]);
      }
    }
  }
  if ((n & 1) && lane == 0) {
    const int k = n - 1;
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      const double g = kb[rr][k] * gb[rr][k];
#pragma unroll
      for (int c = 0; c < D; ++c) acc[rr][c] = fma(g, xi[rr][c] - X0s[(size_t)c * npad + k], acc[rr][c]);
    }
  }
#else
#pragma unroll UNRG
  for (int k = lane; k < n; k += 32) {
    double x0[D];
#pragma unroll
    for (int c = 0; c < D; ++c) x0[c] = X0s[(size_t)c * npad + k];
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      const double g = __ldcs(kb[rr] + k) * __ldcs(gb[rr] + k);
#pragma unroll
      for (int c = 0; c < D; ++c) acc[rr][c] = fma(g, xi[rr][c] - x0[c], acc[rr][c]);
    }
  }
#endif
#pragma unroll
  for (int rr = 0; rr < R; ++rr) {
    const int i = i0 + rr;
    if (i < q) {                                               // warp-uniform
      // self-covariance term (K11 = a^2 kappa(r2(X_s, X_s)); diagonal has zero gradient)
      if (lane < q && lane != i) {
        const double* xo = X + ((size_t)s * q + lane) * D;
        double xj[D];
        double r2 = 0.0;
#pragma unroll
        for (int c = 0; c < D; ++c) {
          xj[c] = xo[c] * mp.inv_ls[c];
          const double df = xi[rr][c] - xj[c];
          r2 += df * df;
        }
        const double* sb = Sigbar + (size_t)s * q * q;
        const double g = (sb[i * q + lane] + sb[lane * q + i]) * kern_dr2_fast<KID>(r2, etab);
#pragma unroll
        for (int c = 0; c < D; ++c) acc[rr][c] += g * (xi[rr][c] - xj[c]);
      }
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const double v = warp_sum(acc[rr][c]);
        if (lane == 0) grad[((size_t)s * (q - p) + (i - p)) * D + c] = 2.0 * mp.amp2 * mp.inv_ls[c] * v;
      }
    }
  }
}

// First pass of the reverse level switch, after cross_bwd_kernel: the truncation error of the cheaper level on row j is
// err x kscale[j] (calibrated per unit row scale of the Vbar planes); restarts where it exceeds tol x max |gradient| are
// listed for the second pass.  One thread per restart.
__global__ void lsw_flag_bwd_kernel(const double* __restrict__ grad, const double* __restrict__ kscale, int S, int q, int p,
                                    int d, LevelSwitch ls, int kq) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  double mx = 0.0, er = 0.0;
  const double* g = grad + (size_t)s * (q - p) * d;
  for (int t = 0; t < (q - p) * d; ++t) mx = fmax(mx, fabs(g[t]));
  for (int i = p; i < q; ++i) er = fmax(er, ls.err * kscale[(size_t)s * kq + i - (q - kq)]);
  if (!(er <= ls.tol * mx)) {
    const int slot = atomicAdd(ls.flag_count, 1);
    ls.flag_sel[slot] = s;
  }
}
