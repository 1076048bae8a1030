// Posterior mean / q x q covariance from V^T = K10 L0^-T, and the reverse-mode
// propagation (mubar, Sigmabar) -> Vbar.
// Reference: gaussian_process._predict_nd (src/models/gaussian_process.py:172-223):
//   Means = m + Beta resid,  Covs = k(X1,X1) - Beta K01   (Beta = (K00^-1 K01)^T).
// With V = L0^-1 K01 and gamma = L0^-1 resid this is  mu = m + V^T gamma,
// Sigma = K11 - V^T V  (identical maths, better conditioned: sqrt(cond(K00))).
#pragma once
#include "common.cuh"

#define POST_TK 256
#define POST_THREADS 256
// register-accumulating kernel (Q <= 8): 64 threads per restart measured best at n = 4096 (256: 4.5, 128: 3.8, 96: 3.9,
// 64: 3.5 ms per step): more CTAs per SM keep more loads in flight and the block reduction of the 44 sums gets shorter
#ifndef POST_REG_THREADS
#define POST_REG_THREADS 64
#endif

// One CTA per restart.  The q x POST_TK slab of V^T is staged in shared memory once and
// every warp accumulates the Gram rows it owns (warp w -> rows w and w+8).
template <int Q, bool DED = false>
__global__ void __launch_bounds__(POST_THREADS)
posterior_kernel(const double* __restrict__ Vt, int npad, const double* __restrict__ gamma,
                 const double* __restrict__ X, ModelParams mp, double mean,
                 double* __restrict__ mu, double* __restrict__ Sig,
                 const double* __restrict__ mu_part = nullptr, int nparts = 0, int Jpad = 0, LevelSwitch ls = LevelSwitch{},
                 int p_rt = 0, const double* __restrict__ Vp = nullptr, const double* __restrict__ mu_part_p = nullptr,
                 int Jpad_p = 0) {
  const int p = DED ? p_rt : 0;                  // DED = false: no shared block, the row selects below fold away
  // p > 0: rows 0..p-1 of every pool are pending points shared by all restarts; their V^T rows come from Vp[p][npad]
  // (contracted once), their mean partial sums from mu_part_p, and Vt holds only the q - p trainable rows per restart.
  __shared__ double tile[Q][POST_TK];
  __shared__ double gs[POST_TK];
  __shared__ double G[Q][Q];
  __shared__ double mus[Q];
  __shared__ double sabs[Q * Q];
  const int r = blockIdx.x;                      // row group of V^T (compact position in the second pass)
  int s = r;                                     // restart the result belongs to
  if (ls.sel != nullptr) {
    if (r >= __ldg(ls.sel_count)) return;
    s = __ldg(ls.sel + r);
  }
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qn = Q - p;
  const double* V = Vt + (size_t)r * qn * npad - (size_t)p * npad;      // trainable row i at V + i npad
  const int i0 = warp, i1 = warp + 8;
  double acc0[Q], acc1[Q], m0 = 0.0, m1 = 0.0;
#pragma unroll
  for (int j = 0; j < Q; ++j) acc0[j] = acc1[j] = 0.0;

  for (int k0 = 0; k0 < npad; k0 += POST_TK) {
    __syncthreads();
    const int k = k0 + tid;
#pragma unroll
    for (int i = 0; i < Q; ++i) tile[i][tid] = (k < npad) ? ((DED && i < p) ? Vp : V)[(size_t)i * npad + k] : 0.0;
    gs[tid] = (k < npad && gamma != nullptr) ? gamma[k] : 0.0;
    __syncthreads();
    for (int kk = lane; kk < POST_TK; kk += 32) {
      if (i0 < Q) {
        double vi = tile[i0][kk];
        m0 += vi * gs[kk];
#pragma unroll
        for (int j = 0; j < Q; ++j)
          if (j <= i0) acc0[j] += vi * tile[j][kk];
      }
      if (Q > 8 && i1 < Q) {
        double vi = tile[i1][kk];
        m1 += vi * gs[kk];
#pragma unroll
        for (int j = 0; j < Q; ++j)
          if (j <= i1) acc1[j] += vi * tile[j][kk];
      }
    }
  }
  if (i0 < Q) {
    m0 = warp_sum(m0);
#pragma unroll
    for (int j = 0; j < Q; ++j) {
      double v = warp_sum(acc0[j]);
      if (lane == 0 && j <= i0) G[i0][j] = v;
    }
    if (lane == 0) mus[i0] = m0;
  }
  if (Q > 8 && i1 < Q) {
    m1 = warp_sum(m1);
#pragma unroll
    for (int j = 0; j < Q; ++j) {
      double v = warp_sum(acc1[j]);
      if (lane == 0 && j <= i1) G[i1][j] = v;
    }
    if (lane == 0) mus[i1] = m1;
  }
  __syncthreads();
  const double* xs = X + (size_t)s * Q * mp.d;
  for (int t = tid; t < Q * Q; t += POST_THREADS) {
    int i = t / Q, j = t - i * Q;
    double r2 = 0.0;
    if (i != j) {
      for (int c = 0; c < mp.d; ++c) {
        double df = xs[i * mp.d + c] * mp.inv_ls[c] - xs[j * mp.d + c] * mp.inv_ls[c];
        r2 += df * df;
      }
    }
    double g = (i >= j) ? G[i][j] : G[j][i];
    const double sv = mp.amp2 * kern_val(mp.kernel_id, r2) - g;
    Sig[(size_t)s * Q * Q + t] = sv;
    sabs[t] = fabs(sv);
  }
  if (tid < Q) {
    double m = mus[tid];
    if (mu_part != nullptr) {                         // mean from K10 . alpha (cross_fwd_planes_kernel), column blocks in order
      m = 0.0;
      for (int b = 0; b < nparts; ++b)
        m += (DED && tid < p) ? mu_part_p[(size_t)b * Jpad_p + tid] : mu_part[(size_t)b * Jpad + (size_t)r * qn + tid - p];
    }
    mu[(size_t)s * Q + tid] = mean + m;
  }
  if (ls.flag_count != nullptr) {                     // level switch: does this restart tolerate the cheaper level?
    __syncthreads();
    if (tid == 0) {
      double mx = 0.0;
      for (int t = 0; t < Q * Q; ++t) mx = fmax(mx, sabs[t]);
      int slot = -1;
      if (!(ls.err <= ls.tol * mx)) {
        slot = atomicAdd(ls.flag_count, 1);
        ls.flag_sel[slot] = s;
      }
      ls.flag_slot[s] = slot;
    }
  }
}

// Same result for Q <= 8 without the shared-memory slab (ncu: posterior_kernel runs at 94 % of the L1/shared
// pipe): one thread owns a strided set of training columns, keeps its Q values of V^T in registers and accumulates the
// Q (Q+1)/2 Gram entries and the Q mean terms there; one block reduction at the end.  Streams V^T once (HBM-bound).
template <int Q, bool DED = false>
__global__ void __launch_bounds__(POST_REG_THREADS)
posterior_reg_kernel(const double* __restrict__ Vt, int npad, const double* __restrict__ gamma,
                     const double* __restrict__ X, ModelParams mp, double mean,
                     double* __restrict__ mu, double* __restrict__ Sig,
                     const double* __restrict__ mu_part = nullptr, int nparts = 0, int Jpad = 0, LevelSwitch ls = LevelSwitch{},
                     int p_rt = 0, const double* __restrict__ Vp = nullptr, const double* __restrict__ mu_part_p = nullptr,
                     int Jpad_p = 0) {
  const int p = DED ? p_rt : 0;                  // DED = false: no shared block, the row selects below fold away
  constexpr int NG = Q * (Q + 1) / 2;
  __shared__ double red[POST_REG_THREADS / 32][NG + Q];
  __shared__ double G[Q][Q];
  __shared__ double mus[Q];
  __shared__ double sabs[Q * Q];
  const int r = blockIdx.x;                      // row group of V^T (compact position in the second pass)
  int s = r;                                     // restart the result belongs to
  if (ls.sel != nullptr) {
    if (r >= __ldg(ls.sel_count)) return;
    s = __ldg(ls.sel + r);
  }
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qn = Q - p;
  const double* V = Vt + (size_t)r * qn * npad - (size_t)p * npad;      // trainable row i at V + i npad
  double acc[NG], m[Q];
#pragma unroll
  for (int t = 0; t < NG; ++t) acc[t] = 0.0;
#pragma unroll
  for (int i = 0; i < Q; ++i) m[i] = 0.0;
  for (int k = tid; k < npad; k += POST_REG_THREADS) {
    double v[Q];
#pragma unroll
    for (int i = 0; i < Q; ++i) v[i] = ((DED && i < p) ? Vp : V)[(size_t)i * npad + k];
    const double g = gamma != nullptr ? gamma[k] : 0.0;      // nullptr: the mean comes from mu_part
#pragma unroll
    for (int i = 0; i < Q; ++i) {
      m[i] += v[i] * g;
#pragma unroll
      for (int j = 0; j <= i; ++j) acc[i * (i + 1) / 2 + j] += v[i] * v[j];
    }
  }
#pragma unroll
  for (int t = 0; t < NG; ++t) {
    const double v = warp_sum(acc[t]);
    if (lane == 0) red[warp][t] = v;
  }
#pragma unroll
  for (int i = 0; i < Q; ++i) {
    const double v = warp_sum(m[i]);
    if (lane == 0) red[warp][NG + i] = v;
  }
  __syncthreads();
  for (int t = tid; t < NG + Q; t += POST_REG_THREADS) {
    double v = 0.0;
    for (int w = 0; w < POST_REG_THREADS / 32; ++w) v += red[w][t];
    if (t < NG) {
      int i = 0;
      while ((i + 1) * (i + 2) / 2 <= t) ++i;
      G[i][t - i * (i + 1) / 2] = v;
    } else {
      mus[t - NG] = v;
    }
  }
  __syncthreads();
  const double* xs = X + (size_t)s * Q * mp.d;
  for (int t = tid; t < Q * Q; t += POST_REG_THREADS) {
    int i = t / Q, j = t - i * Q;
    double r2 = 0.0;
    if (i != j) {
      for (int c = 0; c < mp.d; ++c) {
        double df = xs[i * mp.d + c] * mp.inv_ls[c] - xs[j * mp.d + c] * mp.inv_ls[c];
        r2 += df * df;
      }
    }
    double g = (i >= j) ? G[i][j] : G[j][i];
    const double sv = mp.amp2 * kern_val(mp.kernel_id, r2) - g;
    Sig[(size_t)s * Q * Q + t] = sv;
    sabs[t] = fabs(sv);
  }
  if (tid < Q) {
    double m = mus[tid];
    if (mu_part != nullptr) {
      m = 0.0;
      for (int b = 0; b < nparts; ++b)
        m += (DED && tid < p) ? mu_part_p[(size_t)b * Jpad_p + tid] : mu_part[(size_t)b * Jpad + (size_t)r * qn + tid - p];
    }
    mu[(size_t)s * Q + tid] = mean + m;
  }
  if (ls.flag_count != nullptr) {                     // level switch: does this restart tolerate the cheaper level?
    __syncthreads();
    if (tid == 0) {
      double mx = 0.0;
      for (int t = 0; t < Q * Q; ++t) mx = fmax(mx, sabs[t]);
      int slot = -1;
      if (!(ls.err <= ls.tol * mx)) {
        slot = atomicAdd(ls.flag_count, 1);
        ls.flag_sel[slot] = s;
      }
      ls.flag_slot[s] = slot;
    }
  }
}

// Vbar^T[i][k] = mubar_i gamma_k - sum_j (Sbar_ij + Sbar_ji) V^T[j][k], in place.
template <int Q>
__global__ void __launch_bounds__(256)
vbar_kernel(double* __restrict__ Vt, int npad, const double* __restrict__ gamma,
            const double* __restrict__ mubar, const double* __restrict__ Sigbar) {
  __shared__ double W[Q][Q];
  __shared__ double mb[Q];
  const int s = blockIdx.x, tid = threadIdx.x;
  const double* sb = Sigbar + (size_t)s * Q * Q;
  for (int t = tid; t < Q * Q; t += 256) {
    int i = t / Q, j = t - i * Q;
    W[i][j] = sb[i * Q + j] + sb[j * Q + i];
  }
  if (tid < Q) mb[tid] = mubar[(size_t)s * Q + tid];
  __syncthreads();
  double* V = Vt + (size_t)s * Q * npad;
  for (int k = tid; k < npad; k += 256) {
    double v[Q];
#pragma unroll
    for (int i = 0; i < Q; ++i) v[i] = V[(size_t)i * npad + k];
    const double g = gamma[k];
    // the row loop stays rolled (W, mb from shared memory): fully unrolled it costs 226 registers and leaves one
    // CTA of 8 warps per SM on a kernel whose only job is to stream V^T through HBM
#pragma unroll 1
    for (int i = 0; i < Q; ++i) {
      double o = mb[i] * g;
#pragma unroll
      for (int j = 0; j < Q; ++j) o -= W[i][j] * v[j];
      V[(size_t)i * npad + k] = o;
    }
  }
}
