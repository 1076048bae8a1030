// GP marginal likelihood and its gradient w.r.t. the log hyper-parameters, from the factors
// maf_set_train already caches (L0, L0^-1, gamma = L0^-1 rho).
// Reference: gaussian_process.log_likelihood (src/models/gaussian_process.py:78-106): data term
//   -1/2 (rho^T K^-1 rho + log det K + n log 2 pi),  K = a^2 kappa + (noise + jitter) I.
// Gradient (TF autodiff in the reference; closed form here):
//   d(-llh)/d theta = 1/2 sum_ij (K^-1 - alpha alpha^T)_ij dK_ij/d theta,   alpha = K^-1 rho.
#pragma once
#include "common.cuh"

// out[0] = 2 sum_i log L_ii, out[1] = gamma^T gamma   (single block)
__global__ void nll_scalars_kernel(const double* __restrict__ L, const double* __restrict__ gamma, int n, int npad,
                                   double* __restrict__ out) {
  __shared__ double s0[32], s1[32];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    a += log(L[(size_t)i * npad + i]);
    b += gamma[i] * gamma[i];
  }
  a = warp_sum(a);
  b = warp_sum(b);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { s0[w] = a; s1[w] = b; }
  __syncthreads();
  if (w == 0) {
    a = (l < (blockDim.x >> 5)) ? s0[l] : 0.0;
    b = (l < (blockDim.x >> 5)) ? s1[l] : 0.0;
    a = warp_sum(a);
    b = warp_sum(b);
    if (l == 0) { out[0] = 2.0 * a; out[1] = b; }
  }
}

// alpha_k = sum_{i >= k} LinvT[k][i] gamma_i ; one warp per row.
__global__ void trmv_upper_kernel(const double* __restrict__ LinvT, const double* __restrict__ g, int npad,
                                  double* __restrict__ alpha) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= npad) return;
  double a = 0.0;
  for (int k = row + lane; k < npad; k += 32) a += LinvT[(size_t)row * npad + k] * g[k];
  a = warp_sum(a);
  if (lane == 0) alpha[row] = a;
}

// Residual of the training system in twice the working precision (compensated dot product, Ogita-Rump-Oishi Dot2):
//   res_i = rho_i - sum_k K00[i][k] alpha_k,   K00 recomputed entry by entry exactly as k00_kernel wrote it.
// One warp per row.  Drives the iterative refinement of alpha = K00^-1 rho in maf_set_train: alpha feeds the posterior
// mean (K10 . alpha) and its gradient directly, so it is made accurate to the last bits once per training set instead
// of carrying cond(K00) eps = 4e-10 of relative error.
__device__ __forceinline__ void maf_two_sum(double a, double b, double& s, double& e) {
  s = a + b;
  const double bb = s - a;
  e = (a - (s - bb)) + (b - bb);
}
__global__ void alpha_residual_kernel(const double* __restrict__ X0s, int n, int npad, ModelParams mp,
                                      const double* __restrict__ rho, const double* __restrict__ alpha,
                                      double* __restrict__ res) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= npad) return;
  if (i >= n) {
    if (lane == 0) res[i] = 0.0;
    return;
  }
  double xi[MAF_MAX_D];
  for (int c = 0; c < mp.d; ++c) xi[c] = X0s[(size_t)c * npad + i];
  double sh = 0.0, sl = 0.0;
  for (int k = lane; k < n; k += 32) {
    double r2 = 0.0;
    for (int c = 0; c < mp.d; ++c) {
      const double df = xi[c] - X0s[(size_t)c * npad + k];
      r2 += df * df;
    }
    double v = mp.amp2 * kern_val(mp.kernel_id, r2);
    if (i == k) v = (v + mp.noise) + mp.jitter;
    const double a = alpha[k];
    const double p = v * a, pe = fma(v, a, -p);
    double t, te;
    maf_two_sum(sh, p, t, te);
    sh = t;
    sl += te + pe;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double oh = __shfl_xor_sync(0xffffffffu, sh, o), ol = __shfl_xor_sync(0xffffffffu, sl, o);
    double t, te;
    maf_two_sum(sh, oh, t, te);
    sh = t;
    sl += te + ol;
  }
  if (lane == 0) {
    double t, te;
    maf_two_sum(rho[i], -sh, t, te);
    res[i] = t + (te - sl);
  }
}

__global__ void axpy_kernel(double* __restrict__ y, const double* __restrict__ x, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] += x[i];
}

// out[c] (c < d): d/dlog l_c ; out[d]: d/dlog a ; out[d+1]: d/dlog noise ; out[d+2]: d/dmean.
// grid (ceil(npad/256), n): thread = one (i, j) entry.
__global__ void nll_grad_kernel(const double* __restrict__ X0s, int n, int npad, ModelParams mp,
                                const double* __restrict__ Kinv, const double* __restrict__ alpha,
                                double* __restrict__ out) {
  __shared__ double red[MAF_MAX_D + 3][8];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  double g[MAF_MAX_D + 3];
#pragma unroll
  for (int c = 0; c < MAF_MAX_D + 3; ++c) g[c] = 0.0;
  if (j < n) {
    const double W = Kinv[(size_t)i * npad + j] - alpha[i] * alpha[j];
    double r2 = 0.0, df2[MAF_MAX_D];
#pragma unroll
    for (int c = 0; c < MAF_MAX_D; ++c) {
      if (c < mp.d) {
        double df = X0s[(size_t)c * npad + i] - X0s[(size_t)c * npad + j];
        df2[c] = df * df;
        r2 += df2[c];
      }
    }
    const double kp = mp.amp2 * kern_dr2(mp.kernel_id, r2);
    const double kv = mp.amp2 * kern_val(mp.kernel_id, r2);
#pragma unroll
    for (int c = 0; c < MAF_MAX_D; ++c)
      if (c < mp.d) g[c] = 0.5 * W * kp * (-2.0) * df2[c];        // d r2 / d log l_c = -2 (x_ic - x_jc)^2 / l_c^2
    g[mp.d] = 0.5 * W * 2.0 * kv;                                  // d K / d log a = 2 a^2 kappa
    if (i == j) {
      g[mp.d + 1] = 0.5 * W * mp.noise;                            // d K / d log noise = noise I
      g[mp.d + 2] = -alpha[i];                                     // d(-llh)/d mean = -1^T alpha
    }
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  for (int c = 0; c < mp.d + 3; ++c) {
    double v = warp_sum(g[c]);
    if (l == 0) red[c][w] = v;
  }
  __syncthreads();
  if (threadIdx.x < mp.d + 3) {
    double v = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) v += red[threadIdx.x][k];
    atomicAdd(out + threadIdx.x, v);
  }
}
