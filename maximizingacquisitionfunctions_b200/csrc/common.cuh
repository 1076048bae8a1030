// Shared declarations for the maf_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cmath>
#include "../../include/maf.h"

#define MAF_EPS_F64 2.220446049250313e-16

// Model hyper-parameters, passed by value to kernels (lives in the param bank).
struct ModelParams {
  int kernel_id;
  int d;
  double inv_ls[MAF_MAX_D];   // 1 / lengthscale  (reference multiplies by the reciprocal)
  double amp2;                // amplitude^2
  double noise;               // noise variance (0 if noiseless)
  double jitter;
};

// Acquisition parameters resolved on the host (level/temperature NaNs replaced).
struct AcqDev {
  int acq;
  int lower;
  int soft;          // PI: 1 = sigmoid relaxation
  double level;
  double inv_temp;
  double cb_scale;   // sqrt(pi/2 * beta)
  double beta;
};

__host__ __device__ __forceinline__ int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Level switch (error-bounded choice of the digit planes of a contraction, per restart).  A contraction first runs with
// one plane less (21 instead of 28 exact products); the kernel that consumes its result compares the calibrated absolute
// error of that cheaper level (err, measured against the full level on probe candidates when the training set was
// loaded) with what the restart tolerates (tol x the scale of its own result) and appends the restarts that fail to
// a list.  A second pass recomputes exactly those restarts with all planes on COMPACT rows (position pos of the list
// <-> rows pos q .. pos q + q - 1); later kernels pick the rows of a restart by its slot.  Everything stays on the
// device (no host decision, replayable in a CUDA graph); a restart's result depends only on that restart.
struct LevelSwitch {
  const int* sel;         // second pass: compact position -> restart (null: first pass / switch off)
  const int* sel_count;   // second pass: number of selected restarts
  int* flag_count;        // first pass: counter of the restarts that fail the test (null: no test)
  int* flag_sel;          // first pass: their list
  int* flag_slot;         // first pass: restart -> position in the list, or -1
  double err;             // calibrated error of the cheaper level (absolute, or per unit row scale)
  double tol;             // relative tolerance the result must keep
};

// kappa(r2): src/models/kernels.py:32-49 (eps clamp inside the Matern kernels only).
__device__ __forceinline__ double kern_val(int kid, double r2) {
  if (kid == MAF_KERNEL_SE) return exp(-0.5 * r2);
  if (kid == MAF_KERNEL_MATERN52) {
    double u = 5.0 * fmax(r2, MAF_EPS_F64);
    double r = sqrt(u);
    return (1.0 + r + (1.0 / 3.0) * u) * exp(-r);
  }
  double r = sqrt(3.0 * fmax(r2, MAF_EPS_F64));
  return (1.0 + r) * exp(-r);
}

// d kappa / d r2.  The clamp passes gradient only where r2 >= eps
// (clip_by_value == min(max(x, lo), hi)).
__device__ __forceinline__ double kern_dr2(int kid, double r2) {
  if (kid == MAF_KERNEL_SE) return -0.5 * exp(-0.5 * r2);
  if (r2 < MAF_EPS_F64) return 0.0;
  if (kid == MAF_KERNEL_MATERN52) {
    double r = sqrt(5.0 * r2);
    return -(5.0 / 6.0) * (1.0 + r) * exp(-r);
  }
  double r = sqrt(3.0 * r2);
  return -1.5 * exp(-r);
}

// ---- lean FP64 sqrt / exp for the two FP64-pipe-bound covariance kernels (cross_fwd_planes, cross_bwd) -------------
// Both kernels spend ~60 FP64 instructions per (candidate, training point) pair, a third of it inside the library's
// sqrt (IEEE rounding, slow-path call) and exp (degree-11 polynomial).  The arguments here are confined
// (u in [5 eps, ~1e3], r >= 0), so:
//   sqrt: FP32 rsqrt seed (2^-22) + two coupled Heron steps  r <- r + (u - r^2) y/2 : error eps^2/2 + eps delta < 2^-60
//         before rounding (6 FP64 instructions, <= 1 ulp);
//   exp(-r) = 2^m T[j] e^f,  k = rint(-r 32/ln2) = 32 m + j,  |f| <= ln2/64: degree-6 Taylor (next term 3.5e-18),
//         T[j] = 2^(j/32) from a 32-entry shared-memory table, 2^m folded into the exponent field (11 FP64 instructions,
//         <= 2 ulp).  Results are a few ulp from the library's; the parity bar is 1e-9.
#define MAF_ETAB 32
// The table index is data dependent, so the plain 32-entry table costs shared-memory bank conflicts (ncu round 1: 39 % /
// 49 % of the shared wavefronts of cross_fwd_planes / cross_bwd: entries j and j + 16 share a bank pair).
// MAF_ETAB_REPL=1 replicates the table per lane, tab[j][lane] (8 KB, conflict-free).  Measured on one box
// (profiles/r02_etab_layout_ab.txt): no difference in either direction -- the conflicts are not what limits these kernels
// (FP64 pipe and load latency are) -- so the plain 256-byte table stays the default.
#ifndef MAF_ETAB_REPL
#define MAF_ETAB_REPL 0
#endif
#if MAF_ETAB_REPL
#define MAF_ETAB_WORDS (MAF_ETAB * 32)
#define MAF_ETAB_IDX(j) ((j) << 5)
#define MAF_ETAB_LANE(tid) ((tid) & 31)
#else
#define MAF_ETAB_WORDS MAF_ETAB
#define MAF_ETAB_IDX(j) (j)
#define MAF_ETAB_LANE(tid) 0
#endif
// coefficients live in the constant bank so that every DFMA takes them as a c[][] operand (as literals the compiler
// rebuilds each 64-bit immediate with two moves per use: 18 of the kernel's 126 instructions per element)
__constant__ double MAF_FC[10] = {
    -46.16624130844683,        // 0: -32/ln2
    -0.021660849391992087,     // 1: -(ln2/32) high part 0x1.62e42fef8p-6 (34 bits: k * hi is exact)
    -5.062034433330175e-13,    // 2: -(ln2/32) low part
    1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0,        // 3..6: Taylor coefficients of e^f
    1.0 / 3.0,                 // 7
    6755399441055744.0,        // 8: 1.5 2^52
    1.0e6};                    // 9: cap of the exponent argument
// 2^(j/32), j = 0..31, correctly rounded (hex literals): the fill below is four constant-bank reads per thread, not 1024
// calls of exp2 per CTA (cross_bwd runs one small CTA per restart)
__constant__ double MAF_ETAB_C[MAF_ETAB] = {
    0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0,
    0x1.172b83c7d517bp+0, 0x1.1d4873168b9aap+0, 0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0,
    0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0, 0x1.3dea64c123422p+0, 0x1.44e086061892dp+0,
    0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0, 0x1.6247eb03a5585p+0,
    0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0, 0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0,
    0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0,
    0x1.ae89f995ad3adp+0, 0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0,
    0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0, 0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0};
__device__ __forceinline__ void maf_fill_etab(double* tab, int tid, int nthreads) {
  for (int t = tid; t < MAF_ETAB_WORDS; t += nthreads) tab[t] = MAF_ETAB_C[MAF_ETAB_REPL ? (t >> 5) : t];
}
// max(x, eps) for x >= 0: eps = 2^-52 has a zero low word, so the comparison is one integer compare on the high word
__device__ __forceinline__ double maf_clamp_eps(double x) {
  return (__double2hiint(x) < 0x3CB00000) ? MAF_EPS_F64 : x;
}
__device__ __forceinline__ double maf_sqrt_fast(double u) {   // 2^-126 < u < 2^127
  float yf;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(yf) : "f"((float)u));
  const double y = (double)yf;
  const double hy = 0.5 * y;
  double r = u * y;
  r = fma(fma(-r, r, u), hy, r);
  r = fma(fma(-r, r, u), hy, r);
  return r;
}
__device__ __forceinline__ double maf_exp_neg_fast(double r, const double* __restrict__ tab) {   // r >= 0
  if (__double2hiint(r) > 0x412E8480) r = MAF_FC[9];         // r > 1e6 (result 0 either way): keeps k inside int32
  const double t = fma(r, MAF_FC[0], MAF_FC[8]);             // -r 32/ln2 rounded to an integer in the low word
  const int k = __double2loint(t);
  const double kd = t - MAF_FC[8];
  double f = fma(kd, MAF_FC[1], -r);
  f = fma(kd, MAF_FC[2], f);
  double p = fma(f, MAF_FC[3], MAF_FC[4]);
  p = fma(p, f, MAF_FC[5]);
  p = fma(p, f, MAF_FC[6]);
  p = fma(p, f, 0.5);
  p = fma(p, f, 1.0);
  p = fma(p, f, 1.0);
  const double tj = tab[MAF_ETAB_IDX(k & (MAF_ETAB - 1))];
  const int m = max(k >> 5, -1021);                          // floor(k / 32) <= 0; below 2^-1021 the result is "tiny", not exact
  const double s = __hiloint2double(__double2hiint(tj) + (m << 20), __double2loint(tj));
  return s * p;
}

template <int KID>
__device__ __forceinline__ double kern_val_fast(double r2, const double* __restrict__ tab) {
  if (KID == MAF_KERNEL_SE) return maf_exp_neg_fast(0.5 * r2, tab);
  if (KID == MAF_KERNEL_MATERN52) {
    const double u = 5.0 * maf_clamp_eps(r2);
    const double r = maf_sqrt_fast(u);
    return fma(u, MAF_FC[7], 1.0 + r) * maf_exp_neg_fast(r, tab);
  }
  const double r = maf_sqrt_fast(3.0 * maf_clamp_eps(r2));
  return (1.0 + r) * maf_exp_neg_fast(r, tab);
}
template <int KID>
__device__ __forceinline__ double kern_dr2_fast(double r2, const double* __restrict__ tab) {
  if (KID == MAF_KERNEL_SE) return -0.5 * maf_exp_neg_fast(0.5 * r2, tab);
  const bool below = __double2hiint(r2) < 0x3CB00000;        // r2 < eps: the clamp blocks the gradient
  const double rc = below ? MAF_EPS_F64 : r2;
  if (KID == MAF_KERNEL_MATERN52) {
    const double r = maf_sqrt_fast(5.0 * rc);
    const double g = (1.0 + r) * maf_exp_neg_fast(r, tab);
    return below ? 0.0 : -(5.0 / 6.0) * g;
  }
  const double r = maf_sqrt_fast(3.0 * rc);
  const double g = maf_exp_neg_fast(r, tab);
  return below ? 0.0 : -1.5 * g;
}

// kappa(r2) and d kappa / d r2 from one sqrt and one exp: the same arithmetic as kern_val_fast / kern_dr2_fast, so the
// pair is bit-identical to calling both.  The forward covariance kernel stores the derivative (see cross_bwd_g_kernel).
template <int KID>
__device__ __forceinline__ double kern_val_dr2_fast(double r2, const double* __restrict__ tab, double& g) {
  if (KID == MAF_KERNEL_SE) {
    const double E = maf_exp_neg_fast(0.5 * r2, tab);
    g = -0.5 * E;
    return E;
  }
  const bool below = __double2hiint(r2) < 0x3CB00000;        // r2 < eps: value clamped, gradient blocked
  const double rc = below ? MAF_EPS_F64 : r2;
  if (KID == MAF_KERNEL_MATERN52) {
    const double u = 5.0 * rc;
    const double r = maf_sqrt_fast(u);
    const double E = maf_exp_neg_fast(r, tab);
    const double gg = (1.0 + r) * E;
    g = below ? 0.0 : -(5.0 / 6.0) * gg;
    return fma(u, MAF_FC[7], 1.0 + r) * E;
  }
  const double r = maf_sqrt_fast(3.0 * rc);
  const double E = maf_exp_neg_fast(r, tab);
  g = below ? 0.0 : -1.5 * E;
  return (1.0 + r) * E;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
