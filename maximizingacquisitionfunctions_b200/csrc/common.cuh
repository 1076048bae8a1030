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

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
