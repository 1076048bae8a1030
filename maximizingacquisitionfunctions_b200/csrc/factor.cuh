// Training-set factorisation without library calls: blocked Cholesky  K00 = L L^T  and the explicit triangular
// inverse L^-1 (plus its transpose), all n x n buffers row-major with leading dimension npad (a multiple of 128).
// Reference: gaussian_process.compute_cholesky + utils.jitter_cholesky (src/models/gaussian_process.py:352-360,
// src/utils/linalg_ops.py:112-150) -- tf.cholesky there, once per training set.
//
//   Cholesky, right-looking with 128-wide panels (host loop in maf_api.cu):
//     chol_diag_kernel   : factor the 128 x 128 diagonal block in shared memory, also its inverse D_k^-1
//     gemm_nt (tri = 1)  : panel  L_ik = A_ik D_k^-T                       (FP64 DMMA kernel, gemm_f64.cuh)
//     gemm_nt (epi = 1)  : trailing update  A_ij -= L_ik L_jk^T, lower tiles only
//   Inverse by block doubling: X = [[X11, 0], [-X22 L21 X11, X22]] with X11, X22 from the previous level
//     (level 0: the D_k^-1 blocks), two gemm_nt calls per pair, the transpose maintained alongside.
#pragma once
#include "common.cuh"

#define FAC_NB 128
#define FAC_LD (FAC_NB + 1)
#define FAC_SMEM_BYTES (FAC_NB * FAC_LD * (int)sizeof(double))

// One CTA.  A: top-left of the diagonal block (leading dimension lda), overwritten by its lower Cholesky factor (the
// strict upper triangle of the block is zeroed).  dinv[128][128]: inverse of that factor (lower, row-major).
// info: first failing global row + 1 (0 = success), like LAPACK's potrf.
__global__ void __launch_bounds__(256)
chol_diag_kernel(double* __restrict__ A, int lda, double* __restrict__ dinv, int* __restrict__ info, int row0) {
  extern __shared__ double fsm[];
  __shared__ double pivot;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < FAC_NB * FAC_NB; idx += 256) {
    const int r = idx / FAC_NB, c = idx - r * FAC_NB;
    fsm[r * FAC_LD + c] = (c <= r) ? A[(size_t)r * lda + c] : 0.0;
  }
  __syncthreads();
  for (int j = 0; j < FAC_NB; ++j) {
    if (tid == 0) {
      const double d = fsm[j * FAC_LD + j];
      if (!(d > 0.0)) atomicCAS(info, 0, row0 + j + 1);
      pivot = sqrt(d);
      fsm[j * FAC_LD + j] = pivot;
    }
    __syncthreads();
    const double d = pivot;
    for (int i = j + 1 + tid; i < FAC_NB; i += 256) fsm[i * FAC_LD + j] /= d;
    __syncthreads();
    const int m = FAC_NB - 1 - j;                             // trailing lower triangle: a_ik -= a_ij a_kj, j < k <= i
    for (int idx = tid; idx < m * m; idx += 256) {
      const int ii = idx / m, kk = idx - ii * m;
      if (kk <= ii) {
        const int i = j + 1 + ii, k = j + 1 + kk;
        fsm[i * FAC_LD + k] -= fsm[i * FAC_LD + j] * fsm[k * FAC_LD + j];
      }
    }
    __syncthreads();
  }
  for (int idx = tid; idx < FAC_NB * FAC_NB; idx += 256) {
    const int r = idx / FAC_NB, c = idx - r * FAC_NB;
    A[(size_t)r * lda + c] = fsm[r * FAC_LD + c];             // upper part was loaded as zero
  }
  // inverse of the block by forward substitution, one column per thread: L x = e_t
  if (tid < FAC_NB) {
    const int t = tid;
    for (int i = 0; i < t; ++i) dinv[i * FAC_NB + t] = 0.0;
    dinv[t * FAC_NB + t] = 1.0 / fsm[t * FAC_LD + t];
    for (int i = t + 1; i < FAC_NB; ++i) {
      double acc = 0.0;
      for (int k = t; k < i; ++k) acc += fsm[i * FAC_LD + k] * dinv[k * FAC_NB + t];
      dinv[i * FAC_NB + t] = -acc / fsm[i * FAC_LD + i];
    }
  }
}

// dst[r][c] = src[r][c] for a rows x cols block (different leading dimensions)
__global__ void copy_block_kernel(const double* __restrict__ src, int lds, double* __restrict__ dst, int ldd, int rows, int cols) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c < cols && r < rows) dst[(size_t)r * ldd + c] = src[(size_t)r * lds + c];
}

// dst[c][r] = src[r][c], rows and cols multiples of 32
__global__ void transpose_block_kernel(const double* __restrict__ src, int lds, double* __restrict__ dst, int ldd) {
  __shared__ double t[32][33];
  const int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) t[r][threadIdx.x] = src[(size_t)(y0 + r) * lds + x];
  __syncthreads();
  const int xo = blockIdx.y * 32 + threadIdx.x, yo0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) dst[(size_t)(yo0 + r) * ldd + xo] = t[threadIdx.x][r];
}

// Level 0 of the inverse: Linv = blockdiag(D_k^-1), LinvT = its transpose, zero elsewhere; also clears the strict upper
// triangle of L (the trailing updates only maintain the lower one).
__global__ void inverse_init_kernel(const double* __restrict__ dinv, double* __restrict__ L, double* __restrict__ Linv,
                                    double* __restrict__ LinvT, int npad) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c >= npad) return;
  const size_t o = (size_t)r * npad + c;
  if (c > r) L[o] = 0.0;
  const int br = r / FAC_NB, bc = c / FAC_NB;
  double x = 0.0, xt = 0.0;
  if (br == bc) {
    const double* d = dinv + (size_t)br * FAC_NB * FAC_NB;
    const int rr = r - br * FAC_NB, cc = c - bc * FAC_NB;
    x = d[rr * FAC_NB + cc];                                  // lower
    xt = d[cc * FAC_NB + rr];                                 // upper
  }
  Linv[o] = x;
  LinvT[o] = xt;
}
