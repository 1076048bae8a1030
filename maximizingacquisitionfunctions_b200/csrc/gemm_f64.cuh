// FP64 tensor-core contraction  C[J][N] = A[J][K] * T[N][K]^T  with optional triangular T.
//
// This is the dominant kernel of the path: the solve against the cached training
// factor, V^T = K10 * L0^-T (forward, T = L0^-1 lower-triangular) and
// Kbar10 = Vbar^T * L0^-1 (backward, T = (L0^-1)^T upper-triangular).  It replaces
// utils.cholesky_solve on the [n, S*q] right-hand side (src/utils/linalg_ops.py:184-212,
// src/models/gaussian_process.py:195-202) and its autodiff re-solve.
//
// sm_100a notes: tcgen05 has no FP64 kind, so the FP64 tensor path is the warp-level
// DMMA.8x8x4 (mma.sync.m8n8k4.f64).  Both operands are K-contiguous ("TN"), staged in
// shared memory by a 4-stage cp.async pipeline; rows are padded by 4 doubles so that the
// 64-bit fragment loads of a half-warp hit 16 distinct bank pairs.  CTA tile 128x128x16,
// 8 warps, warp tile 64(J) x 32(N), accumulators in registers (128 per lane).
// Triangular structure: a CTA only walks the K range where its T rows are non-zero.
#pragma once
#include "common.cuh"

#define GEMM_BM 128
#define GEMM_BN 128
#define GEMM_BK 16
#define GEMM_STAGES 4
#define GEMM_LD (GEMM_BK + 4)
#define GEMM_THREADS 256
#define GEMM_SMEM_BYTES (GEMM_STAGES * (GEMM_BM + GEMM_BN) * GEMM_LD * (int)sizeof(double))

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// tri: 0 = full T, 1 = T lower-triangular (T[r][c] = 0 for c > r), 2 = upper (c < r).
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_nt_f64_kernel(const double* __restrict__ A, int lda, const double* __restrict__ T, int ldt,
                   double* __restrict__ C, int ldc, int K, int tri) {
  extern __shared__ __align__(16) double gsm[];
  double* As = gsm;                                         // [STAGES][BM][LD]
  double* Ts = gsm + GEMM_STAGES * GEMM_BM * GEMM_LD;       // [STAGES][BN][LD]

  // heaviest N-blocks first (lower: last block walks all of K; upper: first block does)
  const int rb = (tri == 1) ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;
  const int r0 = rb * GEMM_BN;
  const int j0 = blockIdx.y * GEMM_BM;
  const int c_begin = (tri == 2) ? r0 : 0;
  const int c_end = (tri == 1) ? min(K, r0 + GEMM_BN) : K;
  const int ktiles = (c_end - c_begin) / GEMM_BK;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const int wj = warp & 1;    // 2 warps along J (64 rows each)
  const int wr = warp >> 1;   // 4 warps along N (32 cols each)

  const double* Ag = A + (size_t)j0 * lda + c_begin;
  const double* Tg = T + (size_t)r0 * ldt + c_begin;

  auto load_stage = [&](int stage, int kt) {
    double* as = As + stage * GEMM_BM * GEMM_LD;
    double* ts = Ts + stage * GEMM_BN * GEMM_LD;
    const int c0 = kt * GEMM_BK;
#pragma unroll
    for (int i = 0; i < (GEMM_BM * GEMM_BK / 2) / GEMM_THREADS; ++i) {   // 4 x 16-byte chunks
      int t = tid + i * GEMM_THREADS;
      int row = t >> 3, ch = t & 7;
      cp_async16(as + row * GEMM_LD + ch * 2, Ag + (size_t)row * lda + c0 + ch * 2);
      cp_async16(ts + row * GEMM_LD + ch * 2, Tg + (size_t)row * ldt + c0 + ch * 2);
    }
  };

  double acc[8][4][2];
#pragma unroll
  for (int mi = 0; mi < 8; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

#pragma unroll
  for (int s = 0; s < GEMM_STAGES - 1; ++s) {
    if (s < ktiles) load_stage(s, s);
    cp_async_commit();
  }

  for (int kt = 0; kt < ktiles; ++kt) {
    cp_async_wait<GEMM_STAGES - 2>();
    __syncthreads();
    {
      int nk = kt + GEMM_STAGES - 1;
      if (nk < ktiles) load_stage(nk % GEMM_STAGES, nk);
      cp_async_commit();
    }
    const double* as = As + (kt % GEMM_STAGES) * GEMM_BM * GEMM_LD + (wj * 64 + gid) * GEMM_LD + tig;
    const double* ts = Ts + (kt % GEMM_STAGES) * GEMM_BN * GEMM_LD + (wr * 32 + gid) * GEMM_LD + tig;
#pragma unroll
    for (int ks = 0; ks < GEMM_BK / 4; ++ks) {
      double a[8], b[4];
#pragma unroll
      for (int mi = 0; mi < 8; ++mi) a[mi] = as[mi * 8 * GEMM_LD + ks * 4];
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) b[ni] = ts[ni * 8 * GEMM_LD + ks * 4];
#pragma unroll
      for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
  }
  cp_async_wait<0>();

  // epilogue: each lane owns C[gid][2*tig..2*tig+1] of every 8x8 tile -> 16-byte stores,
  // 4 lanes cover 64 contiguous bytes of a row (full 32-byte sectors).
#pragma unroll
  for (int mi = 0; mi < 8; ++mi) {
    double* crow = C + (size_t)(j0 + wj * 64 + mi * 8 + gid) * ldc + r0 + wr * 32 + 2 * tig;
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      double2 v = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
      *reinterpret_cast<double2*>(crow + ni * 8) = v;
    }
  }
}
