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
// shared memory by a multi-stage cp.async pipeline; rows are padded by 4 doubles so that the
// 64-bit fragment loads of a half-warp hit 16 distinct bank pairs.  CTA tile BM x BN x BK,
// accumulators in registers.  Triangular structure: a CTA only walks the K range where its
// T rows are non-zero.
//
// Template configuration <BM, BN, BK, STAGES, WJ, WR, MINB>: WJ x WR warps, warp tile (BM/WJ) x (BN/WR),
// MINB co-resident CTAs per SM (co-resident CTAs hide each other's barriers, pipeline fill and epilogue).
#pragma once
#include "common.cuh"


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

template <int BM, int BN, int BK, int STAGES, int WJ, int WR, int MINB>
struct GemmCfg {
  static constexpr int LD = BK + 4;                        // row stride == 8 words (mod 32)
  static constexpr int THREADS = WJ * WR * 32;
  static constexpr int MT = BM / WJ / 8;                   // 8x8 tiles per warp along J
  static constexpr int NT = BN / WR / 8;                   // along N
  static constexpr int SMEM_BYTES = STAGES * (BM + BN) * LD * (int)sizeof(double);
};

// tri: 0 = full T, 1 = T lower-triangular (T[r][c] = 0 for c > r), 2 = upper (c < r).
// epi: 0 = store C, 1 = C -= A T^T on the tiles that touch the lower triangle of a square C (symmetric rank-k update
//      of the blocked Cholesky, factor.cuh), 2 = store -A T^T.
template <int BM, int BN, int BK, int STAGES, int WJ, int WR, int MINB>
__global__ void __launch_bounds__(WJ * WR * 32, MINB)
gemm_nt_f64_kernel(const double* __restrict__ A, int lda, const double* __restrict__ T, int ldt,
                   double* __restrict__ C, int ldc, int K, int tri, int epi) {
  using Cfg = GemmCfg<BM, BN, BK, STAGES, WJ, WR, MINB>;
  constexpr int LD = Cfg::LD, THREADS = Cfg::THREADS, MT = Cfg::MT, NT = Cfg::NT;
  extern __shared__ __align__(16) double gsm[];
  double* As = gsm;                                   // [STAGES][BM][LD]
  double* Ts = gsm + STAGES * BM * LD;           // [STAGES][BN][LD]

  // heaviest N-blocks first (lower: the last block walks all of K; upper: the first does)
  const int rb = (tri == 1) ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;
  const int r0 = rb * BN;
  const int j0 = blockIdx.y * BM;
  if (epi == 1 && r0 >= j0 + BM) return;                      // tile strictly above the diagonal
  const int c_begin = (tri == 2) ? r0 : 0;
  const int c_end = (tri == 1) ? min(K, r0 + BN) : K;
  const int ktiles = (c_end - c_begin) / BK;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const int wj = warp % WJ;
  const int wr = warp / WJ;

  const double* Ag = A + (size_t)j0 * lda + c_begin;
  const double* Tg = T + (size_t)r0 * ldt + c_begin;

  auto load_stage = [&](int stage, int kt) {
    double* as = As + stage * BM * LD;
    double* ts = Ts + stage * BN * LD;
    const int c0 = kt * BK;
    constexpr int CPR = BK / 2;                         // 16-byte chunks per row
#pragma unroll
    for (int i = 0; i < (BM * CPR) / THREADS; ++i) {
      int t = tid + i * THREADS;
      int row = t / CPR, ch = t % CPR;
      cp_async16(as + row * LD + ch * 2, Ag + (size_t)row * lda + c0 + ch * 2);
    }
#pragma unroll
    for (int i = 0; i < (BN * CPR) / THREADS; ++i) {
      int t = tid + i * THREADS;
      int row = t / CPR, ch = t % CPR;
      cp_async16(ts + row * LD + ch * 2, Tg + (size_t)row * ldt + c0 + ch * 2);
    }
  };

  double acc[MT][NT][2];
#pragma unroll
  for (int mi = 0; mi < MT; ++mi)
#pragma unroll
    for (int ni = 0; ni < NT; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < ktiles) load_stage(s, s);
    cp_async_commit();
  }

  for (int kt = 0; kt < ktiles; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      int nk = kt + STAGES - 1;
      if (nk < ktiles) load_stage(nk % STAGES, nk);
      cp_async_commit();
    }
    const double* as = As + (kt % STAGES) * BM * LD + (wj * MT * 8 + gid) * LD + tig;
    const double* ts = Ts + (kt % STAGES) * BN * LD + (wr * NT * 8 + gid) * LD + tig;
    double a[2][MT], b[2][NT];                          // fragment double buffer across k4 steps
#pragma unroll
    for (int mi = 0; mi < MT; ++mi) a[0][mi] = as[mi * 8 * LD];
#pragma unroll
    for (int ni = 0; ni < NT; ++ni) b[0][ni] = ts[ni * 8 * LD];
#pragma unroll
    for (int ks = 0; ks < BK / 4; ++ks) {
      const int cur = ks & 1, nxt = cur ^ 1;
      if (ks + 1 < BK / 4) {
#pragma unroll
        for (int mi = 0; mi < MT; ++mi) a[nxt][mi] = as[mi * 8 * LD + (ks + 1) * 4];
#pragma unroll
        for (int ni = 0; ni < NT; ++ni) b[nxt][ni] = ts[ni * 8 * LD + (ks + 1) * 4];
      }
#pragma unroll
      for (int mi = 0; mi < MT; ++mi)
#pragma unroll
        for (int ni = 0; ni < NT; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[cur][mi], b[cur][ni]);
    }
  }
  cp_async_wait<0>();

  // epilogue: each lane owns C[gid][2*tig..2*tig+1] of every 8x8 tile -> 16-byte stores,
  // 4 lanes cover 64 contiguous bytes of a row (full 32-byte sectors).
#pragma unroll
  for (int mi = 0; mi < MT; ++mi) {
    double* crow = C + (size_t)(j0 + wj * MT * 8 + mi * 8 + gid) * ldc + r0 + wr * NT * 8 + 2 * tig;
#pragma unroll
    for (int ni = 0; ni < NT; ++ni) {
      double2 v = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
      if (epi == 1) {
        const double2 old = *reinterpret_cast<const double2*>(crow + ni * 8);
        v = make_double2(old.x - v.x, old.y - v.y);
      } else if (epi == 2) {
        v = make_double2(-v.x, -v.y);
      }
      *reinterpret_cast<double2*>(crow + ni * 8) = v;
    }
  }
}
