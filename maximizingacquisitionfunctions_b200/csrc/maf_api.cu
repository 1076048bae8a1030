// C ABI of the B200-native acquisition-maximisation path (see include/maf.h).
// Host orchestration only: every number is produced by the kernels in the .cuh files.

#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "cov_kernels.cuh"
#include "factor.cuh"
#include "fit_kernels.cuh"
#include "gemm_f64.cuh"
#include "lbfgs.cuh"
#include "mc_kernels.cuh"
#include "ozaki_i8.cuh"
#include "ozaki_i8_v3.cuh"
#include "ozaki_i8_v4.cuh"
#include "posterior.cuh"

#define MAF_VERSION_STR "maf_b200 0.2 (sm_100a; INT8 digit-plane contractions on tcgen05, FP64 DMMA fallback)"

struct ProfSlot {
  std::vector<cudaEvent_t> ev;   // pairs (start, stop)
  size_t used = 0;
  double ms = 0.0;
  int64_t launches = 0;
};

struct Buf {
  double* p = nullptr;
  size_t cap = 0;   // doubles
};

struct maf_ctx {
  int device = 0;
  int dtype = MAF_F64;
  cudaStream_t stream = nullptr;
  std::string err;
  int64_t launch_count = 0;
  // cudaFuncSetAttribute is per device: the opt-ins are tracked per context, never per process
  unsigned attr_gemm_mask = 0;     // launch_gemm_cfg variants
  bool attr_oz = false, attr_chol = false;
  unsigned attr_mc_mask = 0;       // mc_acq_kernel<Q>, bit Q
  int num_sms = 0;

  // model
  bool model_set = false;
  ModelParams mp{};
  double mean_param = 0.0;     // NaN => empirical
  // training set
  bool train_set = false;
  int n = 0, npad = 0;
  double mean_val = 0.0, y_min = 0.0;
  double *X0s = nullptr, *L0 = nullptr, *Linv = nullptr, *LinvT = nullptr, *gamma = nullptr;
  double* alpha = nullptr;         // K00^-1 (y0 - m), refined (maf_set_train); [npad], zero on the pad
  double* linv_rowsum = nullptr;   // L0^-1 1 (ones on the n real columns); [npad]: the add-back of the centred K10 planes
  Buf row_add_fwd;                 // [Jpad] a^2/2 per valid candidate row (0 on pad rows), written by cross_fwd_planes_kernel
  int centre_k10 = 1;              // MAF_CENTRE_K10=0: planes of K10 itself (fixed scale a^2, one bit less of range)
  // Posterior mean by the exact FP64 route mu = m + K10 . alpha inside the cross-covariance kernel, its gradient
  // mubar_j alpha_k added inside cross_bwd_kernel: the mean no longer passes through the digit-sliced contractions
  // (INT8 engine with both fused producers only; MAF_MU_ALPHA=0 restores mu = m + V^T gamma).
  int mu_alpha = 1;
  int vbar_tight = 0;              // MAF_VBAR_TIGHT=1: exact row maxima for the Vbar digit planes (one more sweep)
  Buf mu_part;                     // [column blocks of 1024][Jpad] partial sums of K10 . alpha
  Buf row_add;                     // [Jpad] mubar per row of the Vbar planes (rank-one term of the reverse epilogue)
  // Pending points (rows 0..p-1 of every pool, identical for all restarts: pools = concat(tile(Xpend), new),
  // bayesian_optimization.py:1106-1123) are contracted ONCE per evaluation, on all digit planes: V^T rows, mean partial
  // sums and row maxima of the shared block; the restarts only carry their q - p trainable rows through both contractions.
  Buf Vp, mu_part_p, rowmax_p;
  bool pend_shared = false;        // the caller's pools were checked: identical pending rows (host entry points)
  int pend_dedupe = 1;             // MAF_PEND_DEDUPE=0: every restart contracts its own copy of the pending rows (A/B)
  // Level switch (common.cuh, struct LevelSwitch): both contractions first run with one digit plane less (21 products
  // instead of 28); restarts whose result does not tolerate the calibrated error of that level are recomputed with all
  // planes on compact rows.  MAF_LEVEL_SWITCH=0 / maf_set_level_switch turn it off (every restart on all planes).
  int lsw_on = 1;
  double lsw_tol = 2.5e-10;        // relative error the cheaper level may leave (a quarter of the 1e-9 parity bar)
  int lsw_min_npad = 1024;         // below this the contractions do not dominate and the extra launches cost more than they save
  bool lsw_ready = false;          // calibrated for the current training set
  bool lsw_busy = false;           // calibration in progress (plain passes)
  double lsw_err_fwd = 0.0;        // max |Sigma(cheaper level) - Sigma(all planes)| over the probe candidates
  double lsw_err_bwd = 0.0;        // max |grad(cheaper) - grad(all)| / row scale of Vbar over the probe candidates
  Buf Vt2, rowmax2;                // forward second pass: V^T rows / row maxima of the recomputed restarts (compact)
  int *lsw_cnt = nullptr;          // device: [0] forward count, [1] reverse count of the current chunk
  long long* lsw_stats = nullptr;  // device: {tested forward, recomputed forward, tested reverse, recomputed reverse}
  long long* lsw_host = nullptr;   // pinned mirror of lsw_stats, refreshed (asynchronously) at the end of every evaluation
  // A direction whose cheaper level fails for more than 40 % of the restarts costs more than it saves (21 + 28 share
  // products against 28): it then runs on all planes directly for the next 64 evaluations before the cheaper level is
  // tried again.  Decided on the host from the mirror, between evaluations; the results never depend on it.
  bool lsw_cheap_fwd = true, lsw_cheap_bwd = true;
  long long lsw_base[4] = {0, 0, 0, 0};
  int lsw_cool_fwd = 0, lsw_cool_bwd = 0;
  int *lsw_selF = nullptr, *lsw_slotF = nullptr, *lsw_selB = nullptr;
  size_t lsw_cap = 0;              // restarts the three lists were sized for
  size_t cap_train = 0;        // npad the factor buffers were sized for
  int cap_train_d = 0;
  // GEMM workspace (chunk of candidate rows)
  int gemm_variant = 0;
  size_t chunk_rows = 65536;
  Buf K10, Vt;
  // per-restart buffers
  Buf dX, dz, dw, dinc, dmu, dSig, dchol, dmubar, dSigbar, dloss, dgrad;
  int last_S = 0, last_q = 0;
  bool last_has_chol = false;
  const double* last_dX = nullptr;   // device pools of the last evaluation (maf_best_record_dev)
  double* d_argmin_val = nullptr;
  long long* d_argmin_idx = nullptr;
  // Adam state
  bool adam_on = false;
  maf_acq_params adam_prm{};
  int adam_S = 0, adam_q = 0, adam_p = 0;
  int opt_method = 0;              // MAF_OPT_*
  double opt_momentum = 0.0;
  long long opt_t = 0;             // global_step since maf_adam_begin
  Buf adam_X, adam_m, adam_v, adam_g, adam_z;
  // INT8 digit-sliced contraction (ozaki_i8.cuh)
  int gemm_i8 = 1;                 // requested mode (default: INT8 digit slicing; DMMA when n > 8192)
  bool i8_ready = false;           // factor planes built for the current training set
  int oz_ns = 7, oz_lmax = 8;
  // reverse contraction Kbar10 = Vbar^T L0^-1.  MAF_OZ_NS_BWD=6 (21 products instead of 28) makes it 25 % faster but the
  // gradient error at n=4096 rises from the FP64 noise floor (4e-12 .. 2e-11) to 6e-10 .. 5e-9 for candidates next to
  // training points (profiles/r01_accuracy_n4096.jsonl): above the 1e-9 bar, so the default keeps all 7 planes.
  int oz_ns_bwd = 7, oz_lmax_bwd = 8;
  int oz_ns_fwd = 7;               // digit planes of K10 in the forward contraction (levels <= oz_ns_fwd + 1)
  int oz_fuse = 3;                 // bit 0: cross-covariance, bit 1: Vbar written straight to digit planes (MAF_OZ_FUSE)
  int oz_sj = 2;                   // super-row height of the persistent tile order (MAF_OZ_SJ)
  int oz_hint_a = 2, oz_hint_t = 2; // L2 policy of the A / T plane loads: 0 normal, 1 evict-first, 2 evict-last (MAF_OZ_HINTS=at); evict-last on both: DRAM reads 19 -> 16.5 GB per launch, -3 % time (profiles/r01_i8gemm_v5_l2_policy_sweep.txt)
  int oz_pf = 0;                   // L2 prefetch of plane tiles two k-chunks ahead (MAF_OZ_PF)
  int oz_dbg = 0;                  // MAF_OZ_DBG diagnostics (results invalid)
  int oz_variant = 5;              // 1: generic kernel; 3: static schedule, persistent; 4: 3 on CTA pairs (cta_group::2); 5: 4 + register-stash epilogue
  Buf ozA, ozTL, ozTU;             // digit planes (int8, sized in doubles): candidates / Linv / LinvT
  Buf ozsA, ozsTL, ozsTU;          // row scales
  Buf ozVmax;                      // row maxima of V^T left by the forward contraction's epilogue
  Buf Gk;                          // d kappa / d r2 of every (candidate row, training point) pair, left by the forward covariance kernel
  int g_route = 1;                 // MAF_G_ROUTE=0: the reverse covariance kernel recomputes the derivative (sqrt + exp per pair)
  int oz_rowmax = 1;               // MAF_OZ_ROWMAX=0: the Vbar producer sweeps V^T for them
  struct OzSched { int nN, nJ2, tri, sj, pairs, groups; int* d_list; int* d_begin; };
  std::vector<OzSched> oz_scheds;  // balanced static tile schedules of the CTA-pair kernel, one per launch shape
  int oz_sched = 1;                // MAF_OZ_SCHED=0: round-robin
  double oz_sched_ovh = 1.5;       // per-tile overhead of the cost model, in k-chunks (MAF_OZ_SCHED_OVH)
  // batched L-BFGS (lbfgs.cuh)
  Buf lb_X, lb_ft, lb_gt, lb_x, lb_f, lb_g, lb_dir, lb_step, lb_gd, lb_hs, lb_hy, lb_rho, lb_trace;
  int* lb_int = nullptr;           // device: hcount, hpos, status, iters [S each] + active [1]
  size_t lb_int_cap = 0;
  int* lb_active_host = nullptr;   // pinned mirror of the active counter
  // replayed optimizer steps: one CUDA graph per session (MAF_GRAPH=0: every step launched from the host)
  int graph_on = 1;
  long long generation = 0;        // bumped whenever a device buffer moves or the model / training set / engine changes
  cudaGraphExec_t adam_graph = nullptr;
  long long adam_graph_gen = -1;
  int adam_graph_M = 0, adam_graph_trace = 0;
  int64_t adam_graph_kernels = 0;  // kernels per replay
  Buf adam_zcur, adam_coef, adam_trace;
  long long* adam_ctr = nullptr;   // device: {step being executed, next step} within the current maf_adam_steps call
  int oz_groups = 2;               // column groups of the balanced schedule (MAF_OZ_GROUPS)
  // fantasies (rank-4 observation sets)
  int F = 0, Fpad = 0;
  Buf fGamma, fGammaT, fmeans, flevels, fVG, fmubar;
  double flevel_min = 0.0;         // minimum over ALL fantasised outputs (the level of negative_pi, see run_fantasies)
  double adam_lr = 0, adam_b1 = 0, adam_b2 = 0, adam_eps = 0, adam_b1p = 0, adam_b2p = 0;
  // profiling
  cudaEvent_t timer_ev[2] = {nullptr, nullptr};
  bool prof = false;
  ProfSlot slots[MAF_PROF_NSLOTS];
};

static int fail(maf_ctx* ctx, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  return 1;
}

#define CK(expr)                                                                           \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return fail(ctx, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)
#define CKB(expr)                                                                          \
  do {                                                                                     \
    int _s = (int)(expr);                                                                  \
    if (_s != 0) return fail(ctx, "%s failed with status %d (%s:%d)", #expr, _s, __FILE__, __LINE__); \
  } while (0)
#define CKL() CK(cudaGetLastError())

static int ensure(maf_ctx* ctx, Buf& b, size_t need) {
  if (b.p != nullptr && b.cap >= need) return 0;
  if (b.p) CK(cudaFree(b.p));
  b.p = nullptr;
  b.cap = 0;
  ctx->generation++;                                       // a captured graph holds the old address
  size_t bytes = std::max<size_t>(need, 1) * sizeof(double);
  cudaError_t e = cudaMalloc((void**)&b.p, bytes);
  if (e != cudaSuccess) return fail(ctx, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
  b.cap = need;
  return 0;
}

// ---- profiling helpers ------------------------------------------------------------------
struct ProfScope {
  maf_ctx* ctx;
  int slot;
  cudaEvent_t stop = nullptr;
  ProfScope(maf_ctx* c, int s) : ctx(c), slot(s) {
    ctx->launch_count++;
    ctx->slots[slot].launches++;
    if (!ctx->prof) return;
    ProfSlot& ps = ctx->slots[slot];
    if (ps.used + 2 > ps.ev.size()) {
      cudaEvent_t a, b;
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      ps.ev.push_back(a);
      ps.ev.push_back(b);
    }
    cudaEventRecord(ps.ev[ps.used], ctx->stream);
    stop = ps.ev[ps.used + 1];
    ps.used += 2;
  }
  ~ProfScope() {
    if (stop) cudaEventRecord(stop, ctx->stream);
  }
};

static void prof_collect(maf_ctx* ctx) {
  cudaStreamSynchronize(ctx->stream);
  for (int s = 0; s < MAF_PROF_NSLOTS; ++s) {
    ProfSlot& ps = ctx->slots[s];
    for (size_t i = 0; i + 1 < ps.used; i += 2) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, ps.ev[i], ps.ev[i + 1]) == cudaSuccess) ps.ms += ms;
    }
    ps.used = 0;
  }
}

// ---- small setup kernels ------------------------------------------------------------------
__global__ void tril_identity_kernel(double* L, double* Y, int npad) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y;
  if (j >= npad) return;
  if (j > i) L[(size_t)i * npad + j] = 0.0;
  Y[(size_t)i * npad + j] = (i == j) ? 1.0 : 0.0;
}

__global__ void transpose_kernel(const double* __restrict__ A, double* __restrict__ B, int n) {
  __shared__ double t[32][33];
  int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) t[r][threadIdx.x] = A[(size_t)(y0 + r) * n + x];
  __syncthreads();
  int xo = blockIdx.y * 32 + threadIdx.x, yo0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) B[(size_t)(yo0 + r) * n + xo] = t[threadIdx.x][r];
}

// B[C][R] = A[R][C]^T, R and C multiples of 32
__global__ void transpose_rect_kernel(const double* __restrict__ A, double* __restrict__ B, int R, int Cc) {
  __shared__ double t[32][33];
  int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) t[r][threadIdx.x] = A[(size_t)(y0 + r) * Cc + x];
  __syncthreads();
  int xo = blockIdx.y * 32 + threadIdx.x, yo0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) B[(size_t)(yo0 + r) * R + xo] = t[threadIdx.x][r];
}

// gamma_i = sum_{k<=i} Linv[i][k] rho_k ; one warp per row.
__global__ void trmv_lower_kernel(const double* __restrict__ Linv, const double* __restrict__ rho, int npad,
                                  double* __restrict__ gamma) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= npad) return;
  double a = 0.0;
  for (int k = lane; k <= row; k += 32) a += Linv[(size_t)row * npad + k] * rho[k];
  a = warp_sum(a);
  if (lane == 0) gamma[row] = a;
}

__global__ void diag_extract_kernel(const double* __restrict__ Sig, int S, int q, double* __restrict__ var) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= S * q) return;
  int s = t / q, i = t - s * q;
  var[t] = Sig[(size_t)s * q * q + i * q + i];
}

// ---- dispatch helpers ---------------------------------------------------------------------
#define DISPATCH_Q(q, ...)                                                         \
  switch (q) {                                                                     \
    case 1: { constexpr int Q = 1; __VA_ARGS__; } break;                           \
    case 2: { constexpr int Q = 2; __VA_ARGS__; } break;                           \
    case 3: { constexpr int Q = 3; __VA_ARGS__; } break;                           \
    case 4: { constexpr int Q = 4; __VA_ARGS__; } break;                           \
    case 5: { constexpr int Q = 5; __VA_ARGS__; } break;                           \
    case 6: { constexpr int Q = 6; __VA_ARGS__; } break;                           \
    case 7: { constexpr int Q = 7; __VA_ARGS__; } break;                           \
    case 8: { constexpr int Q = 8; __VA_ARGS__; } break;                           \
    case 9: { constexpr int Q = 9; __VA_ARGS__; } break;                           \
    case 10: { constexpr int Q = 10; __VA_ARGS__; } break;                         \
    case 11: { constexpr int Q = 11; __VA_ARGS__; } break;                         \
    case 12: { constexpr int Q = 12; __VA_ARGS__; } break;                         \
    case 13: { constexpr int Q = 13; __VA_ARGS__; } break;                         \
    case 14: { constexpr int Q = 14; __VA_ARGS__; } break;                         \
    case 15: { constexpr int Q = 15; __VA_ARGS__; } break;                         \
    case 16: { constexpr int Q = 16; __VA_ARGS__; } break;                         \
    default: return fail(ctx, "q=%d outside 1..%d", q, MAF_MAX_Q);                 \
  }

#define DISPATCH_D(d, ...)                                                         \
  switch (d) {                                                                     \
    case 1: { constexpr int D = 1; __VA_ARGS__; } break;                           \
    case 2: { constexpr int D = 2; __VA_ARGS__; } break;                           \
    case 3: { constexpr int D = 3; __VA_ARGS__; } break;                           \
    case 4: { constexpr int D = 4; __VA_ARGS__; } break;                           \
    case 5: { constexpr int D = 5; __VA_ARGS__; } break;                           \
    case 6: { constexpr int D = 6; __VA_ARGS__; } break;                           \
    case 7: { constexpr int D = 7; __VA_ARGS__; } break;                           \
    case 8: { constexpr int D = 8; __VA_ARGS__; } break;                           \
    default: { constexpr int D = 0; __VA_ARGS__; } break;                          \
  }

// compile-time input dimension D and kernel id KID (the two FP64-pipe-bound covariance kernels)
#define DISPATCH_DK(d, kid, ...)                                                   \
  switch (kid) {                                                                   \
    case MAF_KERNEL_SE: { constexpr int KID = MAF_KERNEL_SE; DISPATCH_D(d, __VA_ARGS__) } break;             \
    case MAF_KERNEL_MATERN52: { constexpr int KID = MAF_KERNEL_MATERN52; DISPATCH_D(d, __VA_ARGS__) } break; \
    default: { constexpr int KID = MAF_KERNEL_MATERN32; DISPATCH_D(d, __VA_ARGS__) } break;                  \
  }

// ... and compile-time digit count NSC for the FP64-grade configuration (7 planes); other counts at run time
#define DISPATCH_DKN(d, kid, ns, ...)                                              \
  if ((ns) == 7) { constexpr int NSC = 7; DISPATCH_DK(d, kid, __VA_ARGS__) }       \
  else if ((ns) == 6) { constexpr int NSC = 6; DISPATCH_DK(d, kid, __VA_ARGS__) }  \
  else { constexpr int NSC = 0; DISPATCH_DK(d, kid, __VA_ARGS__) }

// GEMM variants (selected by MAF_GEMM_VARIANT at context creation; 0 is the default)
template <int BM, int BN, int BK, int STAGES, int WJ, int WR, int MINB>
static int launch_gemm_cfg(maf_ctx* ctx, int vbit, int J, int N, const double* A, int lda, const double* T, int ldt, double* C,
                           int ldc, int K, int tri, int epi) {
  using Cfg = GemmCfg<BM, BN, BK, STAGES, WJ, WR, MINB>;
  if (!(ctx->attr_gemm_mask & (1u << vbit))) {
    CK(cudaFuncSetAttribute(gemm_nt_f64_kernel<BM, BN, BK, STAGES, WJ, WR, MINB>,
                            cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    CK(cudaFuncSetAttribute(gemm_nt_f64_kernel<BM, BN, BK, STAGES, WJ, WR, MINB>,
                            cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    ctx->attr_gemm_mask |= 1u << vbit;
  }
  dim3 grid(N / BN, J / BM);
  gemm_nt_f64_kernel<BM, BN, BK, STAGES, WJ, WR, MINB><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, ctx->stream>>>(
      A, lda, T, ldt, C, ldc, K, tri, epi);
  return 0;
}

static int launch_gemm(maf_ctx* ctx, int slot, int J, int N, int K, const double* A, int lda,
                       const double* T, int ldt, double* C, int ldc, int tri, int epi = 0) {
  if (J % 128 || N % 128 || K % 128) return fail(ctx, "gemm dims must be multiples of 128");
  if (tri != 0 && K != N) return fail(ctx, "triangular gemm needs K == N");
  ProfScope ps(ctx, slot);
  int rc;
  switch (ctx->gemm_variant) {
    // <BM, BN, BK, STAGES, WJ, WR, CTAs/SM>; measured on B200 (profiles/gemm_variants_r01.jsonl)
    case 1: rc = launch_gemm_cfg<128, 128, 16, 4, 2, 4, 1>(ctx, 1, J, N, A, lda, T, ldt, C, ldc, K, tri, epi); break;  // v0: 32.0 TF
    case 2: rc = launch_gemm_cfg<128, 64, 16, 3, 2, 2, 2>(ctx, 2, J, N, A, lda, T, ldt, C, ldc, K, tri, epi); break;   // 35.0 TF
    case 3: rc = launch_gemm_cfg<64, 128, 16, 3, 1, 4, 2>(ctx, 3, J, N, A, lda, T, ldt, C, ldc, K, tri, epi); break;   // 33.2 TF
    case 4: rc = launch_gemm_cfg<128, 64, 32, 2, 2, 2, 2>(ctx, 4, J, N, A, lda, T, ldt, C, ldc, K, tri, epi); break;   // 35.0 TF
    case 5: rc = launch_gemm_cfg<64, 64, 16, 3, 2, 2, 3>(ctx, 5, J, N, A, lda, T, ldt, C, ldc, K, tri, epi); break;    // 35.2 TF
    case 6: rc = launch_gemm_cfg<64, 64, 32, 2, 2, 2, 3>(ctx, 6, J, N, A, lda, T, ldt, C, ldc, K, tri, epi); break;    // BK 32
    case 7: rc = launch_gemm_cfg<64, 128, 16, 3, 2, 4, 2>(ctx, 7, J, N, A, lda, T, ldt, C, ldc, K, tri, epi); break;   // 8 warps of 32x32, 2/SM
    default: rc = launch_gemm_cfg<64, 64, 16, 2, 2, 2, 4>(ctx, 0, J, N, A, lda, T, ldt, C, ldc, K, tri, epi); break;   // 4 CTAs/SM: 35.7 TF
  }
  if (rc) return rc;
  CKL();
  return 0;
}

// ---- INT8 digit-sliced contraction ------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int oz_make_tmap(maf_ctx* ctx, CUtensorMap* tm, const int8_t* base, int rows, int K, int ns, int bk, int box_rows = 128) {
  static PFN_encodeTiled encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
      return fail(ctx, "cuTensorMapEncodeTiled is not available from the driver");
    encode = (PFN_encodeTiled)fn;
  }
  cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)ns};
  cuuint64_t gstr[2] = {(cuuint64_t)K, (cuuint64_t)K * (cuuint64_t)rows};
  cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ctx, "cuTensorMapEncodeTiled failed with %d", (int)r);
  return 0;
}

// X[R][K] (ld ldx) -> digit planes [ns][Rpad][K] + row scales
static int oz_slice(maf_ctx* ctx, int slot, const double* X, int ldx, int R, int Rpad, int K, int ns, double fixed_max,
                    Buf& planes, Buf& scales) {
  const size_t bytes = (size_t)ns * Rpad * K;
  if (ensure(ctx, planes, (bytes + 7) / 8) || ensure(ctx, scales, (size_t)Rpad)) return 1;
  ProfScope ps(ctx, slot);
  ozaki_slice_rows_kernel<<<Rpad, OZ_SLICE_THREADS, 0, ctx->stream>>>(X, ldx, R, Rpad, K, ns, fixed_max,
                                                                    (int8_t*)planes.p, scales.p);
  CKL();
  return 0;
}

// Balanced static schedule of the CTA-pair kernel (OzParams::tile_list): list scheduling of the supertiled tile order
// onto `pairs` CTA pairs with cost kchunks + overhead.  Cached per launch shape.
static int oz_schedule(maf_ctx* ctx, int nN, int nJ2, int K, int tri, int sj, int pairs, const int** d_list, const int** d_begin) {
  const int groups = std::max(1, std::min(ctx->oz_groups, nN));
  for (const auto& sc : ctx->oz_scheds)
    if (sc.nN == nN && sc.nJ2 == nJ2 && sc.tri == tri && sc.sj == sj && sc.pairs == pairs && sc.groups == groups) {
      *d_list = sc.d_list;
      *d_begin = sc.d_begin;
      return 0;
    }
  // k-chunks of a column block (same arithmetic as oz3_tile), column blocks heaviest first
  auto kchunks = [&](int rb) {
    const int r0 = rb * OZ_BN, c_begin = (tri == 2) ? r0 : 0, c_end = (tri == 1) ? std::min(K, r0 + OZ_BN) : K;
    return (c_end - c_begin) / OZ3_BK;
  };
  std::vector<int> cols(nN);
  for (int i = 0; i < nN; ++i) cols[i] = (tri == 1) ? (nN - 1 - i) : i;
  // column groups of equal T-plane bytes: a group's T planes stay in L2 while every row block streams past them
  // (groups == 1: the supertiled order of oz3_tile).  A is then read `groups` times, T about once.
  long long total = 0;
  for (int rb : cols) total += kchunks(rb);
  std::vector<int> gend;                                         // end index (in cols) of every group
  {
    long long acc = 0;
    int g = 1;
    for (int i = 0; i < nN; ++i) {
      acc += kchunks(cols[i]);
      if (g < groups && acc * groups >= total * g) gend.push_back(i + 1), ++g;
    }
    gend.push_back(nN);
  }
  std::vector<int> order;                                        // explicit tile entries: row block * nN + column block
  order.reserve((size_t)nN * nJ2);
  for (size_t g = 0, c0 = 0; g < gend.size(); c0 = gend[g], ++g)
    for (int st = 0; st * sj < nJ2; ++st)                        // super-rows of sj row blocks; column outer, row inner
      for (int ci = (int)c0; ci < gend[g]; ++ci)
        for (int jb = st * sj; jb < std::min(nJ2, (st + 1) * sj); ++jb) order.push_back(jb * nN + cols[ci]);
  std::vector<std::vector<int>> per_pair(pairs);
  std::vector<double> busy(pairs, 0.0);
  for (int e : order) {                                          // list scheduling: next tile to the pair that frees first
    const double cost = (double)kchunks(e % nN) + ctx->oz_sched_ovh;
    int best = 0;
    for (int c = 1; c < pairs; ++c)
      if (busy[c] < busy[best]) best = c;
    per_pair[best].push_back(e | OZ_TILE_EXPLICIT);
    busy[best] += cost;
  }
  std::vector<int> list, begin(pairs + 1, 0);
  list.reserve(order.size());
  for (int c = 0; c < pairs; ++c) {
    begin[c] = (int)list.size();
    list.insert(list.end(), per_pair[c].begin(), per_pair[c].end());
  }
  begin[pairs] = (int)list.size();
  maf_ctx::OzSched sc{nN, nJ2, tri, sj, pairs, groups, nullptr, nullptr};
  CK(cudaMalloc((void**)&sc.d_list, sizeof(int) * list.size()));
  CK(cudaMalloc((void**)&sc.d_begin, sizeof(int) * begin.size()));
  CK(cudaMemcpyAsync(sc.d_list, list.data(), sizeof(int) * list.size(), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(sc.d_begin, begin.data(), sizeof(int) * begin.size(), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));                        // the host vectors go out of scope
  ctx->oz_scheds.push_back(sc);
  *d_list = sc.d_list;
  *d_begin = sc.d_begin;
  return 0;
}

static int launch_i8gemm(maf_ctx* ctx, int slot, int J, int N, int K, const Buf& Ap, const Buf& sA, const Buf& Tp,
                         const Buf& sT, double* C, int ldc, int tri, int ns, int lmax, double* rowmax = nullptr,
                         bool* rowmax_done = nullptr, const int* nrows_dev = nullptr, int nrows_mul = 1,
                         const double* row_add = nullptr, const double* col_add = nullptr) {
  if (rowmax_done) *rowmax_done = false;
  if (J % 128 || N % 128 || K % 128) return fail(ctx, "i8 gemm dims must be multiples of 128");
  if (tri != 0 && K != N) return fail(ctx, "triangular gemm needs K == N");
  if (ns < 1 || ns > OZ_MAXS || lmax < 2 || lmax > 2 * ns) return fail(ctx, "bad digit configuration ns=%d lmax=%d", ns, lmax);
  if (K > 8192) return fail(ctx, "i8 gemm: K=%d would overflow the int32 level accumulators (max 8192)", K);
  if (!ctx->attr_oz) {
    CK(cudaFuncSetAttribute(ozaki_i8_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM_BYTES));
#define OZ3_ATTR(NS_, LM_) \
  CK(cudaFuncSetAttribute(ozaki_i8_gemm_v3_kernel<NS_, LM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, Oz3<NS_, LM_>::SMEM_BYTES))
    OZ3_ATTR(7, 8); OZ3_ATTR(6, 7); OZ3_ATTR(5, 6); OZ3_ATTR(4, 5); OZ3_ATTR(3, 4);
#undef OZ3_ATTR
#define OZ4_ATTR(NS_, LM_) \
  CK(cudaFuncSetAttribute(ozaki_i8_gemm_v4_kernel<NS_, LM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, Oz4<NS_, LM_>::SMEM_BYTES))
    OZ4_ATTR(7, 8); OZ4_ATTR(6, 7); OZ4_ATTR(5, 6); OZ4_ATTR(4, 5); OZ4_ATTR(3, 4);
#undef OZ4_ATTR
#define OZ5_ATTR(NS_, LM_) \
  CK(cudaFuncSetAttribute(ozaki_i8_gemm_v4_kernel<NS_, LM_, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Oz4<NS_, LM_>::SMEM_BYTES))
    OZ5_ATTR(7, 8); OZ5_ATTR(6, 7); OZ5_ATTR(5, 6); OZ5_ATTR(4, 5); OZ5_ATTR(3, 4);
#undef OZ5_ATTR
#ifdef MAF_OZ_DIAG
#define OZ3_DATTR(D_) CK(cudaFuncSetAttribute(ozaki_i8_gemm_v3_kernel<7, 8, D_>, cudaFuncAttributeMaxDynamicSharedMemorySize, Oz3<7, 8>::SMEM_BYTES))
    OZ3_DATTR(1); OZ3_DATTR(2); OZ3_DATTR(3); OZ3_DATTR(4); OZ3_DATTR(5); OZ3_DATTR(6); OZ3_DATTR(7);
#undef OZ3_DATTR
#endif
    ctx->attr_oz = true;
  }
  const bool v3ok = lmax == ns + 1 && ns >= 3 && ns <= 7;
  int variant = ctx->oz_variant;
  const bool stash = variant == 5;                                // variant 5 = CTA pairs + register-stash epilogue
  if (variant == 5) variant = 4;
  if (variant == 4 && (!v3ok || J % 256 != 0)) variant = 3;       // CTA pairs need 256-row blocks
  if (variant != 4 && variant != 3) variant = 1;
  if (variant == 3 && !v3ok) variant = 1;
  const int bk = variant == 1 ? OZ_BK : OZ3_BK;
  CUtensorMap tmA, tmT;
  if (oz_make_tmap(ctx, &tmA, (const int8_t*)Ap.p, J, K, ns, bk) ||
      oz_make_tmap(ctx, &tmT, (const int8_t*)Tp.p, N, K, ns, bk, variant == 4 ? 64 : 128))
    return 1;
  static const unsigned long long pol[3] = {OZ3_L2_NORMAL, OZ3_L2_EVICT_FIRST, OZ3_L2_EVICT_LAST};
  OzParams prm{K, ns, lmax, tri, ldc, ctx->oz_sj, pol[ctx->oz_hint_a % 3], pol[ctx->oz_hint_t % 3], ctx->oz_pf, nullptr, ctx->oz_dbg,
               nullptr, nullptr, nullptr, nullptr, 1, nullptr, nullptr};
  if (row_add) {
    if (!(stash && variant == 4)) return fail(ctx, "the rank-one epilogue term needs the register-stash CTA-pair kernel");
    prm.row_add = row_add;
    prm.col_add = col_add;
  }
  if (nrows_dev) {
    if (variant != 4) return fail(ctx, "a device-side row count needs the CTA-pair contraction kernel");
    prm.nrows_dev = nrows_dev;
    prm.nrows_mul = nrows_mul;
  }
  if (rowmax && stash && variant == 4) {                           // only the register-stash epilogue emits row maxima
    CK(cudaMemsetAsync(rowmax, 0, sizeof(double) * J, ctx->stream));
    prm.rowmax = rowmax;
    if (rowmax_done) *rowmax_done = true;
  }
  dim3 grid(N / OZ_BN, J / OZ_BM);
  if (!ctx->num_sms) CK(cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, ctx->device));
  const int num_sms = ctx->num_sms;
  const long long ntiles = (long long)(N / OZ_BN) * (J / OZ_BM);
  dim3 pgrid((unsigned)(ntiles < num_sms ? ntiles : num_sms));   // persistent: one CTA per SM
  ProfScope ps(ctx, slot);
  if (variant == 1) {
    ozaki_i8_gemm_kernel<<<grid, OZ_THREADS, OZ_SMEM_BYTES, ctx->stream>>>(tmA, tmT, sA.p, sT.p, C, prm);
#ifdef MAF_OZ_DIAG
  } else if (variant == 3 && ns == 7 && ctx->oz_dbg > 0 && ctx->oz_dbg < 8) {
#define OZ3_DGO(D_) case D_: ozaki_i8_gemm_v3_kernel<7, 8, D_><<<pgrid, OZ3_THREADS, Oz3<7, 8>::SMEM_BYTES, ctx->stream>>>(tmA, tmT, sA.p, sT.p, C, prm, N / OZ_BN, J / OZ_BM); break
    switch (ctx->oz_dbg) { OZ3_DGO(1); OZ3_DGO(2); OZ3_DGO(3); OZ3_DGO(4); OZ3_DGO(5); OZ3_DGO(6); OZ3_DGO(7); }
#undef OZ3_DGO
#endif
  } else if (variant == 4) {
    const long long ntiles2 = (long long)(N / OZ_BN) * (J / 256);
    const long long pairs = std::min<long long>(ntiles2, num_sms / 2);
    dim3 cgrid((unsigned)(2 * pairs));                           // persistent: one CTA pair per TPC
    if (ctx->oz_sched && ntiles2 > pairs &&
        oz_schedule(ctx, N / OZ_BN, J / 256, K, tri, ctx->oz_sj, (int)pairs, &prm.tile_list, &prm.tile_begin))
      return 1;
#ifdef MAF_OZ_DIAG
    static long long* dstats = nullptr;
    if (getenv("MAF_OZ_STATS")) {
      if (!dstats) CK(cudaMalloc((void**)&dstats, sizeof(long long) * 8 * 128));
      CK(cudaMemsetAsync(dstats, 0, sizeof(long long) * 8 * 128, ctx->stream));
      prm.stats = dstats;
    }
#endif
#define OZ4_GO(NS_, LM_)                                                                                                     \
  case NS_:                                                                                                                  \
    ozaki_i8_gemm_v4_kernel<NS_, LM_><<<cgrid, OZ4_THREADS, Oz4<NS_, LM_>::SMEM_BYTES, ctx->stream>>>(tmA, tmT, sA.p, sT.p, C, prm, N / OZ_BN, J / 256); \
    break
#define OZ5_GO(NS_, LM_)                                                                                                     \
  case NS_:                                                                                                                  \
    ozaki_i8_gemm_v4_kernel<NS_, LM_, 1><<<cgrid, OZ4_THREADS, Oz4<NS_, LM_>::SMEM_BYTES, ctx->stream>>>(tmA, tmT, sA.p, sT.p, C, prm, N / OZ_BN, J / 256); \
    break
    if (stash) {
      switch (ns) {
        OZ5_GO(7, 8);
        OZ5_GO(6, 7);
        OZ5_GO(5, 6);
        OZ5_GO(4, 5);
        OZ5_GO(3, 4);
      }
    } else {
      switch (ns) {
        OZ4_GO(7, 8);
        OZ4_GO(6, 7);
        OZ4_GO(5, 6);
        OZ4_GO(4, 5);
        OZ4_GO(3, 4);
      }
    }
#undef OZ5_GO
#undef OZ4_GO
#ifdef MAF_OZ_DIAG
    if (prm.stats) {
      std::vector<long long> h(8 * 128);
      CK(cudaStreamSynchronize(ctx->stream));
      CK(cudaMemcpy(h.data(), prm.stats, sizeof(long long) * 8 * 128, cudaMemcpyDeviceToHost));
      double tot = 0, full = 0, acc = 0, epi = 0, f[4] = {0, 0, 0, 0}, tmin = 1e30, tmax = 0;
      for (long long c = 0; c < pairs; ++c) {
        tmin = std::min(tmin, (double)h[8 * c]);
        tmax = std::max(tmax, (double)h[8 * c]);
        tot += h[8 * c]; full += h[8 * c + 1]; acc += h[8 * c + 2]; epi += h[8 * c + 3];
        for (int u = 0; u < 4; ++u) f[u] += h[8 * c + 4 + u];
      }
      fprintf(stderr, "[oz4 stats] busiest/mean pair %.4f idlest/mean %.4f\n", tmax * pairs / tot, tmin * pairs / tot);
      fprintf(stderr, "[oz4 stats] tri=%d pairs=%lld tiles/pair=%.1f  issuer cycles: total %.0f  wait-operands %.1f%% (pass0 first chunk %.1f%%, later %.1f%%; pass1 first %.1f%%, later %.1f%%)  wait-epilogue %.1f%%  | epilogue body (1 warp) %.1f%% of total\n",
              tri, pairs, (double)ntiles2 / pairs, tot / pairs, 100 * full / tot, 100 * f[0] / tot, 100 * f[1] / tot, 100 * f[2] / tot,
              100 * f[3] / tot, 100 * acc / tot, 100 * epi / tot);
    }
#endif
  } else {
#define OZ3_GO(NS_, LM_)                                                                                                     \
  case NS_:                                                                                                                  \
    ozaki_i8_gemm_v3_kernel<NS_, LM_><<<pgrid, OZ3_THREADS, Oz3<NS_, LM_>::SMEM_BYTES, ctx->stream>>>(tmA, tmT, sA.p, sT.p, C, prm, N / OZ_BN, J / OZ_BM); \
    break
    switch (ns) {
      OZ3_GO(7, 8);
      OZ3_GO(6, 7);
      OZ3_GO(5, 6);
      OZ3_GO(4, 5);
      OZ3_GO(3, 4);
    }
#undef OZ3_GO
  }
  CKL();
  return 0;
}

// ---- training-set factorisation (factor.cuh): blocked Cholesky + triangular inverse by block doubling ------------
static int factorize(maf_ctx* ctx) {
  const int npad = ctx->npad, nb = npad / FAC_NB;
  if (!ctx->attr_chol) {
    CK(cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FAC_SMEM_BYTES));
    ctx->attr_chol = true;
  }
  double *dinv = nullptr, *panel = nullptr, *scratch = nullptr;
  int* dinfo = nullptr;
  const size_t half = (size_t)((nb + 1) / 2) * FAC_NB;            // largest block of the doubling recursion
  CK(cudaMalloc((void**)&dinv, sizeof(double) * (size_t)nb * FAC_NB * FAC_NB));
  CK(cudaMalloc((void**)&panel, sizeof(double) * (size_t)npad * FAC_NB));
  CK(cudaMalloc((void**)&scratch, sizeof(double) * half * half));
  CK(cudaMalloc((void**)&dinfo, sizeof(int)));
  CK(cudaMemsetAsync(dinfo, 0, sizeof(int), ctx->stream));
  double* L = ctx->L0;
  int rc = 0;
  for (int k = 0; k < nb && rc == 0; ++k) {
    const int k0 = k * FAC_NB, rows = npad - k0 - FAC_NB;
    double* diag = L + (size_t)k0 * npad + k0;
    double* dk = dinv + (size_t)k * FAC_NB * FAC_NB;
    {
      ProfScope ps(ctx, MAF_PROF_CROSS_FWD);
      chol_diag_kernel<<<1, 256, FAC_SMEM_BYTES, ctx->stream>>>(diag, npad, dk, dinfo, k0);
      if (cudaGetLastError() != cudaSuccess) rc = fail(ctx, "chol_diag_kernel launch failed");
    }
    if (rows <= 0 || rc) continue;
    double* below = L + (size_t)(k0 + FAC_NB) * npad + k0;       // A_ik, i > k
    // panel = A_ik D_k^-T  (T = D_k^-1 lower-triangular), then back into place
    rc = launch_gemm(ctx, MAF_PROF_GEMM_FWD, rows, FAC_NB, FAC_NB, below, npad, dk, FAC_NB, panel, FAC_NB, 1);
    if (rc) break;
    {
      ProfScope ps(ctx, MAF_PROF_CROSS_FWD);
      dim3 grid((FAC_NB + 127) / 128, rows);
      copy_block_kernel<<<grid, 128, 0, ctx->stream>>>(panel, FAC_NB, below, npad, rows, FAC_NB);
    }
    // trailing update A_ij -= L_ik L_jk^T (lower tiles)
    rc = launch_gemm(ctx, MAF_PROF_GEMM_FWD, rows, rows, FAC_NB, below, npad, below, npad,
                     L + (size_t)(k0 + FAC_NB) * npad + k0 + FAC_NB, npad, 0, 1);
  }
  int info = 0;
  if (rc == 0) {
    if (cudaMemcpyAsync(&info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess)
      rc = fail(ctx, "Cholesky factorisation failed on the device: %s", cudaGetErrorString(cudaGetLastError()));
    else if (info != 0)
      rc = fail(ctx, "Cholesky decomposition was not successful (leading minor %d of K00 is not positive definite)", info);
  }
  if (rc == 0) {
    {
      ProfScope ps(ctx, MAF_PROF_CROSS_FWD);
      dim3 grid((npad + 255) / 256, npad);
      inverse_init_kernel<<<grid, 256, 0, ctx->stream>>>(dinv, L, ctx->Linv, ctx->LinvT, npad);
    }
    for (int b = FAC_NB; b < npad && rc == 0; b *= 2) {
      for (int o = 0; o + b < npad && rc == 0; o += 2 * b) {
        const int b2 = std::min(b, npad - (o + b));
        const double* L21 = L + (size_t)(o + b) * npad + o;
        const double* Y11 = ctx->LinvT + (size_t)o * npad + o;
        double* X22 = ctx->Linv + (size_t)(o + b) * npad + (o + b);
        double* X21 = ctx->Linv + (size_t)(o + b) * npad + o;
        // scratch[b][b2] = (L21 X11)^T = Y11 L21^T ;  X21 = -X22 (L21 X11)
        rc = launch_gemm(ctx, MAF_PROF_GEMM_FWD, b, b2, b, Y11, npad, L21, npad, scratch, b2, 0);
        if (rc) break;
        rc = launch_gemm(ctx, MAF_PROF_GEMM_FWD, b2, b, b2, X22, npad, scratch, b2, X21, npad, 0, 2);
        if (rc) break;
        ProfScope ps(ctx, MAF_PROF_CROSS_FWD);
        dim3 grid(b / 32, b2 / 32), block(32, 8);
        transpose_block_kernel<<<grid, block, 0, ctx->stream>>>(X21, npad, ctx->LinvT + (size_t)o * npad + (o + b), npad);
      }
    }
    if (rc == 0 && cudaStreamSynchronize(ctx->stream) != cudaSuccess)
      rc = fail(ctx, "triangular inverse failed on the device: %s", cudaGetErrorString(cudaGetLastError()));
  }
  cudaFree(dinv);
  cudaFree(panel);
  cudaFree(scratch);
  cudaFree(dinfo);
  return rc;
}

// ---- life cycle ---------------------------------------------------------------------------
extern "C" const char* maf_version(void) { return MAF_VERSION_STR; }

extern "C" const char* maf_last_error(const maf_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int maf_create(int device, int dtype, maf_ctx** out) {
  if (!out) return 1;
  *out = nullptr;
  maf_ctx* ctx = new maf_ctx();
  *out = ctx;   // returned even on failure so that the caller can read the error
  ctx->device = device;
  ctx->dtype = dtype;
  if (dtype != MAF_F64 && dtype != MAF_F32) return fail(ctx, "unknown dtype %d", dtype);
  if (dtype == MAF_F32) {
    // FP32-grade mode (the reference's dtype='float32', 1e-4 bar): the contractions keep 5 digit planes per operand
    // (39-bit fixed point, 15 exact products instead of 28; 4 planes measured 1.4e-5 on the posterior mean, too
    // close to the bar); everything else, and the ABI, stays FP64.
    ctx->oz_ns = 5;
    ctx->oz_lmax = 6;
    ctx->oz_ns_bwd = 5;
    ctx->oz_lmax_bwd = 6;
    ctx->oz_ns_fwd = 5;
    ctx->lsw_tol = 2.5e-5;         // level switch: 4 planes first (10 products), a quarter of this mode's 1e-4 bar
  }
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(ctx, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
  CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  if (const char* gv = getenv("MAF_GEMM_VARIANT")) ctx->gemm_variant = atoi(gv);
  if (const char* gi = getenv("MAF_GEMM_I8")) ctx->gemm_i8 = atoi(gi);
  if (const char* ov = getenv("MAF_OZ_VARIANT")) ctx->oz_variant = atoi(ov);
  if (const char* od = getenv("MAF_OZ_DBG")) ctx->oz_dbg = atoi(od);
  if (const char* of = getenv("MAF_OZ_FUSE")) ctx->oz_fuse = atoi(of);
  if (const char* h = getenv("MAF_OZ_HINTS")) {
    if (h[0]) ctx->oz_hint_a = h[0] - '0';
    if (h[0] && h[1]) ctx->oz_hint_t = h[1] - '0';
  }
  if (const char* pf = getenv("MAF_OZ_PF")) ctx->oz_pf = atoi(pf);
  if (const char* gr2 = getenv("MAF_G_ROUTE")) ctx->g_route = atoi(gr2);
  if (const char* sc = getenv("MAF_OZ_SCHED")) ctx->oz_sched = atoi(sc);
  if (const char* so = getenv("MAF_OZ_SCHED_OVH")) ctx->oz_sched_ovh = atof(so);
  if (const char* sg = getenv("MAF_OZ_GROUPS")) ctx->oz_groups = atoi(sg);
  if (const char* gr = getenv("MAF_GRAPH")) ctx->graph_on = atoi(gr);
  if (const char* rm = getenv("MAF_OZ_ROWMAX")) ctx->oz_rowmax = atoi(rm);
  if (const char* ma = getenv("MAF_MU_ALPHA")) ctx->mu_alpha = atoi(ma);
  if (const char* vt = getenv("MAF_VBAR_TIGHT")) ctx->vbar_tight = atoi(vt);
  if (const char* lw = getenv("MAF_LEVEL_SWITCH")) ctx->lsw_on = atoi(lw);
  if (const char* pdd = getenv("MAF_PEND_DEDUPE")) ctx->pend_dedupe = atoi(pdd);
  if (const char* ck = getenv("MAF_CENTRE_K10")) ctx->centre_k10 = atoi(ck);
  if (const char* lt = getenv("MAF_LEVEL_SWITCH_TOL")) ctx->lsw_tol = atof(lt);
  if (const char* ln = getenv("MAF_LEVEL_SWITCH_MIN_N")) ctx->lsw_min_npad = atoi(ln);
  if (const char* nf = getenv("MAF_OZ_NS_FWD")) {
    const int v = atoi(nf);
    if (v >= 3 && v <= ctx->oz_ns) ctx->oz_ns_fwd = v;
  }
  if (const char* nb = getenv("MAF_OZ_NS_BWD")) {
    const int v = atoi(nb);
    if (v >= 3 && v <= ctx->oz_ns) ctx->oz_ns_bwd = v, ctx->oz_lmax_bwd = v + 1;
  }
  if (const char* sj = getenv("MAF_OZ_SJ")) ctx->oz_sj = atoi(sj) > 0 ? atoi(sj) : 1;
  CK(cudaMalloc((void**)&ctx->d_argmin_val, sizeof(double)));
  CK(cudaMalloc((void**)&ctx->d_argmin_idx, sizeof(long long)));
  const char* cr = getenv("MAF_CHUNK_ROWS");
  if (cr) {
    long v = atol(cr);
    if (v >= 128) ctx->chunk_rows = (size_t)v / 128 * 128;
  }
  return 0;
}

extern "C" int maf_destroy(maf_ctx* ctx) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  double* raw[] = {ctx->X0s, ctx->L0, ctx->Linv, ctx->LinvT, ctx->gamma, ctx->alpha, ctx->linv_rowsum, ctx->d_argmin_val};
  for (double* b : raw)
    if (b) cudaFree(b);
  Buf* bufs[] = {&ctx->K10, &ctx->Vt, &ctx->dX, &ctx->dz, &ctx->dw, &ctx->dinc, &ctx->dmu, &ctx->dSig, &ctx->dchol,
                 &ctx->dmubar, &ctx->dSigbar, &ctx->dloss, &ctx->dgrad, &ctx->adam_X, &ctx->adam_m, &ctx->adam_v,
                 &ctx->adam_g, &ctx->adam_z, &ctx->fGamma, &ctx->fGammaT, &ctx->fmeans, &ctx->flevels, &ctx->fVG,
                 &ctx->fmubar, &ctx->ozA, &ctx->ozTL, &ctx->ozTU, &ctx->ozsA, &ctx->ozsTL, &ctx->ozsTU, &ctx->adam_zcur,
                 &ctx->lb_X, &ctx->lb_ft, &ctx->lb_gt, &ctx->lb_x, &ctx->lb_f, &ctx->lb_g, &ctx->lb_dir, &ctx->lb_step, &ctx->lb_gd,
                 &ctx->lb_hs, &ctx->lb_hy, &ctx->lb_rho, &ctx->lb_trace,
                 &ctx->adam_coef, &ctx->adam_trace, &ctx->ozVmax, &ctx->mu_part, &ctx->Vt2, &ctx->rowmax2, &ctx->row_add, &ctx->Vp, &ctx->mu_part_p, &ctx->rowmax_p, &ctx->row_add_fwd, &ctx->Gk};
  for (Buf* b : bufs)
    if (b->p) cudaFree(b->p);
  if (ctx->d_argmin_idx) cudaFree(ctx->d_argmin_idx);
  if (ctx->adam_graph) cudaGraphExecDestroy(ctx->adam_graph);
  if (ctx->adam_ctr) cudaFree(ctx->adam_ctr);
  for (void* b : {(void*)ctx->lsw_cnt, (void*)ctx->lsw_stats, (void*)ctx->lsw_selF, (void*)ctx->lsw_slotF, (void*)ctx->lsw_selB})
    if (b) cudaFree(b);
  if (ctx->lsw_host) cudaFreeHost(ctx->lsw_host);
  if (ctx->lb_int) cudaFree(ctx->lb_int);
  if (ctx->lb_active_host) cudaFreeHost(ctx->lb_active_host);
  for (auto& sc : ctx->oz_scheds) {
    cudaFree(sc.d_list);
    cudaFree(sc.d_begin);
  }
  for (auto& ps : ctx->slots)
    for (auto e : ps.ev) cudaEventDestroy(e);
  for (auto e : ctx->timer_ev)
    if (e) cudaEventDestroy(e);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return 0;
}

extern "C" void* maf_stream(maf_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

extern "C" int maf_sync(maf_ctx* ctx) {
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" void* maf_dev_alloc(maf_ctx* ctx, size_t bytes) {
  void* p = nullptr;
  cudaSetDevice(ctx->device);
  if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
    fail(ctx, "maf_dev_alloc(%zu) failed", bytes);
    return nullptr;
  }
  return p;
}
extern "C" int maf_dev_free(maf_ctx* ctx, void* p) {
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaFree(p));
  return 0;
}
extern "C" int maf_h2d(maf_ctx* ctx, void* dst, const void* src, size_t bytes) {
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
extern "C" int maf_d2h(maf_ctx* ctx, void* dst, const void* src, size_t bytes) {
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ---- model + training set -------------------------------------------------------------------
extern "C" int maf_set_model(maf_ctx* ctx, int kernel_id, int d, const double* lenscales, double amplitude,
                             double noise, double mean, double jitter) {
  if (!ctx) return 1;
  ctx->generation++;
  if (d < 1 || d > MAF_MAX_D) return fail(ctx, "d=%d outside 1..%d", d, MAF_MAX_D);
  if (kernel_id < 0 || kernel_id > 2) return fail(ctx, "unknown kernel id %d", kernel_id);
  ModelParams mp{};
  mp.kernel_id = kernel_id;
  mp.d = d;
  for (int c = 0; c < d; ++c) {
    if (!(lenscales[c] > 0.0)) return fail(ctx, "lengthscale %d must be positive", c);
    mp.inv_ls[c] = 1.0 / lenscales[c];
  }
  mp.amp2 = amplitude * amplitude;
  mp.noise = (noise == noise) ? noise : 0.0;
  mp.jitter = jitter;
  ctx->mp = mp;
  ctx->mean_param = mean;
  ctx->model_set = true;
  ctx->train_set = false;   // factors depend on the hyper-parameters
  return 0;
}

static int lsw_calibrate(maf_ctx* ctx, int n, const double* X0);

extern "C" int maf_set_train(maf_ctx* ctx, int n, const double* X0, const double* y0) {
  if (!ctx) return 1;
  ctx->generation++;
  if (!ctx->model_set) return fail(ctx, "maf_set_model must be called first");
  if (n < 1) return fail(ctx, "n must be >= 1");
  CK(cudaSetDevice(ctx->device));
  const int d = ctx->mp.d;
  const int npad = round_up(n, 128);
  ctx->train_set = false;
  if ((size_t)npad != ctx->cap_train || d != ctx->cap_train_d) {
    double** bufs[] = {&ctx->X0s, &ctx->L0, &ctx->Linv, &ctx->LinvT, &ctx->gamma, &ctx->alpha, &ctx->linv_rowsum};
    for (double** b : bufs) {
      if (*b) CK(cudaFree(*b));
      *b = nullptr;
    }
    ctx->cap_train = 0;
    size_t nn = (size_t)npad * npad;
    CK(cudaMalloc((void**)&ctx->X0s, sizeof(double) * (size_t)d * npad));
    CK(cudaMalloc((void**)&ctx->L0, sizeof(double) * nn));
    CK(cudaMalloc((void**)&ctx->Linv, sizeof(double) * nn));
    CK(cudaMalloc((void**)&ctx->LinvT, sizeof(double) * nn));
    CK(cudaMalloc((void**)&ctx->gamma, sizeof(double) * npad));
    CK(cudaMalloc((void**)&ctx->alpha, sizeof(double) * npad));
    CK(cudaMalloc((void**)&ctx->linv_rowsum, sizeof(double) * npad));
    ctx->cap_train = npad;
    ctx->cap_train_d = d;
  }
  ctx->n = n;
  ctx->npad = npad;

  // host-side scalars: prior mean (constant or empirical, gaussian_process.py:226-231) and min(y0)
  double mean = ctx->mean_param;
  if (!(mean == mean)) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i) acc += y0[i];
    mean = acc / n;
  }
  ctx->mean_val = mean;
  double ymin = y0[0];
  for (int i = 1; i < n; ++i) ymin = std::min(ymin, y0[i]);
  ctx->y_min = ymin;
  std::vector<double> rho(npad, 0.0);
  for (int i = 0; i < n; ++i) rho[i] = y0[i] - mean;

  // raw X0 -> LinvT scratch, rho -> gamma scratch (Linv row 0 reused below)
  double* scratch = ctx->LinvT;
  CK(cudaMemcpyAsync(scratch, X0, sizeof(double) * (size_t)n * d, cudaMemcpyHostToDevice, ctx->stream));
  {
    ProfScope ps(ctx, MAF_PROF_CROSS_FWD);
    scale_train_kernel<<<(npad + 255) / 256, 256, 0, ctx->stream>>>(scratch, n, npad, ctx->mp, ctx->X0s);
    CKL();
  }
  {
    ProfScope ps(ctx, MAF_PROF_CROSS_FWD);
    dim3 grid((npad + 255) / 256, npad);
    k00_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->X0s, n, npad, ctx->mp, ctx->L0);
    CKL();
  }
  // L0 (holding K00) -> its lower Cholesky factor, Linv = L0^-1, LinvT = L0^-T: own blocked kernels (factor.cuh)
  if (factorize(ctx)) return 1;
  // gamma = Linv rho  (rho staged in the GEMM-independent dSig-free scratch: reuse X0s? no -> temp)
  double* drho = nullptr;
  CK(cudaMalloc((void**)&drho, sizeof(double) * npad));
  CK(cudaMemcpyAsync(drho, rho.data(), sizeof(double) * npad, cudaMemcpyHostToDevice, ctx->stream));
  {
    ProfScope ps(ctx, MAF_PROF_CROSS_FWD);
    trmv_lower_kernel<<<(npad + 7) / 8, 256, 0, ctx->stream>>>(ctx->Linv, drho, npad, ctx->gamma);
    CKL();
  }
  // alpha = L0^-T gamma, then two steps of iterative refinement with the residual in twice the working precision
  {
    ProfScope ps(ctx, MAF_PROF_CROSS_FWD);
    trmv_upper_kernel<<<(npad + 7) / 8, 256, 0, ctx->stream>>>(ctx->LinvT, ctx->gamma, npad, ctx->alpha);
    CKL();
  }
  {
    double *res = nullptr, *tmp = nullptr;
    CK(cudaMalloc((void**)&res, sizeof(double) * npad));
    CK(cudaMalloc((void**)&tmp, sizeof(double) * npad));
    for (int it = 0; it < 2; ++it) {
      ProfScope ps(ctx, MAF_PROF_CROSS_FWD);
      alpha_residual_kernel<<<(npad + 7) / 8, 256, 0, ctx->stream>>>(ctx->X0s, n, npad, ctx->mp, drho, ctx->alpha, res);
      trmv_lower_kernel<<<(npad + 7) / 8, 256, 0, ctx->stream>>>(ctx->Linv, res, npad, tmp);
      trmv_upper_kernel<<<(npad + 7) / 8, 256, 0, ctx->stream>>>(ctx->LinvT, tmp, npad, res);
      axpy_kernel<<<(npad + 255) / 256, 256, 0, ctx->stream>>>(ctx->alpha, res, npad);
      CKL();
    }
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaFree(res));
    CK(cudaFree(tmp));
  }
  {
    // L0^-1 1 over the n real columns (add-back of the centred cross-covariance planes)
    std::vector<double> ones(npad, 0.0);
    for (int i = 0; i < n; ++i) ones[i] = 1.0;
    CK(cudaMemcpyAsync(drho, ones.data(), sizeof(double) * npad, cudaMemcpyHostToDevice, ctx->stream));
    ProfScope ps(ctx, MAF_PROF_CROSS_FWD);
    trmv_lower_kernel<<<(npad + 7) / 8, 256, 0, ctx->stream>>>(ctx->Linv, drho, npad, ctx->linv_rowsum);
    CKL();
    CK(cudaStreamSynchronize(ctx->stream));                      // `ones` goes out of scope
  }
  ctx->i8_ready = false;
  if (ctx->gemm_i8 && npad <= 8192) {   // int32 level accumulators hold K <= 8192; larger n keeps the DMMA contraction
    if (oz_slice(ctx, MAF_PROF_CROSS_FWD, ctx->Linv, npad, npad, npad, npad, ctx->oz_ns, 0.0, ctx->ozTL, ctx->ozsTL) ||
        oz_slice(ctx, MAF_PROF_CROSS_FWD, ctx->LinvT, npad, npad, npad, npad, ctx->oz_ns, 0.0, ctx->ozTU, ctx->ozsTU))
      return 1;
    ctx->i8_ready = true;
  }
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaFree(drho));
  ctx->train_set = true;
  ctx->adam_on = false;
  ctx->F = 0;
  return lsw_calibrate(ctx, n, X0);
}

extern "C" int maf_get_factors(maf_ctx* ctx, double* L0, double* Linv, double* gamma) {
  if (!ctx || !ctx->train_set) return fail(ctx, "no training set");
  CK(cudaSetDevice(ctx->device));
  const int n = ctx->n, npad = ctx->npad;
  if (L0) CK(cudaMemcpy2DAsync(L0, sizeof(double) * n, ctx->L0, sizeof(double) * npad, sizeof(double) * n, n, cudaMemcpyDeviceToHost, ctx->stream));
  if (Linv) CK(cudaMemcpy2DAsync(Linv, sizeof(double) * n, ctx->Linv, sizeof(double) * npad, sizeof(double) * n, n, cudaMemcpyDeviceToHost, ctx->stream));
  if (gamma) CK(cudaMemcpyAsync(gamma, ctx->gamma, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int maf_gp_nll(maf_ctx* ctx, double* nll, double* grad) {
  if (!ctx) return 1;
  if (!ctx->train_set) return fail(ctx, "maf_set_train must be called first");
  CK(cudaSetDevice(ctx->device));
  const int n = ctx->n, npad = ctx->npad, d = ctx->mp.d;
  const size_t nn = (size_t)npad * npad;
  if (ensure(ctx, ctx->K10, nn) || ensure(ctx, ctx->dmu, (size_t)npad + MAF_MAX_D + 8)) return 1;
  double* Kinv = ctx->K10.p;
  double* alpha = ctx->dmu.p;
  double* scal = ctx->dmu.p + npad;          // [0..1] scalars, [2..] gradient
  CK(cudaMemsetAsync(scal, 0, sizeof(double) * (MAF_MAX_D + 8), ctx->stream));
  {
    ProfScope ps(ctx, MAF_PROF_POSTERIOR);
    nll_scalars_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->L0, ctx->gamma, n, npad, scal);
    CKL();
  }
  {
    ProfScope ps(ctx, MAF_PROF_POSTERIOR);
    trmv_upper_kernel<<<(npad + 7) / 8, 256, 0, ctx->stream>>>(ctx->LinvT, ctx->gamma, npad, alpha);
    CKL();
  }
  // K^-1 = L^-T L^-1 = LinvT LinvT^T  (rows of LinvT are K-contiguous: the NT contraction kernel applies)
  if (launch_gemm(ctx, MAF_PROF_GEMM_FWD, npad, npad, npad, ctx->LinvT, npad, ctx->LinvT, npad, Kinv, npad, 0)) return 1;
  {
    ProfScope ps(ctx, MAF_PROF_POSTERIOR);
    dim3 grid((npad + 255) / 256, n);
    nll_grad_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->X0s, n, npad, ctx->mp, Kinv, alpha, scal + 2);
    CKL();
  }
  double h[MAF_MAX_D + 8];
  CK(cudaMemcpyAsync(h, scal, sizeof(double) * (MAF_MAX_D + 8), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (nll) *nll = 0.5 * (h[1] + h[0] + (double)n * 1.8378770664093453);
  if (grad)
    for (int c = 0; c < d + 3; ++c) grad[c] = h[2 + c];
  return 0;
}

// ---- the pipeline -------------------------------------------------------------------------------
static int resolve_acq(maf_ctx* ctx, const maf_acq_params* prm, AcqDev* out) {
  if (!prm) return fail(ctx, "null acquisition parameters");
  if (prm->acq < 0 || prm->acq > 3) return fail(ctx, "unknown acquisition id %d", prm->acq);
  AcqDev a{};
  a.acq = prm->acq;
  a.lower = prm->lower ? 1 : 0;
  a.level = (prm->level == prm->level) ? prm->level : ctx->y_min;
  a.soft = (prm->temperature == prm->temperature) ? 1 : 0;
  a.inv_temp = a.soft ? 1.0 / prm->temperature : 0.0;
  a.beta = prm->beta;
  a.cb_scale = std::pow(0.5 * 3.141592653589793 * prm->beta, 0.5);
  *out = a;
  return 0;
}

// posterior mean / covariance of one chunk: register-accumulating kernel for Q <= 8, shared-memory slab beyond
template <int Q>
static void launch_posterior(maf_ctx* ctx, int sc, const double* Vt, int npad, const double* dX, double* mu, double* Sig,
                             const double* mu_part, int nparts, int Jpad, LevelSwitch ls = LevelSwitch{}, int pd = 0) {
  const double* gam = mu_part ? nullptr : ctx->gamma;         // mean from K10 . alpha: no gamma sweep
  if constexpr (Q <= 8) {
    if (pd > 0)
      posterior_reg_kernel<Q, true><<<sc, POST_REG_THREADS, 0, ctx->stream>>>(Vt, npad, gam, dX, ctx->mp, ctx->mean_val, mu, Sig, mu_part, nparts, Jpad, ls,
                                                                              pd, ctx->Vp.p, ctx->mu_part_p.p, 256);
    else
      posterior_reg_kernel<Q, false><<<sc, POST_REG_THREADS, 0, ctx->stream>>>(Vt, npad, gam, dX, ctx->mp, ctx->mean_val, mu, Sig, mu_part, nparts, Jpad, ls);
  } else {
    if (pd > 0)
      posterior_kernel<Q, true><<<sc, POST_THREADS, 0, ctx->stream>>>(Vt, npad, gam, dX, ctx->mp, ctx->mean_val, mu, Sig, mu_part, nparts, Jpad, ls,
                                                                      pd, ctx->Vp.p, ctx->mu_part_p.p, 256);
    else
      posterior_kernel<Q, false><<<sc, POST_THREADS, 0, ctx->stream>>>(Vt, npad, gam, dX, ctx->mp, ctx->mean_val, mu, Sig, mu_part, nparts, Jpad, ls);
  }
}

// lists and compact buffers of the level switch for chunks of up to `sc` restarts / `Jpad` candidate rows
static int lsw_ensure(maf_ctx* ctx, int sc, int Jpad, int npad) {
  if (!ctx->lsw_cnt) {
    CK(cudaMalloc((void**)&ctx->lsw_cnt, 4 * sizeof(int)));
    CK(cudaMalloc((void**)&ctx->lsw_stats, 4 * sizeof(long long)));
    CK(cudaMemsetAsync(ctx->lsw_stats, 0, 4 * sizeof(long long), ctx->stream));
    CK(cudaHostAlloc((void**)&ctx->lsw_host, 4 * sizeof(long long), cudaHostAllocDefault));
    memset(ctx->lsw_host, 0, 4 * sizeof(long long));
  }
  if ((size_t)sc > ctx->lsw_cap) {
    for (int** b : {&ctx->lsw_selF, &ctx->lsw_slotF, &ctx->lsw_selB}) {
      if (*b) CK(cudaFree(*b));
      *b = nullptr;
    }
    ctx->generation++;
    CK(cudaMalloc((void**)&ctx->lsw_selF, sizeof(int) * sc));
    CK(cudaMalloc((void**)&ctx->lsw_slotF, sizeof(int) * sc));
    CK(cudaMalloc((void**)&ctx->lsw_selB, sizeof(int) * sc));
    ctx->lsw_cap = (size_t)sc;
  }
  return ensure(ctx, ctx->Vt2, (size_t)Jpad * npad) || ensure(ctx, ctx->rowmax2, (size_t)Jpad);
}

__global__ void lsw_accumulate_kernel(const int* __restrict__ cnt, long long* __restrict__ stats, int tested_fwd, int tested_bwd) {
  stats[0] += tested_fwd;
  stats[1] += cnt[0];
  stats[2] += tested_bwd;
  stats[3] += cnt[1];
}

// forward (+ optional backward) over one chunk of restarts
static int run_chunk(maf_ctx* ctx, const AcqDev& ap, bool posterior_only, bool closed_form, int sc, int q, int p,
                     const double* dX, int M, const double* dz, const double* dw, const double* dinc, double* mu,
                     double* Sig, double* chol, double* mubar, double* Sigbar, double* dloss, double* dgrad, bool ded) {
  // ded: the p pending rows were contracted once by run_all (ctx->Vp ...); only the q - p trainable rows of every pool
  // go through the contractions here
  const int d = ctx->mp.d, n = ctx->n, npad = ctx->npad;
  const int pd = ded ? p : 0, qn = q - pd;
  const int J = sc * qn, Jpad = round_up(J, 256);    // 256: one row block per CTA pair of the INT8 contraction
  double* K10 = ctx->K10.p;
  double* Vt = ctx->Vt.p;
  const bool fused = ctx->i8_ready && (ctx->oz_fuse & 1), fused_bwd = ctx->i8_ready && (ctx->oz_fuse & 2);
  // mean and its gradient outside the contractions (the rank-one gradient term rides the register-stash epilogue)
  const bool mu_alpha = fused && fused_bwd && ctx->mu_alpha && ctx->oz_variant == 5 && ctx->oz_ns >= 3;
  const int need_grad = (!posterior_only && dgrad != nullptr) ? 1 : 0;
  // level switch: cheaper level first, failing restarts again with all planes (see LevelSwitch in common.cuh)
  const bool lsw = mu_alpha && ctx->lsw_on && ctx->lsw_ready && !ctx->lsw_busy && ctx->oz_variant == 5 &&
                   npad >= ctx->lsw_min_npad && ctx->oz_ns_fwd == ctx->oz_ns && ctx->oz_ns_bwd == ctx->oz_ns && ctx->oz_ns >= 4;
  const bool lswF = lsw && ctx->lsw_cheap_fwd, lswB = lsw && ctx->lsw_cheap_bwd && need_grad;
  if ((lswF || lswB) && lsw_ensure(ctx, sc, Jpad, npad)) return 1;
  const int ns_fwd = lswF ? ctx->oz_ns - 1 : std::min(ctx->oz_ns_fwd, ctx->oz_ns);
  const int ns_bwd = lswB ? ctx->oz_ns - 1 : ctx->oz_ns_bwd;
  const int nparts = (npad + OZC_THREADS * 4 - 1) / (OZC_THREADS * 4);
  if (mu_alpha && (ensure(ctx, ctx->mu_part, (size_t)nparts * Jpad) || ensure(ctx, ctx->row_add, (size_t)Jpad) ||
                   ensure(ctx, ctx->row_add_fwd, (size_t)Jpad)))
    return 1;
  const double* const radd = mu_alpha ? ctx->row_add.p : nullptr;
  const double* const cadd = mu_alpha ? ctx->alpha : nullptr;
  const double* mu_part = mu_alpha ? ctx->mu_part.p : nullptr;
  // the Vbar producer needs the row maxima of V^T: the forward contraction's epilogue leaves them behind
  const bool want_rowmax = fused_bwd && ctx->oz_rowmax && need_grad;
  bool have_rowmax = false;
  if (want_rowmax && ensure(ctx, ctx->ozVmax, (size_t)Jpad)) return 1;
  double* const rowmax = want_rowmax ? ctx->ozVmax.p : nullptr;
  if (lswF || lswB) CK(cudaMemsetAsync(ctx->lsw_cnt, 0, 4 * sizeof(int), ctx->stream));
  // kernel derivatives stored by the forward covariance kernel for the reverse one (cross_bwd_g_kernel)
  const bool groute = mu_alpha && ctx->g_route && need_grad && d >= 1 && d <= 8;
  if (groute && ensure(ctx, ctx->Gk, (size_t)Jpad * npad)) return 1;
  double* const gout = groute ? ctx->Gk.p : nullptr;
  if (fused) {
    // K10 goes straight to digit planes (bounded by a^2: fixed row scale), FP64 K10 is never materialised
    const size_t bytes = (size_t)ctx->oz_ns * Jpad * npad;
    if (ensure(ctx, ctx->ozA, (bytes + 7) / 8) || ensure(ctx, ctx->ozsA, (size_t)Jpad)) return 1;
    {
      ProfScope ps(ctx, MAF_PROF_CROSS_FWD);
      dim3 grid(nparts, Jpad / OZC_TJ);
      if (groute) {
        DISPATCH_DKN(d, ctx->mp.kernel_id, ns_fwd, cross_fwd_planes_kernel<D, KID, NSC, 3><<<grid, OZC_THREADS, 0, ctx->stream>>>(dX, J, Jpad, ctx->X0s, n, npad, ctx->mp, ns_fwd,
                                                                                       (int8_t*)ctx->ozA.p, ctx->ozsA.p, ctx->alpha, ctx->mu_part.p, nullptr, nullptr, q, pd, ctx->centre_k10 ? ctx->row_add_fwd.p : nullptr, gout));
      } else if (mu_alpha) {
        DISPATCH_DKN(d, ctx->mp.kernel_id, ns_fwd, cross_fwd_planes_kernel<D, KID, NSC, 1><<<grid, OZC_THREADS, 0, ctx->stream>>>(dX, J, Jpad, ctx->X0s, n, npad, ctx->mp, ns_fwd,
                                                                                       (int8_t*)ctx->ozA.p, ctx->ozsA.p, ctx->alpha, ctx->mu_part.p, nullptr, nullptr, q, pd, ctx->centre_k10 ? ctx->row_add_fwd.p : nullptr));
      } else {
        DISPATCH_DKN(d, ctx->mp.kernel_id, ns_fwd, cross_fwd_planes_kernel<D, KID, NSC, 0><<<grid, OZC_THREADS, 0, ctx->stream>>>(dX, J, Jpad, ctx->X0s, n, npad, ctx->mp, ns_fwd,
                                                                                       (int8_t*)ctx->ozA.p, ctx->ozsA.p));
      }
      CKL();
    }
    if (launch_i8gemm(ctx, MAF_PROF_GEMM_FWD, Jpad, npad, npad, ctx->ozA, ctx->ozsA, ctx->ozTL, ctx->ozsTL, Vt, npad, 1,
                      ns_fwd, ns_fwd + 1, rowmax, &have_rowmax, nullptr, 1, (mu_alpha && ctx->centre_k10) ? ctx->row_add_fwd.p : nullptr,
                      (mu_alpha && ctx->centre_k10) ? ctx->linv_rowsum : nullptr))
      return 1;
  } else {
    {
      ProfScope ps(ctx, MAF_PROF_CROSS_FWD);
      dim3 grid((npad + CROSS_TK - 1) / CROSS_TK, Jpad / CROSS_TJ);
      DISPATCH_D(d, cross_fwd_kernel<D><<<grid, CROSS_TK, 0, ctx->stream>>>(dX, J, ctx->X0s, n, npad, ctx->mp, K10));
      CKL();
    }
    if (ctx->i8_ready) {
      // cross-covariances are bounded by a^2: fixed row scale, no reduction pass
      if (oz_slice(ctx, MAF_PROF_CROSS_FWD, K10, npad, J, Jpad, npad, ctx->oz_ns, ctx->mp.amp2, ctx->ozA, ctx->ozsA)) return 1;
      if (launch_i8gemm(ctx, MAF_PROF_GEMM_FWD, Jpad, npad, npad, ctx->ozA, ctx->ozsA, ctx->ozTL, ctx->ozsTL, Vt, npad, 1,
                        ctx->oz_ns, ctx->oz_lmax, rowmax, &have_rowmax))
        return 1;
    } else if (launch_gemm(ctx, MAF_PROF_GEMM_FWD, Jpad, npad, npad, K10, npad, ctx->Linv, npad, Vt, npad, 1)) {
      return 1;
    }
  }
  {
    ProfScope ps(ctx, MAF_PROF_POSTERIOR);
    LevelSwitch ls{};
    if (lswF) ls = LevelSwitch{nullptr, nullptr, ctx->lsw_cnt, ctx->lsw_selF, ctx->lsw_slotF, ctx->lsw_err_fwd, ctx->lsw_tol};
    DISPATCH_Q(q, launch_posterior<Q>(ctx, sc, Vt, npad, dX, mu, Sig, mu_part, nparts, Jpad, ls, pd));
    CKL();
  }
  bool have_rowmax2 = false;
  if (lswF) {
    // second pass of the forward switch: the listed restarts again with all planes, on compact rows
    const int nsf = ctx->oz_ns;
    {
      ProfScope ps(ctx, MAF_PROF_CROSS_FWD);
      dim3 grid(nparts, Jpad / OZC_TJ);
      if (nsf == 7) {
        DISPATCH_DK(d, ctx->mp.kernel_id, cross_fwd_planes_kernel<D, KID, 7, 2><<<grid, OZC_THREADS, 0, ctx->stream>>>(dX, J, Jpad, ctx->X0s, n, npad, ctx->mp, nsf,
                                                                               (int8_t*)ctx->ozA.p, ctx->ozsA.p, ctx->alpha, ctx->mu_part.p, ctx->lsw_selF, ctx->lsw_cnt, q, pd, ctx->centre_k10 ? ctx->row_add_fwd.p : nullptr));
      } else {                                               // MAF_F32 mode (5 planes): run-time digit count
        DISPATCH_DK(d, ctx->mp.kernel_id, cross_fwd_planes_kernel<D, KID, 0, 2><<<grid, OZC_THREADS, 0, ctx->stream>>>(dX, J, Jpad, ctx->X0s, n, npad, ctx->mp, nsf,
                                                                               (int8_t*)ctx->ozA.p, ctx->ozsA.p, ctx->alpha, ctx->mu_part.p, ctx->lsw_selF, ctx->lsw_cnt, q, pd, ctx->centre_k10 ? ctx->row_add_fwd.p : nullptr));
      }
      CKL();
    }
    if (launch_i8gemm(ctx, MAF_PROF_GEMM_FWD, Jpad, npad, npad, ctx->ozA, ctx->ozsA, ctx->ozTL, ctx->ozsTL, ctx->Vt2.p, npad, 1,
                      nsf, nsf + 1, want_rowmax ? ctx->rowmax2.p : nullptr, &have_rowmax2, ctx->lsw_cnt, qn,
                      ctx->centre_k10 ? ctx->row_add_fwd.p : nullptr, ctx->centre_k10 ? ctx->linv_rowsum : nullptr))
      return 1;
    ProfScope ps(ctx, MAF_PROF_POSTERIOR);
    const LevelSwitch ls{ctx->lsw_selF, ctx->lsw_cnt, nullptr, nullptr, nullptr, 0.0, 0.0};
    DISPATCH_Q(q, launch_posterior<Q>(ctx, sc, ctx->Vt2.p, npad, dX, mu, Sig, mu_part, nparts, Jpad, ls, pd));
    CKL();
  }
  if (posterior_only) {
    if (lswF) {
      ProfScope ps(ctx, MAF_PROF_ADAM);
      lsw_accumulate_kernel<<<1, 1, 0, ctx->stream>>>(ctx->lsw_cnt, ctx->lsw_stats, sc, 0);
    }
    return 0;
  }
  {
    ProfScope ps(ctx, MAF_PROF_MC);
    if (closed_form) {
      closed_form_kernel<<<(sc + 127) / 128, 128, 0, ctx->stream>>>(sc, mu, Sig, ap, need_grad, dloss, mubar, Sigbar);
    } else {
      DISPATCH_Q(q, {
        const size_t smem = sizeof(McSmem<Q>);
        if (smem > 48 * 1024 && !(ctx->attr_mc_mask & (1u << Q))) {
          CK(cudaFuncSetAttribute(mc_acq_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          ctx->attr_mc_mask |= 1u << Q;
        }
        mc_acq_kernel<Q><<<(sc + MC_WARPS - 1) / MC_WARPS, MC_WARPS * 32, smem, ctx->stream>>>(
            sc, M, mu, Sig, dz, dw, dinc, ap, ctx->mp.jitter, need_grad, dloss, mubar, Sigbar, chol);
      });
    }
    CKL();
  }
  if (!need_grad) {
    if (lswF) {
      ProfScope ps(ctx, MAF_PROF_ADAM);
      lsw_accumulate_kernel<<<1, 1, 0, ctx->stream>>>(ctx->lsw_cnt, ctx->lsw_stats, sc, 0);
    }
    return 0;
  }
  if (fused_bwd) {
    // the digit planes of L0^-T were written with oz_ns digits; their first ns_bwd planes are the balanced
    // expansion rounded to that many digits, so the same buffer serves the shorter reverse contraction
    const size_t bytes = (size_t)ctx->oz_ns * Jpad * npad;
    if (ensure(ctx, ctx->ozA, (bytes + 7) / 8) || ensure(ctx, ctx->ozsA, (size_t)Jpad)) return 1;
    const int* vslot = lswF ? ctx->lsw_slotF : nullptr;      // restarts whose V^T rows live in Vt2 (compact)
    const double* vmax2 = (lswF && have_rowmax2) ? ctx->rowmax2.p : nullptr;
    const double* vmax1 = (have_rowmax && (!lswF || have_rowmax2)) ? rowmax : nullptr;
    {
      ProfScope ps(ctx, MAF_PROF_VBAR);
      DISPATCH_Q(q, (pd > 0 ? vbar_planes_kernel<Q, true> : vbar_planes_kernel<Q, false>)<<<sc + (Jpad - J), VBP_THREADS, 0, ctx->stream>>>(Vt, npad, sc, Jpad, ctx->gamma, mu_alpha ? nullptr : mubar, Sigbar, ns_bwd,
                                                                                   (int8_t*)ctx->ozA.p, ctx->ozsA.p, vmax1, ctx->vbar_tight, vslot, ctx->Vt2.p, vmax2,
                                                                                   LevelSwitch{}, mubar, mu_alpha ? ctx->row_add.p : nullptr,
                                                                                   pd, ctx->Vp.p, ctx->rowmax_p.p));
      CKL();
    }
    if (launch_i8gemm(ctx, MAF_PROF_GEMM_BWD, Jpad, npad, npad, ctx->ozA, ctx->ozsA, ctx->ozTU, ctx->ozsTU, K10, npad, 2,
                      ns_bwd, ns_bwd + 1, nullptr, nullptr, nullptr, 1, radd, cadd))
      return 1;
  } else {
    {
      ProfScope ps(ctx, MAF_PROF_VBAR);
      DISPATCH_Q(q, vbar_kernel<Q><<<sc, 256, 0, ctx->stream>>>(Vt, npad, ctx->gamma, mubar, Sigbar));
      CKL();
    }
    if (ctx->i8_ready) {
      if (oz_slice(ctx, MAF_PROF_VBAR, Vt, npad, J, Jpad, npad, ctx->oz_ns_bwd, 0.0, ctx->ozA, ctx->ozsA)) return 1;
      if (launch_i8gemm(ctx, MAF_PROF_GEMM_BWD, Jpad, npad, npad, ctx->ozA, ctx->ozsA, ctx->ozTU, ctx->ozsTU, K10, npad, 2,
                        ctx->oz_ns_bwd, ctx->oz_lmax_bwd))
        return 1;
    } else if (launch_gemm(ctx, MAF_PROF_GEMM_BWD, Jpad, npad, npad, Vt, npad, ctx->LinvT, npad, K10, npad, 2)) {
      return 1;
    }
  }
  {
    ProfScope ps(ctx, MAF_PROF_CROSS_BWD);
    if (groute) {
      constexpr int R = CROSS_BWD_G_R;
      DISPATCH_DK(d, ctx->mp.kernel_id, cross_bwd_g_kernel<(D > 0 ? D : 1), KID, R, false><<<sc, 32 * ((q - p + R - 1) / R), 0, ctx->stream>>>(
                      dX, sc, q, p, ctx->X0s, n, npad, ctx->mp, K10, gout, Sigbar, dgrad, nullptr, nullptr, qn));
    } else {
      DISPATCH_DK(d, ctx->mp.kernel_id, cross_bwd_kernel<D, KID, false><<<sc, 32 * (q - p), 0, ctx->stream>>>(dX, sc, q, p, ctx->X0s, n, npad, ctx->mp, K10, Sigbar, dgrad,
                                                                                                     nullptr, nullptr, qn));
    }
    CKL();
  }
  if (lswB) {
    ProfScope ps(ctx, MAF_PROF_CROSS_BWD);
    const LevelSwitch ls{nullptr, nullptr, ctx->lsw_cnt + 1, ctx->lsw_selB, nullptr, ctx->lsw_err_bwd, ctx->lsw_tol};
    lsw_flag_bwd_kernel<<<(sc + 127) / 128, 128, 0, ctx->stream>>>(dgrad, ctx->ozsA.p, sc, q, p, d, ls, qn);
    CKL();
  }
  if (lswB) {
    // second pass of the reverse switch: Vbar planes of the listed restarts with all planes on compact rows, contraction
    // into the (now free) V^T buffer, gradient of those restarts again
    const int nsb = ctx->oz_ns;
    const LevelSwitch ls{ctx->lsw_selB, ctx->lsw_cnt + 1, nullptr, nullptr, nullptr, 0.0, 0.0};
    const double* vmax2 = (lswF && have_rowmax2) ? ctx->rowmax2.p : nullptr;
    const double* vmax1 = (have_rowmax && (!lswF || have_rowmax2)) ? rowmax : nullptr;
    {
      ProfScope ps(ctx, MAF_PROF_VBAR);
      DISPATCH_Q(q, (pd > 0 ? vbar_planes_kernel<Q, true> : vbar_planes_kernel<Q, false>)<<<sc, VBP_THREADS, 0, ctx->stream>>>(Vt, npad, sc, Jpad, ctx->gamma, nullptr, Sigbar, nsb, (int8_t*)ctx->ozA.p,
                                                                       ctx->ozsA.p, vmax1, ctx->vbar_tight, lswF ? ctx->lsw_slotF : nullptr, ctx->Vt2.p, vmax2, ls,
                                                                       mubar, ctx->row_add.p, pd, ctx->Vp.p, ctx->rowmax_p.p));
      CKL();
    }
    if (launch_i8gemm(ctx, MAF_PROF_GEMM_BWD, Jpad, npad, npad, ctx->ozA, ctx->ozsA, ctx->ozTU, ctx->ozsTU, Vt, npad, 2, nsb, nsb + 1,
                      nullptr, nullptr, ctx->lsw_cnt + 1, qn, radd, cadd))
      return 1;
    {
      ProfScope ps(ctx, MAF_PROF_CROSS_BWD);
      if (groute) {
        constexpr int R = CROSS_BWD_G_R;
        DISPATCH_DK(d, ctx->mp.kernel_id, cross_bwd_g_kernel<(D > 0 ? D : 1), KID, R, true><<<sc, 32 * ((q - p + R - 1) / R), 0, ctx->stream>>>(
                        dX, sc, q, p, ctx->X0s, n, npad, ctx->mp, Vt, gout, Sigbar, dgrad, ctx->lsw_selB, ctx->lsw_cnt + 1, qn));
      } else {
        DISPATCH_DK(d, ctx->mp.kernel_id, cross_bwd_kernel<D, KID, true><<<sc, 32 * (q - p), 0, ctx->stream>>>(dX, sc, q, p, ctx->X0s, n, npad, ctx->mp, Vt, Sigbar, dgrad,
                                                                                                       ctx->lsw_selB, ctx->lsw_cnt + 1, qn));
      }
      CKL();
    }
  }
  if (lswF || lswB) {
    ProfScope ps(ctx, MAF_PROF_ADAM);
    lsw_accumulate_kernel<<<1, 1, 0, ctx->stream>>>(ctx->lsw_cnt, ctx->lsw_stats, lswF ? sc : 0, lswB ? sc : 0);
  }
  return 0;
}

static int check_shape(maf_ctx* ctx, int S, int q, int p) {
  if (!ctx->train_set) return fail(ctx, "maf_set_train must be called first");
  if (S < 1) return fail(ctx, "S must be >= 1");
  if (q < 1 || q > MAF_MAX_Q) return fail(ctx, "q=%d outside 1..%d", q, MAF_MAX_Q);
  if (p < 0 || p >= q) return fail(ctx, "p_pending=%d must satisfy 0 <= p < q=%d", p, q);
  return 0;
}

// Between evaluations: is the cheaper level of each direction still paying (see maf_ctx::lsw_cheap_fwd)?
static void lsw_adapt(maf_ctx* ctx) {
  if (!ctx->lsw_host || !ctx->lsw_ready || ctx->lsw_busy) return;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(ctx->stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) return;
  const long long h[4] = {ctx->lsw_host[0], ctx->lsw_host[1], ctx->lsw_host[2], ctx->lsw_host[3]};
  auto one = [&](bool& cheap, int& cool, int it, int ir) {
    if (cheap) {
      const long long tested = h[it] - ctx->lsw_base[it], redone = h[ir] - ctx->lsw_base[ir];
      if (tested >= 256 && redone * 5 > tested * 2) {            // more than 40 % recomputed
        cheap = false;
        cool = 64;
        ctx->generation++;
      }
      if (tested >= 256) ctx->lsw_base[it] = h[it], ctx->lsw_base[ir] = h[ir];
    } else if (--cool <= 0) {
      cheap = true;
      ctx->lsw_base[it] = h[it];
      ctx->lsw_base[ir] = h[ir];
      ctx->generation++;
    }
  };
  one(ctx->lsw_cheap_fwd, ctx->lsw_cool_fwd, 0, 1);
  one(ctx->lsw_cheap_bwd, ctx->lsw_cool_bwd, 2, 3);
}

// All restarts, in chunks of at most chunk_rows candidate rows.  Device pointers; asynchronous.
// Host pools X[S][q][d]: do all restarts carry the same p pending rows (bit for bit)?
static bool pending_rows_shared(const double* X, int S, int q, int p, int d) {
  if (p <= 0) return false;
  const size_t row = (size_t)q * d, len = (size_t)p * d * sizeof(double);
  for (int s = 1; s < S; ++s)
    if (memcmp(X, X + (size_t)s * row, len) != 0) return false;
  return true;
}

// V^T rows, mean partial sums and row maxima of the p pending points (rows 0..p-1 of the first pool), on all planes
static int run_pending_block(maf_ctx* ctx, const double* dX, int q, int p) {
  const int d = ctx->mp.d, n = ctx->n, npad = ctx->npad, ns = ctx->oz_ns, Jp = 256;
  const int nparts = (npad + OZC_THREADS * 4 - 1) / (OZC_THREADS * 4);
  const size_t bytes = (size_t)ns * Jp * npad;
  if (ensure(ctx, ctx->ozA, (bytes + 7) / 8) || ensure(ctx, ctx->ozsA, (size_t)Jp) || ensure(ctx, ctx->Vp, (size_t)Jp * npad) ||
      ensure(ctx, ctx->mu_part_p, (size_t)nparts * Jp) || ensure(ctx, ctx->rowmax_p, (size_t)Jp) ||
      ensure(ctx, ctx->row_add_fwd, (size_t)Jp))
    return 1;
  {
    ProfScope ps(ctx, MAF_PROF_CROSS_FWD);
    dim3 grid(nparts, Jp / OZC_TJ);
    DISPATCH_DKN(d, ctx->mp.kernel_id, ns, cross_fwd_planes_kernel<D, KID, NSC, 1><<<grid, OZC_THREADS, 0, ctx->stream>>>(dX, p, Jp, ctx->X0s, n, npad, ctx->mp, ns,
                                                                               (int8_t*)ctx->ozA.p, ctx->ozsA.p, ctx->alpha, ctx->mu_part_p.p, nullptr, nullptr, 1, 0,
                                                                               ctx->centre_k10 ? ctx->row_add_fwd.p : nullptr));
    CKL();
  }
  bool done = false;
  if (launch_i8gemm(ctx, MAF_PROF_GEMM_FWD, Jp, npad, npad, ctx->ozA, ctx->ozsA, ctx->ozTL, ctx->ozsTL, ctx->Vp.p, npad, 1, ns, ns + 1,
                    ctx->rowmax_p.p, &done, nullptr, 1, ctx->centre_k10 ? ctx->row_add_fwd.p : nullptr,
                    ctx->centre_k10 ? ctx->linv_rowsum : nullptr))
    return 1;
  if (!done) return fail(ctx, "the shared pending block needs the register-stash contraction kernel");
  return 0;
}

static int run_all(maf_ctx* ctx, const maf_acq_params* prm, bool posterior_only, int S, int q, int p,
                   const double* dX, int M, const double* dz, const double* dw, const double* dinc,
                   double* dloss, double* dgrad, bool pend_shared = false) {
  if (check_shape(ctx, S, q, p)) return 1;
  CK(cudaSetDevice(ctx->device));
  AcqDev ap{};
  bool closed_form = false;
  if (!posterior_only) {
    if (resolve_acq(ctx, prm, &ap)) return 1;
    closed_form = (q == 1 && dz == nullptr);
    if (!closed_form && (dz == nullptr || M < 1))
      return fail(ctx, "base samples z[M][q] are required for the Monte-Carlo path");
  }
  const int d = ctx->mp.d, npad = ctx->npad;
  const size_t sq = (size_t)S * q, sqq = sq * q;
  if (ensure(ctx, ctx->dmu, sq) || ensure(ctx, ctx->dmubar, sq) || ensure(ctx, ctx->dSig, sqq) ||
      ensure(ctx, ctx->dSigbar, sqq) || ensure(ctx, ctx->dchol, sqq))
    return 1;
  lsw_adapt(ctx);
  // pending points contracted once: the INT8 engine with both fused producers and the exact-mean route only
  const bool ded = pend_shared && ctx->pend_dedupe && p > 0 && !posterior_only && ctx->i8_ready && (ctx->oz_fuse & 3) == 3 && ctx->mu_alpha &&
                   ctx->oz_variant == 5 && ctx->oz_ns >= 3 && ctx->oz_rowmax;
  const int qrows = ded ? q - p : q;                              // candidate rows per restart through the contractions
  int sc_max = (int)std::max<size_t>(1, ctx->chunk_rows / qrows);
  sc_max = std::min(sc_max, S);
  const size_t need = (size_t)round_up(sc_max * qrows, 256) * npad;
  if (ensure(ctx, ctx->K10, need) || ensure(ctx, ctx->Vt, need)) return 1;
  if (ded && run_pending_block(ctx, dX, q, p)) return 1;
  for (int s0 = 0; s0 < S; s0 += sc_max) {
    const int sc = std::min(sc_max, S - s0);
    const size_t oq = (size_t)s0 * q, oqq = oq * q;
    if (run_chunk(ctx, ap, posterior_only, closed_form, sc, q, p, dX + oq * d, M, dz, dw, dinc, ctx->dmu.p + oq,
                  ctx->dSig.p + oqq, closed_form ? nullptr : ctx->dchol.p + oqq, ctx->dmubar.p + oq,
                  ctx->dSigbar.p + oqq, dloss ? dloss + s0 : nullptr,
                  dgrad ? dgrad + (size_t)s0 * (q - p) * d : nullptr, ded))
      return 1;
  }
  if (ctx->lsw_stats && ctx->lsw_ready && !ctx->lsw_busy)
    CK(cudaMemcpyAsync(ctx->lsw_host, ctx->lsw_stats, 4 * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  ctx->last_S = S;
  ctx->last_q = q;
  ctx->last_has_chol = !posterior_only && !closed_form;
  ctx->last_dX = dX;
  return 0;
}

// Calibration of the level switch for the loaded training set: the error the cheaper level (one digit plane less)
// leaves on the posterior covariance (absolute) and on the gradient (per unit row scale of Vbar^T), measured against the
// full level on 512 probe candidates -- half of them 1e-3 away from training points, where V has full norm and the
// posterior variance sits at the noise level (the worst case for Sigma = K11 - V^T V), half uniform.  The truncation error
// is a sum over ~n digit products with the fixed row scales of L0^-1, so its size does not depend on which candidates
// are evaluated later; the per-restart test in the kernels compares it with what each restart's own result tolerates.
static int lsw_calibrate(maf_ctx* ctx, int n, const double* X0) {
  ctx->lsw_ready = false;
  const bool eligible = ctx->lsw_on && ctx->i8_ready && (ctx->oz_fuse & 3) == 3 && ctx->mu_alpha &&
                        ctx->oz_variant == 5 && ctx->npad >= ctx->lsw_min_npad && ctx->oz_ns >= 4 &&
                        ctx->oz_ns_fwd == ctx->oz_ns && ctx->oz_ns_bwd == ctx->oz_ns;
  if (!eligible) return 0;
  const int d = ctx->mp.d, S = 512, ns = ctx->oz_ns;
  std::vector<double> X((size_t)S * d);
  unsigned long long st = 0x9E3779B97F4A7C15ull;
  auto u01 = [&st]() {                                       // splitmix64: fixed probe set, independent of any host RNG
    st += 0x9E3779B97F4A7C15ull;
    unsigned long long z = st;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
  };
  for (int s = 0; s < S; ++s)
    for (int c = 0; c < d; ++c) {
      double v = u01();
      if (s < S / 2) {
        const int i = (int)((long long)s * n / (S / 2));
        v = std::min(1.0, std::max(0.0, X0[(size_t)i * d + c] + 2e-3 * (v - 0.5)));
      }
      X[(size_t)s * d + c] = v;
    }
  if (ensure(ctx, ctx->dX, (size_t)S * d) || ensure(ctx, ctx->dloss, (size_t)S) || ensure(ctx, ctx->dgrad, (size_t)S * d)) return 1;
  CK(cudaMemcpyAsync(ctx->dX.p, X.data(), sizeof(double) * S * d, cudaMemcpyHostToDevice, ctx->stream));
  std::vector<double> v6(S), v7(S), g6((size_t)S * d), g7((size_t)S * d), sc6(S);
  maf_acq_params prm{MAF_ACQ_CB, 1, NAN, NAN, 2.0};
  ctx->lsw_busy = true;
  const size_t chunk_rows_saved = ctx->chunk_rows;
  ctx->chunk_rows = std::max<size_t>(ctx->chunk_rows, (size_t)S);    // one chunk: the row scales of all probes are read back
  int rc = 0;
  for (int pass = 0; pass < 2 && rc == 0; ++pass) {
    const int nsp = pass == 0 ? ns - 1 : ns;
    ctx->oz_ns_fwd = nsp;
    ctx->oz_ns_bwd = ns;
    rc = run_all(ctx, nullptr, true, S, 1, 0, ctx->dX.p, 0, nullptr, nullptr, nullptr, nullptr, nullptr);
    if (rc == 0 && cudaMemcpyAsync((pass == 0 ? v6 : v7).data(), ctx->dSig.p, sizeof(double) * S, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
      rc = fail(ctx, "level-switch calibration: copy failed");
    ctx->oz_ns_fwd = ns;
    ctx->oz_ns_bwd = nsp;
    ctx->oz_lmax_bwd = nsp + 1;
    if (rc == 0) rc = run_all(ctx, &prm, false, S, 1, 0, ctx->dX.p, 0, nullptr, nullptr, nullptr, ctx->dloss.p, ctx->dgrad.p);
    if (rc == 0 && cudaMemcpyAsync((pass == 0 ? g6 : g7).data(), ctx->dgrad.p, sizeof(double) * S * d, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
      rc = fail(ctx, "level-switch calibration: copy failed");
    if (rc == 0 && pass == 0 && cudaMemcpyAsync(sc6.data(), ctx->ozsA.p, sizeof(double) * S, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
      rc = fail(ctx, "level-switch calibration: copy failed");
    if (rc == 0 && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, "level-switch calibration failed on the device");
  }
  ctx->oz_ns_fwd = ctx->oz_ns_bwd = ns;
  ctx->oz_lmax_bwd = ns + 1;
  ctx->chunk_rows = chunk_rows_saved;
  ctx->lsw_busy = false;
  ctx->last_S = 0;
  ctx->lsw_cheap_fwd = ctx->lsw_cheap_bwd = true;
  if (rc) return rc;
  double ef = 0.0, eb = 0.0;
  for (int s = 0; s < S; ++s) {
    ef = std::max(ef, std::fabs(v6[s] - v7[s]));
    if (sc6[s] > 0.0)
      for (int c = 0; c < d; ++c) eb = std::max(eb, std::fabs(g6[(size_t)s * d + c] - g7[(size_t)s * d + c]) / sc6[s]);
  }
  ctx->lsw_err_fwd = ef;
  ctx->lsw_err_bwd = eb;
  // share of the uniform probes the per-restart test would send to the second pass: a direction that starts above 90 %
  // (e.g. the squared-exponential kernel at n = 4096, cond(K00) ~ 1e9) begins on all planes.  The probes are single
  // points; pools of q points tolerate more (max |Sigma_s| over q^2 entries), hence the generous threshold here -- the
  // running statistics take over after the first evaluation.
  int ff = 0, fb = 0;
  for (int s = S / 2; s < S; ++s) {
    double gm = 0.0;
    for (int c = 0; c < d; ++c) gm = std::max(gm, std::fabs(g7[(size_t)s * d + c]));
    ff += !(ef <= ctx->lsw_tol * std::fabs(v7[s]));
    fb += !(eb * sc6[s] <= ctx->lsw_tol * gm);
  }
  ctx->lsw_cheap_fwd = ff * 10 <= (S / 2) * 9;
  ctx->lsw_cheap_bwd = fb * 10 <= (S / 2) * 9;
  ctx->lsw_cool_fwd = ctx->lsw_cheap_fwd ? 0 : 64;
  ctx->lsw_cool_bwd = ctx->lsw_cheap_bwd ? 0 : 64;
  if (ctx->lsw_host)
    for (int u = 0; u < 4; ++u) ctx->lsw_base[u] = ctx->lsw_host[u];
  ctx->lsw_ready = true;
  return 0;
}

extern "C" int maf_set_level_switch(maf_ctx* ctx, int on, double tol) {
  if (!ctx) return 1;
  if (tol == tol && !(tol > 0.0)) return fail(ctx, "the tolerance of the level switch must be positive");
  ctx->generation++;
  ctx->lsw_on = on ? 1 : 0;
  if (tol == tol) ctx->lsw_tol = tol;
  if (on && !ctx->lsw_ready) ctx->train_set = false;           // calibrated by the next maf_set_train
  return 0;
}

extern "C" int maf_level_switch_stats(maf_ctx* ctx, int64_t* counts, double* errs, int reset) {
  if (!ctx) return 1;
  CK(cudaSetDevice(ctx->device));
  long long h[4] = {0, 0, 0, 0};
  if (ctx->lsw_stats) {
    CK(cudaMemcpyAsync(h, ctx->lsw_stats, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (reset) {
      CK(cudaMemsetAsync(ctx->lsw_stats, 0, sizeof(h), ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      for (int u = 0; u < 4; ++u) ctx->lsw_host[u] = 0, ctx->lsw_base[u] = 0;
    }
  }
  if (counts)
    for (int u = 0; u < 4; ++u) counts[u] = (int64_t)h[u];
  if (errs) {
    errs[0] = ctx->lsw_ready ? ctx->lsw_err_fwd : NAN;
    errs[1] = ctx->lsw_ready ? ctx->lsw_err_bwd : NAN;
    errs[2] = ctx->lsw_tol;
  }
  return 0;
}

extern "C" int maf_posterior(maf_ctx* ctx, int S, int q, const double* X, int full_cov, double* mu, double* cov) {
  if (!ctx) return 1;
  if (!X) return fail(ctx, "null pointer");
  if (check_shape(ctx, S, q, 0)) return 1;
  CK(cudaSetDevice(ctx->device));
  const size_t nx = (size_t)S * q * ctx->mp.d;
  if (ensure(ctx, ctx->dX, nx)) return 1;
  CK(cudaMemcpyAsync(ctx->dX.p, X, nx * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (run_all(ctx, nullptr, true, S, q, 0, ctx->dX.p, 0, nullptr, nullptr, nullptr, nullptr, nullptr)) return 1;
  if (mu) CK(cudaMemcpyAsync(mu, ctx->dmu.p, sizeof(double) * S * q, cudaMemcpyDeviceToHost, ctx->stream));
  if (cov) {
    if (full_cov) {
      CK(cudaMemcpyAsync(cov, ctx->dSig.p, sizeof(double) * S * q * q, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
      {
        ProfScope ps(ctx, MAF_PROF_POSTERIOR);
        diag_extract_kernel<<<(S * q + 255) / 256, 256, 0, ctx->stream>>>(ctx->dSig.p, S, q, ctx->dSigbar.p);
        CKL();
      }
      CK(cudaMemcpyAsync(cov, ctx->dSigbar.p, sizeof(double) * S * q, cudaMemcpyDeviceToHost, ctx->stream));
    }
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int maf_acq_value_grad_dev(maf_ctx* ctx, const maf_acq_params* prm, int S, int q, int p_pending,
                                      const double* dX, int M, const double* dz, const double* dweights,
                                      const double* dincumbents, double* dout_loss, double* dout_grad) {
  if (!ctx) return 1;
  if (!dX || !dout_loss) return fail(ctx, "null device pointer");
  if (check_shape(ctx, S, q, p_pending)) return 1;
  if (ensure(ctx, ctx->dloss, (size_t)S)) return 1;   // copy kept for maf_argmin_last
  if (run_all(ctx, prm, false, S, q, p_pending, dX, M, dz, dweights, dincumbents, dout_loss, dout_grad)) return 1;
  if (dout_loss != ctx->dloss.p)
    CK(cudaMemcpyAsync(ctx->dloss.p, dout_loss, sizeof(double) * S, cudaMemcpyDeviceToDevice, ctx->stream));
  return 0;
}

extern "C" int maf_acq_value_grad(maf_ctx* ctx, const maf_acq_params* prm, int S, int q, int p_pending,
                                  const double* X, int M, const double* z, const double* weights,
                                  const double* incumbents, double* out_loss, double* out_grad) {
  if (!ctx) return 1;
  if (!X || !out_loss) return fail(ctx, "null pointer");
  if (check_shape(ctx, S, q, p_pending)) return 1;
  CK(cudaSetDevice(ctx->device));
  const int d = ctx->mp.d;
  const size_t nx = (size_t)S * q * d, ng = (size_t)S * (q - p_pending) * d;
  if (ensure(ctx, ctx->dX, nx) || ensure(ctx, ctx->dloss, (size_t)S)) return 1;
  if (out_grad && ensure(ctx, ctx->dgrad, ng)) return 1;
  CK(cudaMemcpyAsync(ctx->dX.p, X, nx * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  const double *dz = nullptr, *dw = nullptr, *dinc = nullptr;
  if (z) {
    if (M < 1) return fail(ctx, "M must be >= 1 when z is given");
    if (ensure(ctx, ctx->dz, (size_t)M * q)) return 1;
    CK(cudaMemcpyAsync(ctx->dz.p, z, sizeof(double) * M * q, cudaMemcpyHostToDevice, ctx->stream));
    dz = ctx->dz.p;
    if (weights) {
      if (ensure(ctx, ctx->dw, (size_t)M)) return 1;
      CK(cudaMemcpyAsync(ctx->dw.p, weights, sizeof(double) * M, cudaMemcpyHostToDevice, ctx->stream));
      dw = ctx->dw.p;
    }
    if (incumbents) {
      if (ensure(ctx, ctx->dinc, (size_t)M)) return 1;
      CK(cudaMemcpyAsync(ctx->dinc.p, incumbents, sizeof(double) * M, cudaMemcpyHostToDevice, ctx->stream));
      dinc = ctx->dinc.p;
    }
  }
  if (run_all(ctx, prm, false, S, q, p_pending, ctx->dX.p, M, dz, dw, dinc, ctx->dloss.p,
              out_grad ? ctx->dgrad.p : nullptr, pending_rows_shared(X, S, q, p_pending, d)))
    return 1;
  CK(cudaMemcpyAsync(out_loss, ctx->dloss.p, sizeof(double) * S, cudaMemcpyDeviceToHost, ctx->stream));
  if (out_grad) CK(cudaMemcpyAsync(out_grad, ctx->dgrad.p, sizeof(double) * ng, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int maf_mc_from_moments(maf_ctx* ctx, const maf_acq_params* prm, int S, int q, const double* mu,
                                   const double* cov, int M, const double* z, const double* weights,
                                   const double* incumbents, double* out_loss, double* out_mubar,
                                   double* out_covbar, double* out_chol) {
  if (!ctx) return 1;
  if (!ctx->model_set) return fail(ctx, "maf_set_model must be called first");
  if (!prm || !mu || !cov || !z) return fail(ctx, "null pointer");
  if (S < 1 || M < 1) return fail(ctx, "S and M must be >= 1");
  if (q < 1 || q > MAF_MAX_Q) return fail(ctx, "q=%d outside 1..%d", q, MAF_MAX_Q);
  CK(cudaSetDevice(ctx->device));
  AcqDev ap{};
  if (prm->level != prm->level && !ctx->train_set && (prm->acq == MAF_ACQ_EI || prm->acq == MAF_ACQ_PI))
    return fail(ctx, "an explicit level is required when no training set is loaded");
  if (resolve_acq(ctx, prm, &ap)) return 1;
  const size_t sq = (size_t)S * q, sqq = sq * q;
  if (ensure(ctx, ctx->dmu, sq) || ensure(ctx, ctx->dmubar, sq) || ensure(ctx, ctx->dSig, sqq) ||
      ensure(ctx, ctx->dSigbar, sqq) || ensure(ctx, ctx->dchol, sqq) || ensure(ctx, ctx->dloss, (size_t)S) ||
      ensure(ctx, ctx->dz, (size_t)M * q))
    return 1;
  CK(cudaMemcpyAsync(ctx->dmu.p, mu, sizeof(double) * sq, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->dSig.p, cov, sizeof(double) * sqq, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->dz.p, z, sizeof(double) * M * q, cudaMemcpyHostToDevice, ctx->stream));
  const double *dw = nullptr, *dinc = nullptr;
  if (weights) {
    if (ensure(ctx, ctx->dw, (size_t)M)) return 1;
    CK(cudaMemcpyAsync(ctx->dw.p, weights, sizeof(double) * M, cudaMemcpyHostToDevice, ctx->stream));
    dw = ctx->dw.p;
  }
  if (incumbents) {
    if (ensure(ctx, ctx->dinc, (size_t)M)) return 1;
    CK(cudaMemcpyAsync(ctx->dinc.p, incumbents, sizeof(double) * M, cudaMemcpyHostToDevice, ctx->stream));
    dinc = ctx->dinc.p;
  }
  const int need_grad = (out_mubar != nullptr || out_covbar != nullptr) ? 1 : 0;
  {
    ProfScope ps(ctx, MAF_PROF_MC);
    DISPATCH_Q(q, {
      const size_t smem = sizeof(McSmem<Q>);
      if (smem > 48 * 1024 && !(ctx->attr_mc_mask & (1u << Q))) {
        CK(cudaFuncSetAttribute(mc_acq_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->attr_mc_mask |= 1u << Q;
      }
      mc_acq_kernel<Q><<<(S + MC_WARPS - 1) / MC_WARPS, MC_WARPS * 32, smem, ctx->stream>>>(
          S, M, ctx->dmu.p, ctx->dSig.p, ctx->dz.p, dw, dinc, ap, ctx->mp.jitter, need_grad, ctx->dloss.p,
          ctx->dmubar.p, ctx->dSigbar.p, ctx->dchol.p);
    });
    CKL();
  }
  if (out_loss) CK(cudaMemcpyAsync(out_loss, ctx->dloss.p, sizeof(double) * S, cudaMemcpyDeviceToHost, ctx->stream));
  if (out_mubar) CK(cudaMemcpyAsync(out_mubar, ctx->dmubar.p, sizeof(double) * sq, cudaMemcpyDeviceToHost, ctx->stream));
  if (out_covbar) CK(cudaMemcpyAsync(out_covbar, ctx->dSigbar.p, sizeof(double) * sqq, cudaMemcpyDeviceToHost, ctx->stream));
  if (out_chol) CK(cudaMemcpyAsync(out_chol, ctx->dchol.p, sizeof(double) * sqq, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->last_S = S;
  ctx->last_q = q;
  ctx->last_has_chol = true;
  ctx->last_dX = nullptr;                                       // moments came from the caller: no candidate sets on the device
  return 0;
}

// ---- fantasies ------------------------------------------------------------------------------------
extern "C" int maf_set_fantasies(maf_ctx* ctx, int F, const double* Y) {
  if (!ctx) return 1;
  ctx->generation++;
  if (!ctx->train_set) return fail(ctx, "maf_set_train must be called first");
  if (F < 1 || !Y) return fail(ctx, "bad arguments");
  CK(cudaSetDevice(ctx->device));
  const int n = ctx->n, npad = ctx->npad;
  const int Fpad = round_up(F, 128);
  std::vector<double> rho((size_t)Fpad * npad, 0.0), means(Fpad, 0.0), levels(Fpad, 0.0);
  const bool empirical = !(ctx->mean_param == ctx->mean_param);
  for (int f = 0; f < F; ++f) {
    const double* y = Y + (size_t)f * n;
    double m = ctx->mean_param, lo = y[0];
    if (empirical) {
      double acc = 0.0;
      for (int i = 0; i < n; ++i) acc += y[i];
      m = acc / n;
    }
    for (int i = 0; i < n; ++i) {
      rho[(size_t)f * npad + i] = y[i] - m;
      lo = std::min(lo, y[i]);
    }
    means[f] = m;
    levels[f] = lo;
  }
  ctx->flevel_min = *std::min_element(levels.begin(), levels.begin() + F);
  const size_t fn = (size_t)Fpad * npad;
  if (ensure(ctx, ctx->fGamma, fn) || ensure(ctx, ctx->fGammaT, fn) || ensure(ctx, ctx->fmeans, Fpad) ||
      ensure(ctx, ctx->flevels, Fpad) || ensure(ctx, ctx->Vt, fn))
    return 1;
  CK(cudaMemcpyAsync(ctx->Vt.p, rho.data(), sizeof(double) * fn, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->fmeans.p, means.data(), sizeof(double) * Fpad, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->flevels.p, levels.data(), sizeof(double) * Fpad, cudaMemcpyHostToDevice, ctx->stream));
  // Gamma[f][r] = sum_c rho_f[c] Linv[r][c]
  if (launch_gemm(ctx, MAF_PROF_GEMM_FWD, Fpad, npad, npad, ctx->Vt.p, npad, ctx->Linv, npad, ctx->fGamma.p, npad, 1)) return 1;
  {
    ProfScope ps(ctx, MAF_PROF_POSTERIOR);
    dim3 grid(npad / 32, Fpad / 32), block(32, 8);
    transpose_rect_kernel<<<grid, block, 0, ctx->stream>>>(ctx->fGamma.p, ctx->fGammaT.p, Fpad, npad);
    CKL();
  }
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->F = F;
  ctx->Fpad = Fpad;
  return 0;
}

// loss/grad (prm != null) or posterior (prm == null) over S single-point candidates, in chunks
// dev: X, out_loss and out_grad are DEVICE pointers, nothing is copied and the stream is not synchronised (the batched
// L-BFGS loop evaluates this way); otherwise host pointers.
static int run_fantasies(maf_ctx* ctx, const maf_acq_params* prm, int S, const double* X, double* out_loss,
                         double* out_grad, double* out_mu, double* out_var, bool dev = false) {
  if (ctx->F < 1) return fail(ctx, "maf_set_fantasies must be called first");
  if (S < 1 || !X) return fail(ctx, "bad arguments");
  CK(cudaSetDevice(ctx->device));
  const int d = ctx->mp.d, n = ctx->n, npad = ctx->npad, F = ctx->F, Fpad = ctx->Fpad;
  AcqDev ap{};
  int use_flevels = 1;
  if (prm) {
    if (resolve_acq(ctx, prm, &ap)) return 1;
    use_flevels = !(prm->level == prm->level);
    // negative_pi's default level is tf.reduce_min(outputs_old) over the WHOLE rank-4 tensor, one scalar for all
    // fantasies (negative_pi.py:32-34); negative_ei reduces along axis -2 only, one level per fantasy
    // (negative_ei.py:31-34).  Checked against the reference's code: tests/golden/ref_acq.npz, fan_negative_pi.
    if (use_flevels && prm->acq == MAF_ACQ_PI) {
      ap.level = ctx->flevel_min;
      use_flevels = 0;
    }
  } else {
    ap.acq = MAF_ACQ_SR;
  }
  const int sc_max = (int)std::min<size_t>(ctx->chunk_rows, (size_t)S);
  const int Jmax = round_up(sc_max, 256);
  if ((!dev && ensure(ctx, ctx->dX, (size_t)S * d)) || ensure(ctx, ctx->K10, (size_t)Jmax * npad) ||
      ensure(ctx, ctx->Vt, (size_t)Jmax * npad) || ensure(ctx, ctx->fVG, (size_t)Jmax * Fpad) ||
      ensure(ctx, ctx->fmubar, (size_t)Jmax * Fpad) || ensure(ctx, ctx->dmu, (size_t)S) ||
      ensure(ctx, ctx->dSig, (size_t)S) || ensure(ctx, ctx->dSigbar, (size_t)S) || ensure(ctx, ctx->dloss, (size_t)S) ||
      ensure(ctx, ctx->dgrad, (size_t)S * d))
    return 1;
  if (!dev) CK(cudaMemcpyAsync(ctx->dX.p, X, sizeof(double) * (size_t)S * d, cudaMemcpyHostToDevice, ctx->stream));
  const double* const Xd = dev ? X : ctx->dX.p;
  double* const lossd = dev ? out_loss : ctx->dloss.p;
  double* const gradd = dev ? out_grad : ctx->dgrad.p;
  const int need_grad = out_grad != nullptr;
  for (int s0 = 0; s0 < S; s0 += sc_max) {
    const int sc = std::min(sc_max, S - s0);
    const int Jpad = round_up(sc, 256);
    const double* dX = Xd + (size_t)s0 * d;
    if (ctx->i8_ready) {
      // same contraction engine as the Monte-Carlo path: K10 straight to digit planes, INT8 products
      const size_t bytes = (size_t)ctx->oz_ns * Jpad * npad;
      if (ensure(ctx, ctx->ozA, (bytes + 7) / 8) || ensure(ctx, ctx->ozsA, (size_t)Jpad)) return 1;
      {
        ProfScope ps(ctx, MAF_PROF_CROSS_FWD);
        dim3 grid((npad + OZC_THREADS * 4 - 1) / (OZC_THREADS * 4), Jpad / OZC_TJ);
        DISPATCH_DKN(d, ctx->mp.kernel_id, ctx->oz_ns, cross_fwd_planes_kernel<D, KID, NSC><<<grid, OZC_THREADS, 0, ctx->stream>>>(dX, sc, Jpad, ctx->X0s, n, npad, ctx->mp, ctx->oz_ns,
                                                                                    (int8_t*)ctx->ozA.p, ctx->ozsA.p));
        CKL();
      }
      if (launch_i8gemm(ctx, MAF_PROF_GEMM_FWD, Jpad, npad, npad, ctx->ozA, ctx->ozsA, ctx->ozTL, ctx->ozsTL, ctx->Vt.p, npad, 1,
                        ctx->oz_ns, ctx->oz_lmax))
        return 1;
    } else {
      {
        ProfScope ps(ctx, MAF_PROF_CROSS_FWD);
        dim3 grid((npad + CROSS_TK - 1) / CROSS_TK, Jpad / CROSS_TJ);
        DISPATCH_D(d, cross_fwd_kernel<D><<<grid, CROSS_TK, 0, ctx->stream>>>(dX, sc, ctx->X0s, n, npad, ctx->mp, ctx->K10.p));
        CKL();
      }
      if (launch_gemm(ctx, MAF_PROF_GEMM_FWD, Jpad, npad, npad, ctx->K10.p, npad, ctx->Linv, npad, ctx->Vt.p, npad, 1)) return 1;
    }
    {
      ProfScope ps(ctx, MAF_PROF_POSTERIOR);
      posterior_kernel<1><<<sc, POST_THREADS, 0, ctx->stream>>>(ctx->Vt.p, npad, ctx->gamma, dX, ctx->mp, ctx->mean_val,
                                                               ctx->dmu.p + s0, ctx->dSig.p + s0);
      CKL();
    }
    // VG[s][f] = v_s . gamma_f
    if (launch_gemm(ctx, MAF_PROF_GEMM_FWD, Jpad, Fpad, npad, ctx->Vt.p, npad, ctx->fGamma.p, npad, ctx->fVG.p, Fpad, 0)) return 1;
    {
      ProfScope ps(ctx, MAF_PROF_MC);
      closed_form_fant_kernel<<<(sc + 3) / 4, 128, 0, ctx->stream>>>(
          sc, F, Fpad, ctx->fVG.p, ctx->fmeans.p, ctx->flevels.p, ctx->dSig.p + s0, ap, use_flevels, need_grad,
          lossd + s0, ctx->fmubar.p, ctx->dSigbar.p + s0, prm ? nullptr : ctx->fVG.p);
      CKL();
    }
    if (!prm && out_mu) {
      // fVG now holds mu[s][f]; transpose into the caller's [F][S] layout on the host
      std::vector<double> h((size_t)sc * Fpad);
      CK(cudaMemcpyAsync(h.data(), ctx->fVG.p, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      for (int f = 0; f < F; ++f)
        for (int s = 0; s < sc; ++s) out_mu[(size_t)f * S + s0 + s] = h[(size_t)s * Fpad + f];
    }
    if (!need_grad) continue;
    if (Jpad > sc) CK(cudaMemsetAsync(ctx->fmubar.p + (size_t)sc * Fpad, 0, sizeof(double) * (size_t)(Jpad - sc) * Fpad, ctx->stream));
    // G = mubar Gamma  -> K10 buffer ; Vbar = G - 2 varbar V ; Kbar = Vbar Linv ; xbar
    if (launch_gemm(ctx, MAF_PROF_GEMM_BWD, Jpad, npad, Fpad, ctx->fmubar.p, Fpad, ctx->fGammaT.p, Fpad, ctx->K10.p, npad, 0)) return 1;
    {
      ProfScope ps(ctx, MAF_PROF_VBAR);
      dim3 grid((npad + 255) / 256, sc);
      fant_vbar_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->Vt.p, ctx->K10.p, ctx->dSigbar.p + s0, sc, npad);
      CKL();
    }
    if (ctx->i8_ready) {
      if (oz_slice(ctx, MAF_PROF_VBAR, ctx->Vt.p, npad, sc, Jpad, npad, ctx->oz_ns_bwd, 0.0, ctx->ozA, ctx->ozsA)) return 1;
      if (launch_i8gemm(ctx, MAF_PROF_GEMM_BWD, Jpad, npad, npad, ctx->ozA, ctx->ozsA, ctx->ozTU, ctx->ozsTU, ctx->K10.p, npad, 2,
                        ctx->oz_ns_bwd, ctx->oz_lmax_bwd))
        return 1;
    } else if (launch_gemm(ctx, MAF_PROF_GEMM_BWD, Jpad, npad, npad, ctx->Vt.p, npad, ctx->LinvT, npad, ctx->K10.p, npad, 2)) {
      return 1;
    }
    {
      ProfScope ps(ctx, MAF_PROF_CROSS_BWD);
      DISPATCH_DK(d, ctx->mp.kernel_id, cross_bwd_kernel<D, KID, false><<<sc, 32, 0, ctx->stream>>>(dX, sc, 1, 0, ctx->X0s, n, npad, ctx->mp, ctx->K10.p,
                                                                     ctx->dSigbar.p + s0, gradd + (size_t)s0 * d));
      CKL();
    }
  }
  ctx->last_S = 0;   // the per-restart extras buffers were reused
  if (dev) return 0;
  if (out_loss) CK(cudaMemcpyAsync(out_loss, ctx->dloss.p, sizeof(double) * S, cudaMemcpyDeviceToHost, ctx->stream));
  if (out_grad) CK(cudaMemcpyAsync(out_grad, ctx->dgrad.p, sizeof(double) * (size_t)S * d, cudaMemcpyDeviceToHost, ctx->stream));
  if (out_var) CK(cudaMemcpyAsync(out_var, ctx->dSig.p, sizeof(double) * S, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int maf_fantasies_value_grad(maf_ctx* ctx, const maf_acq_params* prm, int S, const double* X,
                                        double* out_loss, double* out_grad) {
  if (!ctx) return 1;
  if (!prm || !out_loss) return fail(ctx, "null pointer");
  return run_fantasies(ctx, prm, S, X, out_loss, out_grad, nullptr, nullptr);
}

extern "C" int maf_fantasies_posterior(maf_ctx* ctx, int S, const double* X, double* mu, double* var) {
  if (!ctx) return 1;
  return run_fantasies(ctx, nullptr, S, X, nullptr, nullptr, mu, var);
}

extern "C" int maf_last_extras(maf_ctx* ctx, double* mu, double* cov, double* chol) {
  if (!ctx) return 1;
  if (ctx->last_S == 0) return fail(ctx, "no evaluation to report");
  CK(cudaSetDevice(ctx->device));
  const size_t sq = (size_t)ctx->last_S * ctx->last_q;
  if (mu) CK(cudaMemcpyAsync(mu, ctx->dmu.p, sizeof(double) * sq, cudaMemcpyDeviceToHost, ctx->stream));
  if (cov) CK(cudaMemcpyAsync(cov, ctx->dSig.p, sizeof(double) * sq * ctx->last_q, cudaMemcpyDeviceToHost, ctx->stream));
  if (chol) {
    if (!ctx->last_has_chol) return fail(ctx, "the last evaluation did not factorise the covariance");
    CK(cudaMemcpyAsync(chol, ctx->dchol.p, sizeof(double) * sq * ctx->last_q, cudaMemcpyDeviceToHost, ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int maf_argmin_last(maf_ctx* ctx, int64_t* idx, double* value) {
  if (!ctx) return 1;
  if (ctx->last_S == 0 || !ctx->dloss.p) return fail(ctx, "no evaluation to report");
  CK(cudaSetDevice(ctx->device));
  {
    ProfScope ps(ctx, MAF_PROF_ADAM);
    argmin_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->dloss.p, ctx->last_S, ctx->d_argmin_val, ctx->d_argmin_idx);
    CKL();
  }
  long long hi = -1;
  double hv = 0.0;
  CK(cudaMemcpyAsync(&hi, ctx->d_argmin_idx, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(&hv, ctx->d_argmin_val, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (idx) *idx = (int64_t)hi;
  if (value) *value = hv;
  return 0;
}

extern "C" int maf_best_record_dev(maf_ctx* ctx, int64_t index_offset, double* d_rec) {
  if (!ctx) return 1;
  if (!d_rec) return fail(ctx, "null device pointer");
  if (ctx->last_S == 0 || !ctx->dloss.p || !ctx->last_dX) return fail(ctx, "no evaluation to report");
  CK(cudaSetDevice(ctx->device));
  ProfScope ps(ctx, MAF_PROF_ADAM);
  argmin_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->dloss.p, ctx->last_S, ctx->d_argmin_val, ctx->d_argmin_idx);
  best_record_kernel<<<1, 128, 0, ctx->stream>>>(ctx->d_argmin_val, ctx->d_argmin_idx, ctx->last_dX, ctx->last_q * ctx->mp.d,
                                                 (long long)index_offset, d_rec);
  CKL();
  return 0;
}

// ---- Adam loop ---------------------------------------------------------------------------------
extern "C" int maf_adam_begin(maf_ctx* ctx, const maf_acq_params* prm, int S, int q, int p_pending, const double* X,
                              double lr, double beta1, double beta2, double eps) {
  if (!ctx) return 1;
  if (!prm || !X) return fail(ctx, "null pointer");
  if (check_shape(ctx, S, q, p_pending)) return 1;
  CK(cudaSetDevice(ctx->device));
  const int d = ctx->mp.d;
  const size_t nx = (size_t)S * q * d, nt = (size_t)S * (q - p_pending) * d;
  if (ensure(ctx, ctx->adam_X, nx) || ensure(ctx, ctx->adam_m, nt) || ensure(ctx, ctx->adam_v, nt) ||
      ensure(ctx, ctx->adam_g, nt))
    return 1;
  CK(cudaMemcpyAsync(ctx->adam_X.p, X, nx * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(ctx->adam_m.p, 0, sizeof(double) * nt, ctx->stream));
  CK(cudaMemsetAsync(ctx->adam_v.p, 0, sizeof(double) * nt, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->generation++;                                       // new session: the replayed step is captured again
  ctx->adam_prm = *prm;
  ctx->adam_S = S;
  ctx->adam_q = q;
  ctx->adam_p = p_pending;
  ctx->pend_shared = pending_rows_shared(X, S, q, p_pending, d);   // the update kernels never touch the pending rows
  ctx->adam_lr = lr;
  ctx->adam_b1 = beta1;
  ctx->adam_b2 = beta2;
  ctx->adam_eps = eps;
  ctx->adam_b1p = beta1;
  ctx->adam_b2p = beta2;
  ctx->adam_on = true;
  ctx->opt_method = MAF_OPT_ADAM;
  ctx->opt_momentum = 0.0;
  ctx->opt_t = 0;
  return 0;
}

extern "C" int maf_adam_set_method(maf_ctx* ctx, int method, double momentum) {
  if (!ctx) return 1;
  ctx->generation++;
  if (!ctx->adam_on) return fail(ctx, "maf_adam_begin must be called first");
  if (method != MAF_OPT_ADAM && method != MAF_OPT_GD && method != MAF_OPT_MOMENTUM) return fail(ctx, "unknown optimizer %d", method);
  if (ctx->opt_t != 0) return fail(ctx, "the optimizer cannot change after the first step");
  ctx->opt_method = method;
  ctx->opt_momentum = method == MAF_OPT_MOMENTUM ? momentum : 0.0;
  return 0;
}

// One optimizer step in the replayable form: nothing in the launch arguments depends on the step index (base samples,
// learning-rate coefficient and trace row are picked on the device from ctx->adam_ctr), so the kernel sequence can be
// captured once per session and replayed.  The arithmetic is the host-driven loop's, bit for bit.
static int adam_step_body(maf_ctx* ctx, bool closed, int M, size_t zsz, bool trace) {
  const int S = ctx->adam_S, q = ctx->adam_q, p = ctx->adam_p, d = ctx->mp.d;
  const size_t nt = (size_t)S * (q - p) * d;
  {
    ProfScope ps(ctx, MAF_PROF_ADAM);
    adam_step_in_kernel<<<1, 256, 0, ctx->stream>>>(ctx->adam_ctr, ctx->adam_z.p, zsz, ctx->adam_zcur.p);
    CKL();
  }
  if (run_all(ctx, &ctx->adam_prm, false, S, q, p, ctx->adam_X.p, M, closed ? nullptr : ctx->adam_zcur.p, nullptr, nullptr,
              ctx->dloss.p, ctx->adam_g.p, ctx->pend_shared))
    return 1;
  {
    ProfScope ps(ctx, MAF_PROF_ADAM);
    if (ctx->opt_method == MAF_OPT_ADAM)
      adam_clip_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, ctx->stream>>>(
          ctx->adam_X.p, ctx->adam_g.p, ctx->adam_m.p, ctx->adam_v.p, nt, q, p, d, 0.0, ctx->adam_b1, ctx->adam_b2,
          ctx->adam_eps, ctx->adam_coef.p, ctx->adam_ctr);
    else
      sgd_clip_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, ctx->stream>>>(ctx->adam_X.p, ctx->adam_g.p, ctx->adam_m.p, nt, q, p,
                                                                              d, 0.0, ctx->opt_momentum, ctx->adam_coef.p,
                                                                              ctx->adam_ctr);
    CKL();
  }
  if (trace) {
    ProfScope ps(ctx, MAF_PROF_ADAM);
    adam_step_out_kernel<<<(S + 255) / 256, 256, 0, ctx->stream>>>(ctx->adam_ctr, ctx->dloss.p, S, ctx->adam_trace.p);
    CKL();
  }
  return 0;
}

// Learning-rate coefficient of the next step, then advance the host copy of the optimizer state
// (TF-1 Adam: lr sqrt(1 - b2^t) / (1 - b1^t); GradientDescent / Momentum: lr (global_step + 1)^-0.7, :405-412).
static double adam_next_coef(maf_ctx* ctx) {
  const double c = ctx->opt_method == MAF_OPT_ADAM ? ctx->adam_lr * std::sqrt(1.0 - ctx->adam_b2p) / (1.0 - ctx->adam_b1p)
                                                   : ctx->adam_lr * std::pow((double)(ctx->opt_t + 1), -0.7);
  ctx->adam_b1p *= ctx->adam_b1;
  ctx->adam_b2p *= ctx->adam_b2;
  ctx->opt_t += 1;
  return c;
}

// Replayed loop: step 0 of the first call of a session is launched directly (it sizes every work buffer), one step is
// captured into a CUDA graph, and the graph is launched for the remaining steps of this and the following calls.
// Returns 0 on success, 1 on an error, -1 if the session cannot be replayed (nothing has been executed then).
static int adam_steps_replayed(maf_ctx* ctx, int steps, int M, bool closed, size_t zsz, const double* z_steps,
                               double* out_loss_trace) {
  const int S = ctx->adam_S;
  const bool trace = out_loss_trace != nullptr;
  if (!ctx->adam_ctr) CK(cudaMalloc((void**)&ctx->adam_ctr, 2 * sizeof(long long)));
  if (ensure(ctx, ctx->adam_z, std::max<size_t>(zsz * steps, 1)) || ensure(ctx, ctx->adam_zcur, std::max<size_t>(zsz, 1)) ||
      ensure(ctx, ctx->adam_coef, (size_t)steps) || ensure(ctx, ctx->dloss, (size_t)S) ||
      (trace && ensure(ctx, ctx->adam_trace, (size_t)steps * S)))
    return 1;
  std::vector<double> coef(steps);
  for (int t = 0; t < steps; ++t) coef[t] = adam_next_coef(ctx);
  CK(cudaMemsetAsync(ctx->adam_ctr, 0, 2 * sizeof(long long), ctx->stream));
  CK(cudaMemcpyAsync(ctx->adam_coef.p, coef.data(), sizeof(double) * steps, cudaMemcpyHostToDevice, ctx->stream));
  if (!closed) CK(cudaMemcpyAsync(ctx->adam_z.p, z_steps, sizeof(double) * zsz * steps, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));                        // coef goes out of scope; z_steps belongs to the caller
  int t = 0;
  const bool valid = ctx->adam_graph && ctx->adam_graph_gen == ctx->generation && ctx->adam_graph_M == M &&
                     ctx->adam_graph_trace == (int)trace;
  if (!valid) {
    if (ctx->adam_graph) cudaGraphExecDestroy(ctx->adam_graph), ctx->adam_graph = nullptr;
    if (adam_step_body(ctx, closed, M, zsz, trace)) return 1;    // direct launch: allocates, builds tile schedules
    t = 1;
    const long long gen = ctx->generation;                       // buffers have their final addresses now
    const int64_t before = ctx->launch_count;
    cudaGraph_t graph = nullptr;
    bool ok = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
    if (ok) {
      const int rc = adam_step_body(ctx, closed, M, zsz, trace);
      ok = cudaStreamEndCapture(ctx->stream, &graph) == cudaSuccess && rc == 0 && graph != nullptr && gen == ctx->generation;
    }
    ctx->adam_graph_kernels = ctx->launch_count - before;
    ctx->launch_count = before;                                  // captured, not launched
    if (ok) ok = cudaGraphInstantiate(&ctx->adam_graph, graph, 0) == cudaSuccess;
    if (graph) cudaGraphDestroy(graph);
    if (!ok) {                                                   // keep going from the host, one launch at a time
      cudaGetLastError();
      ctx->adam_graph = nullptr;
      ctx->graph_on = 0;
      for (; t < steps; ++t)
        if (adam_step_body(ctx, closed, M, zsz, trace)) return 1;
    } else {
      ctx->adam_graph_gen = gen;
      ctx->adam_graph_M = M;
      ctx->adam_graph_trace = (int)trace;
    }
  }
  for (; t < steps; ++t) {
    if (cudaGraphLaunch(ctx->adam_graph, ctx->stream) != cudaSuccess) return fail(ctx, "graph launch failed in the optimizer loop");
    ctx->launch_count += ctx->adam_graph_kernels;
  }
  if (trace)
    CK(cudaMemcpyAsync(out_loss_trace, ctx->adam_trace.p, sizeof(double) * (size_t)steps * S, cudaMemcpyDeviceToHost, ctx->stream));
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return fail(ctx, "stream sync failed in adam loop");
  return 0;
}

extern "C" int maf_adam_steps(maf_ctx* ctx, int steps, int M, const double* z_steps, double* out_loss_trace) {
  if (!ctx) return 1;
  if (!ctx->adam_on) return fail(ctx, "maf_adam_begin must be called first");
  const bool closed = (ctx->adam_q == 1 && z_steps == nullptr);      // q == 1: deterministic closed-form loss
  if (steps < 0 || (!closed && (M < 1 || !z_steps))) return fail(ctx, "bad arguments");
  CK(cudaSetDevice(ctx->device));
  const int S = ctx->adam_S, q = ctx->adam_q, p = ctx->adam_p, d = ctx->mp.d;
  const size_t nt = (size_t)S * (q - p) * d;
  const size_t zsz = closed ? 0 : (size_t)M * q;
  // four or more steps per call and no per-kernel timing: replay one captured step (MAF_GRAPH=0 keeps the host loop)
  if (ctx->graph_on && !ctx->prof && steps >= 4) return adam_steps_replayed(ctx, steps, closed ? 0 : M, closed, zsz, z_steps, out_loss_trace);
  if (ensure(ctx, ctx->adam_z, std::max<size_t>(zsz * std::max(steps, 1), 1)) || ensure(ctx, ctx->dloss, (size_t)S)) return 1;
  if (!closed)
    CK(cudaMemcpyAsync(ctx->adam_z.p, z_steps, sizeof(double) * zsz * steps, cudaMemcpyHostToDevice, ctx->stream));
  double* dtrace = nullptr;
  if (out_loss_trace && steps > 0) CK(cudaMalloc((void**)&dtrace, sizeof(double) * (size_t)steps * S));
  int rc = 0;
  for (int t = 0; t < steps && rc == 0; ++t) {
    double* lossp = dtrace ? dtrace + (size_t)t * S : ctx->dloss.p;
    rc = run_all(ctx, &ctx->adam_prm, false, S, q, p, ctx->adam_X.p, M, closed ? nullptr : ctx->adam_z.p + zsz * t,
                 nullptr, nullptr, lossp, ctx->adam_g.p, ctx->pend_shared);
    if (rc) break;
    const double coef = adam_next_coef(ctx);
    ProfScope ps(ctx, MAF_PROF_ADAM);
    if (ctx->opt_method == MAF_OPT_ADAM)
      adam_clip_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, ctx->stream>>>(
          ctx->adam_X.p, ctx->adam_g.p, ctx->adam_m.p, ctx->adam_v.p, nt, q, p, d, coef, ctx->adam_b1, ctx->adam_b2, ctx->adam_eps);
    else
      sgd_clip_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, ctx->stream>>>(ctx->adam_X.p, ctx->adam_g.p, ctx->adam_m.p, nt, q, p,
                                                                              d, coef, ctx->opt_momentum);
    if (cudaGetLastError() != cudaSuccess) rc = fail(ctx, "optimizer update kernel launch failed");
  }
  if (rc == 0 && dtrace) {
    if (cudaMemcpyAsync(out_loss_trace, dtrace, sizeof(double) * (size_t)steps * S, cudaMemcpyDeviceToHost,
                        ctx->stream) != cudaSuccess)
      rc = fail(ctx, "loss trace copy failed");
  }
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && rc == 0) rc = fail(ctx, "stream sync failed in adam loop");
  if (dtrace) cudaFree(dtrace);
  return rc;
}

extern "C" int maf_adam_peek(maf_ctx* ctx, double* X_out) {
  if (!ctx) return 1;
  if (!ctx->adam_on) return fail(ctx, "maf_adam_begin must be called first");
  if (!X_out) return fail(ctx, "null pointer");
  CK(cudaSetDevice(ctx->device));
  const size_t nx = (size_t)ctx->adam_S * ctx->adam_q * ctx->mp.d;
  CK(cudaMemcpyAsync(X_out, ctx->adam_X.p, nx * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int maf_adam_end(maf_ctx* ctx, double* X_out) {
  if (!ctx) return 1;
  if (!ctx->adam_on) return fail(ctx, "maf_adam_begin must be called first");
  CK(cudaSetDevice(ctx->device));
  const size_t nx = (size_t)ctx->adam_S * ctx->adam_q * ctx->mp.d;
  if (X_out) CK(cudaMemcpyAsync(X_out, ctx->adam_X.p, nx * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->adam_on = false;
  return 0;
}

// ---- batched L-BFGS on the device -----------------------------------------------------------------
__global__ void lbfgs_finish_kernel(LbfgsState st, int S, int q, int p, int d, double* __restrict__ Xt) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const int D = (q - p) * d;
  for (int t = 0; t < D; ++t) Xt[((size_t)s * q + p) * d + t] = st.x[(size_t)s * D + t];
}

extern "C" int maf_lbfgs_run(maf_ctx* ctx, const maf_acq_params* prm, int S, int q, int p_pending, double* X, int M,
                             const double* z, int fantasies, int maxiter, double ftol, double pgtol, double time_limit_s,
                             double* out_loss, int32_t* out_info, double* out_trace) {
  if (!ctx) return 1;
  if (!prm || !X) return fail(ctx, "null pointer");
  if (check_shape(ctx, S, q, p_pending)) return 1;
  if (maxiter < 1) return fail(ctx, "maxiter must be >= 1");
  if (fantasies && (q != 1 || ctx->F < 1)) return fail(ctx, "fantasy-conditioned losses are closed-form (q == 1) and need maf_set_fantasies");
  const bool closed = !fantasies && q == 1 && z == nullptr;
  if (!fantasies && !closed && (z == nullptr || M < 1)) return fail(ctx, "base samples z[M][q] are required for the Monte-Carlo path");
  CK(cudaSetDevice(ctx->device));
  const int d = ctx->mp.d, p = p_pending, D = (q - p) * d;
  if (D > LB_MAX_D) return fail(ctx, "(q - p) d = %d exceeds %d", D, LB_MAX_D);
  const size_t nx = (size_t)S * q * d, nd = (size_t)S * D;
  if (ensure(ctx, ctx->lb_X, nx) || ensure(ctx, ctx->lb_ft, (size_t)S) || ensure(ctx, ctx->lb_gt, nd) || ensure(ctx, ctx->lb_x, nd) ||
      ensure(ctx, ctx->lb_f, (size_t)S) || ensure(ctx, ctx->lb_g, nd) || ensure(ctx, ctx->lb_dir, nd) || ensure(ctx, ctx->lb_step, (size_t)S) ||
      ensure(ctx, ctx->lb_gd, (size_t)S) || ensure(ctx, ctx->lb_hs, nd * LB_M) || ensure(ctx, ctx->lb_hy, nd * LB_M) ||
      ensure(ctx, ctx->lb_rho, (size_t)S * LB_M) || (out_trace && ensure(ctx, ctx->lb_trace, (size_t)maxiter * S)))
    return 1;
  if ((size_t)S > ctx->lb_int_cap) {
    if (ctx->lb_int) CK(cudaFree(ctx->lb_int));
    ctx->lb_int = nullptr;
    CK(cudaMalloc((void**)&ctx->lb_int, sizeof(int) * (4 * (size_t)S + 1)));
    ctx->lb_int_cap = (size_t)S;
  }
  if (!ctx->lb_active_host) CK(cudaHostAlloc((void**)&ctx->lb_active_host, sizeof(int), cudaHostAllocDefault));
  int* ib = ctx->lb_int;
  const size_t cap = ctx->lb_int_cap;
  LbfgsState st{ctx->lb_x.p, ctx->lb_f.p, ctx->lb_g.p, ctx->lb_dir.p, ctx->lb_step.p, ctx->lb_gd.p, ctx->lb_hs.p, ctx->lb_hy.p,
                ctx->lb_rho.p, ib, ib + cap, ib + 2 * cap, ib + 3 * cap, ib + 4 * cap};
  CK(cudaMemcpyAsync(ctx->lb_X.p, X, nx * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(st.active, &S, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  const double* dz = nullptr;
  if (!fantasies && !closed) {
    if (ensure(ctx, ctx->dz, (size_t)M * q)) return 1;
    CK(cudaMemcpyAsync(ctx->dz.p, z, sizeof(double) * M * q, cudaMemcpyHostToDevice, ctx->stream));
    dz = ctx->dz.p;
  }
  CK(cudaStreamSynchronize(ctx->stream));                         // &S is a stack address
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  if (time_limit_s == time_limit_s && time_limit_s < 1e30) {
    CK(cudaEventCreate(&t0));
    CK(cudaEventCreate(&t1));
    CK(cudaEventRecord(t0, ctx->stream));
  }
  const bool lb_shared = pending_rows_shared(X, S, q, p, d);
  int it = 0, active = S, rc = 0;
  for (; it < maxiter && rc == 0; ++it) {
    if (fantasies)
      rc = run_fantasies(ctx, prm, S, ctx->lb_X.p, ctx->lb_ft.p, ctx->lb_gt.p, nullptr, nullptr, true);
    else
      rc = run_all(ctx, prm, false, S, q, p, ctx->lb_X.p, M, dz, nullptr, nullptr, ctx->lb_ft.p, ctx->lb_gt.p, lb_shared);
    if (rc) break;
    if (out_trace)
      CK(cudaMemcpyAsync(ctx->lb_trace.p + (size_t)it * S, ctx->lb_ft.p, sizeof(double) * S, cudaMemcpyDeviceToDevice, ctx->stream));
    {
      ProfScope ps(ctx, MAF_PROF_ADAM);
      lbfgs_step_kernel<<<(S + 127) / 128, 128, 0, ctx->stream>>>(st, S, q, p, d, ctx->lb_X.p, ctx->lb_ft.p, ctx->lb_gt.p, it == 0 ? 1 : 0, ftol, pgtol);
      CKL();
    }
    if ((it & 7) == 7 || it + 1 == maxiter) {                     // every 8 iterations: anything left to do? out of time?
      CK(cudaMemcpyAsync(ctx->lb_active_host, st.active, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      if (t0) CK(cudaEventRecord(t1, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      active = *ctx->lb_active_host;
      if (active <= 0) {
        ++it;
        break;
      }
      if (t0) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, t0, t1));
        if ((double)ms * 1e-3 >= time_limit_s) {
          ++it;
          break;
        }
      }
    }
  }
  if (t0) cudaEventDestroy(t0), cudaEventDestroy(t1);
  if (rc) return rc;
  {
    ProfScope ps(ctx, MAF_PROF_ADAM);
    lbfgs_finish_kernel<<<(S + 127) / 128, 128, 0, ctx->stream>>>(st, S, q, p, d, ctx->lb_X.p);
    CKL();
  }
  CK(cudaMemcpyAsync(X, ctx->lb_X.p, nx * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (out_loss) CK(cudaMemcpyAsync(out_loss, ctx->lb_f.p, sizeof(double) * S, cudaMemcpyDeviceToHost, ctx->stream));
  if (out_trace) CK(cudaMemcpyAsync(out_trace, ctx->lb_trace.p, sizeof(double) * (size_t)it * S, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(ctx->lb_active_host, st.active, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (out_info) {
    out_info[0] = it;                                             // value+gradient evaluations of every restart
    out_info[1] = std::max(0, *ctx->lb_active_host);              // restarts that had not converged
  }
  ctx->last_S = 0;
  return 0;
}

// ---- profiling -----------------------------------------------------------------------------------
extern "C" int maf_profile_enable(maf_ctx* ctx, int on) {
  if (!ctx) return 1;
  if (ctx->prof && !on) prof_collect(ctx);
  ctx->prof = on != 0;
  return 0;
}
extern "C" int maf_profile_read(maf_ctx* ctx, double* ms, int64_t* launches) {
  if (!ctx) return 1;
  CK(cudaSetDevice(ctx->device));
  prof_collect(ctx);
  for (int s = 0; s < MAF_PROF_NSLOTS; ++s) {
    if (ms) ms[s] = ctx->slots[s].ms;
    if (launches) launches[s] = ctx->slots[s].launches;
  }
  return 0;
}
extern "C" int maf_profile_reset(maf_ctx* ctx) {
  if (!ctx) return 1;
  prof_collect(ctx);
  for (auto& ps : ctx->slots) {
    ps.ms = 0.0;
    ps.launches = 0;
  }
  return 0;
}
extern "C" int64_t maf_launch_count(const maf_ctx* ctx) { return ctx ? ctx->launch_count : 0; }

extern "C" int maf_timer_start(maf_ctx* ctx) {
  if (!ctx) return 1;
  CK(cudaSetDevice(ctx->device));
  for (auto& e : ctx->timer_ev)
    if (!e) CK(cudaEventCreate(&e));
  CK(cudaEventRecord(ctx->timer_ev[0], ctx->stream));
  return 0;
}
extern "C" int maf_timer_stop(maf_ctx* ctx, double* ms) {
  if (!ctx) return 1;
  if (!ctx->timer_ev[0]) return fail(ctx, "maf_timer_start must be called first");
  CK(cudaSetDevice(ctx->device));
  CK(cudaEventRecord(ctx->timer_ev[1], ctx->stream));
  CK(cudaEventSynchronize(ctx->timer_ev[1]));
  float f = 0.f;
  CK(cudaEventElapsedTime(&f, ctx->timer_ev[0], ctx->timer_ev[1]));
  if (ms) *ms = (double)f;
  return 0;
}
extern "C" void* maf_host_alloc(maf_ctx* ctx, size_t bytes) {
  void* p = nullptr;
  cudaSetDevice(ctx->device);
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
    fail(ctx, "maf_host_alloc(%zu) failed", bytes);
    return nullptr;
  }
  return p;
}
extern "C" int maf_host_free(maf_ctx* ctx, void* p) {
  CK(cudaSetDevice(ctx->device));
  CK(cudaFreeHost(p));
  return 0;
}

extern "C" int maf_set_gemm_mode(maf_ctx* ctx, int mode) {
  if (!ctx) return 1;
  ctx->generation++;
  if (mode != 0 && mode != 1) return fail(ctx, "unknown gemm mode %d", mode);
  ctx->gemm_i8 = mode;
  ctx->i8_ready = false;
  ctx->train_set = false;       // the factor planes are built by maf_set_train
  return 0;
}

extern "C" int maf_set_digits(maf_ctx* ctx, int ns_fwd, int ns_bwd) {
  if (!ctx) return 1;
  if (ns_fwd < 3 || ns_fwd > ctx->oz_ns || ns_bwd < 3 || ns_bwd > ctx->oz_ns)
    return fail(ctx, "digit planes must lie in 3..%d (got %d forward, %d reverse)", ctx->oz_ns, ns_fwd, ns_bwd);
  ctx->generation++;                                            // a captured optimizer step holds the old launch shapes
  ctx->oz_ns_fwd = ns_fwd;
  ctx->oz_ns_bwd = ns_bwd;
  ctx->oz_lmax_bwd = ns_bwd + 1;
  return 0;
}

extern "C" int maf_gemm_nt_i8_dev(maf_ctx* ctx, int J, int N, int K, const double* dA, int lda, const double* dT,
                                  int ldt, double* dC, int ldc, int tri, int ns, int lmax) {
  if (!ctx) return 1;
  CK(cudaSetDevice(ctx->device));
  Buf pa, pt, sa, st;
  int rc = oz_slice(ctx, MAF_PROF_CROSS_FWD, dA, lda, J, J, K, ns, 0.0, pa, sa) ||
           oz_slice(ctx, MAF_PROF_CROSS_FWD, dT, ldt, N, N, K, ns, 0.0, pt, st) ||
           launch_i8gemm(ctx, tri == 2 ? MAF_PROF_GEMM_BWD : MAF_PROF_GEMM_FWD, J, N, K, pa, sa, pt, st, dC, ldc, tri, ns, lmax);
  cudaStreamSynchronize(ctx->stream);
  for (Buf* b : {&pa, &pt, &sa, &st})
    if (b->p) cudaFree(b->p);
  return rc;
}

// ---- raw GEMM ----------------------------------------------------------------------------------------
extern "C" int maf_gemm_nt_dev(maf_ctx* ctx, int J, int N, int K, const double* dA, int lda, const double* dT,
                               int ldt, double* dC, int ldc, int tri) {
  if (!ctx) return 1;
  CK(cudaSetDevice(ctx->device));
  return launch_gemm(ctx, tri == 2 ? MAF_PROF_GEMM_BWD : MAF_PROF_GEMM_FWD, J, N, K, dA, lda, dT, ldt, dC, ldc, tri);
}
