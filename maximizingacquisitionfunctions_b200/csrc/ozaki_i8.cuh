// FP64-grade contraction on the INT8 tensor pipe (tcgen05.mma kind::i8) by exact digit slicing.
//
//   C[J][N] = A[J][K] * T[N][K]^T          (same contract as gemm_nt_f64_kernel, incl. triangular T)
//
// Each row of A and T is scaled by a power of two and written as NS signed base-256 digits
// (int8 planes):  x = 2^e * sum_{i=1..NS} d_i 2^(1-8i),  d_i in [-128,127]  (exact 8*NS-1 bit fixed point).
// Products of digit planes are exact in int32 (|d d'| <= 2^14, K <= 8192, <= 7 pairs per level), so
//   C = 2^(eA_j + eT_r) * sum_{l=2..LMAX} 2^(2-8l) * sum_{i+i'=l} (A_i T_i'^T)
// with the only error being the dropped levels l > LMAX.  NS=7, LMAX=8 (28 pair products) reproduces the
// FP64 posterior to ~2e-13 (tests/test_gpu_ozaki.py), i.e. the same grade as the DMMA path, at the INT8
// rate of the part (4.58 POP/s measured vs 0.036 PFLOP/s FP64).
//
// This file: the digit-plane producers (slicing of an FP64 matrix, fused cross-covariance and Vbar producers) and the
// GENERIC contraction kernel, which accepts any (ns, lmax) at run time.  The configurations the pipeline uses
// (lmax = ns + 1, 3 <= ns <= 7) run on the statically scheduled kernels of ozaki_i8_v3.cuh (single CTA) and
// ozaki_i8_v4.cuh (CTA pairs, default), which are 2x faster; this one is their fallback and their reference in the tests.
//
// Generic kernel: one 128x128 output tile per CTA, 6 warps: warp 0 = TMA producer (cp.async.bulk.tensor 3-D boxes
// {32 B, 128 rows, 1 plane}, SWIZZLE_32B, 3 smem stages, mbarrier full/empty), warp 1 = MMA issuer
// (one thread, tcgen05.mma.cta_group::1.kind::i8 M=128 N=128 K=32, accumulators in TMEM, one per level),
// warps 2-5 = epilogue (tcgen05.ld -> int32 -> FP64 recombination -> global).  TMEM holds 4 level
// accumulators of 128 columns (512 columns), so the levels are processed in passes of <= 4
// (levels 2-5: 10 pairs, levels 6-8: 18 pairs); pass 2 adds into C.
#pragma once
#include <cuda.h>
#include "common.cuh"

#define OZ_MAXS 7
#define OZ_BM 128
#define OZ_BN 128
#define OZ_BK 32                                   // bytes (= int8 elements) per k-chunk = one MMA K
#define OZ_STAGES 3
#define OZ_TILE_BYTES (128 * OZ_BK)                // one digit plane tile
#define OZ_STAGE_BYTES (2 * OZ_MAXS * OZ_TILE_BYTES)
#define OZ_SMEM_BYTES (OZ_STAGES * OZ_STAGE_BYTES + 1024)
#define OZ_THREADS 192

__device__ __forceinline__ uint32_t oz_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void oz_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(oz_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void oz_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tOZ_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra OZ_DONE;\n\tbra OZ_WAIT;\n\tOZ_DONE:\n\t}" ::"r"(oz_smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void oz_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(oz_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void oz_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(oz_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void oz_tma_load_3d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(oz_smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(oz_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// K-major, SWIZZLE_32B canonical tile [rows][32 B]: 8-row groups are 256 B apart
__device__ __forceinline__ uint64_t oz_desc_sw32(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                 // LBO: unused for swizzled K-major
  d |= (uint64_t)(256 >> 4) << 32;        // SBO
  d |= (uint64_t)1 << 46;                 // descriptor version (sm_100)
  d |= (uint64_t)6 << 61;                 // SWIZZLE_32B
  return d;
}
__device__ __forceinline__ void oz_mma_i8(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void oz_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(oz_smem_u32(bar)) : "memory");
}

struct OzParams {
  int K;            // contraction length (bytes per digit-plane row), multiple of 128
  int ns;           // digit planes in use (<= OZ_MAXS)
  int lmax;         // highest level kept (2 .. 2*ns)
  int tri;          // 0 full, 1 T lower-triangular, 2 upper
  int ldc;
  int sj;           // persistent kernel: row blocks per super-row of the tile order
  unsigned long long hintA, hintT;   // L2 eviction policies of the A / T plane loads (persistent kernel)
  int pf;           // persistent kernel: TMA-prefetch plane tiles into L2 two k-chunks ahead
  long long* stats; // MAF_OZ_DIAG builds: per-CTA cycle counters of the issuer / epilogue (or null)
  int dbg;          // diagnostics only (results invalid): 1 = issue no MMAs, 2 = issue no TMA loads, 4 = skip the epilogue body
  // CTA-pair kernel: balanced static schedule.  Pair c walks tile_list[tile_begin[c] .. tile_begin[c + 1]) (positions in
  // the supertiled order); null = round-robin t = c, c + pairs, ...  Built on the host by list scheduling (each tile of
  // the global order goes to the pair that becomes free first under the cost model kchunks + overhead), so the pairs
  // still work inside a moving window of the tile order, but none is left with 1.5 % more k-chunks than the mean.
  // Entries with OZ_TILE_EXPLICIT set name the tile directly (row block * nN + column block): the host is then free to
  // choose any order, e.g. column groups whose T planes stay in L2 while the A planes stream past (MAF_OZ_GROUPS).
  const int* tile_list;
  const int* tile_begin;
  // register-stash epilogue only: row maxima of |C| (atomicMax on the bit patterns; the caller zeroes the array).  The
  // Vbar producer takes its row scales from them instead of sweeping V^T once more (17 GB per step at the headline size).
  double* rowmax;
  // CTA-pair kernel, second pass of the level switch: only the first *nrows_dev rows of A (a device-side count, rounded up
  // to whole 256-row blocks) are contracted; tiles of later row blocks are skipped by all three roles.  null = all rows.
  const int* nrows_dev;
  int nrows_mul;    // rows = *nrows_dev x nrows_mul (restarts x pool size)
  // register-stash epilogue only: C[j][r] += row_add[j] * col_add[r] (one FMA on the rounded product, exactly what a
  // consumer adding the rank-one term to the stored C would compute).  The reverse contraction uses it for the mean part
  // of d loss / d K10, mubar_j alpha_r, which never passes through the digit planes.  null = nothing added.
  const double* row_add;
  const double* col_add;
};
#define OZ_TILE_EXPLICIT 0x40000000

__global__ void __launch_bounds__(OZ_THREADS, 1)
ozaki_i8_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmT,
                     const double* __restrict__ sA /*[J] row scales 2^eA*/, const double* __restrict__ sT /*[N]*/,
                     double* __restrict__ C, OzParams p) {
  extern __shared__ uint8_t oz_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(oz_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[OZ_STAGES], empty_bar[OZ_STAGES], acc_full, acc_empty;
  __shared__ uint32_t tmem_base_sh;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rb = (p.tri == 1) ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;
  const int r0 = rb * OZ_BN, j0 = blockIdx.y * OZ_BM;
  const int c_begin = (p.tri == 2) ? r0 : 0;
  const int c_end = (p.tri == 1) ? min(p.K, r0 + OZ_BN) : p.K;
  const int kchunks = (c_end - c_begin) / OZ_BK;
  const int npass = (p.lmax - 1 + 3) / 4;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < OZ_STAGES; ++s) {
      oz_mbar_init(&full_bar[s], 1);
      oz_mbar_init(&empty_bar[s], 1);
    }
    oz_mbar_init(&acc_full, 1);
    oz_mbar_init(&acc_empty, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmT)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem_u32(&tmem_base_sh)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base_sh;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int pass = 0; pass < npass; ++pass) {
        const int lo = 2 + 4 * pass, hi = min(p.lmax, lo + 3);
        const int smax = min(p.ns, hi - 1);                 // planes 1..smax take part in levels <= hi
        for (int kc = 0; kc < kchunks; ++kc) {
          oz_mbar_wait(&empty_bar[stage], phase ^ 1);
          if (p.dbg & 2) { oz_mbar_arrive(&full_bar[stage]); if (++stage == OZ_STAGES) { stage = 0; phase ^= 1; } continue; }
          oz_mbar_expect_tx(&full_bar[stage], 2u * smax * OZ_TILE_BYTES);
          uint8_t* st = smem + stage * OZ_STAGE_BYTES;
          const int c = c_begin + kc * OZ_BK;
          for (int i = 0; i < smax; ++i) {
            oz_tma_load_3d(st + i * OZ_TILE_BYTES, &tmA, &full_bar[stage], c, j0, i);
            oz_tma_load_3d(st + (OZ_MAXS + i) * OZ_TILE_BYTES, &tmT, &full_bar[stage], c, r0, i);
          }
          if (++stage == OZ_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      // instruction descriptor: D = S32, A = B = signed int8, K-major, N = 128, M = 128
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BN >> 3) << 17) | ((uint32_t)(OZ_BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      for (int pass = 0; pass < npass; ++pass) {
        const int lo = 2 + 4 * pass, hi = min(p.lmax, lo + 3);
        if (pass > 0) {
          oz_mbar_wait(&acc_empty, (uint32_t)((pass - 1) & 1));   // epilogue has drained the accumulators
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        for (int kc = 0; kc < kchunks; ++kc) {
          oz_mbar_wait(&full_bar[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sbase = oz_smem_u32(smem + stage * OZ_STAGE_BYTES);
          for (int l = lo; l <= hi; ++l) {
            const uint32_t d_tmem = tbase + (uint32_t)((l - lo) * OZ_BN);
            bool first = (kc == 0);
            for (int i = max(1, l - p.ns); i <= min(p.ns, l - 1); ++i) {
              const int ip = l - i;
              const uint64_t da = oz_desc_sw32(sbase + (uint32_t)((i - 1) * OZ_TILE_BYTES));
              const uint64_t db = oz_desc_sw32(sbase + (uint32_t)((OZ_MAXS + ip - 1) * OZ_TILE_BYTES));
              if (!(p.dbg & 1)) oz_mma_i8(d_tmem, da, db, idesc, first ? 0u : 1u);
              first = false;
            }
          }
          oz_commit(&empty_bar[stage]);                     // frees the smem stage when the MMAs retire
          if (++stage == OZ_STAGES) { stage = 0; phase ^= 1; }
        }
        oz_commit(&acc_full);                               // accumulators of this pass are complete
      }
    }
  } else {
    // ================= epilogue: TMEM -> FP64 -> global =================
    const int quarter = warp & 3;                           // TMEM lanes [32*quarter, +32) belong to this warp
    const int row = quarter * 32 + lane;
    const size_t j = (size_t)(j0 + row);
    const double srow = sA[j];
    double* crow = C + j * p.ldc + r0;
    const uint32_t lane_addr = tbase + ((uint32_t)(quarter * 32) << 16);
    for (int pass = 0; pass < npass; ++pass) {
      const int lo = 2 + 4 * pass, hi = min(p.lmax, lo + 3);
      oz_mbar_wait(&acc_full, (uint32_t)(pass & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int c0 = 0; c0 < ((p.dbg & 4) ? 0 : OZ_BN); c0 += 8) {
        double v[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) v[t] = 0.0;
        for (int l = hi; l >= lo; --l) {                    // smallest contributions first
          uint32_t r[8];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                       : "r"(lane_addr + (uint32_t)((l - lo) * OZ_BN + c0)));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const double w = exp2((double)(2 - 8 * l));
#pragma unroll
          for (int t = 0; t < 8; ++t) v[t] += (double)(int)r[t] * w;
        }
#pragma unroll
        for (int t = 0; t < 8; t += 2) {
          double2 o = make_double2(v[t] * srow * sT[r0 + c0 + t], v[t + 1] * srow * sT[r0 + c0 + t + 1]);
          double2* dst = reinterpret_cast<double2*>(crow + c0 + t);
          if (pass > 0) {
            double2 old = *dst;
            o.x += old.x;
            o.y += old.y;
          }
          *dst = o;
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      oz_mbar_arrive(&acc_empty);
    }
  }
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Digit slicing:  row r of X[R][K] (FP64, leading dimension ldx) -> OZ planes [ns][Rpad][K] of int8 digits
// and the row scale 2^e.  e is chosen from the row maximum (|x| 2^-e <= 0.98) or, when `fixed_max` > 0,
// from that bound (cross-covariances are bounded by a^2, no reduction pass needed).
// One CTA per row; the row is read once (kept in registers across the two phases: 16 values per thread).
// ---------------------------------------------------------------------------------------------
// Balanced base-256 digits of 4 fixed-point values -> one packed int per plane.
// I = sum_i d_i 256^i with d_i in [-128,127]  <=>  I + sum_i 128 256^i = sum_i (d_i + 128) 256^i, the ordinary
// unsigned bytes; so one 64-bit add and an XOR with 0x80 per byte give all digits at once, and three PRMTs per
// plane gather byte b of the four values.  The value is left-aligned to 7 digits so that plane i is byte 6 - i.
__device__ __forceinline__ void oz_emit_digits4(const long long (&I)[4], int ns, int8_t* __restrict__ dst, size_t plane_stride) {
  const unsigned long long bias = 0x0080808080808080ull >> (8 * (OZ_MAXS - ns));
  const int sh = 8 * (OZ_MAXS - ns);
  unsigned lo[4], hi[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const unsigned long long u = (((unsigned long long)I[t] + bias) << sh) ^ 0x0080808080808080ull;
    lo[t] = (unsigned)u;
    hi[t] = (unsigned)(u >> 32);
  }
#pragma unroll
  for (int i = 0; i < OZ_MAXS; ++i) {
    if (i < ns) {
      const int b = 6 - i;                                   // byte of the left-aligned value
      const unsigned sel = (unsigned)((b & 3) | (((b & 3) + 4) << 4));
      const unsigned p01 = (b < 4) ? __byte_perm(lo[0], lo[1], sel) : __byte_perm(hi[0], hi[1], sel);
      const unsigned p23 = (b < 4) ? __byte_perm(lo[2], lo[3], sel) : __byte_perm(hi[2], hi[3], sel);
      *reinterpret_cast<unsigned*>(dst + i * plane_stride) = __byte_perm(p01, p23, 0x5410);
    }
  }
}
// compile-time digit count: plane addresses by pointer increments, no run-time shifts
template <int NS>
__device__ __forceinline__ void oz_emit_digits4_t(const long long (&I)[4], int8_t* __restrict__ dst, size_t plane_stride) {
  constexpr unsigned long long bias = 0x0080808080808080ull >> (8 * (OZ_MAXS - NS));
  constexpr int sh = 8 * (OZ_MAXS - NS);
  unsigned lo[4], hi[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const unsigned long long u = (((unsigned long long)I[t] + bias) << sh) ^ 0x0080808080808080ull;
    lo[t] = (unsigned)u;
    hi[t] = (unsigned)(u >> 32);
  }
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    const int b = 6 - i;
    const unsigned sel = (unsigned)((b & 3) | (((b & 3) + 4) << 4));
    const unsigned p01 = (b < 4) ? __byte_perm(lo[0], lo[1], sel) : __byte_perm(hi[0], hi[1], sel);
    const unsigned p23 = (b < 4) ? __byte_perm(lo[2], lo[3], sel) : __byte_perm(hi[2], hi[3], sel);
    *reinterpret_cast<unsigned*>(dst) = __byte_perm(p01, p23, 0x5410);
    dst += plane_stride;
  }
}
// two values -> one packed 16-bit store per plane (same digits as oz_emit_digits4)
__device__ __forceinline__ void oz_emit_digits2(const long long (&I)[2], int ns, int8_t* __restrict__ dst, size_t plane_stride) {
  const unsigned long long bias = 0x0080808080808080ull >> (8 * (OZ_MAXS - ns));
  const int sh = 8 * (OZ_MAXS - ns);
  unsigned lo[2], hi[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const unsigned long long u = (((unsigned long long)I[t] + bias) << sh) ^ 0x0080808080808080ull;
    lo[t] = (unsigned)u;
    hi[t] = (unsigned)(u >> 32);
  }
#pragma unroll
  for (int i = 0; i < OZ_MAXS; ++i) {
    if (i < ns) {
      const int b = 6 - i;
      const unsigned sel = (unsigned)((b & 3) | (((b & 3) + 4) << 4));
      const unsigned p01 = (b < 4) ? __byte_perm(lo[0], lo[1], sel) : __byte_perm(hi[0], hi[1], sel);
      *reinterpret_cast<unsigned short*>(dst + i * plane_stride) = (unsigned short)(p01 & 0xffffu);
    }
  }
}
__host__ __device__ __forceinline__ int oz_exponent_for(double mx) {
  int e = 0;
  if (mx > 0.0) frexp(mx / 0.98, &e);                        // mx/0.98 = f 2^e, f in [0.5,1)  =>  |x| 2^-e <= 0.98
  return e;
}

#define OZ_SLICE_THREADS 256
#define OZ_SLICE_PER_THREAD 4

__global__ void __launch_bounds__(OZ_SLICE_THREADS)
ozaki_slice_rows_kernel(const double* __restrict__ X, int ldx, int R, int Rpad, int K, int ns, double fixed_max,
                        int8_t* __restrict__ planes, double* __restrict__ scales) {
  __shared__ double red[OZ_SLICE_THREADS / 32];
  __shared__ double s_scale_inv;
  const int r = blockIdx.x;
  const int tid = threadIdx.x;
  const size_t plane_stride = (size_t)Rpad * K;
  if (r >= R) {                                            // pad rows: zero digits, unit scale
    for (int i = 0; i < ns; ++i)
      for (int k = tid * 4; k < K; k += OZ_SLICE_THREADS * 4)
        *reinterpret_cast<int*>(planes + i * plane_stride + (size_t)r * K + k) = 0;
    if (tid == 0) scales[r] = 1.0;
    return;
  }
  const double* x = X + (size_t)r * ldx;
  double mx = fixed_max;
  if (!(fixed_max > 0.0)) {
    mx = 0.0;
    for (int k = tid; k < K; k += OZ_SLICE_THREADS) mx = fmax(mx, fabs(x[k]));
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < OZ_SLICE_THREADS / 32; ++w) mx = fmax(mx, red[w]);
  }
  if (tid == 0) {
    int e = 0;
    if (mx > 0.0) {
      frexp(mx / 0.98, &e);                               // mx/0.98 = f 2^e, f in [0.5,1)  =>  |x| 2^-e <= 0.98
    }
    scales[r] = ldexp(1.0, e);
    s_scale_inv = ldexp(1.0, -e + 8 * ns - 1);             // x -> fixed point with 8 ns - 1 fractional bits
  }
  __syncthreads();
  const double sc = s_scale_inv;
  for (int k = tid * 4; k < K; k += OZ_SLICE_THREADS * 4) {
    long long I[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) I[t] = __double2ll_rn(x[k + t] * sc);
    oz_emit_digits4(I, ns, planes + (size_t)r * K + k, plane_stride);
  }
}

// ---------------------------------------------------------------------------------------------
// Fused producers of digit planes (INT8 mode): the FP64 operand of a contraction is never written to HBM.
// Both emit exactly the digits ozaki_slice_rows_kernel would produce from the FP64 matrix.
// ---------------------------------------------------------------------------------------------
// Cross-covariance rows K10[j][:] = a^2 kappa(|xs_j - x0s_k|^2) straight to digit planes (cross_fwd_kernel +
// ozaki_slice_rows_kernel with the fixed bound a^2).  One thread owns 4 consecutive training points (their
// scaled coordinates stay in registers), OZC_TJ candidate rows per CTA from shared memory; each store is one
// packed int per plane, 128 contiguous bytes per warp.
#ifndef OZC_TJ
#define OZC_TJ 32
#endif
#define OZC_THREADS 256
#ifndef OZC_MINB
#define OZC_MINB 2
#endif
#ifndef OZC_UNROLL
#define OZC_UNROLL 2
#endif
#ifndef OZC_UNROLL
#define OZC_UNROLL 2
#endif
// NS > 0: compile-time digit count (the FP64-grade configurations); NS == 0: run-time ns.
// MODE 0: digit planes only.
// MODE 1: also the posterior mean the exact FP64 way, mu_j = m + sum_k K10[j][k] alpha_k with alpha = K00^-1 (y0 - m)
//         (Means = mean_fn + Beta . resid, gaussian_process.py:204-208): every kernel value is in a register here anyway.
//         The partial sum of this CTA's 1024 columns goes to mu_part[blockIdx.x][j].  Every thread parks its 4-column
//         partial in shared memory (no dependent shuffle chain in the element loop: the chain cost 2.1 ms per step at the
//         headline size); after every OZC_RG rows warp w folds row w of the group.  Fixed summation order (thread partials
//         strided by lane, shuffle tree, column blocks in index order): results do not depend on scheduling.
// MODE 3: MODE 1 that also stores the kernel derivative of every pair for the reverse covariance kernel (gout).
// MODE 2: MODE 1 on the compact rows of the level switch's second pass: row r = pos (q - p) + i holds candidate row
//         sel[pos] q + p + i of X for pos < *sel_count; J is ignored.
// p > 0 (pending points, shared by every restart and contracted once elsewhere): only the q - p trainable rows of
//         each pool are produced; row j = s (q - p) + i holds candidate row s q + p + i of X.
// MODE >= 1 also CENTRES the operand: cross-covariances lie in (0, a^2], so the planes hold K10 - a^2/2 in (-a^2/2, a^2/2]
//         and the fixed row scale halves -- one more bit of the fixed-point range, i.e. half the truncation error of every
//         level, for free.  The contraction's epilogue adds the rank-one term back, (a^2/2) (L0^-1 1)_r: row_add[j] = a^2/2
//         (written here, 0 on pad rows), col_add = ctx->linv_rowsum.  L0^-1 1 is the whitened constant function: O(1)
//         entries, so the add-back rounds at the FP64 floor.
template <int D, int KID, int NS, int MODE = 0>
__global__ void __launch_bounds__(OZC_THREADS, (D > 0 && D <= 8) ? OZC_MINB : 1)
cross_fwd_planes_kernel(const double* __restrict__ X, int J, int Jpad, const double* __restrict__ X0s, int n, int npad,
                        ModelParams mp, int ns_rt, int8_t* __restrict__ planes, double* __restrict__ scales,
                        const double* __restrict__ alpha = nullptr, double* __restrict__ mu_part = nullptr,
                        const int* __restrict__ sel = nullptr, const int* __restrict__ sel_count = nullptr, int q = 1,
                        int p = 0, double* __restrict__ row_add = nullptr, double* __restrict__ gout = nullptr) {
  // (row_add == nullptr: not centred)
  // MODE 3: MODE 1 + gout, d kappa / d r2 of every (row, training point) pair, [Jpad][npad] FP64, written once and read once
  // by cross_bwd_g_kernel: the reverse kernel then neither recomputes r2 nor takes a square root or an exponential.
  constexpr int DD = D > 0 ? D : MAF_MAX_D;
  constexpr bool ALPHA = MODE >= 1, SEL = MODE == 2, GST = MODE == 3;
  const int qn = q - p;
  const int d = D > 0 ? D : mp.d;
  const int ns = NS > 0 ? NS : ns_rt;
  __shared__ double xs[OZC_TJ][DD];
  __shared__ double etab_sh[MAF_ETAB_WORDS];
  constexpr int OZC_RG = OZC_THREADS / 32;                   // rows per reduction group = warps per CTA
  static_assert(OZC_TJ % OZC_RG == 0, "row groups");
  __shared__ double psum[ALPHA ? OZC_RG : 1][ALPHA ? OZC_THREADS : 1];
  __shared__ double msum[ALPHA ? OZC_TJ : 1];
  maf_fill_etab(etab_sh, threadIdx.x, blockDim.x);
  const double* etab = etab_sh + MAF_ETAB_LANE(threadIdx.x);
  const int k = (blockIdx.x * OZC_THREADS + threadIdx.x) * 4;
  const int j0 = blockIdx.y * OZC_TJ;
  if (SEL) {
    J = __ldg(sel_count) * qn;
    if (j0 >= J) return;                                     // nothing selected in this row block (rows stay stale, unread)
  }
  for (int t = threadIdx.x; t < OZC_TJ * d; t += OZC_THREADS) {
    int jj = t / d, c = t - jj * d;
    int j = j0 + jj;
    if ((SEL || p > 0) && j < J) {
      const int pos = j / qn;
      j = (SEL ? __ldg(sel + pos) : pos) * q + p + (j - pos * qn);
    }
    xs[jj][c] = (j0 + jj < J) ? X[(size_t)j * d + c] * mp.inv_ls[c] : 0.0;
  }
  const bool centred = ALPHA && row_add != nullptr;
  const double centre = centred ? 0.5 * mp.amp2 : 0.0;
  const int e = oz_exponent_for(centred ? centre : mp.amp2);
  if (blockIdx.x == 0 && threadIdx.x < OZC_TJ) {
    scales[j0 + threadIdx.x] = (j0 + (int)threadIdx.x < J) ? ldexp(1.0, e) : 1.0;
    if (centred) row_add[j0 + threadIdx.x] = (j0 + (int)threadIdx.x < J) ? centre : 0.0;
  }
  __syncthreads();
  const int jend = min(OZC_TJ, J - j0);                      // rows >= J of the last row block: zero digits
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool active = k < npad;                              // whole warps (npad is a multiple of 128)
  if (!ALPHA && !active) return;
  {
    // a^2 2^(8 ns - 1 - e) per column, 0 on the pad columns k >= n (their digits must be zero): no per-element predicate
    const double sc = mp.amp2 * ldexp(1.0, -e + 8 * ns - 1);
    const double coff = centre * ldexp(1.0, -e + 8 * ns - 1);
    double x0[4][DD], cs[4], co[ALPHA ? 4 : 1], al[ALPHA ? 4 : 1];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      cs[t] = (active && (k + t) < n) ? sc : 0.0;
      if (ALPHA) co[t] = (active && (k + t) < n) ? coff : 0.0;   // pad columns: zero digits
      if (ALPHA) al[t] = active ? alpha[k + t] : 0.0;
#pragma unroll
      for (int c = 0; c < DD; ++c) x0[t][c] = (c < d && active) ? X0s[(size_t)c * npad + k + t] : 0.0;
    }
    const size_t plane_stride = (size_t)Jpad * npad;
    constexpr int UNRJ = OZC_UNROLL;
    const int jloop = ALPHA ? OZC_TJ : jend;                 // ALPHA: whole groups (barriers inside); rows >= jend add 0
#pragma unroll UNRJ
    for (int jj = 0; jj < jloop; ++jj) {
      long long I[4];
      double pm = 0.0;
      double gv[GST ? 4 : 1];
      if (!ALPHA || (active && jj < jend)) {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        double r2 = 0.0;
#pragma unroll
        for (int c = 0; c < DD; ++c) {
          if (c < d) {
            double df = xs[jj][c] - x0[t][c];
            r2 += df * df;
          }
        }
        double v;
        if constexpr (GST) v = kern_val_dr2_fast<KID>(r2, etab, gv[t]) * cs[t];
        else v = kern_val_fast<KID>(r2, etab) * cs[t];
        if (ALPHA) pm = fma(v, al[t], pm);
        I[t] = __double2ll_rn(ALPHA ? v - co[t] : v);
      }
      if constexpr (GST) {
        double* gd = gout + (size_t)(j0 + jj) * npad + k;
        __stcs(reinterpret_cast<double2*>(gd), make_double2(gv[0], gv[1]));
        __stcs(reinterpret_cast<double2*>(gd + 2), make_double2(gv[2], gv[3]));
      }
      int8_t* dst = planes + (size_t)(j0 + jj) * npad + k;
      if constexpr (NS > 0) oz_emit_digits4_t<(NS > 0 ? NS : 1)>(I, dst, plane_stride);
      else oz_emit_digits4(I, ns, dst, plane_stride);
      }
      if (ALPHA) {
        psum[jj % OZC_RG][threadIdx.x] = pm;
        if (jj % OZC_RG == OZC_RG - 1) {                     // group complete: warp w folds the group's row w
          __syncthreads();
          double a = 0.0;
#pragma unroll
          for (int u = 0; u < OZC_THREADS / 32; ++u) a += psum[warp][lane + 32 * u];
          a = warp_sum(a);
          if (lane == 0) msum[jj - (OZC_RG - 1) + warp] = a;
          __syncthreads();
        }
      }
    }
    if (active)
      for (int jj = (jend > 0 ? jend : 0); jj < OZC_TJ; ++jj)
        for (int i = 0; i < ns; ++i) *reinterpret_cast<unsigned*>(planes + i * plane_stride + (size_t)(j0 + jj) * npad + k) = 0u;
  }
  if (ALPHA) {
    __syncthreads();
    if (threadIdx.x < OZC_TJ)                                // undo the fixed-point factor (exact)
      mu_part[(size_t)blockIdx.x * Jpad + j0 + threadIdx.x] = ((int)threadIdx.x < jend ? msum[threadIdx.x] : 0.0) * ldexp(1.0, e - 8 * ns + 1);
  }
}

// Vbar^T[i][k] = mubar_i gamma_k - sum_j (Sbar_ij + Sbar_ji) V^T[j][k]  (vbar_kernel) straight to digit planes.
// One CTA per restart.  Sweep 1 bounds the row maxima of Vbar^T from the row maxima of V^T, sweep 2 (the q x n slab
// of V^T re-read from L2) computes Vbar^T two columns per thread and emits its digits.  Rows >= J of the last
// row block are zeroed by the CTAs with s >= S.
#ifndef VBP_THREADS
#define VBP_THREADS 128
#endif
#ifndef VBP_MINB
#define VBP_MINB 4
#endif
template <int Q, bool DED = false>
__global__ void __launch_bounds__(VBP_THREADS, Q <= 8 ? VBP_MINB : 1)
vbar_planes_kernel(const double* __restrict__ Vt, int npad, int S, int Jpad, const double* __restrict__ gamma,
                   const double* __restrict__ mubar, const double* __restrict__ Sigbar, int ns,
                   int8_t* __restrict__ planes, double* __restrict__ scales, const double* __restrict__ vmax = nullptr,
                   int tight = 0, const int* __restrict__ vslot = nullptr, const double* __restrict__ Vt2 = nullptr,
                   const double* __restrict__ vmax2 = nullptr, LevelSwitch ls = LevelSwitch{},
                   const double* __restrict__ mubar_src = nullptr, double* __restrict__ row_add = nullptr,
                   int p_rt = 0, const double* __restrict__ Vp = nullptr, const double* __restrict__ vmaxp = nullptr) {
  const int p = DED ? p_rt : 0;                             // DED = false: no shared block, the row selects fold away
  // p > 0: the first p rows of every pool are pending points; their V^T rows are shared (Vp[p][npad], row maxima vmaxp)
  // and only the q - p trainable rows exist in Vt / Vt2 and are written here (rows rc (Q - p) + i - p).
  // row_add != nullptr: row_add[row] = mubar_src[s][i] for the rows this CTA writes (the reverse contraction's epilogue adds
  // row_add[j] alpha[r]: the mean part of d loss / d K10); 0 on the pad rows.
  // vslot != nullptr: restarts whose forward contraction was redone with all planes (level switch) keep their V^T rows
  // at compact position vslot[s] of Vt2 (row maxima in vmax2).  ls.sel != nullptr: this is the second pass of the REVERSE
  // switch; CTA r serves restart ls.sel[r] and writes compact rows r q .. r q + q - 1.
  // mubar == nullptr: the mean part of the gradient does not pass through the reverse contraction (cross_bwd_kernel
  // adds mubar_j alpha_k); Vbar^T = -(Sbar + Sbar^T) V^T only.
  // tight != 0: the row scales come from the TRUE row maxima of Vbar^T (one more sweep over the slab, which the L2
  // mostly still holds) instead of the bound below: up to log2(q + 1) bits more of the fixed-point range are used.
  __shared__ double W[Q][Q];
  __shared__ double mb[Q];
  __shared__ double red[VBP_THREADS / 32][Q];
  __shared__ double redg[VBP_THREADS / 32];
  __shared__ double sinv[Q];
  const int tid = threadIdx.x;
  const int rc = blockIdx.x;                                // row group written (compact position in the second pass)
  int s = rc;                                               // restart served
  const int qn = Q - p;                                     // rows per restart in Vt / the planes
  const size_t plane_stride = (size_t)Jpad * npad;
  if (ls.sel != nullptr) {
    if (rc >= __ldg(ls.sel_count)) return;                  // (pad rows of the compact block stay stale: never read back)
    s = __ldg(ls.sel + rc);
  } else if (s >= S) {                                      // pad rows [S qn, Jpad): zero digits, unit scale
    const int r = S * qn + (s - S);
    if (r >= Jpad) return;
    for (int i = 0; i < ns; ++i)
      for (int k = tid * 4; k < npad; k += VBP_THREADS * 4) *reinterpret_cast<int*>(planes + i * plane_stride + (size_t)r * npad + k) = 0;
    if (tid == 0) {
      scales[r] = 1.0;
      if (row_add != nullptr) row_add[r] = 0.0;
    }
    return;
  }
  if (row_add != nullptr && tid >= p && tid < Q) row_add[(size_t)rc * qn + tid - p] = mubar_src[(size_t)s * Q + tid];
  const double* sb = Sigbar + (size_t)s * Q * Q;
  for (int t = tid; t < Q * Q; t += VBP_THREADS) {
    int i = t / Q, j = t - i * Q;
    W[i][j] = sb[i * Q + j] + sb[j * Q + i];
  }
  if (tid < Q) mb[tid] = mubar != nullptr ? mubar[(size_t)s * Q + tid] : 0.0;
  __syncthreads();
  const int vs = vslot != nullptr ? __ldg(vslot + s) : -1;
  // row j of this restart: pending rows from the shared block, trainable rows from Vt (or Vt2 after a forward redo)
  const double* Vn = (vs >= 0 ? Vt2 + (size_t)vs * qn * npad : Vt + (size_t)s * qn * npad) - (size_t)p * npad;
  const double* vmn = vmax == nullptr ? nullptr : (vs >= 0 ? vmax2 + (size_t)vs * qn : vmax + (size_t)s * qn) - p;
#define VBP_ROW(j) ((DED && (j) < p) ? Vp + (size_t)(j) * npad : Vn + (size_t)(j) * npad)
#define VBP_VMAX(j) ((DED && (j) < p) ? vmaxp[j] : vmn[j])
  double mx[Q], gx = 0.0;
#pragma unroll
  for (int i = 0; i < Q; ++i) mx[i] = 0.0;
  if (tight) {
    // Sweep 1 (tight): the values themselves, |.| maxima per row.  The row loop is rolled like sweep 2 (register
    // budget), so the running maxima are selected with a compile-time unrolled compare instead of a dynamic index.
    for (int k = tid * 2; k < npad; k += VBP_THREADS * 2) {
      double2 v[Q];
#pragma unroll
      for (int i = 0; i < Q; ++i) v[i] = *reinterpret_cast<const double2*>(VBP_ROW(i) + k);
      double2 g = make_double2(0.0, 0.0);
      if (mubar != nullptr) g = *reinterpret_cast<const double2*>(gamma + k);
#pragma unroll 1
      for (int i = 0; i < Q; ++i) {
        double o0 = mb[i] * g.x, o1 = mb[i] * g.y;
#pragma unroll
        for (int j = 0; j < Q; ++j) {
          o0 -= W[i][j] * v[j].x;
          o1 -= W[i][j] * v[j].y;
        }
        const double a = fmax(fabs(o0), fabs(o1));
#pragma unroll
        for (int u = 0; u < Q; ++u) mx[u] = (u == i) ? fmax(mx[u], a) : mx[u];
      }
    }
  } else if (vmax) {
    // Sweep 1 bounds the row maxima of Vbar^T from the row maxima of V^T and of gamma only (no q x q work):
    // |Vbar_ik| <= |mubar_i| max|gamma| + sum_j |W_ij| max_k |V_jk|, at most a factor q + 1 above the true maximum.
    // (vmax: the forward contraction's epilogue has already left the row maxima of V^T there.)
    if (mubar != nullptr)
      for (int k = tid; k < npad; k += VBP_THREADS) gx = fmax(gx, fabs(gamma[k]));
  } else {
    for (int k = tid; k < npad; k += VBP_THREADS) {
#pragma unroll
      for (int i = 0; i < Q; ++i) mx[i] = fmax(mx[i], fabs(VBP_ROW(i)[k]));
      if (mubar != nullptr) gx = fmax(gx, fabs(gamma[k]));
    }
  }
#pragma unroll
  for (int i = 0; i < Q; ++i) {
    double m = mx[i];
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0) red[tid >> 5][i] = m;
  }
  for (int o = 16; o > 0; o >>= 1) gx = fmax(gx, __shfl_xor_sync(0xffffffffu, gx, o));
  if ((tid & 31) == 0) redg[tid >> 5] = gx;
  __syncthreads();
  if (tid < Q) {
    double bound;
    if (tight) {
      bound = red[0][tid];
      for (int w = 1; w < VBP_THREADS / 32; ++w) bound = fmax(bound, red[w][tid]);
      bound *= 1.0 + 1e-12;                                 // sweep 2 recomputes the values: a last-bit difference stays inside
    } else {
      double gm = redg[0];
      for (int w = 1; w < VBP_THREADS / 32; ++w) gm = fmax(gm, redg[w]);
      bound = fabs(mb[tid]) * gm;
      for (int j = 0; j < Q; ++j) {
        double vm = red[0][j];
        for (int w = 1; w < VBP_THREADS / 32; ++w) vm = fmax(vm, red[w][j]);
        if (vmax) vm = VBP_VMAX(j);
        bound += fabs(W[tid][j]) * vm;
      }
    }
    const int e = oz_exponent_for(bound);
    // an all-zero row (no gradient reaches it) gets scale 0: its product is exactly 0 and carries no truncation error
    if (tid >= p) scales[(size_t)rc * qn + tid - p] = bound > 0.0 ? ldexp(1.0, e) : 0.0;
    sinv[tid] = ldexp(1.0, -e + 8 * ns - 1);
  }
  __syncthreads();
  // two columns per thread (q values of V^T each stay in registers); sweep 2 walks the slab backwards, so the columns
  // sweep 1 read last (the ones most likely still in L2) are re-read first
  for (int k = ((npad - 1) / (2 * VBP_THREADS)) * (2 * VBP_THREADS) + tid * 2; k >= 0; k -= VBP_THREADS * 2) {
    if (k >= npad) continue;
    double2 v[Q];
#pragma unroll
    for (int i = 0; i < Q; ++i) v[i] = *reinterpret_cast<const double2*>(VBP_ROW(i) + k);
    double2 g = make_double2(0.0, 0.0);
    if (mubar != nullptr) g = *reinterpret_cast<const double2*>(gamma + k);
#pragma unroll 1
    for (int i = p; i < Q; ++i) {                             // trainable rows only: pending rows have no gradient
      double o0 = mb[i] * g.x, o1 = mb[i] * g.y;
#pragma unroll
      for (int j = 0; j < Q; ++j) {
        o0 -= W[i][j] * v[j].x;
        o1 -= W[i][j] * v[j].y;
      }
      const long long I[2] = {__double2ll_rn(o0 * sinv[i]), __double2ll_rn(o1 * sinv[i])};
      oz_emit_digits2(I, ns, planes + ((size_t)rc * qn + i - p) * npad + k, plane_stride);
    }
  }
#undef VBP_ROW
#undef VBP_VMAX
}
