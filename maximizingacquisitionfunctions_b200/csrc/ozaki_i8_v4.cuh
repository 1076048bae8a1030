// INT8 digit-sliced contraction on CTA pairs (v4): tcgen05.mma.cta_group::2, M = 256 over two SMs.
//
// Same statically scheduled pair walk as v3 (ozaki_i8_v3.cuh); what changes is who holds what:
//   * a cluster of two CTAs (one TPC) owns a 256 x 128 output tile; CTA r keeps rows [128 r, 128 r + 128) of every
//     level accumulator in ITS tensor memory and loads ITS 128 rows of each A plane tile (16 KB);
//   * the T plane tile (128 rows of L0^-1) is split: each CTA loads 64 rows (8 KB) and the MMA reads both halves,
//     so a CTA ingests A + T/2 instead of A + T per product group (-25 % L2->SM traffic, half the B-operand
//     shared-memory reads) and the freed shared memory double-buffers every T tile by k-chunk parity;
//   * only the leader CTA (cluster rank 0) issues MMAs; its tcgen05.commit multicasts the "slot free" and
//     "accumulators complete" arrivals to both CTAs; every TMA load of either CTA completes on the LEADER's full
//     barrier (address with the peer bit cleared), which the leader arms with the byte count of both halves;
//   * both epilogues drain their own tensor memory and arrive on the leader's acc_empty barrier.
#pragma once
#include "ozaki_i8_v3.cuh"

#define OZ4_A_BYTES (128 * 128)
#define OZ4_T_BYTES (64 * 128)
#define OZ4_THREADS 576

template <int NS, int LMAX>
struct Oz4 : Oz3<NS, LMAX> {
  using B = Oz3<NS, LMAX>;
  static constexpr int NA = NS, NT = 2 * NS, NBAR = NA + NT;
  static constexpr int SMEM_BYTES = NA * OZ4_A_BYTES + NT * OZ4_T_BYTES + 1024;
  static constexpr bool A_DBL0 = (B::NPASS > 1) && (2 * B::smax(0) <= NS);   // first pass: A tiles doubled too
  __host__ __device__ static constexpr int bufA(int pass, int i, int par) {
    return (pass == 0 && A_DBL0) ? par * B::smax(0) + (i - 1) : i - 1;
  }
  __host__ __device__ static constexpr int bufT(int ip, int par) { return par * NS + (ip - 1); }
  __host__ __device__ static constexpr int offA(int b) { return b * OZ4_A_BYTES; }
  __host__ __device__ static constexpr int offT(int b) { return NA * OZ4_A_BYTES + b * OZ4_T_BYTES; }
  __host__ __device__ static constexpr int barA(int b) { return b; }
  __host__ __device__ static constexpr int barT(int b) { return NA + b; }
};

__device__ __forceinline__ uint32_t oz4_cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void oz4_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void oz4_mma(uint32_t d_tmem, uint32_t alo, uint32_t blo, uint32_t idesc, uint32_t acc) {
  const uint64_t da = ((uint64_t)OZ3_DESC_HI << 32) | alo, db = ((uint64_t)OZ3_DESC_HI << 32) | blo;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
// arrival on the barrier at this shared-memory offset in BOTH CTAs of the pair once all prior MMAs have completed
__device__ __forceinline__ void oz4_commit2(uint32_t bar_addr) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar_addr),
               "h"((uint16_t)3)
               : "memory");
}
// TMA load of this CTA's part of a plane tile; the bytes are accounted on the leader's barrier
__device__ __forceinline__ void oz4_load(uint32_t dst, const CUtensorMap* tm, uint32_t leader_bar, int c0, int c1, int c2,
                                         uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], "
      "[%2], %6;" ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void oz4_expect(uint32_t bar_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(bytes) : "memory");
}
// wait with cluster-scope acquire: the arrivals on acc_empty come from the peer CTA as well (release.cluster)
__device__ __forceinline__ void oz4_wait_cluster(uint32_t bar_addr, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tOZ4_WAIT:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1, %2;\n\t"
      "@P1 bra OZ4_DONE;\n\tbra OZ4_WAIT;\n\tOZ4_DONE:\n\t}" ::"r"(bar_addr), "r"(parity), "r"(0x989680u)
      : "memory");
}
__device__ __forceinline__ void oz4_arrive_leader(uint32_t local_bar_addr) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(local_bar_addr)
      : "memory");
}

// tile walk of one CTA pair: balanced list (p.tile_list) or round-robin
struct Oz4Walk {
  int i, end, stride;
  const int* list;
  __device__ __forceinline__ Oz4Walk(const OzParams& p, int cluster_id, int nclusters, int ntiles) {
    if (p.tile_list) {
      list = p.tile_list;
      i = __ldg(p.tile_begin + cluster_id);
      end = __ldg(p.tile_begin + cluster_id + 1);
      stride = 1;
    } else {
      list = nullptr;
      i = cluster_id;
      end = ntiles;
      stride = nclusters;
    }
  }
  __device__ __forceinline__ bool next(int& t) {
    if (i >= end) return false;
    t = list ? __ldg(list + i) : i;
    i += stride;
    return true;
  }
};

// ---- one k-chunk of one pass: producer (both CTAs) ------------------------------------------------------
template <int NS, int LMAX, int PASS, int PAR>
__device__ __forceinline__ void oz4_produce_chunk(const CUtensorMap* tmA, const CUtensorMap* tmT, uint32_t smem0, uint32_t full0,
                                                  uint32_t full0_leader, uint32_t empty0, uint32_t& ph_empty, int c, int jrow,
                                                  int trow, bool leader, int dbg, uint64_t hintA, uint64_t hintT) {
  using Z = Oz4<NS, LMAX>;
#pragma unroll
  for (int i = 1; i <= Z::smax(PASS); ++i) {
#pragma unroll
    for (int ip = Z::whi(PASS, i); ip >= Z::wlo(PASS, i); --ip) {
      if (i == 1 || ip < Z::wlo(PASS, i - 1)) {
        const int b = Z::bufT(ip, PAR), s = Z::barT(b);
        oz3_wait(empty0 + 8u * s, (ph_empty >> s) & 1u);
        ph_empty ^= 1u << s;
        if (oz_elect_one()) {
#ifdef MAF_OZ_DIAG
          if (dbg & 16) {                                      // diagnostics: no T loads (operands stale)
            if (leader) oz3_arrive(full0 + 8u * s);
          } else
#endif
          {
            if (leader) oz4_expect(full0 + 8u * s, 2u * OZ4_T_BYTES);
            oz4_load(smem0 + (uint32_t)Z::offT(b), tmT, full0_leader + 8u * s, c, trow, ip - 1, hintT);
          }
        }
        __syncwarp();
      }
    }
    {
      const int b = Z::bufA(PASS, i, PAR), s = Z::barA(b);
      oz3_wait(empty0 + 8u * s, (ph_empty >> s) & 1u);
      ph_empty ^= 1u << s;
      if (oz_elect_one()) {
#ifdef MAF_OZ_DIAG
        if (dbg & 8) {                                         // diagnostics: no A loads
          if (leader) oz3_arrive(full0 + 8u * s);
        } else
#endif
        {
          if (leader) oz4_expect(full0 + 8u * s, 2u * OZ4_A_BYTES);
          oz4_load(smem0 + (uint32_t)Z::offA(b), tmA, full0_leader + 8u * s, c, jrow, i - 1, hintA);
        }
      }
      __syncwarp();
    }
  }
}

// ---- one k-chunk of one pass: MMA issuer (leader CTA) ------------------------------------------------------
#ifdef MAF_OZ_DIAG
#define OZ4_T0() const long long t0_ = clock64()
#define OZ4_ACC(x) (x) += clock64() - t0_
#else
#define OZ4_T0()
#define OZ4_ACC(x)
#endif

template <int NS, int LMAX, int PASS, int PAR>
__device__ __forceinline__ void oz4_issue_chunk(uint32_t desc0, uint32_t full0, uint32_t empty0, uint32_t& ph_full, uint32_t tbase,
                                                uint32_t idesc, uint32_t not_first, long long& t_full) {   // t_full: waiting for operand tiles
  using Z = Oz4<NS, LMAX>;
#pragma unroll
  for (int i = 1; i <= Z::smax(PASS); ++i) {
    OZ4_T0();
#pragma unroll
    for (int ip = Z::whi(PASS, i); ip >= Z::wlo(PASS, i); --ip) {
      if (i == 1 || ip < Z::wlo(PASS, i - 1)) {
        const int s = Z::barT(Z::bufT(ip, PAR));
        oz3_wait(full0 + 8u * s, (ph_full >> s) & 1u);
        ph_full ^= 1u << s;
      }
    }
    {
      const int s = Z::barA(Z::bufA(PASS, i, PAR));
      oz3_wait(full0 + 8u * s, (ph_full >> s) & 1u);
      ph_full ^= 1u << s;
    }
    OZ4_ACC(t_full);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (oz_elect_one()) {
      const uint32_t alo = desc0 + (uint32_t)(Z::offA(Z::bufA(PASS, i, PAR)) >> 4);
#pragma unroll
      for (int ip = Z::wlo(PASS, i); ip <= Z::whi(PASS, i); ++ip) {
        const int l = i + ip;
        const uint32_t d_tmem = tbase + (uint32_t)((l - Z::lo(PASS)) * OZ_BN);
        const uint32_t blo = desc0 + (uint32_t)(Z::offT(Z::bufT(ip, PAR)) >> 4);
        const bool level_start = (i == ((l - NS > 1) ? l - NS : 1));
#pragma unroll
        for (int k = 0; k < OZ3_BK / 32; ++k) oz4_mma(d_tmem, alo + 2u * k, blo + 2u * k, idesc, (level_start && k == 0) ? not_first : 1u);
      }
      oz4_commit2(empty0 + 8u * Z::barA(Z::bufA(PASS, i, PAR)));
      if (i < Z::smax(PASS)) {
        if (Z::hi(PASS) - i <= NS) oz4_commit2(empty0 + 8u * Z::barT(Z::bufT(Z::hi(PASS) - i, PAR)));
      } else {
#pragma unroll
        for (int ip = Z::wlo(PASS, i); ip <= Z::whi(PASS, i); ++ip) oz4_commit2(empty0 + 8u * Z::barT(Z::bufT(ip, PAR)));
      }
    }
    __syncwarp();
  }
}

template <int NS, int LMAX, int PASS>
__device__ __forceinline__ void oz4_produce_pass(const CUtensorMap* tmA, const CUtensorMap* tmT, uint32_t smem0, uint32_t full0,
                                                 uint32_t full0_leader, uint32_t empty0, uint32_t& ph_empty, int c_begin,
                                                 int kchunks, int jrow, int trow, bool leader, int dbg, uint64_t hintA, uint64_t hintT) {
  for (int kc = 0; kc < kchunks; kc += 2) {
    oz4_produce_chunk<NS, LMAX, PASS, 0>(tmA, tmT, smem0, full0, full0_leader, empty0, ph_empty, c_begin + kc * OZ3_BK, jrow, trow, leader, dbg, hintA, hintT);
    if (kc + 1 < kchunks)
      oz4_produce_chunk<NS, LMAX, PASS, 1>(tmA, tmT, smem0, full0, full0_leader, empty0, ph_empty, c_begin + (kc + 1) * OZ3_BK, jrow, trow, leader, dbg, hintA, hintT);
  }
}
template <int NS, int LMAX, int PASS>
__device__ __forceinline__ void oz4_issue_pass(uint32_t desc0, uint32_t full0, uint32_t empty0, uint32_t& ph_full, uint32_t tbase,
                                               uint32_t idesc, int kchunks, long long* t_full /*[first chunk, later chunks]*/) {
  for (int kc = 0; kc < kchunks; kc += 2) {
    oz4_issue_chunk<NS, LMAX, PASS, 0>(desc0, full0, empty0, ph_full, tbase, idesc, kc > 0 ? 1u : 0u, t_full[kc > 0]);
    if (kc + 1 < kchunks) oz4_issue_chunk<NS, LMAX, PASS, 1>(desc0, full0, empty0, ph_full, tbase, idesc, 1u, t_full[1]);
  }
}

// ---- epilogue with the pass-0 partial result held in registers (EPI = 1, variant 5, default) ----------------------
// Each epilogue lane owns the same 32 elements of the 128 x 128 tile in both passes (fragment layout of
// tcgen05.ld.16x256b: lane t holds rows t/4 and t/4 + 8, columns 2 (t%4) + {0,1} of every 16 x 8 block).  A pass is
// DRAINED into registers (tcgen05.ld + exact int64 Horner recombination of its <= 4 level accumulators) and tensor
// memory goes back to the issuer at once; FP64 conversion (in place), scaling and the global stores run after the
// hand-over, under the MMAs of the next pass / tile.  C is written exactly once and never read back:
//   sum_l 2^(2-8l) S_l = 2^(2 - 8 hi1) (H0 2^(8 nl1) + H1),   H_p = sum_{l in pass p} S_l 2^(8 (hi_p - l)),
// |H0| < 2^46 is exact in FP64, and the FMA rounds once (the EPI = 0 path rounds every pass: last-bit differences).
template <int NS, int LMAX>
__device__ __forceinline__ void oz4_epilogue_stash(const double* __restrict__ sA, const double* __restrict__ sT,
                                                   double* __restrict__ C, const OzParams& p, int nN, int nJ2, int cluster_id,
                                                   int nclusters, uint32_t rank, int warp, int lane, uint32_t tbase, uint32_t accf,
                                                   uint32_t acce) {
  using Z = Oz4<NS, LMAX>;
  static_assert(Z::NPASS <= 2, "register-stash epilogue: at most two passes");
  constexpr int EW = (OZ4_THREADS - 64) / 128;
  constexpr int CW = OZ_BN / EW;
  constexpr int NG = CW / 16;
  constexpr int NE = NG * 2 * 8;                             // elements per lane
  const int quarter = warp & 3;
  const int part = (warp - 2) >> 2;
  const int fr = lane >> 2, fc = (lane & 3) * 2;
  const int ntiles = nN * nJ2;
  uint32_t g = 0;
  long long H[NE];                                           // [gq][h][n][e][c]; int64 after pass 0, FP64 bits afterwards
  const int nrows = p.nrows_dev ? __ldg(p.nrows_dev) * p.nrows_mul : 0x7fffffff;
  Oz4Walk walk(p, cluster_id, nclusters, ntiles);
  for (int t; walk.next(t);) {
    const Oz3Tile x = oz3_tile(t, nN, nJ2, p);
    if (2 * x.j0 >= nrows) continue;
    const int j0 = 2 * x.j0 + 128 * (int)rank;
#pragma unroll
    for (int pass = 0; pass < Z::NPASS; ++pass) {
      const int nl = Z::hi(pass) - Z::lo(pass) + 1;
      oz3_wait(accf, g & 1u);
      ++g;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#ifdef MAF_OZ_DIAG
      const long long te0 = clock64();
#endif
#pragma unroll
      for (int gq = 0; gq < NG; ++gq) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int n = 0; n < 2; ++n) {                      // one load = 16 rows x 8 columns: 4 registers per level
            const uint32_t taddr = tbase + ((uint32_t)(quarter * 32 + h * 16) << 16) + (uint32_t)(part * CW + gq * 16 + n * 8);
            uint32_t r[4][4];
#pragma unroll
            for (int tt = 0; tt < 4; ++tt) {
              if (tt < nl) {
                asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(r[tt][0]), "=r"(r[tt][1]), "=r"(r[tt][2]), "=r"(r[tt][3])
                             : "r"(taddr + (uint32_t)(tt * OZ_BN)));
              }
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int u = 0; u < 4; ++u) {                    // u = 2 e + c
              long long acc = (long long)(int)r[0][u];
#pragma unroll
              for (int tt = 1; tt < 4; ++tt)
                if (tt < nl) acc = acc * 256 + (long long)(int)r[tt][u];
              const int idx = ((gq * 2 + h) * 2 + n) * 4 + u;
              if (pass == 0) {
                H[idx] = acc;
              } else {
                const double tot = fma(__longlong_as_double(H[idx]), (double)(1ull << (8 * nl)), oz3_ll2double(acc));
                H[idx] = __double_as_longlong(tot);
              }
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) oz4_arrive_leader(acce);                // tensor memory is free again
#ifdef MAF_OZ_DIAG
      if (p.stats && rank == 0 && warp == 2 && lane == 0) p.stats[8 * cluster_id + 3] += clock64() - te0;   // drain (one warp)
#endif
      if (pass + 1 < Z::NPASS) {
#pragma unroll
        for (int idx = 0; idx < NE; ++idx) H[idx] = __double_as_longlong(oz3_ll2double(H[idx]));
      } else {
        const double w = exp2((double)(2 - 8 * Z::hi(pass)));
        double rm[2][2] = {{0.0, 0.0}, {0.0, 0.0}};          // |C| maxima of this lane's four rows (p.rowmax)
#pragma unroll
        for (int gq = 0; gq < NG; ++gq) {
          double2 sc[2], ca[2];
#pragma unroll
          for (int n = 0; n < 2; ++n) {
            sc[n] = __ldg(reinterpret_cast<const double2*>(sT + x.r0 + part * CW + gq * 16 + fc + 8 * n));
            ca[n] = p.col_add ? __ldg(reinterpret_cast<const double2*>(p.col_add + x.r0 + part * CW + gq * 16 + fc + 8 * n)) : make_double2(0.0, 0.0);
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const size_t j = (size_t)(j0 + quarter * 32 + h * 16 + fr + 8 * e);
              const double sa = __ldg(sA + j) * w;
              const double ra = p.row_add ? __ldg(p.row_add + j) : 0.0;
              double* crow = C + j * p.ldc + x.r0 + part * CW + gq * 16 + fc;
#pragma unroll
              for (int n = 0; n < 2; ++n) {
                const int idx = ((gq * 2 + h) * 2 + n) * 4 + 2 * e;
                double v0, v1;
                if (Z::NPASS == 1) {
                  v0 = oz3_ll2double(H[idx]);
                  v1 = oz3_ll2double(H[idx + 1]);
                } else {
                  v0 = __longlong_as_double(H[idx]);
                  v1 = __longlong_as_double(H[idx + 1]);
                }
                // written once, read by the next kernel only after the whole 2 GB result: streaming store (evict-first)
                const double o0 = fma(ra, ca[n].x, v0 * sc[n].x * sa), o1 = fma(ra, ca[n].y, v1 * sc[n].y * sa);
                __stcs(reinterpret_cast<double2*>(crow + 8 * n), make_double2(o0, o1));
                if (p.rowmax) rm[h][e] = fmax(rm[h][e], fmax(fabs(o0), fabs(o1)));
              }
            }
          }
        }
        if (p.rowmax) {                                      // four lanes share a row: two shuffles, one atomic per row
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              double m = rm[h][e];
              m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 1));
              m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 2));
              if ((lane & 3) == 0)
                atomicMax(reinterpret_cast<unsigned long long*>(p.rowmax + (size_t)(j0 + quarter * 32 + h * 16 + fr + 8 * e)),
                          (unsigned long long)__double_as_longlong(m));
            }
          }
        }
      }
    }
  }
}

// nJ2 = row blocks of 256 (one per CTA pair); tile order as in v3 with 256-row blocks
template <int NS, int LMAX, int EPI = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(OZ4_THREADS, 1)
ozaki_i8_gemm_v4_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmT,
                        const double* __restrict__ sA, const double* __restrict__ sT, double* __restrict__ C, OzParams p,
                        int nN, int nJ2) {
  using Z = Oz4<NS, LMAX>;
  extern __shared__ uint8_t oz_raw[];
  __shared__ uint64_t full_bar[Z::NBAR], empty_bar[Z::NBAR], acc_full, acc_empty;
  __shared__ uint32_t tmem_base_sh;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = oz4_cta_rank();
  const bool leader = rank == 0;
  const uint32_t smem0 = (oz_smem_u32(oz_raw) + 1023u) & ~1023u;
  const int ntiles = nN * nJ2;
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < Z::NBAR; ++s) {
      oz_mbar_init(&full_bar[s], 1);
      oz_mbar_init(&empty_bar[s], 1);
    }
    oz_mbar_init(&acc_full, 1);
    oz_mbar_init(&acc_empty, 2 * (OZ4_THREADS - 64) / 32);    // every epilogue warp of both CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmT)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem_u32(&tmem_base_sh)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  oz4_cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base_sh;
  const uint32_t full0 = oz_smem_u32(&full_bar[0]), empty0 = oz_smem_u32(&empty_bar[0]);
  uint32_t full0_leader;                                       // shared::cluster address of the leader's full barriers
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(full0_leader) : "r"(full0));

  if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    uint32_t ph_empty = 0xffffffffu;
    const int nrows = p.nrows_dev ? __ldg(p.nrows_dev) * p.nrows_mul : 0x7fffffff;
    Oz4Walk walk(p, cluster_id, nclusters, ntiles);
    for (int t; walk.next(t);) {
      Oz3Tile x = oz3_tile(t, nN, nJ2, p);                     // j0 counts 128-row blocks: rescale to 256
      if (2 * x.j0 >= nrows) continue;
      const int jrow = 2 * x.j0 + 128 * (int)rank, trow = x.r0 + 64 * (int)rank;
      oz4_produce_pass<NS, LMAX, 0>(&tmA, &tmT, smem0, full0, full0_leader, empty0, ph_empty, x.c_begin, x.kchunks, jrow, trow, leader, p.dbg, p.hintA, p.hintT);
      if constexpr (Z::NPASS > 1)
        oz4_produce_pass<NS, LMAX, 1>(&tmA, &tmT, smem0, full0, full0_leader, empty0, ph_empty, x.c_begin, x.kchunks, jrow, trow, leader, p.dbg, p.hintA, p.hintT);
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (leader) {
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t desc0 = ((smem0 & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t accf = oz_smem_u32(&acc_full), acce = oz_smem_u32(&acc_empty);
      uint32_t ph_full = 0u, g = 0;
      long long t_full[4] = {0, 0, 0, 0}, t_acc = 0;
      (void)t_acc;
#ifdef MAF_OZ_DIAG
      const long long t_begin = clock64();
#endif
      const int nrows = p.nrows_dev ? __ldg(p.nrows_dev) * p.nrows_mul : 0x7fffffff;
      Oz4Walk walk(p, cluster_id, nclusters, ntiles);
      for (int t; walk.next(t);) {
        const Oz3Tile x = oz3_tile(t, nN, nJ2, p);
        if (2 * x.j0 >= nrows) continue;
        if (g > 0) {
          OZ4_T0();
          oz4_wait_cluster(acce, (g - 1) & 1u);
          OZ4_ACC(t_acc);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        oz4_issue_pass<NS, LMAX, 0>(desc0, full0, empty0, ph_full, tbase, idesc, x.kchunks, t_full);
        if (oz_elect_one()) oz4_commit2(accf);
        __syncwarp();
        ++g;
        if constexpr (Z::NPASS > 1) {
          {
            OZ4_T0();
            oz4_wait_cluster(acce, (g - 1) & 1u);
            OZ4_ACC(t_acc);
          }
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          oz4_issue_pass<NS, LMAX, 1>(desc0, full0, empty0, ph_full, tbase, idesc, x.kchunks, t_full + 2);
          if (oz_elect_one()) oz4_commit2(accf);
          __syncwarp();
          ++g;
        }
      }
#ifdef MAF_OZ_DIAG
      if (p.stats && lane == 0) {
        p.stats[8 * cluster_id + 0] = clock64() - t_begin;       // issuer: first wait to last commit
        p.stats[8 * cluster_id + 1] = t_full[0] + t_full[1] + t_full[2] + t_full[3];   // ... waiting for operand tiles
        p.stats[8 * cluster_id + 2] = t_acc;                     // ... waiting for the epilogues to drain TMEM
        for (int u = 0; u < 4; ++u) p.stats[8 * cluster_id + 4 + u] = t_full[u];        // pass 0 first/later chunks, pass 1 first/later
      }
#endif
    }
  } else if constexpr (EPI == 1) {
    oz4_epilogue_stash<NS, LMAX>(sA, sT, C, p, nN, nJ2, cluster_id, nclusters, rank, warp, lane, tbase, oz_smem_u32(&acc_full),
                                 oz_smem_u32(&acc_empty));
  } else {
    // ================= epilogue (both CTAs, own 128 rows): as v3 =================
    constexpr int EW = (OZ4_THREADS - 64) / 128;
    constexpr int CW = OZ_BN / EW;
    constexpr int NG = CW / 16;
    const int quarter = warp & 3;
    const int part = (warp - 2) >> 2;
    const int fr = lane >> 2, fc = (lane & 3) * 2;
    const uint32_t accf = oz_smem_u32(&acc_full), acce = oz_smem_u32(&acc_empty);
    uint32_t g = 0;
    const int nrows = p.nrows_dev ? __ldg(p.nrows_dev) * p.nrows_mul : 0x7fffffff;
    Oz4Walk walk(p, cluster_id, nclusters, ntiles);
    for (int t; walk.next(t);) {
      const Oz3Tile x = oz3_tile(t, nN, nJ2, p);
      if (2 * x.j0 >= nrows) continue;
      const int j0 = 2 * x.j0 + 128 * (int)rank;
      double sa[2][2];
      double2 sc[NG][2];
      double* crow[2][2];
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const size_t j = (size_t)(j0 + quarter * 32 + h * 16 + fr + 8 * e);
          sa[h][e] = __ldg(sA + j);
          crow[h][e] = C + j * p.ldc + x.r0 + part * CW + fc;
        }
#pragma unroll
      for (int gq = 0; gq < NG; ++gq)
#pragma unroll
        for (int n = 0; n < 2; ++n)
          sc[gq][n] = __ldg(reinterpret_cast<const double2*>(sT + x.r0 + part * CW + gq * 16 + fc + 8 * n));
#pragma unroll
      for (int pass = 0; pass < Z::NPASS; ++pass) {
        const int lo = Z::lo(pass), nl = Z::hi(pass) - Z::lo(pass) + 1;
        const double w = exp2((double)(2 - 8 * (lo + nl - 1)));
        oz3_wait(accf, g & 1u);
        ++g;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#ifdef MAF_OZ_DIAG
        const long long te0 = clock64();
#endif
#pragma unroll
        for (int gq = 0; gq < NG; ++gq) {
          double2 old[2][2][2];
          if (pass > 0) {
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int n = 0; n < 2; ++n) old[h][e][n] = *reinterpret_cast<const double2*>(crow[h][e] + gq * 16 + 8 * n);
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t taddr = tbase + ((uint32_t)(quarter * 32 + h * 16) << 16) + (uint32_t)(part * CW + gq * 16);
            uint32_t r[4][8];
#pragma unroll
            for (int tt = 0; tt < 4; ++tt) {
              if (tt < nl) {
                asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(r[tt][0]), "=r"(r[tt][1]), "=r"(r[tt][2]), "=r"(r[tt][3]), "=r"(r[tt][4]), "=r"(r[tt][5]),
                               "=r"(r[tt][6]), "=r"(r[tt][7])
                             : "r"(taddr + (uint32_t)(tt * OZ_BN)));
              }
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int n = 0; n < 2; ++n) {
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                double v[2];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                  long long acc = (long long)(int)r[0][4 * n + 2 * e + c];
#pragma unroll
                  for (int tt = 1; tt < 4; ++tt)
                    if (tt < nl) acc = acc * 256 + (long long)(int)r[tt][4 * n + 2 * e + c];
                  v[c] = oz3_ll2double(acc) * (c ? sc[gq][n].y : sc[gq][n].x) * (sa[h][e] * w);
                }
                if (pass > 0) {
                  v[0] += old[h][e][n].x;
                  v[1] += old[h][e][n].y;
                }
                *reinterpret_cast<double2*>(crow[h][e] + gq * 16 + 8 * n) = make_double2(v[0], v[1]);
              }
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
#ifdef MAF_OZ_DIAG
        if (p.stats && leader && warp == 2 && lane == 0) p.stats[8 * cluster_id + 3] += clock64() - te0;   // epilogue body (one warp)
#endif
        if (lane == 0) oz4_arrive_leader(acce);                // one arrival per epilogue warp, on the leader's barrier
      }
    }
  }
  // neither CTA may leave (or free its tensor memory) while the pair still reads its shared memory / signals its barriers
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  oz4_cluster_sync();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512) : "memory");
}
