// INT8 digit-sliced contraction, statically scheduled (v3).
//
// Measured on B200 (gpurun_out/gemm_i8_dbg.jsonl, microbench_tcgen05_v2.jsonl): a tcgen05.mma of shape
// 128x128x32 (kind::i8) occupies the tensor pipe for 64 clocks, and ONE thread issues every MMA of the CTA, so
// the issue loop may spend ~10 instructions per MMA.  The generic kernel (ozaki_i8.cuh) spends ~25 (descriptor
// construction, run-time loop bounds, divergent-code ELECT loops) and runs at 145 clocks per MMA with loads and
// epilogue switched off.  This version removes all of that:
//
//   * <NS, LMAX> are template parameters; the (A plane, T plane) pair schedule of a k-chunk is unrolled at
//     compile time, every plane tile has a FIXED shared-memory slot (16 KB, [128 rows][128 B], SWIZZLE_128B)
//     with its own full/empty mbarrier, so a descriptor is `base + constant` and a phase is one bit of a mask;
//   * producer and issuer warps stay converged (barrier waits by all lanes, elect.sync for the asynchronous
//     instructions), so UTCIMMA / UTMALDG are issued straight from uniform registers;
//   * within a k-chunk the pairs are walked by A plane i = 1..smax; the T planes A_i meets form a window
//     [lo-i, hi-i] that slides down, each tile is released (tcgen05.commit) right after its last MMA and is
//     refilled for the next k-chunk while the remaining pairs run.  The short first pass (levels 2..4) is
//     double-buffered by chunk parity (its k-chunk is shorter than a TMA round trip);
//   * the epilogue recombines the <= 4 level accumulators exactly in int64 (one I2F per element instead of
//     four), on 16 warps instead of 4;
//   * the grid is persistent (one CTA per SM, static round-robin over a supertiled, heaviest-first tile list):
//     the producer prefetches the next pass / tile while the epilogue drains TMEM.
#pragma once
#include "ozaki_i8.cuh"

#define OZ3_BK 128
#define OZ3_TILE_BYTES (128 * OZ3_BK)
#define OZ3_THREADS 576                       // warp 0: TMA producer, warp 1: MMA issuer, warps 2-17: epilogue
#define OZ3_DESC_HI 0x40004040u               // SBO = 1024 B, descriptor version 1, SWIZZLE_128B

template <int NS, int LMAX>
struct Oz3 {
  // LMAX - 1 levels, 4 TMEM accumulators: the LAST pass takes 4 levels, the first one the remainder (FP64-grade:
  // levels 2-4 = 6 products on planes 1-3, then levels 5-8 = 22 products on planes 1-7).  A k-chunk of the short
  // first pass lasts only ~0.8 us, less than a TMA round trip, so ALL its tiles are double-buffered by chunk
  // parity (4 smax(0) <= NSLOT); when that does not fit (single-pass configurations) only the T tiles are.
  static constexpr int NLEV = LMAX - 1;
  static constexpr int NPASS = (NLEV + 3) / 4;
  // levels of the first pass.  OZ3_N0_6: override for the six-level configuration <6,7> (default 2: levels 2-3 | 4-7;
  // 3: levels 2-4 | 5-7), measured in profiles/r02_i8gemm_pass_split_ns6.txt
#ifndef OZ3_N0_6
#define OZ3_N0_6 2
#endif
  static constexpr int N0 = (NLEV == 6) ? OZ3_N0_6 : NLEV - 4 * (NPASS - 1);
  static constexpr int NSLOT = (2 * NS > 12) ? 2 * NS : 12;
  static constexpr int SMEM_BYTES = NSLOT * OZ3_TILE_BYTES + 1024;
  __host__ __device__ static constexpr int lo(int pass) { return pass == 0 ? 2 : 2 + N0 + 4 * (pass - 1); }
  __host__ __device__ static constexpr int hi(int pass) { return pass == 0 ? 2 + N0 - 1 : (lo(pass) + 3 < LMAX ? lo(pass) + 3 : LMAX); }
  __host__ __device__ static constexpr int smax(int pass) { return (NS < hi(pass) - 1) ? NS : hi(pass) - 1; }
  __host__ __device__ static constexpr int wlo(int pass, int i) { return (lo(pass) - i > 1) ? lo(pass) - i : 1; }
  __host__ __device__ static constexpr int whi(int pass, int i) { return (hi(pass) - i < NS) ? hi(pass) - i : NS; }
  static constexpr bool DBL_ALL = (4 * smax(0) <= NSLOT);     // first pass: every tile has two buffers
  // second buffer of T plane ip in pass 0 when only the T tiles are doubled: taken from the slots pass 0 leaves idle
  __host__ __device__ static constexpr int alt(int ip) {
    const int s0 = smax(0), nfree = NS - s0;
    int idx = ip - 1;
    if (idx < nfree) return s0 + idx;
    idx -= nfree;
    if (idx < nfree) return NS + s0 + idx;
    idx -= nfree;
    return 2 * NS + idx;
  }
  __host__ __device__ static constexpr int slotA(int pass, int i, int par) {
    return (pass == 0 && DBL_ALL) ? par * 2 * smax(0) + (i - 1) : i - 1;
  }
  __host__ __device__ static constexpr int slotT(int pass, int ip, int par) {
    if (pass == 0 && DBL_ALL) return par * 2 * smax(0) + smax(0) + (ip - 1);
    return (pass == 0 && par == 1) ? alt(ip) : NS + ip - 1;
  }
};

__device__ __forceinline__ bool oz_elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void oz3_mma(uint32_t d_tmem, uint32_t alo, uint32_t blo, uint32_t idesc, uint32_t acc) {
  const uint64_t da = ((uint64_t)OZ3_DESC_HI << 32) | alo, db = ((uint64_t)OZ3_DESC_HI << 32) | blo;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void oz3_commit(uint32_t bar_addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
}
// try_wait with a suspend-time hint (as CUTLASS' ClusterBarrier::wait): a waiting thread sleeps in hardware until the
// phase completes instead of re-polling shared memory; 512 epilogue threads otherwise poll acc_full for a whole pass.
__device__ __forceinline__ void oz3_wait(uint32_t bar_addr, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tOZ3_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
      "@P1 bra OZ3_DONE;\n\tbra OZ3_WAIT;\n\tOZ3_DONE:\n\t}" ::"r"(bar_addr), "r"(parity), "r"(0x989680u)
      : "memory");
}
__device__ __forceinline__ void oz3_arrive(uint32_t bar_addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
// L2 eviction-priority policies for TMA loads (the encodings createpolicy.fractional.L2::evict_* produces)
#define OZ3_L2_NORMAL 0x1000000000000000ull
#define OZ3_L2_EVICT_FIRST 0x12F0000000000000ull
#define OZ3_L2_EVICT_LAST 0x14F0000000000000ull
__device__ __forceinline__ void oz3_load(uint32_t dst, const CUtensorMap* tm, uint32_t bar_addr, int c0, int c1, int c2,
                                         uint64_t hint) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(OZ3_TILE_BYTES) : "memory");
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_addr), "r"(c0), "r"(c1), "r"(c2), "l"(hint)
      : "memory");
}

// TMA prefetch of one plane tile into L2 (no shared memory, no barrier): issued OZ3_PF k-chunks ahead of the load so
// that first-touch (DRAM) latency is not exposed when a k-chunk of the first pass lasts only 1.3 us.
#define OZ3_PF 2
__device__ __forceinline__ void oz3_prefetch(const CUtensorMap* tm, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0),
               "r"(c1), "r"(c2)
               : "memory");
}

// ---- one k-chunk of one pass: producer side -----------------------------------------------------
template <int NS, int LMAX, int PASS, int PAR, int DBG>
__device__ __forceinline__ void oz3_produce_chunk(const CUtensorMap* tmA, const CUtensorMap* tmT, uint32_t smem0,
                                                  uint32_t full0, uint32_t empty0, uint32_t& ph_empty, int c, int j0,
                                                  int r0, uint64_t hintA, uint64_t hintT, int c_pf) {
  using Z = Oz3<NS, LMAX>;
  if (c_pf >= 0 && oz_elect_one()) {
#pragma unroll
    for (int i = 1; i <= Z::smax(PASS); ++i) {
      oz3_prefetch(tmA, c_pf, j0, i - 1);
      oz3_prefetch(tmT, c_pf, r0, i - 1);
    }
  }
  __syncwarp();
#pragma unroll
  for (int i = 1; i <= Z::smax(PASS); ++i) {
    // T planes entering the window at this step (all of it at i == 1), then A_i
#pragma unroll
    for (int ip = Z::whi(PASS, i); ip >= Z::wlo(PASS, i); --ip) {
      if (i == 1 || ip < Z::wlo(PASS, i - 1)) {
        const int s = Z::slotT(PASS, ip, PAR);
        oz3_wait(empty0 + 8u * s, (ph_empty >> s) & 1u);
        ph_empty ^= 1u << s;
        if (oz_elect_one()) {
          if (DBG & 2) oz3_arrive(full0 + 8u * s);
          else oz3_load(smem0 + (uint32_t)s * OZ3_TILE_BYTES, tmT, full0 + 8u * s, c, r0, ip - 1, hintT);
        }
        __syncwarp();
      }
    }
    {
      const int s = Z::slotA(PASS, i, PAR);
      oz3_wait(empty0 + 8u * s, (ph_empty >> s) & 1u);
      ph_empty ^= 1u << s;
      if (oz_elect_one()) {
        if (DBG & 2) oz3_arrive(full0 + 8u * s);
        else oz3_load(smem0 + (uint32_t)s * OZ3_TILE_BYTES, tmA, full0 + 8u * s, c, j0, i - 1, hintA);
      }
      __syncwarp();
    }
  }
}

// ---- one k-chunk of one pass: MMA issuer side -----------------------------------------------------
template <int NS, int LMAX, int PASS, int PAR, int DBG>
__device__ __forceinline__ void oz3_issue_chunk(uint32_t desc0 /* (smem0 >> 4) | LBO */, uint32_t full0, uint32_t empty0,
                                                uint32_t& ph_full, uint32_t tbase, uint32_t idesc, uint32_t not_first) {
  using Z = Oz3<NS, LMAX>;
#pragma unroll
  for (int i = 1; i <= Z::smax(PASS); ++i) {
#pragma unroll
    for (int ip = Z::whi(PASS, i); ip >= Z::wlo(PASS, i); --ip) {
      if (i == 1 || ip < Z::wlo(PASS, i - 1)) {
        const int s = Z::slotT(PASS, ip, PAR);
        oz3_wait(full0 + 8u * s, (ph_full >> s) & 1u);
        ph_full ^= 1u << s;
      }
    }
    {
      const int s = Z::slotA(PASS, i, PAR);
      oz3_wait(full0 + 8u * s, (ph_full >> s) & 1u);
      ph_full ^= 1u << s;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (oz_elect_one()) {
      const uint32_t alo = desc0 + (uint32_t)(Z::slotA(PASS, i, PAR) * (OZ3_TILE_BYTES >> 4));
#pragma unroll
      for (int ip = Z::wlo(PASS, i); ip <= Z::whi(PASS, i); ++ip) {
        const int l = i + ip;
        const uint32_t d_tmem = tbase + (uint32_t)((l - Z::lo(PASS)) * OZ_BN);
        const uint32_t blo = desc0 + (uint32_t)(Z::slotT(PASS, ip, PAR) * (OZ3_TILE_BYTES >> 4));
        const bool level_start = (i == ((l - NS > 1) ? l - NS : 1));
#pragma unroll
        for (int k = 0; k < OZ3_BK / 32; ++k)
          if (!(DBG & 1)) oz3_mma(d_tmem, alo + 2u * k, blo + 2u * k, idesc, (level_start && k == 0) ? not_first : 1u);
      }
      oz3_commit(empty0 + 8u * Z::slotA(PASS, i, PAR));
      if (i < Z::smax(PASS)) {
        if (Z::hi(PASS) - i <= NS) oz3_commit(empty0 + 8u * Z::slotT(PASS, Z::hi(PASS) - i, PAR));
      } else {
#pragma unroll
        for (int ip = Z::wlo(PASS, i); ip <= Z::whi(PASS, i); ++ip) oz3_commit(empty0 + 8u * Z::slotT(PASS, ip, PAR));
      }
    }
    __syncwarp();
  }
}

template <int NS, int LMAX, int PASS, int DBG>
__device__ __forceinline__ void oz3_produce_pass(const CUtensorMap* tmA, const CUtensorMap* tmT, uint32_t smem0,
                                                 uint32_t full0, uint32_t empty0, uint32_t& ph_empty, int c_begin,
                                                 int kchunks, int j0, int r0, uint64_t hintA, uint64_t hintT, int pf) {
  for (int kc = 0; kc < kchunks; kc += 2) {
    oz3_produce_chunk<NS, LMAX, PASS, 0, DBG>(tmA, tmT, smem0, full0, empty0, ph_empty, c_begin + kc * OZ3_BK, j0, r0, hintA, hintT,
                                           (pf && kc + OZ3_PF < kchunks) ? c_begin + (kc + OZ3_PF) * OZ3_BK : -1);
    if (kc + 1 < kchunks)
      oz3_produce_chunk<NS, LMAX, PASS, 1, DBG>(tmA, tmT, smem0, full0, empty0, ph_empty, c_begin + (kc + 1) * OZ3_BK, j0, r0,
                                             hintA, hintT, (pf && kc + 1 + OZ3_PF < kchunks) ? c_begin + (kc + 1 + OZ3_PF) * OZ3_BK : -1);
  }
}
template <int NS, int LMAX, int PASS, int DBG>
__device__ __forceinline__ void oz3_issue_pass(uint32_t desc0, uint32_t full0, uint32_t empty0, uint32_t& ph_full,
                                               uint32_t tbase, uint32_t idesc, int kchunks) {
  for (int kc = 0; kc < kchunks; kc += 2) {
    oz3_issue_chunk<NS, LMAX, PASS, 0, DBG>(desc0, full0, empty0, ph_full, tbase, idesc, kc > 0 ? 1u : 0u);
    if (kc + 1 < kchunks) oz3_issue_chunk<NS, LMAX, PASS, 1, DBG>(desc0, full0, empty0, ph_full, tbase, idesc, 1u);
  }
}

// int64 -> double (round to nearest, as I2F.F64.S64) from the two 32-bit halves with the 2^52 exponent trick: two
// LOPs, two DADDs and one DFMA on the full-rate FP64 pipe instead of one slow 64-bit conversion per element.
__device__ __forceinline__ double oz3_ll2double(long long a) {
  const unsigned lo = (unsigned)a;
  const int hi = (int)(a >> 32);
  const double dlo = __hiloint2double(0x43300000, (int)lo) - 4503599627370496.0;                        // 2^52 + lo
  const double dhi = __hiloint2double(0x43300000, hi ^ (int)0x80000000) - 4503601774854144.0;           // 2^52 + 2^31 + hi
  return fma(dhi, 4294967296.0, dlo);
}

// Tile order of the persistent grid.  Tiles are grouped in super-rows of p.sj row blocks (the A planes of a
// super-row, sj x 128 x K x NS bytes, stay in L2 while every column block passes over them); inside a super-row
// the column blocks go heaviest first (triangular T: the K range of a column block is proportional to its
// index) with the row block as the fast index, so CTAs running together share T tiles and the static round-robin
// t = blockIdx.x + i gridDim.x hands every CTA the same mix of heavy and light tiles.
struct Oz3Tile {
  int r0, j0, c_begin, kchunks;
};
__device__ __forceinline__ Oz3Tile oz3_tile(int t, int nN, int nJ, const OzParams& p) {
  int rb, jb;
  if (t & OZ_TILE_EXPLICIT) {                                  // host-built order: entry = row block * nN + column block
    const int e = t & ~OZ_TILE_EXPLICIT;
    jb = e / nN;
    rb = e - jb * nN;
  } else {                                                     // position in the supertiled order
    const int OZ3_SJ = p.sj;
    const int per = OZ3_SJ * nN;
    const int st = t / per, rem = t - st * per;
    const int jcount = min(OZ3_SJ, nJ - st * OZ3_SJ);
    const int rbi = rem / jcount;
    jb = st * OZ3_SJ + (rem - rbi * jcount);
    rb = (p.tri == 1) ? (nN - 1 - rbi) : rbi;
  }
  Oz3Tile x;
  x.r0 = rb * OZ_BN;
  x.j0 = jb * OZ_BM;
  x.c_begin = (p.tri == 2) ? x.r0 : 0;
  const int c_end = (p.tri == 1) ? min(p.K, x.r0 + OZ_BN) : p.K;
  x.kchunks = (c_end - x.c_begin) / OZ3_BK;
  return x;
}

template <int NS, int LMAX, int DBG = 0>
__global__ void __launch_bounds__(OZ3_THREADS, 1)
ozaki_i8_gemm_v3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmT,
                        const double* __restrict__ sA, const double* __restrict__ sT, double* __restrict__ C, OzParams p,
                        int nN, int nJ) {
  using Z = Oz3<NS, LMAX>;
  extern __shared__ uint8_t oz_raw[];
  __shared__ uint64_t full_bar[Z::NSLOT], empty_bar[Z::NSLOT], acc_full, acc_empty;
  __shared__ uint32_t tmem_base_sh;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);      // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const uint32_t smem0 = (oz_smem_u32(oz_raw) + 1023u) & ~1023u;
  const int ntiles = nN * nJ;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < Z::NSLOT; ++s) {
      oz_mbar_init(&full_bar[s], 1);
      oz_mbar_init(&empty_bar[s], 1);
    }
    oz_mbar_init(&acc_full, 1);
    oz_mbar_init(&acc_empty, (OZ3_THREADS - 64) / 32);
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
  const uint32_t full0 = oz_smem_u32(&full_bar[0]), empty0 = oz_smem_u32(&empty_bar[0]);

  if (warp == 0) {
    // ================= TMA producer: runs ahead across passes and tiles =================
    uint32_t ph_empty = 0xffffffffu;                          // first wait on every slot passes
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const Oz3Tile x = oz3_tile(t, nN, nJ, p);
      oz3_produce_pass<NS, LMAX, 0, DBG>(&tmA, &tmT, smem0, full0, empty0, ph_empty, x.c_begin, x.kchunks, x.j0, x.r0, p.hintA, p.hintT, p.pf);
      if constexpr (Z::NPASS > 1)
        oz3_produce_pass<NS, LMAX, 1, DBG>(&tmA, &tmT, smem0, full0, empty0, ph_empty, x.c_begin, x.kchunks, x.j0, x.r0, p.hintA, p.hintT, p.pf);
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BN >> 3) << 17) | ((uint32_t)(OZ_BM >> 4) << 24);
    const uint32_t desc0 = ((smem0 & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t accf = oz_smem_u32(&acc_full), acce = oz_smem_u32(&acc_empty);
    uint32_t ph_full = 0u;
    uint32_t g = 0;                                           // passes issued so far (accumulator hand-overs)
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const Oz3Tile x = oz3_tile(t, nN, nJ, p);
      if (g > 0) {
        oz3_wait(acce, (g - 1) & 1u);                         // epilogue has drained the accumulators
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      oz3_issue_pass<NS, LMAX, 0, DBG>(desc0, full0, empty0, ph_full, tbase, idesc, x.kchunks);
      if (oz_elect_one()) oz3_commit(accf);
      __syncwarp();
      ++g;
      if constexpr (Z::NPASS > 1) {
        oz3_wait(acce, (g - 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        oz3_issue_pass<NS, LMAX, 1, DBG>(desc0, full0, empty0, ph_full, tbase, idesc, x.kchunks);
        if (oz_elect_one()) oz3_commit(accf);
        __syncwarp();
        ++g;
      }
    }
  } else {
    // ================= epilogue: TMEM -> int64 recombination -> FP64 -> global =================
    // tcgen05.ld.16x256b hands a warp 16 accumulator rows x 8 columns per repetition in the mma C-fragment
    // layout (lane t: row t/4 and t/4+8, columns 2 (t%4) + {0,1}), so four neighbouring lanes write 64
    // contiguous bytes of one C row and a warp-wide 16-byte store touches 8 lines instead of the 32 that the
    // row-per-lane shape (32x32b) costs; the L1 tag stage, one line per clock, was the epilogue's limiter.
    constexpr int EW = (OZ3_THREADS - 64) / 128;              // warps per TMEM lane quarter
    constexpr int CW = OZ_BN / EW;                            // columns per warp
    static_assert(CW % 16 == 0, "epilogue column split");
    const int quarter = warp & 3;                             // TMEM lanes [32 quarter, +32) belong to this warp
    const int part = (warp - 2) >> 2;                         // columns [CW part, +CW)
    const int fr = lane >> 2, fc = (lane & 3) * 2;            // fragment row / column of this lane
    const uint32_t accf = oz_smem_u32(&acc_full);
    constexpr int NG = CW / 16;                               // 16-column groups per warp
    uint32_t g = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const Oz3Tile x = oz3_tile(t, nN, nJ, p);
      // row / column scales of this lane's fragment positions: fetched while the MMAs of the tile run
      double sa[2][2];
      double2 sc[NG][2];
      double* crow[2][2];
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const size_t j = (size_t)(x.j0 + quarter * 32 + h * 16 + fr + 8 * e);
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
        // sum_t r_t 2^(2 - 8 (lo + t)) = 2^(2 - 8 (lo + nl - 1)) * sum_t r_t 2^(8 (nl - 1 - t))
        const double w = exp2((double)(2 - 8 * (lo + nl - 1)));
        oz3_wait(accf, g & 1u);
        ++g;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (!(DBG & 4)) {
#pragma unroll
          for (int gq = 0; gq < NG; ++gq) {
            // this 16-column group: both 16-row halves; the read-modify-write operands of pass > 0 are requested first
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
                               : "=r"(r[tt][0]), "=r"(r[tt][1]), "=r"(r[tt][2]), "=r"(r[tt][3]), "=r"(r[tt][4]),
                                 "=r"(r[tt][5]), "=r"(r[tt][6]), "=r"(r[tt][7])
                               : "r"(taddr + (uint32_t)(tt * OZ_BN)));
                }
              }
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
              for (int n = 0; n < 2; ++n) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {               // e: row fr (0) / fr + 8 (1); registers 4n + 2e + {0,1}
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
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) oz_mbar_arrive(&acc_empty);            // one arrival per epilogue warp
      }
    }
  }
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512) : "memory");
}
