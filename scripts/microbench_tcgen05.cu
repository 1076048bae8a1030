// Raw tcgen05.mma issue-rate probe on sm_100a: kind::i8 (int8 x int8 -> int32 in TMEM) and kind::tf32,
// operands in shared memory (K-major, SWIZZLE_128B canonical layout, contents irrelevant), one CTA per SM,
// one thread issues, tcgen05.commit -> mbarrier, all threads wait.  Prints achieved TOP/s per shape.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/microbench_tcgen05 scripts/microbench_tcgen05.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);            // start address  [0,14)
  d |= (uint64_t)1 << 16;                              // LBO (ignored for swizzled K-major) [16,30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;    // SBO [32,46)
  d |= (uint64_t)1 << 46;                              // version = 1 (sm_100)
  d |= (uint64_t)layout_type << 61;                    // 2 = SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity));
}

// kind: 0 = i8 (K = 32 bytes per MMA), 1 = tf32 (K = 8 elements = 32 bytes per MMA)
template <int KIND, int N>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, int mmas_per_iter, uint32_t* out, int rot, int commit_every, int layout) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x == 32) {
    mbar_init(&bar, 1);
    mbar_init(&bar2, 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tbase = tmem_base;
  // instruction descriptor
  uint32_t idesc;
  if (KIND == 0) idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);   // S32, s8, s8
  else           idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);   // F32, tf32, tf32
  const uint32_t a_addr = smem_u32(smem);                    // A: 128 rows x 128 B
  const uint32_t b_addr = smem_u32(smem + 128 * 128);        // B: N rows x 128 B
  if (threadIdx.x == 0) {
    for (int it = 0; it < iters; ++it) {
      for (int m = 0; m < mmas_per_iter; ++m) {
        const int ks = m & 3;                                 // 4 k-steps of 32 B inside the 128-byte swizzle atom
        const uint64_t da = layout == 2 ? make_desc(a_addr, 1024, 2) + (uint64_t)(ks * 2) : make_desc(a_addr + ks * 4096, 256, 6);
        const uint64_t db = layout == 2 ? make_desc(b_addr, 1024, 2) + (uint64_t)(ks * 2) : make_desc(b_addr + ks * 8192, 256, 6);
        const uint32_t d_tmem = tbase + (uint32_t)((m % rot) * N);   // rotate over `rot` accumulators
        const uint32_t acc = (it | m) ? 1u : 0u;
        if (KIND == 0)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc));
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc));
        if (commit_every && (m + 1) % commit_every == 0)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)));
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;");
  // touch the accumulator so that nothing is optimised away
  uint32_t v0;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v0) : "r"(tbase + ((uint32_t)(warp * 32) << 16)));
  asm volatile("tcgen05.wait::ld.sync.aligned;");
  if (threadIdx.x == 0) out[blockIdx.x] = v0;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

template <int KIND, int N>
int run(const char* name, int sms, uint32_t* out, int rot = 512 / N, int commit_every = 0, int layout = 2) {
  const int smem = 128 * 128 + 256 * 128 + 1024;
  CK(cudaFuncSetAttribute(mma_rate_kernel<KIND, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 2000, per = 28;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    mma_rate_kernel<KIND, N><<<sms, 128, smem>>>(iters, per, out, rot, commit_every, layout);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
  }
  CK(cudaGetLastError());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double kdim = KIND == 0 ? 32.0 : 8.0;
  double ops = 2.0 * 128 * N * kdim * (double)iters * per * sms;
  printf("{\"bench\": \"tcgen05_%s\", \"M\": 128, \"N\": %d, \"rot\": %d, \"commit_every\": %d, \"layout\": %d, \"ms\": %.3f, \"tops\": %.1f}\n", name, N, rot, commit_every, layout, ms, ops / ms * 1e-9);
  return 0;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  uint32_t* out;
  CK(cudaMalloc(&out, sizeof(uint32_t) * sms));
  if (run<0, 64>("i8", sms, out)) return 1;
  if (run<0, 128>("i8", sms, out)) return 1;
  for (int rot = 1; rot <= 4; rot *= 2) run<0, 128>("i8", sms, out, rot, 0, 2);
  run<0, 128>("i8", sms, out, 4, 7, 2);
  run<0, 128>("i8", sms, out, 4, 1, 2);
  run<0, 128>("i8", sms, out, 1, 7, 2);
  run<0, 128>("i8", sms, out, 4, 0, 6);
  run<0, 128>("i8", sms, out, 1, 0, 6);
  run<0, 128>("i8", sms, out, 1, 7, 6);
  run<0, 256>("i8", sms, out, 1, 0, 2);
  run<0, 256>("i8", sms, out, 2, 0, 2);
  if (run<0, 256>("i8", sms, out)) return 1;
  if (run<1, 64>("tf32", sms, out)) return 1;
  if (run<1, 128>("tf32", sms, out)) return 1;
  if (run<1, 256>("tf32", sms, out)) return 1;
  return 0;
}
