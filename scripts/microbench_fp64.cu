// FP64 roofline denominators on this B200: DFMA issue rate, DMMA.8x8x4 issue rate,
// cuBLAS DGEMM (the "measured peak" the FP64 roofline fractions are quoted against).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/microbench_fp64 scripts/microbench_fp64.cu -lcublas
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void dfma_kernel(double* out, int iters) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-9 + i;
  double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_kernel(double* out, int iters) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, prop.clockRate);
  double* out;
  CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int warps = 4; warps <= 32; warps *= 2) {
    int iters = 20000;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      dfma_kernel<<<sms, warps * 32>>>(out, iters);
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 16 * iters * (double)sms * warps * 32;
    printf("{\"bench\": \"dfma\", \"warps_per_sm\": %d, \"tflops\": %.3f}\n", warps, fl / ms * 1e-9);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      dmma_kernel<<<sms, warps * 32>>>(out, iters);
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
    }
    cudaEventElapsedTime(&ms, e0, e1);
    fl = 2.0 * 256 * 16 * iters * (double)sms * warps;
    printf("{\"bench\": \"dmma884\", \"warps_per_sm\": %d, \"tflops\": %.3f}\n", warps, fl / ms * 1e-9);
  }
  // cuBLAS DGEMM
  cublasHandle_t h;
  cublasCreate(&h);
  for (int n : {4096, 8192}) {
    double *A, *B, *C;
    size_t bytes = sizeof(double) * (size_t)n * n;
    CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
    std::vector<double> hbuf((size_t)n * n);
    for (size_t i = 0; i < hbuf.size(); ++i) hbuf[i] = (double)((i * 2654435761u) % 1000) * 1e-3 - 0.5;
    CK(cudaMemcpy(A, hbuf.data(), bytes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(B, hbuf.data(), bytes, cudaMemcpyHostToDevice));
    double one = 1.0, zero = 0.0;
    float best = 1e30f, total = 0;
    for (int rep = 0; rep < 8; ++rep) {
      cudaEventRecord(e0);
      cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n);
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep >= 2) { best = ms < best ? ms : best; total += ms; }
    }
    printf("{\"bench\": \"cublas_dgemm_tn\", \"n\": %d, \"best_tflops\": %.3f, \"mean_tflops\": %.3f}\n", n,
           2.0 * n * (double)n * n / best * 1e-9, 2.0 * n * (double)n * n / (total / 6) * 1e-9);
    // DTRMM: B := tril(A) * B  (n^3 flops)
    best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      cublasDtrmm(h, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, n, n, &one, A, n, B, n, C, n);
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep >= 1) best = ms < best ? ms : best;
    }
    printf("{\"bench\": \"cublas_dtrmm\", \"n\": %d, \"best_tflops_n3\": %.3f}\n", n, (double)n * n * n / best * 1e-9);
    cudaFree(A); cudaFree(B); cudaFree(C);
  }
  return 0;
}
