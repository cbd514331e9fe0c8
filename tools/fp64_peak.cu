// fp64_peak.cu — microbenchmark: sustained DMMA.8x8x4 vs DFMA issue rate per SM on B200.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int NACC>
__global__ void dmma_loop(double* out, double a0, double b0) {
  double c[NACC][2];
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void dfma_loop(double* out, double a0, double b0) {
  double c[NACC];
  for (int i = 0; i < NACC; ++i) c[i] = i;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(a, c[i], b);
  }
  double s = 0;
  for (int i = 0; i < NACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F>
float time_it(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  printf("device %s SMs %d clock %d kHz\n", p.name, sms, p.clockRate);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  for (int warps : {1, 2, 4, 8, 16}) {
    int threads = warps * 32;
    for (int rep = 0; rep < 2; ++rep) {
      float ms = time_it([&] { dmma_loop<16><<<sms, threads>>>(out, 1.0, 1.0); });
      double flops = 2.0 * 256 * 16 * (double)ITERS * warps * sms;
      if (rep) printf("DMMA  warps/SM %2d  %.3f ms  %.2f TFLOP/s\n", warps, ms, flops / ms * 1e-9);
    }
  }
  for (int warps : {4, 8, 16, 32}) {
    int threads = warps * 32;
    for (int rep = 0; rep < 2; ++rep) {
      float ms = time_it([&] { dfma_loop<16><<<sms, threads>>>(out, 1.0000001, 1e-9); });
      double flops = 2.0 * 32 * 16 * (double)ITERS * warps * sms;
      if (rep) printf("DFMA  warps/SM %2d  %.3f ms  %.2f TFLOP/s\n", warps, ms, flops / ms * 1e-9);
    }
  }
  // sustained ~2 s DMMA run to see the power-capped clock
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    int n = 0;
    for (; n < 400; ++n) dmma_loop<16><<<sms, 256>>>(out, 1.0, 1.0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 256 * 16 * (double)ITERS * 8 * sms * n;
    printf("DMMA sustained: %d launches %.1f ms  %.2f TFLOP/s\n", n, ms, flops / ms * 1e-9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
