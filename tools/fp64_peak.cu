// FP64 throughput probe for the roofline denominator of the basis rotation ("FP64 peak must be
// measured on the box"): dependent-chain-free DFMA and DMMA.8x8x4 (mma.sync.m8n8k4.f64) loops, CUDA events.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak.bin tools/fp64_peak.cu
//   tools/fp64_peak.bin  ->  one JSON line {"dfma_tflops": .., "dmma_tflops": ..}
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a, double b) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dmma_kernel(double *out, int iters, double a, double b) {
  double acc[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                   : "+d"(acc[i][0]), "+d"(acc[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *out;
  cudaMalloc(&out, sizeof(double) * 256 * sms * 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000, grid = sms * 8;
  double best[2] = {0, 0};
  for (int which = 0; which < 2; ++which)
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      if (which == 0)
        dfma_kernel<<<grid, 256>>>(out, iters, 0.999999, 1e-9);
      else
        dmma_kernel<<<grid, 256>>>(out, iters, 0.999999, 1e-9);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double flops = which == 0 ? (double)grid * 256 * 16 * 2.0 * iters : (double)grid * 8 /*warps*/ * 8 * 512.0 * iters;
      const double tf = flops / (ms * 1e-3) / 1e12;
      if (rep > 0 && tf > best[which]) best[which] = tf;
    }
  cudaError_t e = cudaGetLastError();
  printf("{\"sms\": %d, \"dfma_tflops\": %.2f, \"dmma_tflops\": %.2f, \"status\": \"%s\"}\n", sms, best[0], best[1],
         cudaGetErrorString(e));
  return 0;
}
