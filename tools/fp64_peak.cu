// FP64 throughput probe for the roofline denominator of the basis rotation ("FP64 peak must be
// measured on the box"): dependent-chain-free DFMA and DMMA.8x8x4 (mma.sync.m8n8k4.f64) loops, CUDA events.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak.bin tools/fp64_peak.cu
//   tools/fp64_peak.bin  ->  one JSON line: peaks, plus DMMA throughput against resident warps per SM sub-partition
//   and independent accumulator chains per warp (how many warps a DMMA kernel needs to fill the pipe).
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

template <int ILP>
__global__ void __launch_bounds__(256) dmma_kernel(double *out, int iters, double a, double b) {
  double acc[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i][0] = acc[i][1] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                   : "+d"(acc[i][0]), "+d"(acc[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i][0] + acc[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static cudaEvent_t e0, e1;
template <class F> static double best_tflops(F launch, double flops) {
  double best = 0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  return best;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *out;
  cudaMalloc(&out, sizeof(double) * 256 * sms * 8);
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  const double dfma = best_tflops([&] { dfma_kernel<<<sms * 8, 256>>>(out, iters, 0.999999, 1e-9); },
                                  (double)sms * 8 * 256 * 16 * 2.0 * iters);
  const double dmma = best_tflops([&] { dmma_kernel<8><<<sms * 8, 256>>>(out, iters, 0.999999, 1e-9); },
                                  (double)sms * 8 * 8 * 8 * 512.0 * iters);
  printf("{\"sms\": %d, \"dfma_tflops\": %.2f, \"dmma_tflops\": %.2f, \"dmma_by_warps_per_smsp\": {", sms, dfma, dmma);
  // one CTA per SM with 4 * w warps = w warps per sub-partition
  const int wps[] = {1, 2, 4, 8};
  for (int wi = 0; wi < 4; ++wi) {
    const int w = wps[wi], threads = 128 * w;
    const double f8 = best_tflops([&] { dmma_kernel<8><<<sms, threads>>>(out, iters, 0.999999, 1e-9); },
                                  (double)sms * 4 * w * 8 * 512.0 * iters);
    const double f4 = best_tflops([&] { dmma_kernel<4><<<sms, threads>>>(out, iters, 0.999999, 1e-9); },
                                  (double)sms * 4 * w * 4 * 512.0 * iters);
    const double f1 = best_tflops([&] { dmma_kernel<1><<<sms, threads>>>(out, iters, 0.999999, 1e-9); },
                                  (double)sms * 4 * w * 1 * 512.0 * iters);
    printf("%s\"%d\": {\"ilp8\": %.2f, \"ilp4\": %.2f, \"ilp1\": %.2f}", wi ? ", " : "", w, f8, f4, f1);
  }
  printf("}, \"status\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
