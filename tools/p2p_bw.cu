// Peer-store / peer-load bandwidth microbenchmark (single process, 2 GPUs, cudaDeviceEnablePeerAccess).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/p2p_bw tools/p2p_bw.cu && /tmp/p2p_bw
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int VEC> __global__ void push(const double *__restrict__ src, double *__restrict__ dst, size_t n) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  if (VEC == 2) {
    const double2 *s = (const double2 *)src; double2 *d = (double2 *)dst;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n / 2; i += stride) d[i] = s[i];
  } else if (VEC == 4) {
    const double4 *s = (const double4 *)src; double4 *d = (double4 *)dst;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n / 4; i += stride) d[i] = s[i];
  } else {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
  }
}
__global__ void pull(const double2 *__restrict__ src, double2 *__restrict__ dst, size_t n2) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) dst[i] = src[i];
}

int main() {
  int nd = 0; CK(cudaGetDeviceCount(&nd));
  if (nd < 2) { printf("need 2 GPUs\n"); return 0; }
  int can = 0; CK(cudaDeviceCanAccessPeer(&can, 0, 1)); printf("canAccessPeer(0,1) = %d\n", can);
  CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0));
  CK(cudaSetDevice(1)); CK(cudaDeviceEnablePeerAccess(0, 0));
  for (size_t mb : {8, 64}) {
    size_t n = mb * 1024 * 1024 / 8;
    double *a, *b, *loc;
    CK(cudaSetDevice(0)); CK(cudaMalloc(&a, n * 8)); CK(cudaMalloc(&loc, n * 8)); CK(cudaMemset(a, 1, n * 8));
    CK(cudaSetDevice(1)); CK(cudaMalloc(&b, n * 8));
    CK(cudaSetDevice(0));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int grid : {148, 592, 1184, 4736}) {
      for (int vec : {1, 2, 4}) {
        for (int it = 0; it < 3; ++it) {
          cudaEventRecord(e0);
          for (int r = 0; r < 10; ++r) {
            if (vec == 1) push<1><<<grid, 256>>>(a, b, n);
            else if (vec == 2) push<2><<<grid, 256>>>(a, b, n);
            else push<4><<<grid, 256>>>(a, b, n);
          }
          cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        }
        printf("push  %3zu MB grid %5d vec %d: %7.1f us  %6.1f GB/s\n", mb, grid, vec, ms * 100, n * 8 / (ms / 10 * 1e-3) / 1e9);
      }
    }
    for (int grid : {592, 4736}) {
      for (int it = 0; it < 3; ++it) {
        cudaEventRecord(e0);
        for (int r = 0; r < 10; ++r) pull<<<grid, 256>>>((const double2 *)b, (double2 *)loc, n / 2);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
      }
      printf("pull  %3zu MB grid %5d      : %7.1f us  %6.1f GB/s\n", mb, grid, ms * 100, n * 8 / (ms / 10 * 1e-3) / 1e9);
    }
    for (int it = 0; it < 3; ++it) {
      cudaEventRecord(e0);
      for (int r = 0; r < 10; ++r) cudaMemcpyPeerAsync(b, 1, a, 0, n * 8);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("memcpyPeer %3zu MB         : %7.1f us  %6.1f GB/s\n", mb, ms * 100, n * 8 / (ms / 10 * 1e-3) / 1e9);
    cudaFree(a); cudaFree(loc); cudaSetDevice(1); cudaFree(b);
  }
  return 0;
}
