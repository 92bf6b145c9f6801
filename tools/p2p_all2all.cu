// All-to-all push microbenchmark: every GPU stores `mb` MB into every other GPU at the same time
// (single process, peer access) - the traffic pattern of the fused x exchange.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/p2p_all2all.bin tools/p2p_all2all.cu
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

struct Peers { double2 *p[8]; int n; };

// mode 0: thread loads once, stores to every peer (what cgs_finish does); mode 1: one peer per blockIdx.y
__global__ void push_all(const double2 *__restrict__ src, Peers peers, size_t n2, size_t my_off, int self, int mode) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  if (mode == 0) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
      double2 v = src[i];
      for (int p = 0; p < peers.n; ++p) if (p != self) peers.p[p][my_off + i] = v;
    }
  } else {
    int p = blockIdx.y; if (p == self) return;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) peers.p[p][my_off + i] = src[i];
  }
}

int main() {
  int nd = 0; CK(cudaGetDeviceCount(&nd));
  if (nd < 2) { printf("need >= 2 GPUs\n"); return 0; }
  if (nd > 8) nd = 8;
  for (int a = 0; a < nd; ++a) { CK(cudaSetDevice(a)); for (int b = 0; b < nd; ++b) if (a != b) CK(cudaDeviceEnablePeerAccess(b, 0)); }
  const size_t mb = 8, n2 = mb * 1024 * 1024 / 16;
  std::vector<double2 *> src(nd), dst(nd);
  std::vector<cudaStream_t> st(nd);
  std::vector<cudaEvent_t> e0(nd), e1(nd);
  for (int a = 0; a < nd; ++a) {
    CK(cudaSetDevice(a)); CK(cudaMalloc(&src[a], n2 * 16)); CK(cudaMalloc(&dst[a], n2 * 16 * nd)); CK(cudaMemset(src[a], 1, n2 * 16));
    CK(cudaStreamCreate(&st[a])); cudaEventCreate(&e0[a]); cudaEventCreate(&e1[a]);
  }
  Peers peers; peers.n = nd; for (int a = 0; a < nd; ++a) peers.p[a] = dst[a];
  for (int mode = 0; mode < 2; ++mode) for (int grid : {148, 592, 1184}) {
    float worst = 0;
    for (int it = 0; it < 4; ++it) {
      for (int a = 0; a < nd; ++a) { cudaSetDevice(a); cudaDeviceSynchronize(); }
      for (int a = 0; a < nd; ++a) {
        cudaSetDevice(a); cudaEventRecord(e0[a], st[a]);
        for (int r = 0; r < 10; ++r) {
          dim3 g(grid, mode == 1 ? nd : 1);
          push_all<<<g, 256, 0, st[a]>>>(src[a], peers, n2, (size_t)a * n2, a, mode);
        }
        cudaEventRecord(e1[a], st[a]);
      }
      worst = 0;
      for (int a = 0; a < nd; ++a) { cudaSetDevice(a); CK(cudaEventSynchronize(e1[a])); float ms; cudaEventElapsedTime(&ms, e0[a], e1[a]); if (ms > worst) worst = ms; }
    }
    double bytes = (double)n2 * 16 * (nd - 1);
    printf("N=%d mode %d grid %4d: %7.1f us per all-to-all push of %zu MB/peer, egress %6.1f GB/s per GPU\n", nd, mode, grid,
           worst * 100, mb, bytes / (worst / 10 * 1e-3) / 1e9);
  }
  return 0;
}
