// kernels_blocks.cuh - one-time operator preparation on the device (upload path, not the hot loop):
//   * validate_csr      index-range / monotonicity check of an uploaded CSR or CSC structure (the reference gets
//                       this from the SparseMatrixCSC constructor; a malformed array must be an ArgumentError,
//                       not an out-of-bounds gather)
//   * owner-group build reorder the entries of a row shard BLOCK-MAJOR by the GROUP of ranks that owns the column:
//                       owner = col / W, block = ((owner - me) mod P) / G, i.e. block 0 = this rank's own columns
//                       and those of the next G - 1 ranks in the order in which the staged exchange delivers their
//                       slices, block 1 the next G ranks, ... - stable within a row - so that the row-sharded
//                       mat-vec runs one L2-sized pass per group behind the arrivals (b2a.cu enqueue_matvec):
//                       count -> flat inclusive scan -> scatter.  No library calls (no thrust / cub).
#pragma once

#include "device_common.cuh"

namespace b2a {

// err bits: 1 = pointer array not monotone / wrong ends, 2 = index out of range
__global__ void __launch_bounds__(256)
    validate_csr_kernel(int64_t n_ptr, const int64_t *__restrict__ ptr, int64_t nnz, const int32_t *__restrict__ idx,
                        int64_t idx_bound, int *err) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int bad = 0;
  for (int64_t r = tid; r < n_ptr; r += stride) {
    const int64_t a = ptr[r], b = ptr[r + 1];
    if (a > b || a < 0 || b > nnz) bad |= 1;
  }
  if (tid == 0 && (ptr[0] != 0 || ptr[n_ptr] != nnz)) bad |= 1;
  for (int64_t i = tid; i < nnz; i += stride) {
    const int32_t c = idx[i];
    if (c < 0 || (int64_t)c >= idx_bound) bad |= 2;
  }
  if (bad) atomicOr(err, bad);
}

constexpr int kMaxOwnerBlocks = 16;
struct OwnerGroups {
  int64_t W;  // rows (= columns) per rank of the uniform partition
  int me, P, G;
};
__device__ __forceinline__ int owner_block(const OwnerGroups &og, int32_t col) {
  const int owner = (int)((int64_t)col / og.W);
  return ((owner - og.me + og.P) % og.P) / og.G;
}

// bptr: nblocks x (n_rows + 1) int64, zero on entry; on exit bptr[b][r + 1] = entries of row r in block b
__global__ void __launch_bounds__(256)
    blk_count_kernel(int64_t n_rows, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
                     OwnerGroups og, int64_t *bptr) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t ld = n_rows + 1;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += stride) {
    const int64_t s = rowptr[r], e = rowptr[r + 1];
    for (int64_t i = s; i < e; ++i) {
      const int64_t b = owner_block(og, colind[i]);
      bptr[b * ld + r + 1] += 1;  // slot private to this thread
    }
  }
}

// ---- flat inclusive scan of an int64 array: per-chunk sums -> scan of the sums (one CTA) -> apply
constexpr int kScanThreads = 256;
constexpr int kScanPer = 8;
constexpr int kScanChunk = kScanThreads * kScanPer;

__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t *total) {
  // exclusive scan of one value per thread over the CTA (kScanThreads threads)
  __shared__ int64_t wsum[kScanThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) wsum[warp] = x;
  __syncthreads();
  int64_t base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w) {
    if (w < warp) base += wsum[w];
    tot += wsum[w];
  }
  __syncthreads();
  if (total) *total = tot;
  return base + x - v;
}

__global__ void __launch_bounds__(kScanThreads)
    scan_partial_kernel(const int64_t *__restrict__ a, int64_t L, int64_t *__restrict__ sums) {
  const int64_t base = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanPer;
  int64_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanPer; ++k)
    if (base + k < L) s += a[base + k];
  int64_t tot;
  (void)block_exclusive_scan(s, &tot);
  if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

// exclusive scan of sums[0 .. m) in place, one CTA, sequential over slabs of kScanThreads with a running carry
__global__ void __launch_bounds__(kScanThreads) scan_sums_kernel(int64_t *sums, int64_t m) {
  int64_t carry = 0;
  for (int64_t b0 = 0; b0 < m; b0 += kScanThreads) {
    const int64_t i = b0 + threadIdx.x;
    const int64_t v = i < m ? sums[i] : 0;
    int64_t tot;
    const int64_t ex = block_exclusive_scan(v, &tot);
    if (i < m) sums[i] = carry + ex;
    carry += tot;
  }
}

__global__ void __launch_bounds__(kScanThreads)
    scan_apply_kernel(int64_t *__restrict__ a, int64_t L, const int64_t *__restrict__ sums) {
  const int64_t base = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanPer;
  int64_t v[kScanPer];
  int64_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanPer; ++k) {
    v[k] = base + k < L ? a[base + k] : 0;
    s += v[k];
  }
  int64_t run = sums[blockIdx.x] + block_exclusive_scan(s, nullptr);
#pragma unroll
  for (int k = 0; k < kScanPer; ++k) {
    run += v[k];
    if (base + k < L) a[base + k] = run;
  }
}

// bptr (scanned): bptr[b][r] = first position of row r in block b.  Stable within (block, row).
template <class T>
__global__ void __launch_bounds__(256)
    blk_scatter_kernel(int64_t n_rows, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
                       const T *__restrict__ vals, OwnerGroups og, int nblocks, const int64_t *__restrict__ bptr,
                       int32_t *__restrict__ bcol, T *__restrict__ bval) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t ld = n_rows + 1;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += stride) {
    int64_t cur[kMaxOwnerBlocks];
    for (int b = 0; b < nblocks; ++b) cur[b] = bptr[(int64_t)b * ld + r];
    const int64_t s = rowptr[r], e = rowptr[r + 1];
    for (int64_t i = s; i < e; ++i) {
      const int32_t c = colind[i];
      const int b = owner_block(og, c);
      const int64_t dst = cur[b]++;
      bcol[dst] = c;
      bval[dst] = vals[i];
    }
  }
}

}  // namespace b2a
