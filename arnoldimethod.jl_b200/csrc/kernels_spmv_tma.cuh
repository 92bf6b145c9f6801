// kernels_spmv_tma.cuh - CSR mat-vec as a persistent TMA-fed stream ("CSR-stream").
//
//   y = A x  for `mul!(view(V,:,j+1), A, view(V,:,j))`  (src/expansion.jl:121)
//
// At upload the rows are cut into TILES of at most kSpmvMaxRows rows / kSpmvMaxNnz non-zeros
// (host pass over rowptr; tile boundaries on even rows).  One CTA per SM walks a contiguous
// range of tiles:
//   * the PRODUCER thread streams the tile's slice of A - colind, vals and the rowptr slice -
//     into a shared-memory ring with three 1-D TMA bulk copies (16-byte aligned supersets of
//     the slices; tens of KB each, so the per-operation cost of the TMA unit is amortised);
//   * phase A (8 consumer warps, nnz-parallel): p[i] = vals[i] * x[colind[i]] with the column
//     indices coming from shared memory, so the only long-latency operation is the gather of
//     x and each thread keeps 8 independent gathers in flight whatever the row lengths are;
//   * phase B (row-parallel): LPR lanes per row add the row's products from shared memory,
//     combine with warp shuffles, and lane 0 stores y[row] - fixed order, no atomics.
// The A stream is read exactly once, coalesced, by the copy engine; x is gathered through the
// read-only path (L1/L2); y is written once.  Algorithmic bytes: nnz (s+4) + 8 (n+1) + 2 n s.
#pragma once

#include "device_common.cuh"
#include "kernels_cgs_tma.cuh"
#include "kernels_spmv.cuh"

namespace b2a {

constexpr int kSpmvMaxRows = 512;
constexpr int kSpmvMaxNnz = 4096;  // per tile; ComplexF64 tiles use half (same bytes per stage)
constexpr int kSpmvMaxStages = 8;
template <class T> struct SpmvTile { static constexpr int max_nnz = kSpmvMaxNnz * 8 / (int)sizeof(T); };

struct SpmvSmem {
  uint64_t full[kSpmvMaxStages];
  uint64_t empty[kSpmvMaxStages];
};

template <class T> struct SpmvStage {
  static constexpr size_t vals_bytes = (size_t)(SpmvTile<T>::max_nnz + 8) * sizeof(T);
  static constexpr size_t cols_bytes = (size_t)(SpmvTile<T>::max_nnz + 8) * sizeof(int32_t);
  static constexpr size_t rp_bytes = (size_t)(kSpmvMaxRows + 4) * sizeof(int64_t);
  static constexpr size_t bytes = vals_bytes + cols_bytes + rp_bytes;
};

// tile_row[t] = first row of tile t (tile_row[ntiles] = n_rows)
template <class T, int LPR>
__global__ void __launch_bounds__(kTmaThreads, 2)
    spmv_csr_tma_kernel(int64_t n_rows, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
                        const T *__restrict__ vals, const int32_t *__restrict__ tile_row, int ntiles, int tiles_per_cta,
                        int stages, const T *__restrict__ x, T *__restrict__ y, const int *poison) {
  if (*poison) return;
  extern __shared__ __align__(128) unsigned char spmv_smem_raw[];
  SpmvSmem *sm = reinterpret_cast<SpmvSmem *>(spmv_smem_raw);
  unsigned char *ring = spmv_smem_raw + 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&sm->full[s], 1);
      mbar_init(&sm->empty[s], kTmaConsumerWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const int t_first = blockIdx.x * tiles_per_cta;
  const int ntl = max(0, min(tiles_per_cta, ntiles - t_first));

  if (warp == kTmaConsumerWarps) {
    if (lane == 0) {
      for (int t = 0; t < ntl; ++t) {
        const int s = t % stages;
        mbar_wait(&sm->empty[s], (((uint32_t)(t / stages)) & 1u) ^ 1u);
        const int r0 = __ldg(tile_row + t_first + t), r1 = __ldg(tile_row + t_first + t + 1);
        const int64_t nz0 = __ldg(rowptr + r0), nz1 = __ldg(rowptr + r1);
        const int64_t a0 = nz0 & ~(int64_t)3, a1 = (nz1 + 3) & ~(int64_t)3;  // 16-byte aligned superset
        const uint32_t cnt = (uint32_t)(a1 - a0);
        const uint32_t nrp = (uint32_t)((r1 - r0 + 1 + 1) & ~1);  // even number of rowptr entries
        unsigned char *st = ring + (size_t)s * SpmvStage<T>::bytes;
        const uint32_t bytes = cnt * (uint32_t)(sizeof(T) + 4) + nrp * 8u;
        mbar_expect_tx(&sm->full[s], bytes);
        if (cnt) {
          tma_load_1d(st, vals + a0, cnt * (uint32_t)sizeof(T), &sm->full[s]);
          tma_load_1d(st + SpmvStage<T>::vals_bytes, colind + a0, cnt * 4u, &sm->full[s]);
        }
        tma_load_1d(st + SpmvStage<T>::vals_bytes + SpmvStage<T>::cols_bytes, rowptr + r0, nrp * 8u, &sm->full[s]);
      }
    }
    return;
  }

  const int tid = threadIdx.x;  // 0..255 (consumers)
  constexpr int NC = kTmaConsumerWarps * 32;
  for (int t = 0; t < ntl; ++t) {
    const int s = t % stages;
    mbar_wait(&sm->full[s], ((uint32_t)(t / stages)) & 1u);
    unsigned char *st = ring + (size_t)s * SpmvStage<T>::bytes;
    T *pv = reinterpret_cast<T *>(st);
    const int32_t *pc = reinterpret_cast<const int32_t *>(st + SpmvStage<T>::vals_bytes);
    const int64_t *rp = reinterpret_cast<const int64_t *>(st + SpmvStage<T>::vals_bytes + SpmvStage<T>::cols_bytes);
    const int r0 = __ldg(tile_row + t_first + t), r1 = __ldg(tile_row + t_first + t + 1);
    const int nrows = r1 - r0;
    const int64_t nz0 = rp[0];
    const int64_t a0 = nz0 & ~(int64_t)3;
    const int shift = (int)(nz0 - a0);
    const int nt = (int)(rp[nrows] - nz0);

    // ---- phase A: products.  Blocks of 8 entries per thread, fully unrolled: the column indices come
    // from shared memory and the gathers are UNCONDITIONAL (slots past the tile read x[0]), so the
    // compiler issues all 8 gathers back to back; only the product store is predicated.
    for (int base = tid; base < nt; base += 8 * NC) {
      int32_t c[8];
      T xv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int idx = base + k * NC;
        c[k] = idx < nt ? pc[shift + idx] : 0;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) xv[k] = ld_ro<T>(x + c[k]);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int idx = base + k * NC;
        if (idx < nt) pv[shift + idx] = Scalar<T>::mul(pv[shift + idx], xv[k]);
      }
    }
    consumer_bar_sync();

    // ---- phase B: LPR lanes per row.  The loop bound is uniform over the CTA so that every lane
    // of a warp takes part in the full-mask shuffles (rows past the tile are predicated off).
    const int sub = tid & (LPR - 1);
    for (int rb = 0; rb < nrows; rb += NC / LPR) {
      const int r = rb + tid / LPR;
      const bool live = r < nrows;
      T acc = Scalar<T>::zero();
      if (live) {
        const int b = (int)(rp[r] - a0), e = (int)(rp[r + 1] - a0);
        for (int k = b + sub; k < e; k += LPR) acc = Scalar<T>::add(acc, pv[k]);
      }
      if (Scalar<T>::is_complex) {
        double2 *ap = reinterpret_cast<double2 *>(&acc);
        ap->x = group_sum_d<LPR>(ap->x);
        ap->y = group_sum_d<LPR>(ap->y);
      } else {
        double *ap = reinterpret_cast<double *>(&acc);
        *ap = group_sum_d<LPR>(*ap);
      }
      if (live && sub == 0) y[r0 + r] = acc;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm->empty[s]);
  }
}

}  // namespace b2a
