// kernels_spmv.cuh - sparse mat-vec y = A x for the operator call of the Arnoldi step
// (`mul!(view(V,:,j+1), A, view(V,:,j))`, src/expansion.jl:121; Julia's stdlib does this
// with a single-threaded CSC loop).
//
//   spmv_csr_vector : CSR, LPR lanes per row (LPR = 2..32 chosen from nnz/row at upload),
//                     coalesced row-segment loads of (colind, vals), gathers of x through
//                     the read-only path, warp-shuffle partial sums, U rows in flight per
//                     lane group for memory-level parallelism.  Deterministic.
//   spmv_csr_scalar : one thread per row (nnz/row <= ~2, e.g. diagonal operators).
//   spmv_csc_scatter: Julia's native CSC layout without a transpose: one lane group per
//                     COLUMN, y[row] += val * x[col] with red.global.add.f64 (result sums
//                     in arrival order - not bit-reproducible; mode 1 of b2a_csc_create).
//
//   Column blocking (x larger than L2, scattered columns): the operator is stored block-major - one CSR
//   per column block of ~32 MB of x - and the vector kernel runs once per block (first pass writes y,
//   the others accumulate), with L2 eviction hints: the A stream is evict_first, the gathers of x
//   evict_last, so the block of x stays L2-resident while A streams through.  Measured on a cfg-5 shard
//   (n = 1.25e7, 15 nnz/row): unblocked 2089 us (every gather misses L2), see DESIGN.md for blocked.
//
// Algorithmic bytes per launch: nnz (s + 4) + 8 (n + 1) + 2 n s.
#pragma once

#include "device_common.cuh"
#include "peer_comm.cuh"

namespace b2a {

template <class T> __device__ __forceinline__ T ld_ro(const T *p);
template <> __device__ __forceinline__ double ld_ro<double>(const double *p) { return __ldg(p); }
template <> __device__ __forceinline__ cdouble ld_ro<cdouble>(const cdouble *p) { return __ldg(p); }

// ---- L2 eviction-priority hints (createpolicy + ld ... .L2::cache_hint) ------------------------
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ int32_t ld_hint(const int32_t *p, uint64_t pol) {
  int32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ double ld_hint(const double *p, uint64_t pol) {
  double v;
  asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ cdouble ld_hint(const cdouble *p, uint64_t pol) {
  cdouble v;
  asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  return v;
}

// ---- coherent gathers for the row-sharded mat-vec -----------------------------------------------
// When x is the NVLink exchange buffer, peers (or the copy engines) write it while earlier launches of this
// very kernel may still sit in L1 with the previous step's lines, and a kernel launched early spins on the
// arrival flags while the data lands.  ld.global.nc requires read-only data for the kernel's lifetime and sits
// outside the memory model, so those instantiations gather x with L2-coherent loads (ld.global.cg: never served
// from L1; peer and copy-engine writes land in L2) ordered after the flag's ld.acquire.sys by the CTA barrier.
__device__ __forceinline__ double ld_coh(const double *p) { return __ldcg(p); }
__device__ __forceinline__ cdouble ld_coh(const cdouble *p) { return __ldcg(p); }
__device__ __forceinline__ double ld_coh_hint(const double *p, uint64_t pol) {
  double v;
  asm volatile("ld.global.cg.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol) : "memory");
  return v;
}
__device__ __forceinline__ cdouble ld_coh_hint(const cdouble *p, uint64_t pol) {
  cdouble v;
  asm volatile("ld.global.cg.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol) : "memory");
  return v;
}

// What a mat-vec launch waits for before it gathers from the exchange buffer (peer_comm.cuh):
//   mode 0  nothing (single GPU, or the rank's own column block read straight from the workspace column)
//   mode 1  every rank's slice, sequence number taken from the device counter seq_x (push model: the normalising
//           kernel of the previous step stored the slices)
//   mode 2  the slice of ONE owner rank, sequence number given by the host (staged exchange: one launch per owner
//           block, so the mat-vec on the blocks that have arrived overlaps the transfer of the others)
//   mode 3  the slices of all other ranks, host sequence number (staged exchange, operator not stored by owner group)
//   mode 4  the slices of the `count` ranks rank+owner, rank+owner+1, ... (mod P), this rank itself excepted (staged
//           exchange, one launch per owner GROUP: the gathers on the groups that have arrived overlap the transfer
//           of the others, and every pass gathers from an L2-sized part of x)
struct XWait {
  PeerView pv;
  int mode = 0;
  int owner = 0, count = 1;
  unsigned long long want = 0;
};
__device__ __forceinline__ void spmv_wait_x(const XWait &xw) {
  if (xw.mode == 0) return;
  if (threadIdx.x == 0) {
    if (xw.mode == 1)
      peer_x_wait(xw.pv);
    else if (xw.mode == 2)
      peer_x_wait_one(xw.pv, xw.owner, xw.want);
    else if (xw.mode == 3)
      peer_x_wait_others(xw.pv, xw.want);
    else
      for (int i = 0; i < xw.count; ++i) {
        const int o = (xw.pv.rank + xw.owner + i) % xw.pv.P;
        if (o != xw.pv.rank) peer_x_wait_one(xw.pv, o, xw.want);
      }
  }
  __syncthreads();
}

template <int LPR> __device__ __forceinline__ double group_sum_d(double v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// HINT: column-blocked mode - A loads evict_first, x gathers evict_last (see header)
// COH : x is the NVLink exchange buffer - coherent gathers, and the launch waits for its slice(s) first
// E   : entries in flight per lane and row.  The gathers are latency-bound (ncu r1b: long-scoreboard 35 of 44 stall
//       cycles per issue, L2 at 70 % of its sector rate), and a lane that walks its row one entry at a time serialises
//       (column index -> gather) round trips; with E = 4 a lane group of LPR = 4 has a whole 16-entry row in flight
//       at once.  Entries are still accumulated in ascending order, so the result is bit-identical for every E.
template <class T, int LPR, int U, bool HINT = false, bool COH = false, int E = 4>
__global__ void __launch_bounds__(256)
    spmv_csr_vector_kernel(int64_t n_rows, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
                           const T *__restrict__ vals, const T *x, T *__restrict__ y,
                           const int *poison, const __grid_constant__ XWait xw, int accumulate) {
  pdl_wait();
  if (*poison) return;
  if (COH) spmv_wait_x(xw);  // multi-GPU: wait until the slice(s) of x this launch gathers from have landed
  const int sub = threadIdx.x & (LPR - 1);
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / LPR;
  uint64_t pol_a = 0, pol_x = 0;
  if (HINT) {
    pol_a = l2_policy_evict_first();
    pol_x = l2_policy_evict_last();
  }
  auto ld_col = [&](const int32_t *p) { return HINT ? ld_hint(p, pol_a) : __ldg(p); };
  auto ld_val = [&](const T *p) { return HINT ? ld_hint(p, pol_a) : ld_ro<T>(p); };
  auto ld_x = [&](const T *p) {
    if (COH) return HINT ? ld_coh_hint(p, pol_x) : ld_coh(p);
    return HINT ? ld_hint(p, pol_x) : ld_ro<T>(p);
  };

  // loop bound uniform over the grid (rbase0 is the same for every thread) so that all lanes of a
  // warp reach the full-mask shuffles; rows past the end are predicated off
  for (int64_t rbase0 = 0; rbase0 < n_rows; rbase0 += ngroups * U) {
    const int64_t rbase = rbase0 + group;
    int64_t rs[U], re[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = rbase + (int64_t)u * ngroups;
      if (r < n_rows) {
        rs[u] = __ldg(rowptr + r);
        re[u] = __ldg(rowptr + r + 1);
      } else {
        rs[u] = re[u] = 0;
      }
    }
    // first E * LPR entries of each of the U rows: all index / value loads issued before any gather
    int32_t c[U][E];
    T a[U][E];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int64_t i = rs[u] + sub + e * LPR;
        const bool ok = i < re[u];
        c[u][e] = ok ? ld_col(colind + i) : -1;
        a[u][e] = ok ? ld_val(vals + i) : Scalar<T>::zero();
      }
    T acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      T xv[E];
#pragma unroll
      for (int e = 0; e < E; ++e) xv[e] = c[u][e] >= 0 ? ld_x(x + c[u][e]) : Scalar<T>::zero();
      acc[u] = Scalar<T>::mul(a[u][0], xv[0]);
#pragma unroll
      for (int e = 1; e < E; ++e) acc[u] = Scalar<T>::fma_(a[u][e], xv[e], acc[u]);
    }
    // rows longer than E * LPR: further chunks of E entries per lane
#pragma unroll
    for (int u = 0; u < U; ++u) {
      for (int64_t i0 = rs[u] + sub + E * LPR; i0 < re[u]; i0 += E * LPR) {
        int32_t cc[E];
        T aa[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int64_t i = i0 + e * LPR;
          const bool ok = i < re[u];
          cc[e] = ok ? ld_col(colind + i) : -1;
          aa[e] = ok ? ld_val(vals + i) : Scalar<T>::zero();
        }
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const T xv = cc[e] >= 0 ? ld_x(x + cc[e]) : Scalar<T>::zero();
          acc[u] = Scalar<T>::fma_(aa[e], xv, acc[u]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      T s = acc[u];
      if (Scalar<T>::is_complex) {
        double2 *sp = reinterpret_cast<double2 *>(&s);
        sp->x = group_sum_d<LPR>(sp->x);
        sp->y = group_sum_d<LPR>(sp->y);
      } else {
        double *sp = reinterpret_cast<double *>(&s);
        *sp = group_sum_d<LPR>(*sp);
      }
      const int64_t r = rbase + (int64_t)u * ngroups;
      if (sub == 0 && r < n_rows) y[r] = accumulate ? Scalar<T>::add(y[r], s) : s;
    }
  }
}

template <class T, bool COH = false>
__global__ void __launch_bounds__(256)
    spmv_csr_scalar_kernel(int64_t n_rows, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
                           const T *__restrict__ vals, const T *x, T *__restrict__ y,
                           const int *poison, const __grid_constant__ XWait xw, int accumulate) {
  if (*poison) return;
  if (COH) spmv_wait_x(xw);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += stride) {
    const int64_t s = __ldg(rowptr + r), e = __ldg(rowptr + r + 1);
    T acc = Scalar<T>::zero();
    for (int64_t i = s; i < e; ++i) {
      const T xv = COH ? ld_coh(x + __ldg(colind + i)) : ld_ro<T>(x + __ldg(colind + i));
      acc = Scalar<T>::fma_(ld_ro<T>(vals + i), xv, acc);
    }
    y[r] = accumulate ? Scalar<T>::add(y[r], acc) : acc;
  }
}

// ---- CSC scatter (K2) ----------------------------------------------------------
__device__ __forceinline__ void red_add(double *p, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

template <class T>
__global__ void __launch_bounds__(256) zero_vector_kernel(T *__restrict__ y, int64_t n, const int *poison) {
  if (*poison) return;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) y[r] = Scalar<T>::zero();
}

template <class T, int LPC>
__global__ void __launch_bounds__(256)
    spmv_csc_scatter_kernel(int64_t n_cols, const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowind,
                            const T *__restrict__ vals, const T *__restrict__ x, T *__restrict__ y,
                            const int *poison) {
  if (*poison) return;
  const int sub = threadIdx.x & (LPC - 1);
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPC;
  const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / LPC;
  for (int64_t col = group; col < n_cols; col += ngroups) {
    const int64_t s = __ldg(colptr + col), e = __ldg(colptr + col + 1);
    const T xv = ld_ro<T>(x + col);
    for (int64_t i = s + sub; i < e; i += LPC) {
      const T p = Scalar<T>::mul(ld_ro<T>(vals + i), xv);
      double *dst = reinterpret_cast<double *>(y + __ldg(rowind + i));
      if (Scalar<T>::is_complex) {
        const double2 *pp = reinterpret_cast<const double2 *>(&p);
        red_add(dst, pp->x);
        red_add(dst + 1, pp->y);
      } else {
        red_add(dst, *reinterpret_cast<const double *>(&p));
      }
    }
  }
}

}  // namespace b2a
