// kernels_cgs.cuh - iterated classical Gram-Schmidt on the n x j Krylov panel.
//
// Replaces the BLAS-2 calls of orthogonalize! / reinitialize!
// (src/expansion.jl:37-48 and :81-107):
//     rnorm = norm(v);  h = Vj' v;  v -= Vj h;  wnorm = norm(v)
//     [if wnorm < eta*rnorm:  c = Vj' v;  v -= Vj c;  h += c;  wnorm = norm(v)]
//     breakdown ? H[j+1,j] = 0 : (H[j+1,j] = wnorm; v /= wnorm)
//
// Layout: V is column-major n_loc x (maxdim+1), leading dimension ld (multiple of 16
// elements, 128-byte aligned columns); rows [n_loc, ld) of every column are kept ZERO so
// that 128-bit loads may run over the tail.
//
// All three kernels are HBM-bound streaming kernels:
//   cgs_dots   reads the panel once + v once        -> (j+1) n s bytes
//   cgs_update reads the panel once, v once, writes v -> (j+2) n s bytes
//   cgs_finish reads + writes v                      -> 2 n s bytes
// Reductions are two-stage with a fixed summation order (no FP64 atomics) so results are
// bit-reproducible for a given GPU count.
#pragma once

#include "device_common.cuh"
#include "peer_comm.cuh"

namespace b2a {

constexpr int kCgsThreads = 256;
constexpr int kCgsWarps = kCgsThreads / 32;

template <class T> __device__ __forceinline__ T dot_acc(double2 a, double2 v, T acc);
template <> __device__ __forceinline__ double dot_acc<double>(double2 a, double2 v, double acc) {
  return fma(a.y, v.y, fma(a.x, v.x, acc));
}
template <> __device__ __forceinline__ cdouble dot_acc<cdouble>(double2 a, double2 v, cdouble acc) {
  return Scalar<cdouble>::fma_conj(a, v, acc);
}

template <class T> __device__ __forceinline__ double2 axpy_neg(double2 a, T h, double2 x);
template <> __device__ __forceinline__ double2 axpy_neg<double>(double2 a, double h, double2 x) {
  x.x = fma(-a.x, h, x.x);
  x.y = fma(-a.y, h, x.y);
  return x;
}
template <> __device__ __forceinline__ double2 axpy_neg<cdouble>(double2 a, cdouble h, double2 x) {
  return Scalar<cdouble>::fnma(a, h, x);
}

__device__ __forceinline__ double vec_abs2(double2 x) { return fma(x.x, x.x, x.y * x.y); }

// Gate shared by the kernels of the conditional second pass: run only if the DGKS test
// of src/expansion.jl:91 fired.  rsq / w1sq are the all-reduced squared norms.
__device__ __forceinline__ bool dgks_fired(const double *rsq, const double *w1sq) {
  return sqrt(*w1sq) < kEta * sqrt(*rsq);
}

// ---------------------------------------------------------------------------------
// cgs_dots:  h[c] = sum_r conj(V[r,c]) v[r]  (c < ncols),  nrm2 = sum_r |v[r]|^2
// Warp w of each CTA owns columns [w*CPW, (w+1)*CPW); lanes own rows and accumulate
// privately over the CTA's whole row range; one shuffle reduction at the end.  The CTA
// that finishes last sums the per-CTA partials in a fixed order.
// ---------------------------------------------------------------------------------
template <class T, int CPW, int U>
__global__ void __launch_bounds__(kCgsThreads)
    cgs_dots_kernel(const T *__restrict__ V, int64_t ld, const T *__restrict__ v, int64_t n, int ncols,
                    int64_t rows_per_cta, T *__restrict__ partials, T *__restrict__ hout,
                    double *__restrict__ nrm2_out, unsigned int *ticket, const int *poison,
                    const double *gate_rsq, const double *gate_w1sq) {
  if (*poison) return;
  if (gate_rsq && !dgks_fired(gate_rsq, gate_w1sq)) return;

  constexpr int PV = Scalar<T>::per_vec;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = min(n, r0 + rows_per_cta);
  const int c0 = warp * CPW;

  T acc[CPW];
#pragma unroll
  for (int i = 0; i < CPW; ++i) acc[i] = Scalar<T>::zero();
  double nacc = 0.0;

  const T *colp[CPW];
#pragma unroll
  for (int i = 0; i < CPW; ++i) colp[i] = V + (int64_t)min(c0 + i, max(ncols - 1, 0)) * ld;

  for (int64_t r = r0 + (int64_t)lane * PV; r < r1; r += (int64_t)32 * PV * U) {
    double2 vv[U];
    double2 a[CPW][U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t rr = r + (int64_t)u * 32 * PV;
      vv[u] = rr < r1 ? *reinterpret_cast<const double2 *>(v + rr) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int i = 0; i < CPW; ++i) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t rr = r + (int64_t)u * 32 * PV;
        a[i][u] = (c0 + i < ncols && rr < r1) ? ldg_stream(reinterpret_cast<const double2 *>(colp[i] + rr))
                                              : make_double2(0.0, 0.0);
      }
    }
#pragma unroll
    for (int i = 0; i < CPW; ++i) {
#pragma unroll
      for (int u = 0; u < U; ++u) acc[i] = dot_acc<T>(a[i][u], vv[u], acc[i]);
    }
    if (warp == 0) {
#pragma unroll
      for (int u = 0; u < U; ++u) nacc += vec_abs2(vv[u]);
    }
  }

  const int grid = gridDim.x;
#pragma unroll
  for (int i = 0; i < CPW; ++i) {
    const T s = warp_sum(acc[i]);
    if (lane == 0 && c0 + i < ncols) partials[(int64_t)(c0 + i) * grid + blockIdx.x] = s;
  }
  if (warp == 0) {
    const double s = warp_sum(nacc);
    if (lane == 0) partials[(int64_t)ncols * grid + blockIdx.x] = Scalar<T>::from_real(s);
  }

  // ---- last CTA: deterministic final reduction
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == (unsigned)grid - 1u);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int c = warp; c <= ncols; c += kCgsWarps) {
    T s = Scalar<T>::zero();
    const T *p = partials + (int64_t)c * grid;
    for (int b = lane; b < grid; b += 32) s = Scalar<T>::add(s, __ldcg(p + b));
    s = warp_sum(s);
    if (lane == 0) {
      if (c < ncols)
        hout[c] = s;
      else if (nrm2_out)
        *nrm2_out = *reinterpret_cast<const double *>(&s);
    }
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

// ---------------------------------------------------------------------------------
// cgs_update:  v[r] -= sum_c V[r,c] h[c];  nrm2 = sum_r |v[r]|^2 (after the update)
// Threads own rows (one 128-bit vector each per grid-stride step) and stream the columns
// with 8 independent loads in flight.
// ---------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(kCgsThreads)
    cgs_update_kernel(const T *__restrict__ V, int64_t ld, T *__restrict__ v, int64_t n, int ncols,
                      const T *__restrict__ h, double *__restrict__ partials, double *__restrict__ nrm2_out,
                      unsigned int *ticket, const int *poison, const double *gate_rsq,
                      const double *gate_w1sq) {
  if (*poison) return;
  if (gate_rsq && !dgks_fired(gate_rsq, gate_w1sq)) return;

  constexpr int PV = Scalar<T>::per_vec;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *hs = reinterpret_cast<T *>(smem_raw);
  for (int c = threadIdx.x; c < ncols; c += blockDim.x) hs[c] = h[c];
  __syncthreads();

  double nacc = 0.0;
  const int64_t nvec = (n + PV - 1) / PV;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t iv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; iv < nvec; iv += stride) {
    const int64_t r = iv * PV;
    double2 x = *reinterpret_cast<const double2 *>(v + r);
    int c = 0;
    for (; c + 8 <= ncols; c += 8) {
      double2 a[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] = ldg_stream(reinterpret_cast<const double2 *>(V + (int64_t)(c + k) * ld + r));
#pragma unroll
      for (int k = 0; k < 8; ++k) x = axpy_neg<T>(a[k], hs[c + k], x);
    }
    if (c < ncols) {
      double2 a[8];
#pragma unroll
      for (int k = 0; k < 8; ++k)
        a[k] = (c + k < ncols) ? ldg_stream(reinterpret_cast<const double2 *>(V + (int64_t)(c + k) * ld + r))
                               : make_double2(0.0, 0.0);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (c + k < ncols) x = axpy_neg<T>(a[k], hs[c + k], x);
    }
    if (PV == 2 && r + 1 >= n) x.y = 0.0;  // keep the padding row zero
    *reinterpret_cast<double2 *>(v + r) = x;
    nacc += vec_abs2(x);
  }

  // block reduction of the squared norm, then last-CTA final sum
  __shared__ double wsum[kCgsWarps];
  __shared__ int is_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  nacc = warp_sum(nacc);
  if (lane == 0) wsum[warp] = nacc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kCgsWarps; ++w) s += wsum[w];
    partials[blockIdx.x] = s;
    __threadfence();
    is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1u);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (warp == 0) {
    double s = 0.0;
    for (int b = lane; b < (int)gridDim.x; b += 32) s += __ldcg(partials + b);
    s = warp_sum(s);
    if (lane == 0) {
      *nrm2_out = s;
      *ticket = 0u;
    }
  }
}

// ---------------------------------------------------------------------------------
// cgs_finish: the tail of orthogonalize! / reinitialize! (src/expansion.jl:49-58,91-107)
//   decides second pass / breakdown from the (all-reduced) squared norms, writes column j
//   of the device copy of H (h1 [+ h2], then wnorm or 0), and normalises v (v ./= wnorm).
// mode: 0 = Arnoldi step (writes H, raises `poison = step` on breakdown)
//       1 = re-seed (no H, no poison; on failure v is left as is - its Bool is ignored,
//           src/expansion.jl:128)
//       2 = plain normalisation (reinitialize! with j == 0, src/expansion.jl:27-30)
// ---------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256)
    cgs_finish_kernel(T *__restrict__ v, int64_t n, int j, const T *__restrict__ h1, const T *__restrict__ h2,
                      const double *rsq_p, const double *w1sq_p, const double *w2sq_p, T *__restrict__ Hcol,
                      int *info_col, SweepState *state, int step, int mode, const __grid_constant__ PeerView pv,
                      int64_t row_offset, int push) {
  pdl_wait();
  if (state->poison) return;
  constexpr int PV = Scalar<T>::per_vec;
  double rnorm = sqrt(*rsq_p);
  double wnorm;
  bool second = false;
  if (mode == 2) {
    wnorm = rnorm;
  } else {
    wnorm = sqrt(*w1sq_p);
    if (wnorm < kEta * rnorm) {  // expansion.jl:91 (strict)
      second = true;
      rnorm = wnorm;
      wnorm = sqrt(*w2sq_p);
    }
  }
  const bool breakdown = (mode != 2) && (wnorm <= kEta * rnorm);  // expansion.jl:99 (non-strict)

  if (blockIdx.x == 0) {
    if (mode == 0) {
      for (int c = threadIdx.x; c < j; c += blockDim.x)
        Hcol[c] = second ? Scalar<T>::add(h1[c], h2[c]) : h1[c];  // expansion.jl:95
      if (threadIdx.x == 0) {
        Hcol[j] = Scalar<T>::from_real(breakdown ? 0.0 : wnorm);  // expansion.jl:100,104
        info_col[0] = (second ? 1 : 0) | (breakdown ? 2 : 0);
        if (second) atomicAdd(&state->second_passes, 1ull);
      }
    } else if (threadIdx.x == 0 && info_col) {
      info_col[0] = (second ? 1 : 0) | (breakdown ? 2 : 0);
    }
  }
  if (breakdown) {
    // every CTA takes this branch (same scalars); only mode 0 raises the flag.  CTAs that
    // start after the store see `poison` above and return - equivalent, nothing is scaled.
    if (mode == 0 && blockIdx.x == 0 && threadIdx.x == 0) state->poison = step;
    return;
  }
  const int64_t nvec = (n + PV - 1) / PV;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const bool do_push = push && pv.P > 1;
  const bool vec_ok = PV == 1 || (row_offset & 1) == 0;
  constexpr int U = 4;  // vectors in flight per thread
  for (int64_t iv0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; iv0 < nvec; iv0 += stride * U) {
    double2 x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t iv = iv0 + u * stride;
      x[u] = iv < nvec ? *reinterpret_cast<const double2 *>(v + iv * PV) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      x[u].x /= wnorm;  // v ./= wnorm (expansion.jl:106): a true division, like the reference
      x[u].y /= wnorm;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t iv = iv0 + u * stride;
      if (iv < nvec) *reinterpret_cast<double2 *>(v + iv * PV) = x[u];
    }
    if (do_push) {
      // fused x-exchange: this rank's slice of the next mat-vec input goes straight into every
      // rank's x buffer over NVLink (one 16-byte store per lane when the row offset allows it)
      for (int p = 0; p < pv.P; ++p) {
        T *pxb = reinterpret_cast<T *>(pv.peer[p] + pv.off_x) + row_offset;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t iv = iv0 + u * stride;
          if (iv >= nvec) continue;
          const int64_t r = iv * PV;
          if (vec_ok && (PV == 1 || r + 1 < n)) {
            *reinterpret_cast<double2 *>(pxb + r) = x[u];
          } else {
            reinterpret_cast<double *>(pxb + r)[0] = x[u].x;
            if (r + 1 < n) reinterpret_cast<double *>(pxb + r)[1] = x[u].y;
          }
        }
      }
    }
  }
  if (do_push) {
    __shared__ int last_cta;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last_cta = (atomicAdd(&state->ticket[5], 1u) == gridDim.x - 1u);
    __syncthreads();
    if (last_cta && threadIdx.x == 0) {
      state->ticket[5] = 0u;
      peer_x_publish(pv);
    }
  }
}

// v ./= alpha (fine-grained API, b2a_ws_scal_div)
template <class T>
__global__ void __launch_bounds__(256) scal_div_kernel(T *__restrict__ v, int64_t n, double alpha) {
  constexpr int PV = Scalar<T>::per_vec;
  const int64_t nvec = (n + PV - 1) / PV;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t iv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; iv < nvec; iv += stride) {
    double2 x = *reinterpret_cast<const double2 *>(v + iv * PV);
    x.x /= alpha;
    x.y /= alpha;
    *reinterpret_cast<double2 *>(v + iv * PV) = x;
  }
}

// Explicit push of a column that no fused finish kernel produced (start of a sweep, after a
// rotation / re-seed / set_col).
template <class T>
__global__ void __launch_bounds__(256)
    xpush_kernel(const T *__restrict__ v, int64_t n, SweepState *state, const __grid_constant__ PeerView pv,
                 int64_t row_offset) {
  if (state->poison) return;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
    const T x = v[r];
    for (int p = 0; p < pv.P; ++p) (reinterpret_cast<T *>(pv.peer[p] + pv.off_x) + row_offset)[r] = x;
  }
  __shared__ int last_cta;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) last_cta = (atomicAdd(&state->ticket[5], 1u) == gridDim.x - 1u);
  __syncthreads();
  if (last_cta && threadIdx.x == 0) {
    state->ticket[5] = 0u;
    peer_x_publish(pv);
  }
}

}  // namespace b2a
