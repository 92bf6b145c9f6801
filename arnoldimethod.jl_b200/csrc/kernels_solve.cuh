// kernels_solve.cuh - device-side inner solver for shift-and-invert operators.
//
// The reference leaves spectral transformations to the user (docs/src/index.md:234-262, "Shift-and-invert with
// LinearMaps.jl": a linear map that applies inv(A) through a factorisation, handed to partialschur as the operator;
// bench/partial_schur.jl:11-35 does the same).  On the device there is no sparse LU, so the map
//     y = (A - sigma I)^{-1} x
// is a Jacobi-preconditioned CONJUGATE GRADIENT solve (A - sigma I Hermitian positive definite: the smallest modes of
// a Laplacian-type operator) built on the existing CSR mat-vec.  One iteration =
//     mat-vec            q = A p
//     cg_pq_kernel       q -= sigma p;  pq = p' q;                       last CTA: alpha = rz / pq
//     cg_update_kernel   x += alpha p;  r -= alpha q;  z = r ./ d;  rz' = r' z, rr = r' r;
//                        last CTA: beta = rz' / rz, rz = rz', done = (rr <= rtol^2 * bb)
//     cg_p_kernel        p = z + beta p
// All scalars live in device memory (CgState), so a chunk of iterations is enqueued without a host round trip; once
// `done` is raised the remaining launches of the chunk are no-ops and the host reads the flag once per chunk.
// Reductions: per-CTA partial sums in a fixed order, final sum by the CTA that finishes last - deterministic.
#pragma once

#include "device_common.cuh"

namespace b2a {

struct CgState {
  double rz_re, rz_im;      // r' z (complex in general; real for Hermitian positive definite systems)
  double alpha_re, alpha_im;
  double beta_re, beta_im;
  double rr, bb;            // ||r||^2, ||b||^2
  double rtol2;
  int done;                 // 1: converged, 2: breakdown (p' q == 0)
  int iters;
  unsigned int ticket[2];
};

constexpr int kSolveThreads = 256;

template <class T> __device__ __forceinline__ T cmul(T a, T b);
template <> __device__ __forceinline__ double cmul<double>(double a, double b) { return a * b; }
template <> __device__ __forceinline__ cdouble cmul<cdouble>(cdouble a, cdouble b) { return Scalar<cdouble>::mul(a, b); }
template <class T> __device__ __forceinline__ T make_scalar(double re, double im);
template <> __device__ __forceinline__ double make_scalar<double>(double re, double) { return re; }
template <> __device__ __forceinline__ cdouble make_scalar<cdouble>(double re, double im) { return make_double2(re, im); }
template <class T> __device__ __forceinline__ double re_of(T a);
template <> __device__ __forceinline__ double re_of<double>(double a) { return a; }
template <> __device__ __forceinline__ double re_of<cdouble>(cdouble a) { return a.x; }
template <class T> __device__ __forceinline__ double im_of(T a);
template <> __device__ __forceinline__ double im_of<double>(double) { return 0.0; }
template <> __device__ __forceinline__ double im_of<cdouble>(cdouble a) { return a.y; }
// conj(a) * b
template <class T> __device__ __forceinline__ T conj_mul(T a, T b) { return Scalar<T>::fma_conj(a, b, Scalar<T>::zero()); }
template <class T> __device__ __forceinline__ T cdiv_scalar(T a, T b);
template <> __device__ __forceinline__ double cdiv_scalar<double>(double a, double b) { return a / b; }
template <> __device__ __forceinline__ cdouble cdiv_scalar<cdouble>(cdouble a, cdouble b) {
  const double d = b.x * b.x + b.y * b.y;
  return make_double2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}

// block-wide sum of (re, im) pairs in a fixed order; result valid in thread 0
__device__ __forceinline__ double2 block_sum2(double2 v) {
  __shared__ double2 ws2[kSolveThreads / 32];
  v.x = warp_sum(v.x);
  v.y = warp_sum(v.y);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) ws2[warp] = v;
  __syncthreads();
  double2 s = make_double2(0.0, 0.0);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < kSolveThreads / 32; ++w) {
      s.x += ws2[w].x;
      s.y += ws2[w].y;
    }
  }
  __syncthreads();
  return s;
}

// d[r] = A[r, r] - sigma  (Jacobi preconditioner of the shifted operator); rows without a diagonal entry get -sigma
template <class T>
__global__ void __launch_bounds__(256)
    shifted_diag_kernel(int64_t n, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
                        const T *__restrict__ vals, int64_t row_offset, T sigma, T *__restrict__ d) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
    T acc = Scalar<T>::zero();
    for (int64_t i = rowptr[r]; i < rowptr[r + 1]; ++i)
      if ((int64_t)colind[i] == row_offset + r) acc = Scalar<T>::add(acc, vals[i]);
    d[r] = Scalar<T>::add(acc, Scalar<T>::scale(sigma, -1.0));
  }
}

// start of a solve with x0 = 0:  r = b, z = r ./ d, p = z, x = 0;  rz = r' z, bb = rr = ||b||^2
template <class T>
__global__ void __launch_bounds__(kSolveThreads)
    cg_init_kernel(int64_t n, const T *__restrict__ b, const T *__restrict__ d, T *__restrict__ x, T *__restrict__ r,
                   T *__restrict__ z, T *__restrict__ p, double2 *__restrict__ partials, CgState *st, double rtol2,
                   const int *poison) {
  if (*poison) return;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double2 acc = make_double2(0.0, 0.0);  // (re r'z, rr)
  double im = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const T bi = b[i];
    const T zi = cdiv_scalar<T>(bi, d[i]);
    x[i] = Scalar<T>::zero();
    r[i] = bi;
    z[i] = zi;
    p[i] = zi;
    const T t = conj_mul<T>(bi, zi);
    acc.x += re_of<T>(t);
    im += im_of<T>(t);
    acc.y += Scalar<T>::abs2(bi);
  }
  const double2 s = block_sum2(acc);
  const double2 s2 = block_sum2(make_double2(im, 0.0));
  __shared__ int last;
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = s;
    partials[gridDim.x + blockIdx.x] = s2;
    __threadfence();
    last = atomicAdd(&st->ticket[0], 1u) == gridDim.x - 1u;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  double2 a = make_double2(0.0, 0.0), c = a;
  for (int k = threadIdx.x; k < (int)gridDim.x; k += blockDim.x) {
    const double2 u = __ldcg(partials + k), w = __ldcg(partials + gridDim.x + k);
    a.x += u.x;
    a.y += u.y;
    c.x += w.x;
  }
  a = block_sum2(a);
  c = block_sum2(c);
  if (threadIdx.x == 0) {
    st->rz_re = a.x;
    st->rz_im = c.x;
    st->rr = st->bb = a.y;
    st->rtol2 = rtol2;
    st->done = a.y == 0.0 ? 1 : 0;  // b == 0: x = 0 is the answer
    st->iters = 0;
    st->ticket[0] = 0u;
  }
}

// q -= sigma p;  pq = p' q;  alpha = rz / pq
template <class T>
__global__ void __launch_bounds__(kSolveThreads)
    cg_pq_kernel(int64_t n, const T *__restrict__ p, T *__restrict__ q, T sigma, int has_sigma,
                 double2 *__restrict__ partials, CgState *st, const int *poison) {
  if (*poison || st->done) return;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double2 acc = make_double2(0.0, 0.0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const T pi = p[i];
    T qi = q[i];
    if (has_sigma) {
      qi = Scalar<T>::fnma(sigma, pi, qi);
      q[i] = qi;
    }
    const T t = conj_mul<T>(pi, qi);
    acc.x += re_of<T>(t);
    acc.y += im_of<T>(t);
  }
  const double2 s = block_sum2(acc);
  __shared__ int last;
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = s;
    __threadfence();
    last = atomicAdd(&st->ticket[0], 1u) == gridDim.x - 1u;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  double2 a = make_double2(0.0, 0.0);
  for (int k = threadIdx.x; k < (int)gridDim.x; k += blockDim.x) {
    const double2 u = __ldcg(partials + k);
    a.x += u.x;
    a.y += u.y;
  }
  a = block_sum2(a);
  if (threadIdx.x == 0) {
    st->ticket[0] = 0u;
    if (a.x == 0.0 && a.y == 0.0) {
      st->done = 2;
    } else {
      const cdouble al = cdiv_scalar<cdouble>(make_double2(st->rz_re, st->rz_im), a);
      st->alpha_re = al.x;
      st->alpha_im = al.y;
    }
  }
}

// x += alpha p;  r -= alpha q;  z = r ./ d;  rz' = r' z;  rr = ||r||^2;  beta = rz' / rz;  convergence test
template <class T>
__global__ void __launch_bounds__(kSolveThreads)
    cg_update_kernel(int64_t n, const T *__restrict__ p, const T *__restrict__ q, const T *__restrict__ d,
                     T *__restrict__ x, T *__restrict__ r, T *__restrict__ z, double2 *__restrict__ partials,
                     CgState *st, const int *poison) {
  if (*poison || st->done) return;
  const T alpha = make_scalar<T>(st->alpha_re, st->alpha_im);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double2 acc = make_double2(0.0, 0.0);
  double im = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    x[i] = Scalar<T>::fma_(alpha, p[i], x[i]);
    const T ri = Scalar<T>::fnma(alpha, q[i], r[i]);
    r[i] = ri;
    const T zi = cdiv_scalar<T>(ri, d[i]);
    z[i] = zi;
    const T t = conj_mul<T>(ri, zi);
    acc.x += re_of<T>(t);
    im += im_of<T>(t);
    acc.y += Scalar<T>::abs2(ri);
  }
  const double2 s = block_sum2(acc);
  const double2 s2 = block_sum2(make_double2(im, 0.0));
  __shared__ int last;
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = s;
    partials[gridDim.x + blockIdx.x] = s2;
    __threadfence();
    last = atomicAdd(&st->ticket[1], 1u) == gridDim.x - 1u;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  double2 a = make_double2(0.0, 0.0), c = a;
  for (int k = threadIdx.x; k < (int)gridDim.x; k += blockDim.x) {
    const double2 u = __ldcg(partials + k), w = __ldcg(partials + gridDim.x + k);
    a.x += u.x;
    a.y += u.y;
    c.x += w.x;
  }
  a = block_sum2(a);
  c = block_sum2(c);
  if (threadIdx.x == 0) {
    const cdouble rz_new = make_double2(a.x, c.x);
    const cdouble be = cdiv_scalar<cdouble>(rz_new, make_double2(st->rz_re, st->rz_im));
    st->beta_re = be.x;
    st->beta_im = be.y;
    st->rz_re = rz_new.x;
    st->rz_im = rz_new.y;
    st->rr = a.y;
    st->iters += 1;
    if (a.y <= st->rtol2 * st->bb) st->done = 1;
    st->ticket[1] = 0u;
  }
}

// p = z + beta p
template <class T>
__global__ void __launch_bounds__(kSolveThreads)
    cg_p_kernel(int64_t n, const T *__restrict__ z, T *__restrict__ p, const CgState *st, const int *poison) {
  if (*poison || st->done) return;
  const T beta = make_scalar<T>(st->beta_re, st->beta_im);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    p[i] = Scalar<T>::fma_(beta, p[i], z[i]);
}

}  // namespace b2a
