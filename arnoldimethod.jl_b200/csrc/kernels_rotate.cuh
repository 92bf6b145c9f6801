// kernels_rotate.cuh - the Krylov-Schur restart's change of basis, in place.
//
// Replaces src/run.jl:363-365 (and :382-383):
//     V_tmp[:, purge:k] = V[:, purge:maxdim] * Q[purge:maxdim, purge:k]
//     V[:, purge:k] <- V_tmp[:, purge:k];   V[:, k+1] <- V[:, maxdim+1]
// The reference needs V_tmp because BLAS gemm cannot alias.  Here one CTA owns a tile of
// rows: it stages V[rows, purge:maxdim+1] in shared memory BEFORE storing anything, so the
// product is written back over its own input and V_tmp does not exist (half the memory,
// B_rot = n s [(m-purge+1) + (k-purge+1) + 2] bytes instead of the reference's
// (m-purge+1) + 3 (k-purge+1) + 2 columns of traffic).
//
// tcgen05 has no FP64 MMA kind (kinds: tf32/f16/bf16/i8/f8f6f4/mx*), so the Float64 /
// ComplexF64 contraction runs on the FP64 FMA pipe; at the BASELINE shapes
// (K <= 60 inputs, N <= 45 outputs) it is HBM-bound (AI ~ 4 flop/B), see DESIGN.md.
#pragma once

#include "device_common.cuh"

namespace b2a {

// Thread mapping: row = tid % R, output group = tid / R; each thread produces its outputs
// four at a time (one tile value feeds 4 FMAs; Q is read as broadcast vectors).
template <class T, int R>
__global__ void __launch_bounds__(256)
    rotate_basis_kernel(T *__restrict__ V, int64_t ld, int64_t n, int col0, int K, int N,
                        const T *__restrict__ Qd /* K x N, column-major, ld = K */, int move_src,
                        int move_dst) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Npad = (N + 3) & ~3;
  T *tile = reinterpret_cast<T *>(smem_raw);  // (K+1) x R, tile[c * R + row]
  T *Qs = tile + (size_t)(K + 1) * R;         // K x Npad, Qs[c * Npad + o]

  const int64_t r0 = (int64_t)blockIdx.x * R;
  const int tid = threadIdx.x;

  for (int idx = tid; idx < K * Npad; idx += 256) {
    const int c = idx / Npad, o = idx % Npad;
    Qs[idx] = o < N ? Qd[(size_t)o * K + c] : Scalar<T>::zero();
  }
  const int ncopy = (move_dst >= 0) ? K + 1 : K;
  for (int idx = tid; idx < ncopy * R; idx += 256) {
    const int c = idx / R, row = idx % R;
    const int64_t gr = r0 + row;
    const int src = (c == K) ? move_src : col0 + c;
    tile[idx] = gr < n ? V[(int64_t)src * ld + gr] : Scalar<T>::zero();
  }
  __syncthreads();

  constexpr int G = 256 / R;
  const int row = tid % R, og = tid / R;
  const int64_t gr = r0 + row;
  for (int ob = og * 4; ob < N; ob += G * 4) {
    T acc0 = Scalar<T>::zero(), acc1 = acc0, acc2 = acc0, acc3 = acc0;
    const T *q = Qs + ob;
#pragma unroll 4
    for (int c = 0; c < K; ++c) {
      const T vv = tile[c * R + row];
      acc0 = Scalar<T>::fma_(vv, q[c * Npad + 0], acc0);
      acc1 = Scalar<T>::fma_(vv, q[c * Npad + 1], acc1);
      acc2 = Scalar<T>::fma_(vv, q[c * Npad + 2], acc2);
      acc3 = Scalar<T>::fma_(vv, q[c * Npad + 3], acc3);
    }
    if (gr < n) {
      T *out = V + (int64_t)(col0 + ob) * ld + gr;
      out[0] = acc0;
      if (ob + 1 < N) out[ld] = acc1;
      if (ob + 2 < N) out[2 * ld] = acc2;
      if (ob + 3 < N) out[3 * ld] = acc3;
    }
  }
  if (move_dst >= 0 && og == 0 && gr < n) V[(int64_t)move_dst * ld + gr] = tile[K * R + row];
}


// ---------------------------------------------------------------------------------------------
// Register-blocked variant (default): tile of 128 rows, 256 threads = 64 row pairs x 4 output groups.
// Each thread owns TWO rows and NB outputs (NB = outputs per group, 1..8, template) and keeps the
// 2 x NB accumulators in registers: per panel column one vector load of the two tile values and
// NB broadcast loads of Q feed 2*NB FMAs (the 1-row x 4-output kernel above needs 3 shared-memory
// loads per 4 FMAs and is shared-memory bound).  Outputs are split evenly over the four groups
// (per = ceil(N/4), in `nchunks` chunks of NB), so no group idles.  Rows past n are the zero padding
// of the workspace (ld is a multiple of 1024), so neither loads nor stores need masks.
// ---------------------------------------------------------------------------------------------
constexpr int kRot2Rows = 128;

template <class T, int NB>
__global__ void __launch_bounds__(256)
    rotate_basis2_kernel(T *__restrict__ V, int64_t ld, int col0, int K, int N,
                         const T *__restrict__ Qd /* K x N, column-major, ld = K */, int move_src, int move_dst,
                         int per, int nchunks) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int R = kRot2Rows;
  T *tile = reinterpret_cast<T *>(smem_raw);   // (K+1) x R
  T *Qs = tile + (size_t)(K + 1) * R;          // [4 groups][nchunks][K][8]
  const int tid = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * R;

  const int qtotal = 4 * nchunks * K * 8;
  for (int idx = tid; idx < qtotal; idx += 256) {
    const int i = idx & 7;
    const int c = (idx >> 3) % K;
    const int ch = ((idx >> 3) / K) % nchunks;
    const int g = (idx >> 3) / (K * nchunks);
    const int o = g * per + ch * NB + i;
    const bool ok = i < NB && o < N && o < (g + 1) * per;
    Qs[idx] = ok ? Qd[(size_t)o * K + c] : Scalar<T>::zero();
  }
  const int ncopy = (move_dst >= 0) ? K + 1 : K;
  for (int idx = tid; idx < ncopy * R; idx += 256) {
    const int c = idx / R, row = idx % R;
    const int src = (c == K) ? move_src : col0 + c;
    tile[idx] = V[(int64_t)src * ld + r0 + row];
  }
  __syncthreads();

  const int rp = tid & 63, g = tid >> 6;
  const int row = 2 * rp;
  for (int ch = 0; ch < nchunks; ++ch) {
    T acc0[NB], acc1[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) acc0[i] = acc1[i] = Scalar<T>::zero();
    const T *q = Qs + (size_t)(g * nchunks + ch) * K * 8;
#pragma unroll 2
    for (int c = 0; c < K; ++c) {
      const T v0 = tile[c * R + row], v1 = tile[c * R + row + 1];
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        const T qi = q[c * 8 + i];
        acc0[i] = Scalar<T>::fma_(v0, qi, acc0[i]);
        acc1[i] = Scalar<T>::fma_(v1, qi, acc1[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int o = g * per + ch * NB + i;
      if (o < N && o < (g + 1) * per) {
        T *out = V + (int64_t)(col0 + o) * ld + r0 + row;
        out[0] = acc0[i];
        out[1] = acc1[i];
      }
    }
  }
  if (move_dst >= 0 && g == 0) {
    T *out = V + (int64_t)move_dst * ld + r0 + row;
    out[0] = tile[K * R + row];
    out[1] = tile[K * R + row + 1];
  }
}

// X[:, 0:N) = V[:, 0:K) * Y   with Y complex (partialeigen, src/eigvals.jl:94); out of place.
template <class T>
__global__ void __launch_bounds__(256)
    basis_times_kernel(const T *__restrict__ V, int64_t ld, int64_t n, int K, int N,
                       const cdouble *__restrict__ Y /* K x N col-major */, cdouble *__restrict__ X, int64_t ldx) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
    for (int o = 0; o < N; ++o) {
      cdouble acc = make_double2(0.0, 0.0);
      for (int c = 0; c < K; ++c) {
        const T vv = V[(int64_t)c * ld + r];
        const cdouble yy = __ldg(Y + (size_t)o * K + c);
        if (Scalar<T>::is_complex) {
          acc = Scalar<cdouble>::fma_(*reinterpret_cast<const cdouble *>(&vv), yy, acc);
        } else {
          const double vr = *reinterpret_cast<const double *>(&vv);
          acc.x = fma(vr, yy.x, acc.x);
          acc.y = fma(vr, yy.y, acc.y);
        }
      }
      X[(int64_t)o * ldx + r] = acc;
    }
  }
}

// ---------------------------------------------------------------------------------
// Counter-based uniform fill for reinitialize! (`rand!`, src/expansion.jl:21): keyed by the
// GLOBAL row so that the vector does not depend on how rows are sharded over GPUs.
// ---------------------------------------------------------------------------------
__host__ __device__ inline uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ inline double unit_uniform(uint64_t key, uint64_t index) {
  const uint64_t bits = splitmix64(key + index * 0xD1342543DE82EF95ull);
  return (double)(bits >> 11) * (1.0 / 9007199254740992.0);  // [0, 1)
}

template <class T>
__global__ void __launch_bounds__(256)
    fill_uniform_kernel(T *__restrict__ v, int64_t n, int64_t row_offset, uint64_t key, const int *poison) {
  if (*poison) return;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
    const uint64_t g = (uint64_t)(row_offset + r);
    if (Scalar<T>::is_complex) {
      cdouble z = make_double2(unit_uniform(key, 2 * g), unit_uniform(key, 2 * g + 1));
      *reinterpret_cast<cdouble *>(v + r) = z;
    } else {
      *reinterpret_cast<double *>(v + r) = unit_uniform(key, 2 * g);
    }
  }
}

}  // namespace b2a
