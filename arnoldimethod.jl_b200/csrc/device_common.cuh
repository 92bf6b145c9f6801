// device_common.cuh - scalar-type helpers shared by all kernels (Float64 / ComplexF64).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace b2a {

constexpr int kSMs = 148;          // B200: 2 dies x 74 SMs
constexpr double kEta = 0.70710678118654752440;  // sqrt(2)/2, src/expansion.jl:33,74

// ComplexF64 on the device is a double2 (re = x, im = y): one element per 128-bit load.
using cdouble = double2;

template <class T> struct Scalar;

template <> struct Scalar<double> {
  static constexpr bool is_complex = false;
  static constexpr int per_vec = 2;  // elements per 128-bit vector
  __host__ __device__ static inline double zero() { return 0.0; }
  __device__ static inline double mul(double a, double b) { return a * b; }
  // acc + conj(a) * b
  __device__ static inline double fma_conj(double a, double b, double acc) { return fma(a, b, acc); }
  // acc + a * b
  __device__ static inline double fma_(double a, double b, double acc) { return fma(a, b, acc); }
  // acc - a * b
  __device__ static inline double fnma(double a, double b, double acc) { return fma(-a, b, acc); }
  __device__ static inline double add(double a, double b) { return a + b; }
  __device__ static inline double abs2(double a) { return a * a; }
  __device__ static inline double divr(double a, double r) { return a / r; }
  __device__ static inline double scale(double a, double r) { return a * r; }
  __device__ static inline double from_real(double r) { return r; }
};

template <> struct Scalar<cdouble> {
  static constexpr bool is_complex = true;
  static constexpr int per_vec = 1;
  __host__ __device__ static inline cdouble zero() { return make_double2(0.0, 0.0); }
  __device__ static inline cdouble mul(cdouble a, cdouble b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
  }
  __device__ static inline cdouble fma_conj(cdouble a, cdouble b, cdouble acc) {
    // conj(a) * b = (a.x b.x + a.y b.y) + i (a.x b.y - a.y b.x)
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(-a.y, b.x, acc.y);
    return acc;
  }
  __device__ static inline cdouble fma_(cdouble a, cdouble b, cdouble acc) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
    return acc;
  }
  __device__ static inline cdouble fnma(cdouble a, cdouble b, cdouble acc) {
    acc.x = fma(-a.x, b.x, acc.x);
    acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(-a.x, b.y, acc.y);
    acc.y = fma(-a.y, b.x, acc.y);
    return acc;
  }
  __device__ static inline cdouble add(cdouble a, cdouble b) { return make_double2(a.x + b.x, a.y + b.y); }
  __device__ static inline double abs2(cdouble a) { return fma(a.x, a.x, a.y * a.y); }
  __device__ static inline cdouble divr(cdouble a, double r) { return make_double2(a.x / r, a.y / r); }
  __device__ static inline cdouble scale(cdouble a, double r) { return make_double2(a.x * r, a.y * r); }
  __device__ static inline cdouble from_real(double r) { return make_double2(r, 0.0); }
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ cdouble warp_sum(cdouble v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
  }
  return v;
}

// streaming 128-bit loads/stores (read-once panel data: do not pollute L1)
__device__ __forceinline__ double2 ldg_stream(const double2 *p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

// Programmatic dependent launch (PDL): kernels of a sweep are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so the launch and prologue of kernel N+1 overlap
// the tail of kernel N.  pdl_wait() must precede the first read of anything a predecessor wrote;
// pdl_trigger() lets the successor's CTAs be scheduled as soon as SMs free up.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Sweep state shared by the kernels of one workspace.  All decisions of src/expansion.jl:91,99 are taken
// on the device from (all-reduced) scalars so that a whole sweep can be enqueued without a host round trip.
struct SweepState {
  int poison;          // != 0: step index (1-based) whose orthogonalisation broke down
  int error;           // != 0: a grid barrier of the fused sweep kernel timed out (kernels_cgs_sweep.cuh)
  unsigned long long second_passes;
  unsigned int ticket[8];  // last-block-done counters (one per reduction kernel kind)
};

}  // namespace b2a
