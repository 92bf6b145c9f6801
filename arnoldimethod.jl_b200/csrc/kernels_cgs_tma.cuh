// kernels_cgs_tma.cuh - Gram-Schmidt sweeps as persistent, warp-specialised TMA pipelines.
//
// The Krylov panel is streamed exactly once per sweep through a shared-memory ring:
//   * one PRODUCER thread per CTA issues ONE 2-D tensor-map TMA per tile
//     (cp.async.bulk.tensor.2d ... mbarrier::complete_tx): the box is RT rows x (j+1) columns,
//     i.e. the panel tile AND the matching segment of v (column j sits right after the panel),
//     so the bytes in flight are bounded by the ring (up to ~200 KB per SM), not by registers
//     or by how the compiler schedules loads.  (Measured on B200: the per-SM TMA unit costs
//     ~70-100 cycles per operation, so 41 one-column bulk copies per tile top out at
//     3.5 TB/s with 1 KB copies; one boxed tensor copy per tile removes that limit.)
//   * eight CONSUMER warps wait on the stage's `full` mbarrier, reduce from shared memory and
//     release the stage through its `empty` mbarrier.
// One CTA per SM (persistent, grid = #SMs), static contiguous tile ranges per CTA, fixed-order
// two-stage reductions => bit-reproducible results.
//
// Sweeps (j panel columns, n rows, s = sizeof(T)):
//   S1  cgs_dots_tma      h = V' v, ||v||^2                       reads (j+1) n s
//   S2  cgs_update_tma    v -= V h, ||v||^2  AND, in the same pass over the tile that is
//                         already on chip, the SPECULATIVE second-pass coefficients
//                         c = V' v_new (src/expansion.jl:93)       reads (j+1) n s, writes n s
//   S3  cgs_update_tma    (gated, no speculation) v -= V c, ||v||^2   only if the DGKS test fired
// so an Arnoldi step whose DGKS test fires reads the panel 3 times instead of the 4 times of
// "dots, update, dots, update"; a step without second pass reads it twice.  S2 walks the tiles
// in REVERSE order so that the tail of the panel S1 left in the 126 MB L2 is hit first.
//
// Requires the workspace invariant: ld is a multiple of 1024 rows and rows [n, ld) of every
// column are zero, so every tile is loaded full-size without masks.
#pragma once

#include <cuda.h>  // CUtensorMap (types only; the encoder is resolved at run time)

#include "device_common.cuh"
#include "kernels_cgs.cuh"
#include "peer_comm.cuh"

namespace b2a {

constexpr int kTmaConsumerWarps = 8;
constexpr int kTmaThreads = (kTmaConsumerWarps + 1) * 32;
constexpr int kTmaMaxStages = 8;
constexpr int kTmaMaxCols = 64;  // panel columns per launch (8 per consumer warp)

// ---- mbarrier / bulk-copy PTX -------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// 2-D tensor-map TMA global -> shared (SASS: UTMALDG); c0 = inner (row) coordinate, c1 = column
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tmap, int c0, int c1, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void consumer_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kTmaConsumerWarps * 32) : "memory"); }

struct TmaGeom {
  int RT;             // rows per tile (power of two, divides 1024)
  int stages;         // ring depth
  int tiles_per_cta;  // contiguous tiles per CTA
  int ntiles;         // ceil(n / RT)
  int reverse;        // walk tiles from the end of the vector
};

// tile index handled at local step `t` of CTA `b`
__device__ __forceinline__ int tile_of(const TmaGeom &g, int b, int t) {
  const int first = b * g.tiles_per_cta;
  const int idx = first + t;
  return g.reverse ? (g.ntiles - 1 - idx) : idx;
}
__device__ __forceinline__ int my_tile_count(const TmaGeom &g, int b) {
  const int first = b * g.tiles_per_cta;
  return max(0, min(g.tiles_per_cta, g.ntiles - first));
}

// producer: stream tiles [V[:, 0:ncols) | v] = columns 0..ncols of the tensor map into the ring
template <class T>
__device__ __forceinline__ void tma_producer(const CUtensorMap *tmap, int ncols, const TmaGeom &g, T *ring,
                                             uint64_t *full, uint64_t *empty, int lane) {
  if (lane != 0) return;
  const int ntl = my_tile_count(g, blockIdx.x);
  const size_t stage_elems = (size_t)(ncols + 1) * g.RT;
  const uint32_t stage_bytes = (uint32_t)(stage_elems * sizeof(T));
  constexpr int kInnerPerRow = sizeof(T) / sizeof(double);  // ComplexF64 rows are 2 doubles in the map
  for (int t = 0; t < ntl; ++t) {
    const int s = t % g.stages;
    const uint32_t round = (uint32_t)(t / g.stages);
    mbar_wait(&empty[s], (round & 1u) ^ 1u);  // first round passes immediately
    mbar_expect_tx(&full[s], stage_bytes);
    const int r0 = tile_of(g, blockIdx.x, t) * g.RT;
    tma_load_2d(ring + (size_t)s * stage_elems, tmap, r0 * kInnerPerRow, 0, &full[s]);
  }
}

// consumer: acc[i] += sum_r conj(tile[col_i][r]) * x[r] over one tile (x = last column of the stage)
template <class T, int CPW>
__device__ __forceinline__ void tile_dots(const T *tile, const T *xt, int RT, int ncols, int warp, int lane,
                                          T (&acc)[CPW], double &nacc, bool want_norm) {
  constexpr int PV = Scalar<T>::per_vec;
  for (int rr = lane * PV; rr < RT; rr += 32 * PV) {
    const double2 xv = *reinterpret_cast<const double2 *>(xt + rr);
#pragma unroll
    for (int i = 0; i < CPW; ++i) {
      const int c = warp + i * kTmaConsumerWarps;
      if (c < ncols) {
        const double2 a = *reinterpret_cast<const double2 *>(tile + (size_t)c * RT + rr);
        acc[i] = dot_acc<T>(a, xv, acc[i]);
      }
    }
    if (want_norm) nacc += vec_abs2(xv);
  }
}

// write per-CTA partial sums, then the last CTA reduces them in a fixed order
template <class T, int CPW>
__device__ __forceinline__ void publish_and_reduce(T (&acc)[CPW], double nacc, bool have_cols, bool have_norm, int ncols,
                                                   int warp, int lane, T *partials, T *hout, double *nrm2_out,
                                                   unsigned int *ticket, int *is_last_smem, const PeerView &pv) {
  const int grid = gridDim.x;
  if (warp < kTmaConsumerWarps) {
    if (have_cols) {
#pragma unroll
      for (int i = 0; i < CPW; ++i) {
        const int c = warp + i * kTmaConsumerWarps;
        const T s = warp_sum(acc[i]);
        if (lane == 0 && c < ncols) partials[(int64_t)c * grid + blockIdx.x] = s;
      }
    }
    if (have_norm) {
      // nacc: warp 0 holds the norm in dots mode; in update mode every consumer warp holds a
      // partial over its own rows -> the caller has already combined them into warp 0
      if (warp == 0) {
        const double s = warp_sum(nacc);
        if (lane == 0) partials[(int64_t)ncols * grid + blockIdx.x] = Scalar<T>::from_real(s);
      }
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) *is_last_smem = (atomicAdd(ticket, 1u) == (unsigned)grid - 1u);
  __syncthreads();
  if (!*is_last_smem) return;
  __threadfence();
  const int c_lo = have_cols ? 0 : ncols;
  const int c_hi = have_norm ? ncols : ncols - 1;
  for (int c = c_lo + warp; c <= c_hi; c += kTmaConsumerWarps + 1) {
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
  if (pv.P > 1) {
    // fused cross-GPU all-reduce of [h | nrm2] (contiguous by construction) over NVLink peer memory
    __syncthreads();
    if (warp == 0) {
      double *vals = have_cols ? reinterpret_cast<double *>(hout) : nrm2_out;
      const int cnt = (have_cols ? ncols * (int)(sizeof(T) / sizeof(double)) : 0) + (have_norm && nrm2_out ? 1 : 0);
      if (vals && cnt > 0) peer_allreduce_warp(pv, vals, cnt);
    }
  }
}

struct TmaSmem {
  uint64_t full[kTmaMaxStages];
  uint64_t empty[kTmaMaxStages];
  double wnorm[kTmaConsumerWarps];
  int is_last;
  int pad;
};

// ---------------------------------------------------------------------------------------
// S1: h[c] = sum_r conj(V[r,c]) v[r], nrm2 = ||v||^2
// ---------------------------------------------------------------------------------------
template <class T, int CPW>
__global__ void __launch_bounds__(kTmaThreads, 1)
    cgs_dots_tma_kernel(const __grid_constant__ CUtensorMap tmap, int ncols, TmaGeom g,
                        T *__restrict__ partials, T *__restrict__ hout, double *__restrict__ nrm2_out,
                        unsigned int *ticket, const int *poison, const double *gate_rsq, const double *gate_w1sq,
                        const __grid_constant__ PeerView pv) {
  extern __shared__ __align__(128) unsigned char tma_smem_raw[];
  TmaSmem *sm = reinterpret_cast<TmaSmem *>(tma_smem_raw);
  T *ring = reinterpret_cast<T *>(tma_smem_raw + 256);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // prologue that does not depend on the previous kernel: overlaps its tail under PDL
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap);
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(&sm->full[s], 1);
      mbar_init(&sm->empty[s], kTmaConsumerWarps);
    }
    mbar_fence_init();
  }
  pdl_wait();
  if (*poison) return;
  if (gate_rsq && !dgks_fired(gate_rsq, gate_w1sq)) return;
  __syncthreads();

  T acc[CPW];
#pragma unroll
  for (int i = 0; i < CPW; ++i) acc[i] = Scalar<T>::zero();
  double nacc = 0.0;

  if (warp == kTmaConsumerWarps) {
    tma_producer<T>(&tmap, ncols, g, ring, sm->full, sm->empty, lane);
  } else {
    const int ntl = my_tile_count(g, blockIdx.x);
    const size_t stage_elems = (size_t)(ncols + 1) * g.RT;
    for (int t = 0; t < ntl; ++t) {
      const int s = t % g.stages;
      mbar_wait(&sm->full[s], (uint32_t)(t / g.stages) & 1u);
      const T *tile = ring + (size_t)s * stage_elems;
      tile_dots<T, CPW>(tile, tile + (size_t)ncols * g.RT, g.RT, ncols, warp, lane, acc, nacc, warp == 0);
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm->empty[s]);
    }
  }
  pdl_trigger();  // main loop done: let the next kernel's launch overlap the final reduction
  publish_and_reduce<T, CPW>(acc, nacc, true, true, ncols, warp, lane, partials, hout, nrm2_out, ticket, &sm->is_last, pv);
}

// ---------------------------------------------------------------------------------------
// S2 / S3: v[r] -= sum_c V[r,c] h[c]; nrm2 = ||v_new||^2; if SPEC also cout[c] = V[:,c]' v_new
// Phase 1: each consumer warp owns 32-row slabs of the tile (lane = row), walks all columns
//          with h broadcast from shared memory, stores the new v to global AND back into the
//          stage.  Phase 2 (SPEC): the warp-owns-columns dot of S1 on the updated tile.
// ---------------------------------------------------------------------------------------
template <class T, int CPW, bool SPEC>
__global__ void __launch_bounds__(kTmaThreads, 1)
    cgs_update_tma_kernel(const __grid_constant__ CUtensorMap tmap, T *__restrict__ v, int ncols, TmaGeom g,
                          const T *__restrict__ h, T *__restrict__ partials, T *__restrict__ cout,
                          double *__restrict__ nrm2_out, unsigned int *ticket, const int *poison,
                          const double *gate_rsq, const double *gate_w1sq, const __grid_constant__ PeerView pv) {
  extern __shared__ __align__(128) unsigned char tma_smem_raw[];
  TmaSmem *sm = reinterpret_cast<TmaSmem *>(tma_smem_raw);
  T *hs = reinterpret_cast<T *>(tma_smem_raw + 256);       // kTmaMaxCols coefficients
  T *ring = hs + kTmaMaxCols;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap);
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(&sm->full[s], 1);
      mbar_init(&sm->empty[s], kTmaConsumerWarps);
    }
    mbar_fence_init();
  }
  pdl_wait();
  if (*poison) return;
  if (gate_rsq && !dgks_fired(gate_rsq, gate_w1sq)) return;
  for (int c = threadIdx.x; c < ncols; c += blockDim.x) hs[c] = h[c];
  __syncthreads();

  T acc[CPW];
#pragma unroll
  for (int i = 0; i < CPW; ++i) acc[i] = Scalar<T>::zero();
  double nacc = 0.0;

  if (warp == kTmaConsumerWarps) {
    tma_producer<T>(&tmap, ncols, g, ring, sm->full, sm->empty, lane);
  } else {
    const int ntl = my_tile_count(g, blockIdx.x);
    const size_t stage_elems = (size_t)(ncols + 1) * g.RT;
    for (int t = 0; t < ntl; ++t) {
      const int s = t % g.stages;
      mbar_wait(&sm->full[s], (uint32_t)(t / g.stages) & 1u);
      T *tile = ring + (size_t)s * stage_elems;
      T *xt = tile + (size_t)ncols * g.RT;
      const int64_t r0 = (int64_t)tile_of(g, blockIdx.x, t) * g.RT;
      // ---- phase 1: lane = row
      for (int rr = warp * 32 + lane; rr < g.RT; rr += kTmaConsumerWarps * 32) {
        T x0 = xt[rr];
        T x1 = Scalar<T>::zero();  // two accumulation chains for ILP
        int c = 0;
        for (; c + 2 <= ncols; c += 2) {
          x0 = Scalar<T>::fnma(tile[(size_t)c * g.RT + rr], hs[c], x0);
          x1 = Scalar<T>::fnma(tile[(size_t)(c + 1) * g.RT + rr], hs[c + 1], x1);
        }
        if (c < ncols) x0 = Scalar<T>::fnma(tile[(size_t)c * g.RT + rr], hs[c], x0);
        x0 = Scalar<T>::add(x0, x1);
        v[r0 + rr] = x0;
        if (SPEC) xt[rr] = x0;
        nacc += Scalar<T>::abs2(x0);
      }
      if (SPEC) {
        consumer_bar_sync();  // the updated x tile is complete
        double dummy = 0.0;
        tile_dots<T, CPW>(tile, xt, g.RT, ncols, warp, lane, acc, dummy, false);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm->empty[s]);
    }
    // combine the per-warp row partials of ||v_new||^2 in warp order (deterministic)
    nacc = warp_sum(nacc);
    if (lane == 0) sm->wnorm[warp] = nacc;
    consumer_bar_sync();
    if (warp == 0) {
      nacc = 0.0;
      if (lane == 0) {
#pragma unroll
        for (int w = 0; w < kTmaConsumerWarps; ++w) nacc += sm->wnorm[w];
      }
    }
  }
  pdl_trigger();
  publish_and_reduce<T, CPW>(acc, nacc, SPEC, true, ncols, warp, lane, partials, cout, nrm2_out, ticket, &sm->is_last, pv);
}

}  // namespace b2a
