// kernels_cgs_sweep.cuh - one whole orthogonalisation (src/expansion.jl:69-109) as ONE persistent kernel.
//
// The four kernels of kernels_cgs_tma.cuh / kernels_cgs.cuh (S1 dots, S2 update + speculative dots,
// gated S3 update, finish) cost a kernel boundary each: drain, last-CTA reduction, launch, pipeline
// refill - about 8-10 us per boundary on B200, i.e. ~17 % of an Arnoldi step at n = 1e6.  This kernel
// runs the same four phases back to back inside one persistent grid (one CTA per SM):
//
//   P1  h = V' v, ||v||^2            tiles forward      -> grid barrier A (last CTA reduces, all-reduce)
//   P2  v -= V h, ||v||^2, c = V' v  tiles BACKWARD     -> grid barrier B
//   P3  v -= V c, ||v||^2            tiles forward, only if the DGKS test fired (expansion.jl:91)
//                                                        -> grid barrier C
//   P4  H[:, j], breakdown test, v ./= wnorm (expansion.jl:95-107), x push to the peers
//
// and keeps the shared-memory ring ALIVE across the phases: local tile l always lives in ring slot
// l % stages, so when the walking direction flips, the `stages` tiles the previous phase touched last
// are still on chip and are re-used without a reload (no pipeline bubble after a barrier, and
// stages x tile bytes per SM less HBM traffic per phase change).  A grid barrier is a ticket + a
// release flag (st.release.gpu / ld.acquire.gpu on a monotone epoch); all CTAs are co-resident by
// construction (grid <= #SMs, one CTA per SM by shared memory).  A spin limit turns a lost barrier
// into an error flag instead of a hang.
//
// Determinism: static contiguous tile ranges per CTA, fixed-order two-stage reductions, as in the
// unfused kernels (bit-reproducible per GPU count).
#pragma once

#include "kernels_cgs_tma.cuh"

namespace b2a {

constexpr int kSweepTraceSlots = 8;  // start, P1 end, A done, P2 end, B done, P3 end, C done, end
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
constexpr unsigned long long kGridSpinLimit = 1ull << 24;  // ~10 s of polling; then flag an error

__device__ __forceinline__ void st_release_gpu(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// generic-proxy writes (st.global of the new v) -> later async-proxy reads (TMA loads of the same rows)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <class T> __device__ __forceinline__ T ld_cg(const T *p);
template <> __device__ __forceinline__ double ld_cg<double>(const double *p) { return __ldcg(p); }
template <> __device__ __forceinline__ cdouble ld_cg<cdouble>(const cdouble *p) { return __ldcg(p); }

constexpr int kSweepPartStride = 160;  // row stride of the per-CTA partial sums (>= #SMs, multiple of 32)
constexpr int kSweepCoefs = kTmaMaxCols + 2;  // coefficient vector + its squared norm, per buffer
// Every grid barrier of a launch (A, B, C) owns its OWN region of per-CTA partial sums: in the single-GPU mode every
// CTA sums all partials itself AFTER the release flag, so a fast CTA that already stores its partials of the next
// barrier must not touch what a slow CTA is still reading (write-after-read across barriers).  Regions are re-used
// only by the next launch, which stream order (griddepcontrol.wait included) separates from this one.
constexpr int kSweepBarriers = 3;
constexpr size_t kSweepPartRegion = (size_t)(kTmaMaxCols + 1) * kSweepPartStride;  // elements per barrier region

template <class T> __host__ __device__ constexpr size_t sweep_header_bytes() {
  return 256 + (2 * kSweepCoefs * sizeof(T) + 127) / 128 * 128;  // TmaSmem | ha | hb, ring 128-byte aligned
}

// Grid-wide reduction + barrier.  Every thread of every CTA calls it.  Per-CTA partial sums of the columns
// [c_lo, ncols) and of the squared norm (row `ncols`) go to `partials`; after the barrier red[c] (shared
// memory of every CTA) holds the grid-wide sums, red[ncols] the squared norm.
//   single GPU: the CTA that arrives last publishes `epoch`; then EVERY CTA sums the partials itself, in the
//               same fixed order (no serial reduce-then-broadcast hop through one CTA);
//   multi GPU : the last CTA sums, all-reduces [h | nrm2] over NVLink peer memory (peer_comm.cuh), writes the
//               result to hout / nrm2_out and publishes `epoch`; the others fetch it from there.
template <class T, int CPW>
__device__ __forceinline__ void sweep_reduce_barrier(T (&acc)[CPW], double nacc, bool have_cols, int ncols, int warp,
                                                     int lane, T *partials, T *red, T *hout, double *nrm2_out,
                                                     unsigned int *ticket, unsigned long long *flag,
                                                     unsigned long long epoch, int *is_last_smem, int *error,
                                                     const PeerView &pv) {
  const int grid = gridDim.x;
  const int c_lo = have_cols ? 0 : ncols;
  if (warp < kTmaConsumerWarps) {
    if (have_cols) {
#pragma unroll
      for (int i = 0; i < CPW; ++i) {
        const int c = warp + i * kTmaConsumerWarps;
        const T s = warp_sum(acc[i]);
        if (lane == 0 && c < ncols) partials[(size_t)c * kSweepPartStride + blockIdx.x] = s;
      }
    }
    if (warp == 0) {  // nacc: already combined into warp 0 by the caller
      const double s = warp_sum(nacc);
      if (lane == 0) partials[(size_t)ncols * kSweepPartStride + blockIdx.x] = Scalar<T>::from_real(s);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();  // cumulative: covers the partials the other threads stored before the barrier above
    *is_last_smem = (atomicAdd(ticket, 1u) == (unsigned)grid - 1u);
  }
  __syncthreads();
  const bool last = *is_last_smem != 0;
  const bool everyone_reduces = pv.P == 1;
  if (everyone_reduces && last && threadIdx.x == 0) {
    *ticket = 0u;
    __threadfence();
    st_release_gpu(flag, epoch);
  }
  if (everyone_reduces || last) {
    if (everyone_reduces && threadIdx.x == 0) {
      unsigned long long spins = 0;
      while (ld_acquire_gpu(flag) < epoch) {
        if (++spins > kGridSpinLimit) {
          *error = 1;
          break;
        }
      }
    }
    if (everyone_reduces) __syncthreads(); else __threadfence();
    // rows c_lo + warp, + 9, ... ; four rows per batch so that 20 independent L2 loads are in flight per lane
    for (int k0 = 0; c_lo + warp + (kTmaConsumerWarps + 1) * k0 <= ncols; k0 += 4) {
      T val[4][5];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = c_lo + warp + (kTmaConsumerWarps + 1) * (k0 + q);
#pragma unroll
        for (int m = 0; m < 5; ++m) {
          const int b = lane + 32 * m;
          val[q][m] = (c <= ncols && b < grid) ? ld_cg<T>(partials + (size_t)c * kSweepPartStride + b) : Scalar<T>::zero();
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = c_lo + warp + (kTmaConsumerWarps + 1) * (k0 + q);
        T s = val[q][0];
#pragma unroll
        for (int m = 1; m < 5; ++m) s = Scalar<T>::add(s, val[q][m]);
        s = warp_sum(s);
        if (lane == 0 && c <= ncols) red[c] = s;
      }
    }
  }
  if (everyone_reduces) {
    __syncthreads();
    return;
  }
  // ---- multi-GPU: the last CTA owns the cross-GPU all-reduce and the publication
  if (last) {
    __syncthreads();
    for (int c = c_lo + threadIdx.x; c < ncols; c += blockDim.x) hout[c] = red[c];
    if (threadIdx.x == 0) {
      *nrm2_out = *reinterpret_cast<const double *>(&red[ncols]);
      *ticket = 0u;
    }
    __syncthreads();
    if (warp == 0) {  // [h | nrm2] is contiguous by construction (hb1 / hb2 layout)
      double *vals = have_cols ? reinterpret_cast<double *>(hout) : nrm2_out;
      const int cnt = (have_cols ? ncols * (int)(sizeof(T) / sizeof(double)) : 0) + 1;
      peer_allreduce_warp(pv, vals, cnt);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      st_release_gpu(flag, epoch);
    }
  }
  if (threadIdx.x == 0) {
    unsigned long long spins = 0;
    while (ld_acquire_gpu(flag) < epoch) {
      if (++spins > kGridSpinLimit) {
        *error = 1;
        break;
      }
    }
  }
  __syncthreads();
  for (int c = c_lo + threadIdx.x; c < ncols; c += blockDim.x) red[c] = ld_cg<T>(hout + c);
  if (threadIdx.x == 0) red[ncols] = Scalar<T>::from_real(__ldcg(nrm2_out));
  __syncthreads();
}

template <class T, int CPW>
__global__ void __launch_bounds__(kTmaThreads, 1)
    cgs_sweep_tma_kernel(const __grid_constant__ CUtensorMap tmap, T *__restrict__ v, int64_t n, int ncols, TmaGeom g,
                         T *__restrict__ partials, T *__restrict__ h1, T *__restrict__ h2, double *rsq_p,
                         double *w1sq_p, double *w2sq_p, T *__restrict__ Hcol, int *info_col, SweepState *state,
                         unsigned long long *flag, unsigned long long epoch, int step,
                         const __grid_constant__ PeerView pv, int64_t row_offset, int push, int early_trigger,
                         unsigned long long *trace) {
  extern __shared__ __align__(128) unsigned char tma_smem_raw[];
  TmaSmem *sm = reinterpret_cast<TmaSmem *>(tma_smem_raw);
  T *ha = reinterpret_cast<T *>(tma_smem_raw + 256);  // pass-1 coefficients h, ha[ncols] = rnorm^2
  T *hb = ha + kSweepCoefs;                           // pass-2 coefficients c, hb[ncols] = wnorm^2, hb[last] = w2^2
  T *ring = reinterpret_cast<T *>(tma_smem_raw + sweep_header_bytes<T>());
  // optional phase trace (B2A_SWEEP_TRACE=1): globaltimer of thread 0 of every CTA at the phase boundaries
  auto mark = [&](int k) {
    if (trace && threadIdx.x == 0) trace[(size_t)blockIdx.x * kSweepTraceSlots + k] = globaltimer_ns();
  };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool producer_warp = warp == kTmaConsumerWarps;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap);
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(&sm->full[s], 1);
      mbar_init(&sm->empty[s], kTmaConsumerWarps);
    }
    mbar_fence_init();
  }
  pdl_wait();
  if (state->poison) return;
  __syncthreads();
  mark(0);

  const int S = g.stages;
  const int first = blockIdx.x * g.tiles_per_cta;
  const int ntl = my_tile_count(g, blockIdx.x);
  const int keep = min(S, ntl);  // tiles that stay in the ring across a phase change
  const size_t stage_elems = (size_t)(ncols + 1) * g.RT;
  const uint32_t stage_bytes = (uint32_t)(stage_elems * sizeof(T));
  constexpr int kInnerPerRow = sizeof(T) / sizeof(double);
  uint32_t cmask = 0;  // consumer: parity of the next fill of each slot it has not yet observed
  uint32_t pmask = 0;  // producer: parity of the number of fills issued per slot

  // producer (one thread): fill slot l % S with local tile l once its previous content has been released
  auto produce = [&](int l) {
    const int s = l % S;
    mbar_wait(&sm->empty[s], ((pmask >> s) & 1u) ^ 1u);  // first fill of a slot passes immediately
    pmask ^= 1u << s;
    mbar_expect_tx(&sm->full[s], stage_bytes);
    tma_load_2d(ring + (size_t)s * stage_elems, &tmap, (first + l) * g.RT * kInnerPerRow, 0, &sm->full[s]);
  };
  // consumer: tile l is either still resident from the previous phase or arrives through its full barrier
  auto acquire = [&](int l, bool resident) -> T * {
    const int s = l % S;
    if (!resident) {
      mbar_wait(&sm->full[s], (cmask >> s) & 1u);
      cmask ^= 1u << s;
    }
    return ring + (size_t)s * stage_elems;
  };
  auto release = [&](int l) {
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm->empty[l % S]);
  };
  // per-warp row partials of ||v||^2 -> warp 0, in warp order (deterministic)
  auto combine_norm = [&](double nacc) -> double {
    nacc = warp_sum(nacc);
    if (lane == 0) sm->wnorm[warp] = nacc;
    consumer_bar_sync();
    double r = 0.0;
    if (warp == 0 && lane == 0) {
#pragma unroll
      for (int w = 0; w < kTmaConsumerWarps; ++w) r += sm->wnorm[w];
    }
    consumer_bar_sync();  // wnorm[] may be rewritten by the next phase
    return r;
  };
  // v[rows of tile] -= tile * hs, lane = row; returns the partial of ||v_new||^2
  auto update_rows = [&](T *tile, T *xt, const T *hs, int64_t r0, bool write_back) -> double {
    double part = 0.0;
    for (int rr = warp * 32 + lane; rr < g.RT; rr += kTmaConsumerWarps * 32) {
      T x0 = xt[rr];
      T x1 = Scalar<T>::zero();
      int c = 0;
      for (; c + 2 <= ncols; c += 2) {
        x0 = Scalar<T>::fnma(tile[(size_t)c * g.RT + rr], hs[c], x0);
        x1 = Scalar<T>::fnma(tile[(size_t)(c + 1) * g.RT + rr], hs[c + 1], x1);
      }
      if (c < ncols) x0 = Scalar<T>::fnma(tile[(size_t)c * g.RT + rr], hs[c], x0);
      x0 = Scalar<T>::add(x0, x1);
      v[r0 + rr] = x0;
      if (write_back) xt[rr] = x0;
      part += Scalar<T>::abs2(x0);
    }
    return part;
  };

  T acc[CPW];
#pragma unroll
  for (int i = 0; i < CPW; ++i) acc[i] = Scalar<T>::zero();
  double nacc = 0.0;

  // ------------------------------------------------------------------ P1: h = V' v, rnorm^2 (forward)
  if (producer_warp) {
    if (lane == 0)
      for (int l = 0; l < ntl; ++l) produce(l);
  } else {
    for (int l = 0; l < ntl; ++l) {
      const T *tile = acquire(l, false);
      tile_dots<T, CPW>(tile, tile + (size_t)ncols * g.RT, g.RT, ncols, warp, lane, acc, nacc, warp == 0);
      if (l < ntl - keep) release(l);  // the last `keep` tiles stay on chip for P2
    }
  }
  __syncwarp();
  mark(1);
  sweep_reduce_barrier<T, CPW>(acc, nacc, true, ncols, warp, lane, partials, ha, h1, rsq_p, &state->ticket[2], flag,
                               epoch + 1, &sm->is_last, &state->error, pv);
  mark(2);

  // ------------------------------------------- P2: v -= V h, wnorm^2, speculative c = V' v_new (backward)
#pragma unroll
  for (int i = 0; i < CPW; ++i) acc[i] = Scalar<T>::zero();
  nacc = 0.0;
  if (producer_warp) {
    if (lane == 0)
      for (int l = ntl - 1 - keep; l >= 0; --l) produce(l);
  } else {
    for (int l = ntl - 1; l >= 0; --l) {
      T *tile = acquire(l, l >= ntl - keep);
      T *xt = tile + (size_t)ncols * g.RT;
      nacc += update_rows(tile, xt, ha, (int64_t)(first + l) * g.RT, true);
      consumer_bar_sync();  // the updated x tile is complete
      double dummy = 0.0;
      tile_dots<T, CPW>(tile, xt, g.RT, ncols, warp, lane, acc, dummy, false);
      if (l >= keep) release(l);  // the first `keep` tiles stay on chip (with v_new in place) for P3
    }
    fence_proxy_async();  // P3's TMA loads read the rows of v written above
    nacc = combine_norm(nacc);
  }
  __syncwarp();
  mark(3);
  sweep_reduce_barrier<T, CPW>(acc, nacc, true, ncols, warp, lane, partials + kSweepPartRegion, hb, h2, w1sq_p,
                               &state->ticket[3], flag, epoch + 2, &sm->is_last, &state->error, pv);
  mark(4);

  const double rsq = *reinterpret_cast<const double *>(&ha[ncols]);
  const double w1sq = *reinterpret_cast<const double *>(&hb[ncols]);
  double rnorm = sqrt(rsq), wnorm = sqrt(w1sq);
  const bool second = wnorm < kEta * rnorm;  // expansion.jl:91 (strict); identical on every CTA and rank

  // ------------------------------------------------------- P3 (gated): v -= V c, wnorm^2 (forward)
  if (second) {
    nacc = 0.0;
    if (producer_warp) {
      if (lane == 0)
        for (int l = keep; l < ntl; ++l) produce(l);
    } else {
      for (int l = 0; l < ntl; ++l) {
        T *tile = acquire(l, l < keep);
        nacc += update_rows(tile, tile + (size_t)ncols * g.RT, hb, (int64_t)(first + l) * g.RT, false);
        release(l);
      }
      nacc = combine_norm(nacc);
    }
    __syncwarp();
    mark(5);
    // only the norm is reduced; its result lands in the spare slot hb[kSweepCoefs - 1]
    sweep_reduce_barrier<T, CPW>(acc, nacc, false, ncols, warp, lane, partials + 2 * kSweepPartRegion,
                                 hb + (kSweepCoefs - 1 - ncols), h2, w2sq_p, &state->ticket[4], flag, epoch + 3, &sm->is_last, &state->error, pv);
    mark(6);
    rnorm = wnorm;
    wnorm = sqrt(*reinterpret_cast<const double *>(&hb[kSweepCoefs - 1]));
  }

  // --------------------------------------- P4: H column, breakdown test, normalisation (expansion.jl:95-107)
  const bool breakdown = wnorm <= kEta * rnorm;  // expansion.jl:99 (non-strict)
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < ncols; c += blockDim.x)
      Hcol[c] = second ? Scalar<T>::add(ha[c], hb[c]) : ha[c];  // expansion.jl:95
    if (threadIdx.x == 0) {
      Hcol[ncols] = Scalar<T>::from_real(breakdown ? 0.0 : wnorm);  // expansion.jl:100,104
      info_col[0] = (second ? 1 : 0) | (breakdown ? 2 : 0);
      if (second) atomicAdd(&state->second_passes, 1ull);
      if (breakdown) state->poison = step;
    }
  }
  if (breakdown) return;
  if (early_trigger) pdl_trigger();  // experiment: let the next mat-vec's CTAs queue up behind the normalisation

  constexpr int PV = Scalar<T>::per_vec;
  constexpr int U = 4;  // vectors in flight per thread
  const int64_t rb = (int64_t)first * g.RT;
  const int64_t re = rb + (int64_t)ntl * g.RT;  // <= ld: rows past n are the zero padding of the workspace
  const int64_t stride = (int64_t)blockDim.x * PV;
  const bool do_push = push && pv.P > 1;
  const bool vec_ok = PV == 1 || (row_offset & 1) == 0;
  for (int64_t r0 = rb + (int64_t)threadIdx.x * PV; r0 < re; r0 += stride * U) {
    double2 x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = r0 + u * stride;
      x[u] = r < re ? *reinterpret_cast<const double2 *>(v + r) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      x[u].x /= wnorm;  // v ./= wnorm (expansion.jl:106): a true division, like the reference
      x[u].y /= wnorm;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = r0 + u * stride;
      if (r < re) *reinterpret_cast<double2 *>(v + r) = x[u];
    }
    if (do_push) {
      for (int p = 0; p < pv.P; ++p) {
        T *pxb = reinterpret_cast<T *>(pv.peer[p] + pv.off_x) + row_offset;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t r = r0 + u * stride;
          if (r >= re || r >= n) continue;
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
  mark(7);
  pdl_trigger();
  if (do_push) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) sm->is_last = (atomicAdd(&state->ticket[5], 1u) == gridDim.x - 1u);
    __syncthreads();
    if (sm->is_last && threadIdx.x == 0) {
      state->ticket[5] = 0u;
      peer_x_publish(pv);
    }
  }
}

}  // namespace b2a
