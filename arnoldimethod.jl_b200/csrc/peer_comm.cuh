// peer_comm.cuh - one-shot collectives over NVLink peer memory, callable from inside a kernel.
//
// Row-sharded Arnoldi needs, per orthogonalisation sweep, an all-reduce of the (j+1)-vector
// [h | ||v||^2].  These are <= ~1 KB and purely latency-bound; a host-launched
// ncclAllReduce costs 10-25 us each.  Instead, the CTA that finishes a reduction kernel last
// performs the all-reduce itself before the kernel ends:
//
//   every rank owns a communication block (cudaMalloc + CUDA IPC, mapped into every peer process)
//       data [kPeerBufs][P][slot] doubles,  flag [kPeerBufs][P] u64
//   rank r:  for all peers p:  p.data[buf][r][:] = my partial     (plain stores over NVLink)
//            __threadfence_system();  p.flag[buf][r] = seq         (st.release.sys)
//            wait until my.flag[buf][p] == seq for every p          (ld.acquire.sys)
//            result[i] = sum_p my.data[buf][p][i]  in rank order    => identical bits on every rank
//
// `seq` is a DEVICE counter that advances only when a collective really executes, so kernels that
// gate themselves off (second Gram-Schmidt pass, poisoned sweep) stay consistent across ranks: all
// ranks gate on the same all-reduced scalars.  kPeerBufs rotating buffers: a rank can be at most one
// collective ahead of a peer that is still reading the previous one.
#pragma once

#include <stdint.h>

namespace b2a {

constexpr int kPeerBufs = 4;
constexpr int kPeerMaxRanks = 16;
constexpr unsigned long long kPeerSpinLimit = 1ull << 31;  // ~seconds; then flag an error instead of hanging

struct PeerView {
  int P = 1, rank = 0;
  int slot = 0;                         // doubles per rank slot
  char *peer[kPeerMaxRanks] = {nullptr};  // base of every rank's communication block (peer[rank] = local)
  unsigned long long off_flag_ar = 0, off_data_ar = 0, off_flag_x = 0, off_x = 0;
  // staged exchange (copy engines / per-owner flags): its own flag array, and a second x buffer (exchange number
  // parity selects the buffer, so a rank that runs one mat-vec ahead never overwrites what a peer still reads)
  unsigned long long off_flag_xs = 0, x_stride = 0;
  // low-latency all-reduce cells (16 bytes per double: lo | flag | hi | flag), see peer_allreduce_ll_warp
  unsigned long long off_ll = 0;
  int ll = 0;
  unsigned long long *seq_ar = nullptr;  // local device counters
  unsigned long long *seq_x = nullptr;
  int *err = nullptr;
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double *p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// Low-latency variant (the default): every double travels as ONE 16-byte store {lo32, flag, hi32, flag} with
// flag = the collective's sequence number, the way NCCL's LL protocol packs data and flag into 8-byte units that the
// fabric delivers atomically.  The receiver polls the cell itself until both flags match: no __threadfence_system, no
// separate flag round trip - one NVLink store latency per all-reduce instead of three dependent hops.  Cells of
// buffer `seq % kPeerBufs` are rewritten four collectives later, by which time every reader has long passed (a rank
// can only start collective s + 1 after all ranks have contributed to s).  Sum in rank order: identical bits everywhere.
__device__ __forceinline__ void st_ll(uint4 *p, double v, unsigned flag) {
  const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(lo), "r"(flag), "r"(hi), "r"(flag) : "memory");
}
__device__ __forceinline__ bool ld_ll(const uint4 *p, unsigned flag, double *v) {
  unsigned lo, f1, hi, f2;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f1), "=r"(hi), "=r"(f2) : "l"(p) : "memory");
  *v = __hiloint2double((int)hi, (int)lo);
  return f1 == flag && f2 == flag;
}
__device__ __forceinline__ void peer_allreduce_ll_warp(const PeerView &pv, double *vals, int cnt) {
  const int lane = threadIdx.x & 31;
  unsigned long long seq = 0;
  if (lane == 0) seq = atomicAdd(pv.seq_ar, 1ull) + 1ull;
  seq = __shfl_sync(0xffffffffu, seq, 0);
  const int buf = (int)(seq % kPeerBufs);
  const unsigned flag = (unsigned)seq;
  for (int i = lane; i < cnt; i += 32) {
    const double v = vals[i];
    for (int p = 0; p < pv.P; ++p) {
      uint4 *dst = reinterpret_cast<uint4 *>(pv.peer[p] + pv.off_ll) + ((size_t)buf * pv.P + pv.rank) * pv.slot;
      st_ll(dst + i, v, flag);
    }
  }
  const uint4 *my = reinterpret_cast<const uint4 *>(pv.peer[pv.rank] + pv.off_ll) + (size_t)buf * pv.P * pv.slot;
  for (int i = lane; i < cnt; i += 32) {
    double s = 0.0;
    for (int p = 0; p < pv.P; ++p) {
      double v;
      unsigned long long spins = 0;
      while (!ld_ll(my + (size_t)p * pv.slot + i, flag, &v)) {
        if (++spins > kPeerSpinLimit) {
          *pv.err = 4;
          break;
        }
      }
      s += v;
    }
    vals[i] = s;
  }
  __syncwarp();
}

// In-place all-reduce (sum) of vals[0..cnt) across the P ranks.  Call with ONE full warp; vals is
// local global memory already visible to the calling warp.
__device__ __forceinline__ void peer_allreduce_warp(const PeerView &pv, double *vals, int cnt) {
  if (pv.ll) {
    peer_allreduce_ll_warp(pv, vals, cnt);
    return;
  }
  const int lane = threadIdx.x & 31;
  unsigned long long seq = 0;
  if (lane == 0) {
    // atomic at L2: inside the fused sweep kernel successive all-reduces may run on different SMs, and a
    // plain load could return a stale L1 copy of the counter
    seq = atomicAdd(pv.seq_ar, 1ull) + 1ull;
  }
  seq = __shfl_sync(0xffffffffu, seq, 0);
  const int buf = (int)(seq % kPeerBufs);
  // 1. scatter my partial into every rank's block (my own included)
  for (int p = 0; p < pv.P; ++p) {
    double *dst = reinterpret_cast<double *>(pv.peer[p] + pv.off_data_ar) + ((size_t)buf * pv.P + pv.rank) * pv.slot;
    for (int i = lane; i < cnt; i += 32) dst[i] = vals[i];
  }
  __threadfence_system();
  __syncwarp();
  // 2. publish
  for (int p = lane; p < pv.P; p += 32) {
    unsigned long long *f = reinterpret_cast<unsigned long long *>(pv.peer[p] + pv.off_flag_ar) + (size_t)buf * pv.P + pv.rank;
    st_release_sys(f, seq);
  }
  // 3. wait for every rank's contribution
  const unsigned long long *myf = reinterpret_cast<const unsigned long long *>(pv.peer[pv.rank] + pv.off_flag_ar) + (size_t)buf * pv.P;
  for (int p = lane; p < pv.P; p += 32) {
    unsigned long long spins = 0;
    while (ld_acquire_sys(myf + p) != seq) {
      if (++spins > kPeerSpinLimit) {
        *pv.err = 1;
        break;
      }
    }
  }
  __syncwarp();
  // 4. sum in rank order
  const double *myd = reinterpret_cast<const double *>(pv.peer[pv.rank] + pv.off_data_ar) + (size_t)buf * pv.P * pv.slot;
  for (int i = lane; i < cnt; i += 32) {
    double s = 0.0;
    for (int p = 0; p < pv.P; ++p) s += ld_relaxed_sys_f64(myd + (size_t)p * pv.slot + i);
    vals[i] = s;
  }
  __syncwarp();
}

// ---- x exchange: push model ----------------------------------------------------------------
// The kernel that produces this rank's slice of the next mat-vec input stores it into EVERY rank's
// x buffer (peer[p] + off_x, indexed by global row).  Each CTA fences its remote stores; the CTA that
// finishes last bumps the device counter seq_x and publishes it in every peer's flag_x[rank].
// The consumer (SpMV) waits until all P flags reached its own seq_x.
__device__ __forceinline__ void peer_x_publish(const PeerView &pv) {  // one thread of the last CTA
  const unsigned long long seq = atomicAdd(pv.seq_x, 1ull) + 1ull;
  __threadfence_system();
  for (int p = 0; p < pv.P; ++p)
    st_release_sys(reinterpret_cast<unsigned long long *>(pv.peer[p] + pv.off_flag_x) + pv.rank, seq);
}
__device__ __forceinline__ void peer_x_wait(const PeerView &pv) {  // one thread per CTA, then __syncthreads
  const unsigned long long want = *reinterpret_cast<volatile unsigned long long *>(pv.seq_x);
  const unsigned long long *f = reinterpret_cast<const unsigned long long *>(pv.peer[pv.rank] + pv.off_flag_x);
  for (int p = 0; p < pv.P; ++p) {
    unsigned long long spins = 0;
    while (ld_acquire_sys(f + p) < want) {
      if (++spins > kPeerSpinLimit) {
        *pv.err = 2;
        break;
      }
    }
  }
}

// staged exchange: wait for the slice of ONE owner rank; `want` is the host's exchange counter (every exchange
// publishes, poisoned sweeps included, so host and device never drift)
__device__ __forceinline__ void peer_x_wait_one(const PeerView &pv, int owner, unsigned long long want) {
  const unsigned long long *f = reinterpret_cast<const unsigned long long *>(pv.peer[pv.rank] + pv.off_flag_xs) + owner;
  unsigned long long spins = 0;
  while (ld_acquire_sys(f) < want) {
    if (++spins > kPeerSpinLimit) {
      *pv.err = 3;
      break;
    }
  }
}
// ... or for the slices of all other ranks (operators that are not stored by owner group)
__device__ __forceinline__ void peer_x_wait_others(const PeerView &pv, unsigned long long want) {
  for (int p = 0; p < pv.P; ++p)
    if (p != pv.rank) peer_x_wait_one(pv, p, want);
}

}  // namespace b2a
