// kernels_rotate_mma.cuh - the Krylov-Schur change of basis V[:, col0 : col0+N) <- V[:, col0 : col0+K) * Q, in place,
// on the FP64 tensor pipe (src/run.jl:363-365 and :382-383 of the reference; see kernels_rotate.cuh for the in-place
// argument).
//
// tcgen05.mma has no FP64 kind, but sm_100a keeps the FP64 warp-level MMA (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4):
// one instruction = 256 FMAs fed by ONE 8-byte operand per lane and side, so the shared-memory : FMA ratio that
// limits a register-blocked DFMA kernel (kernels_rotate.cuh: 58 % l1tex, 22 % fp64 pipe at 0.46 of the HBM bound)
// drops by an order of magnitude.  Structure:
//
//   * persistent grid (one CTA per SM), contiguous row-tile ranges per CTA;
//   * one PRODUCER thread streams the input tile [R rows x K columns] with 2-D tensor-map TMA into a multi-stage
//     shared-memory ring (full / empty mbarriers) - the same pipeline as the Gram-Schmidt sweeps; columns past K are
//     zero-filled by the TMA unit (out-of-bounds fill), so the k loop needs no mask;
//   * up to eight CONSUMER warps each own a 16-row slab of the tile = two 8-row MMA blocks made of the EVEN and the
//     ODD rows, so one 16-byte shared-memory load yields the A fragments of both blocks and the results of both
//     blocks form one 16-byte store per output column (full 128-byte lines per warp);
//   * Q is pre-arranged ON THE HOST in B-fragment order (one coalesced 8-byte load per lane and MMA, conflict-free)
//     and kept in shared memory (or read through L1 when it does not fit);
//   * ComplexF64 runs as the real product [Vre Vim] * [[Qre Qim], [-Qim Qre]] on the interleaved (re, im) tile:
//     the re / im parts of one 16-byte element feed two MMAs against the two matching B fragments, and an
//     accumulator pair IS one complex output element.
//   * the column move V[:, k+1] <- V[:, maxdim+1] of run.jl:365 rides along in a dedicated MOVER warp (after the
//     tile is in shared memory: the destination column is one of the inputs).  First version: the consumer warps
//     did it at the top of every tile, and the dependent LDG -> STG pair (HBM latency, ~1 us) was 20 % of all
//     stall samples and starved the tensor pipe (ncu r2: dmma pipe 43 %, DRAM 46 %).
//
// Everything is read once and written once: n s (K + N + 2) bytes, B_rot of SURVEY 8(d).
#pragma once

#include "kernels_cgs_tma.cuh"

namespace b2a {

constexpr int kRotConsumerWarps = 8;
constexpr int kRotThreads = (kRotConsumerWarps + 2) * 32;  // + producer warp + mover warp

struct RotGeom {
  int R;              // rows per tile = 16 * warps
  int warps;          // active consumer warps
  int stages;
  int tiles_per_cta;
  int ntiles;
  int nbox;           // TMA boxes per tile (<= 256 columns each)
  int box_cols;
  int kpad;           // columns held by a stage = nbox * box_cols (multiple of 4, >= K)
};

__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(d[0]), "+d"(d[1])
               : "d"(a), "d"(b));
}

// V is addressed in doubles: element (row, col, part) of the workspace at (col * ld + row) * inner + part.
// The product is written to Out[:, col0 : col0 + N) with leading dimension ldo (in elements of the OUTPUT type);
// the rotation passes Out = V, ldo = ld (in place).  CPLX_OUT with a real input (CPLX = false) is the product of a
// real basis with a complex coefficient matrix (partialeigen, src/eigvals.jl:94: X = Q * Y): B then holds Y as
// interleaved (re, im) columns and an accumulator pair is one complex output element, exactly as in the
// all-complex case.
template <bool CPLX, bool CPLX_OUT, int NT, int MINB>
__global__ void __launch_bounds__(kRotThreads, MINB)
    rotate_mma_kernel(const __grid_constant__ CUtensorMap tmap, double *__restrict__ V, int64_t ld,
                      double *__restrict__ Out, int64_t ldo, int col0, int N, const double *__restrict__ Bg, int KS,
                      int NCH, RotGeom g, int move_src, int move_dst, int b_in_smem) {
  static_assert(CPLX_OUT || !CPLX, "a complex input has a complex output");
  extern __shared__ __align__(128) unsigned char rot_smem_raw[];
  constexpr int INNER = CPLX ? 2 : 1;
  constexpr int PARTS = CPLX ? 2 : 1;
  TmaSmem *sm = reinterpret_cast<TmaSmem *>(rot_smem_raw);
  const int NTT = NCH * NT;
  const size_t b_elems = (size_t)KS * NTT * PARTS * 32;
  double *Bs = reinterpret_cast<double *>(rot_smem_raw + 256);
  const size_t b_bytes = b_in_smem ? (b_elems * sizeof(double) + 127) / 128 * 128 : 0;
  double *ring = reinterpret_cast<double *>(rot_smem_raw + 256 + b_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t stage_elems = (size_t)g.kpad * g.R * INNER;
  const uint32_t stage_bytes = (uint32_t)(stage_elems * sizeof(double));

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap);
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(&sm->full[s], 1);
      mbar_init(&sm->empty[s], g.warps + 1);  // consumer warps + the mover
    }
    mbar_fence_init();
  }
  if (b_in_smem)
    for (size_t i = threadIdx.x; i < b_elems; i += blockDim.x) Bs[i] = Bg[i];
  __syncthreads();

  const int first = blockIdx.x * g.tiles_per_cta;
  const int ntl = max(0, min(g.tiles_per_cta, g.ntiles - first));

  if (warp == kRotConsumerWarps) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      for (int l = 0; l < ntl; ++l) {
        const int s = l % g.stages;
        mbar_wait(&sm->empty[s], (((uint32_t)(l / g.stages)) & 1u) ^ 1u);
        mbar_expect_tx(&sm->full[s], stage_bytes);
        double *dst = ring + (size_t)s * stage_elems;
        const int r0 = (first + l) * g.R * INNER;
        for (int b = 0; b < g.nbox; ++b)
          tma_load_2d(dst + (size_t)b * g.box_cols * g.R * INNER, &tmap, r0, b * g.box_cols, &sm->full[s]);
      }
    }
    return;
  }
  if (warp == kRotConsumerWarps + 1) {
    // ------------------------------------------------------------------ mover: V[rows, move_dst] <- V[rows, move_src]
    // The destination column is an INPUT of the product, so its rows may be overwritten only once the tile that
    // holds them sits in shared memory (full barrier).  The mover releases the stage at once (it never reads it) and
    // takes the HBM latency of the copy off the consumers' critical path.
    const int pieces = g.R * INNER / 2;  // 16-byte pieces per tile column
    for (int l = 0; l < ntl; ++l) {
      const int s = l % g.stages;
      mbar_wait(&sm->full[s], ((uint32_t)(l / g.stages)) & 1u);
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm->empty[s]);
      if (move_dst >= 0) {
        const int64_t off = (int64_t)(first + l) * g.R * INNER;
        const double2 *src = reinterpret_cast<const double2 *>(V + (int64_t)move_src * ld * INNER + off);
        double2 *dst = reinterpret_cast<double2 *>(V + (int64_t)move_dst * ld * INNER + off);
        double2 x[4];  // R * INNER / 2 <= 128 pieces = 4 per lane
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (lane + 32 * u < pieces) x[u] = src[lane + 32 * u];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (lane + 32 * u < pieces) dst[lane + 32 * u] = x[u];
      }
    }
    return;
  }
  if (warp >= g.warps) {
    return;
  }

  // ---------------------------------------------------------------------- consumers
  const int gq = lane >> 2, tq = lane & 3;  // MMA group id (row of A / column of B) and thread-in-group (k slot)
  const int rw = 16 * warp + 2 * gq;        // first of this lane's two tile rows (even row; the odd one follows)
  const double *Bp = b_in_smem ? Bs : Bg;
  for (int l = 0; l < ntl; ++l) {
    const int s = l % g.stages;
    mbar_wait(&sm->full[s], ((uint32_t)(l / g.stages)) & 1u);
    const double *tile = ring + (size_t)s * stage_elems;
    const int64_t row0 = (int64_t)(first + l) * g.R;
    for (int ch = 0; ch < NCH; ++ch) {
      double acc[2][NT][2];
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[m][nt][0] = acc[m][nt][1] = 0.0;
      const double *bch = Bp + (size_t)ch * NT * PARTS * 32 + lane;
      for (int ks = 0; ks < KS; ++ks) {
        const int k = 4 * ks + tq;
        const double *bk = bch + (size_t)ks * NTT * PARTS * 32;
        if (!CPLX) {
          const double2 a = *reinterpret_cast<const double2 *>(tile + (size_t)k * g.R + rw);
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const double b = bk[nt * 32];
            dmma884(acc[0][nt], a.x, b);
            dmma884(acc[1][nt], a.y, b);
          }
        } else {
          const double2 *zp = reinterpret_cast<const double2 *>(tile) + (size_t)k * g.R + rw;
          const double2 z0 = zp[0], z1 = zp[1];
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const double bre = bk[nt * 64], bim = bk[nt * 64 + 32];
            dmma884(acc[0][nt], z0.x, bre);
            dmma884(acc[0][nt], z0.y, bim);
            dmma884(acc[1][nt], z1.x, bre);
            dmma884(acc[1][nt], z1.y, bim);
          }
        }
      }
      // results: lane holds rows rw, rw+1 and outputs 2 tq, 2 tq + 1 of every 8-wide n tile
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int n0 = (ch * NT + nt) * 8 + 2 * tq;
        if (!CPLX_OUT) {
#pragma unroll
          for (int i = 0; i < 2; ++i)
            if (n0 + i < N)
              *reinterpret_cast<double2 *>(Out + (int64_t)(col0 + n0 + i) * ldo + row0 + rw) =
                  make_double2(acc[0][nt][i], acc[1][nt][i]);
        } else {
          const int o = n0 >> 1;  // complex output column: the accumulator pair is (re, im)
          if (o < N) {
            double2 *out = reinterpret_cast<double2 *>(Out) + (int64_t)(col0 + o) * ldo + row0 + rw;
            out[0] = make_double2(acc[0][nt][0], acc[0][nt][1]);
            out[1] = make_double2(acc[1][nt][0], acc[1][nt][1]);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm->empty[s]);
  }
}

}  // namespace b2a
