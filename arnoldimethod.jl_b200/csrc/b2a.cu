// b2a.cu - engine and C ABI of libb200arnoldi.so (see include/b200arnoldi.h).
//
// One translation unit: the CUDA kernels (kernels_*.cuh), the host m x m algebra
// (host_dense.hpp), the device engine that strings them into Arnoldi sweeps, the C++
// restatement of the restart driver `_partialschur` (src/run.jl:224-392 of the reference)
// and the extern "C" entry points.
//
// Execution model: a whole `iterate_arnoldi!(A, arnoldi, from:to)` sweep is enqueued on one
// CUDA stream without any host round trip.  Every data-dependent decision of
// src/expansion.jl (second Gram-Schmidt pass at :91, breakdown at :99) is taken on the
// device from all-reduced scalars; the kernels of the conditional second pass gate
// themselves, and a breakdown raises a `poison` flag that turns the rest of the sweep into
// no-ops.  The host synchronises once per sweep, reads back the new H columns, and - only if
// a breakdown happened - re-seeds that column (reinitialize!) and resumes the sweep.

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "b200arnoldi.h"
#include "device_common.cuh"
#include "host_dense.hpp"
#include "kernels_cgs.cuh"
#include "kernels_cgs_tma.cuh"
#include "kernels_blocks.cuh"
#include "kernels_cgs_sweep.cuh"
#include "kernels_rotate.cuh"
#include "kernels_rotate_mma.cuh"
#include "kernels_solve.cuh"
#include "kernels_spmv.cuh"
#include "kernels_spmv_tma.cuh"
#include "peer_comm.cuh"

using b2a::cdouble;
using b2a::host::cplx;

// =============================================================================== errors
static thread_local std::string g_err;
static int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
#define CUDA_TRY(expr)                                                                     \
  do {                                                                                     \
    cudaError_t e__ = (expr);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      (void)cudaGetLastError();                                                            \
      return fail(e__ == cudaErrorMemoryAllocation ? B2A_ERR_OOM : B2A_ERR_CUDA,           \
                  std::string(#expr) + ": " + cudaGetErrorString(e__));                    \
    }                                                                                      \
  } while (0)
#define B2A_TRY(expr)            \
  do {                           \
    int s__ = (expr);            \
    if (s__ != B2A_OK) return s__; \
  } while (0)
#define ARG_CHECK(cond, msg) \
  do {                       \
    if (!(cond)) return fail(B2A_ERR_ARGUMENT, msg); \
  } while (0)

// ================================================================================ NCCL
// NCCL is resolved at run time (dlopen) so that the library loads on hosts without it and
// re-uses the libnccl.so.2 the host program (PyTorch) has already mapped.
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static int load_nccl() {
  if (g_nccl.handle) return B2A_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *nm : names) {
    h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return fail(B2A_ERR_NCCL, std::string("dlopen(libnccl.so.2) failed: ") + dlerror());
#define LOAD(sym)                                                                  \
  g_nccl.sym = reinterpret_cast<decltype(g_nccl.sym)>(dlsym(h, "nccl" #sym));      \
  if (!g_nccl.sym) return fail(B2A_ERR_NCCL, "dlsym(nccl" #sym ") failed")
  LOAD(GetUniqueId);
  LOAD(CommInitRank);
  LOAD(CommDestroy);
  LOAD(AllReduce);
  LOAD(AllGather);
  LOAD(Broadcast);
  LOAD(GroupStart);
  LOAD(GroupEnd);
  LOAD(GetErrorString);
#undef LOAD
  g_nccl.handle = h;
  return B2A_OK;
}
#define NCCL_TRY(expr)                                                                           \
  do {                                                                                           \
    ncclResult_t r__ = (expr);                                                                   \
    if (r__ != ncclSuccess)                                                                      \
      return fail(B2A_ERR_NCCL, std::string(#expr) + ": " + g_nccl.GetErrorString(r__));         \
  } while (0)

// ============================================================================== handles
struct b2a_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int rank = 0, world = 1;
  ncclComm_t comm = nullptr;
  int64_t launches = 0;
  int num_sms = b2a::kSMs;
  int *d_zero = nullptr;  // a device int that is always 0 (poison stand-in outside sweeps)
  // optional per-kernel-kind timing (b2a_ctx_profile_*)
  struct ProfRec {
    int kind;
    double bytes;
    int gate_step;  // > 0: launch belongs to the gated second pass of that step
    cudaEvent_t e0, e1;
    double extra_bytes = 0.0;  // fused sweep: bytes of its gated phase, counted if step `extra_step` ran it
    int extra_step = 0;
  };
  bool prof_on = false;
  std::vector<ProfRec> prof_pending;
  std::vector<cudaEvent_t> prof_pool;
  int64_t prof_n[B2A_K_COUNT] = {0};
  double prof_ms[B2A_K_COUNT] = {0}, prof_bytes[B2A_K_COUNT] = {0};
  // pinned host staging buffers are expensive to create: recycled across workspaces
  std::vector<std::pair<size_t, char *>> pinned_cache;
  // NVLink peer communication block (cudaMalloc + CUDA IPC): created once, re-used by every workspace
  // whose needs fit (creating it is a collective with ~ms cost); see peer_setup()
  char *peer_local = nullptr;
  std::vector<void *> peer_opened;
  b2a::PeerView peer_view;   // P == 1: not available
  int peer_slot = 0;
  size_t peer_x_bytes = 0;
  bool peer_busy = false;    // handed to a live workspace
  // staged x exchange (row-sharded mat-vec): side streams that carry the copy-engine transfers + per-slice flag
  // publications, and the exchange counter every rank advances in lockstep (see enqueue_matvec)
  std::vector<cudaStream_t> xchg_streams;
  std::vector<cudaEvent_t> xchg_done;   // one per side stream: its last transfer + publication
  cudaEvent_t xchg_data_ev[2] = {nullptr, nullptr};  // chained exchange: data transfer of the previous stage done
  cudaEvent_t xchg_ready = nullptr;     // main stream: the column to exchange is final
  unsigned long long x_seq = 0;
  // device table of consecutive exchange numbers: the per-slice flag is published by an 8-byte COPY-ENGINE transfer
  // from this table (a kernel would need an SM slot, and the mat-vec launches that spin on the flags may hold them all)
  unsigned long long *xchg_seq_table = nullptr;
  unsigned long long xchg_table_base = 0;
  // kernels whose opt-in dynamic shared memory limit has been raised on THIS device (the attribute is per device
  // and per function; a process may drive several GPUs through several contexts)
  std::vector<const void *> smem_attr_done;
};

// Device memory comes from the device's stream-ordered pool (cudaMallocAsync) with an unlimited
// release threshold: after the first solve, creating and destroying operators / workspaces costs
// microseconds instead of the milliseconds of cudaMalloc / cudaFree (which also synchronise).
static cudaError_t dev_alloc(b2a_ctx *c, void **p, size_t bytes) {
  return cudaMallocAsync(p, std::max<size_t>(bytes, 16), c->stream);
}
static void dev_free(b2a_ctx *c, void *p) {
  if (p) cudaFreeAsync(p, c->stream);
}
static cudaError_t pinned_get(b2a_ctx *c, size_t bytes, char **out, size_t *got) {
  for (size_t i = 0; i < c->pinned_cache.size(); ++i)
    if (c->pinned_cache[i].first >= bytes) {
      *out = c->pinned_cache[i].second;
      *got = c->pinned_cache[i].first;
      c->pinned_cache.erase(c->pinned_cache.begin() + i);
      return cudaSuccess;
    }
  *got = bytes;
  return cudaMallocHost(reinterpret_cast<void **>(out), bytes);
}
static void pinned_put(b2a_ctx *c, char *p, size_t bytes) {
  if (!p) return;
  if (c->pinned_cache.size() < 8)
    c->pinned_cache.emplace_back(bytes, p);
  else
    cudaFreeHost(p);
}

static cudaEvent_t prof_event(b2a_ctx *c) {
  cudaEvent_t e;
  if (!c->prof_pool.empty()) {
    e = c->prof_pool.back();
    c->prof_pool.pop_back();
  } else {
    cudaEventCreate(&e);
  }
  return e;
}
static inline void prof_begin(b2a_ctx *c, int kind, double bytes, int gate_step = 0, cudaStream_t st = nullptr) {
  if (!c->prof_on) return;
  b2a_ctx::ProfRec r{kind, bytes, gate_step, prof_event(c), prof_event(c), 0.0, 0};
  cudaEventRecord(r.e0, st ? st : c->stream);
  c->prof_pending.push_back(r);
}
static inline void prof_end(b2a_ctx *c, cudaStream_t st = nullptr) {
  if (!c->prof_on) return;
  cudaEventRecord(c->prof_pending.back().e1, st ? st : c->stream);
}
// call after a stream synchronisation; info[step - info_base] bit0 tells whether the gated
// second pass of `step` really ran (gated-off launches are dropped from the statistics)
static void prof_collect(b2a_ctx *c, const int *info, int info_base, int info_count) {
  if (!c->prof_on) return;
  for (auto &r : c->prof_pending) {
    bool executed = true;
    if (r.gate_step > 0) {
      const int i = r.gate_step - info_base;
      executed = info && i >= 0 && i < info_count && (info[i] & 1);
    }
    if (executed) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
        c->prof_n[r.kind] += 1;
        c->prof_ms[r.kind] += ms;
        c->prof_bytes[r.kind] += r.bytes;
        if (r.extra_step > 0) {
          const int i = r.extra_step - info_base;
          if (info && i >= 0 && i < info_count && (info[i] & 1)) c->prof_bytes[r.kind] += r.extra_bytes;
        }
      } else {
        (void)cudaGetLastError();
      }
    }
    c->prof_pool.push_back(r.e0);
    c->prof_pool.push_back(r.e1);
  }
  c->prof_pending.clear();
}

enum OpKind { OP_CSR = 0, OP_CSC_SCATTER = 1, OP_CALLBACK = 2, OP_SHIFT_INVERT = 3 };

struct b2a_op {
  b2a_ctx *ctx = nullptr;
  int dtype = B2A_F64;
  int kind = OP_CSR;
  int64_t n_local = 0, n_global = 0, row_offset = 0, nnz = 0;
  int64_t *d_ptr = nullptr;  // rowptr (CSR) / colptr (CSC), 0-based
  int32_t *d_idx = nullptr;  // colind (CSR) / rowind (CSC), 0-based, global
  void *d_vals = nullptr;
  bool owns = true;
  int lpr = 8;  // lanes per row (column)
  int rows_in_flight = 2, grid_mult = 16;  // SpMV tuning (env B2A_SPMV_U / B2A_SPMV_GRID)
  int entries_in_flight = 4;               // entries per lane and row issued together (B2A_SPMV_E = 1, 2, 4)
  // TMA-stream kernel: row tiles built at upload (kernels_spmv_tma.cuh)
  int32_t *d_tile_row = nullptr;
  int ntiles = 0, tma_stages = 0;
  bool use_tma = false;
  // column blocking (x larger than L2 and scattered columns): block-major CSR, d_ptr holds nblocks row-pointer
  // arrays of n_local+1 entries each (absolute positions in d_idx / d_vals)
  int nblocks = 1;
  // row-sharded operators: the column blocks are OWNER GROUPS of x (block b = columns of the ranks rank + b G ..
  // rank + b G + G - 1, in the order the staged exchange delivers their slices), about 32 MB of x each, so that one
  // launch per group starts as soon as its slices have arrived and gathers from an L2-resident part of x
  bool owner_blocks = false;
  int64_t owner_W = 0;  // rows per rank of the uniform partition the blocks were cut for
  int owner_G = 1;      // ranks per block
  b2a_matvec_fn fn = nullptr;
  void *user = nullptr;
  // shift-and-invert (kernels_solve.cuh): inner operator, shift, work vectors d | r | z | p | q (n_local each)
  b2a_op *inner = nullptr;
  double sigma_re = 0.0, sigma_im = 0.0, solve_rtol = 1e-13;
  int solve_maxit = 10000;
  void *solve_work = nullptr;
  b2a::CgState *cg = nullptr;
  double2 *cg_partials = nullptr;
  b2a::CgState *cg_host = nullptr;  // pinned read-back of the solver state
  int64_t solves = 0, solve_iters = 0;
  double solve_worst = 0.0;
};

struct b2a_ws {
  b2a_ctx *ctx = nullptr;
  int dtype = B2A_F64;
  int64_t n_local = 0, n_global = 0, row_offset = 0, ld = 0;
  int maxdim = 0;
  size_t esz = 8;
  void *dV = nullptr;  // ld x (maxdim+1)
  void *arena = nullptr;  // one allocation behind all the small device scratch arrays below
  std::vector<char> H, Q;  // host, column-major: (maxdim+1) x maxdim and maxdim x maxdim
  // device scratch
  void *dH = nullptr;        // (maxdim+1) x maxdim, filled column by column by cgs_finish
  int *dinfo = nullptr;      // per column: bit0 = second pass, bit1 = breakdown
  char *hb1 = nullptr;       // [h1 (j T) | rsq]
  char *hb2 = nullptr;       // [h2 (j T) | w1sq]
  double *w2sq = nullptr;
  void *partials = nullptr;  // T x (maxdim+2) x dots_grid
  double *partials2 = nullptr;
  b2a::SweepState *state = nullptr;
  void *dQ = nullptr;     // maxdim x maxdim
  void *xfull = nullptr;  // n_global (multi-GPU operator input)
  char *pinned = nullptr;  // host staging for read-backs
  size_t pinned_bytes = 0;
  int dots_grid_max = 0, upd_grid_max = 0;
  uint64_t reseed_counter = 0;
  std::vector<int64_t> all_offsets, all_counts;  // row partition over ranks
  bool uniform_partition = true;
  bool use_tma = true;  // TMA-pipelined Gram-Schmidt sweeps (B2A_NO_TMA=1 selects the LDG kernels)
  // in-kernel collectives over NVLink peer memory (peer_comm.cuh); B2A_NO_PEER=1 falls back to NCCL
  b2a::PeerView peer;            // P == 1 when disabled; the block itself is owned by the context
  bool peer_borrowed = false;
  int finish_grid_mult = 4;
  bool push_separate = false;  // experiment: push x with its own kernel instead of inside cgs_finish
  bool peer_x = true;  // fused x push (B2A_PEER_X=0: NCCL all-gather for x, in-kernel all-reduce kept)
  // fused orthogonalisation (kernels_cgs_sweep.cuh): 0 = four kernels per step, 1 (default) = one persistent
  // kernel on single-GPU workspaces, 2 = also on row-sharded ones (B2A_FUSED_SWEEP)
  int fused_sweep = 1;
  unsigned long long *sweep_flag = nullptr;  // release flag of the in-kernel grid barriers (monotone epoch)
  unsigned long long sweep_epoch = 0;
  int sweep_early_trigger = 0;               // experiment (B2A_SWEEP_TRIGGER=1)
  // bit 0: the fused sweep is launched with the PDL attribute (its prologue overlaps the mat-vec's tail), bit 1: the
  // mat-vec after it too (B2A_SWEEP_PDL).  Measured per Arnoldi step at cfg 2: 0 -> 216.7 us, 1 -> 214.9, 2 -> 225.2,
  // 3 -> 222.9 (four-kernel path with PDL: 230.6)
  int sweep_pdl = 1;
  unsigned long long *sweep_trace = nullptr;  // per-CTA phase timestamps of the last fused launch (B2A_SWEEP_TRACE=1)
  int x_pushed_col = -1;  // 0-based column whose normalised content currently sits in every rank's x buffer
  // staged exchange (default on row-sharded workspaces that own the peer block; B2A_XCHG=0 selects the push from
  // inside the normalising kernel): copy-engine transfers on side streams, one flag per owner slice, the mat-vec
  // runs owner group by owner group behind the arrivals.
  bool xchg_staged = false;
  // rotation: Q travels host -> device through two alternating pinned slots (no stream synchronisation per
  // rotation: a slot is re-used only after the event behind its last copy)
  size_t q_bytes = 0, qpin_off = 0;
  cudaEvent_t qev[2] = {nullptr, nullptr};
  bool q_used[2] = {false, false};
  int q_slot = 0;
  int rot_ctas = 0;     // CTAs per SM of the DMMA rotation: 0 = automatic (B2A_ROT_CTAS)
  int rotate_mode = 1;  // 1 = TMA + DMMA kernel (kernels_rotate_mma.cuh), 0 = shared-memory DFMA kernels (B2A_ROTATE=0)
  int tune_rt_dots = 0, tune_rt_upd = 0, tune_stages = 0, tune_ctas = 1, tune_l2promo = 2;  // experiment overrides (env)
};

template <class HT> struct Dev;
template <> struct Dev<double> { using type = double; };
template <> struct Dev<cplx> { using type = cdouble; };

static inline size_t dtype_size(int dtype) { return dtype == B2A_C64 ? 16 : 8; }
static inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ======================================================================= device engine
namespace eng {

// Launch with the programmatic-stream-serialization attribute (PDL, see device_common.cuh).
static bool g_pdl = !(getenv("B2A_PDL") && getenv("B2A_PDL")[0] == '0');
// Scoped override: the mat-vec that follows a fused sweep kernel is a plain stream-ordered launch (measured on
// B200, tools/sweepbench.py: its CTAs scheduled early behind the persistent kernel cost ~10 us per step).
static thread_local bool g_pdl_off = false;
struct PdlScope {
  bool saved;
  explicit PdlScope(bool off) : saved(g_pdl_off) { g_pdl_off = g_pdl_off || off; }
  ~PdlScope() { g_pdl_off = saved; }
};
template <class... KArgs, class... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st,
                              Args &&...args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (g_pdl && !g_pdl_off) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

template <class DT> static inline DT *col(b2a_ws *ws, int c0) {
  return reinterpret_cast<DT *>(ws->dV) + (int64_t)c0 * ws->ld;
}

static int allreduce_f64(b2a_ctx *ctx, void *buf, size_t count) {
  if (ctx->world == 1) return B2A_OK;
  NCCL_TRY(g_nccl.AllReduce(buf, buf, count, ncclFloat64, ncclSum, ctx->comm, ctx->stream));
  return B2A_OK;
}

// ---- Gram-Schmidt launches ------------------------------------------------------
template <class DT, int CPW, int U>
static int launch_dots_inst(b2a_ws *ws, const DT *V, const DT *v, int ncols, DT *hout, double *nrm2,
                            const double *g_rsq, const double *g_w1sq, int gate_step) {
  constexpr int PV = b2a::Scalar<DT>::per_vec;
  const int64_t n = ws->n_local;
  const int64_t quantum = 32 * PV * U;
  int64_t grid = std::max<int64_t>(1, std::min<int64_t>(ws->dots_grid_max, cdiv(n, quantum)));
  const int64_t rows_per_cta = round_up(cdiv(n, grid), quantum);
  grid = std::max<int64_t>(1, cdiv(n, rows_per_cta));
  prof_begin(ws->ctx, B2A_K_DOTS, (double)(ncols + 1) * n * sizeof(DT), gate_step);
  b2a::cgs_dots_kernel<DT, CPW, U><<<(unsigned)grid, b2a::kCgsThreads, 0, ws->ctx->stream>>>(
      V, ws->ld, v, n, ncols, rows_per_cta, reinterpret_cast<DT *>(ws->partials), hout, nrm2,
      &ws->state->ticket[0], &ws->state->poison, g_rsq, g_w1sq);
  prof_end(ws->ctx);
  ws->ctx->launches++;
  CUDA_TRY(cudaGetLastError());
  return B2A_OK;
}

// dots over panel columns [0, ncols) in blocks of at most 64 columns
template <class DT>
static int launch_dots(b2a_ws *ws, int ncols, const DT *v, DT *hout, double *nrm2, const double *g_rsq,
                       const double *g_w1sq, int gate_step = 0) {
  const DT *V = col<DT>(ws, 0);
  int done = 0;
  bool first = true;
  do {
    const int nc = std::min(64, ncols - done);
    const int cpw = std::max(1, (nc + b2a::kCgsWarps - 1) / b2a::kCgsWarps);
    double *nrm = first ? nrm2 : nullptr;
    const DT *Vb = V + (int64_t)done * ws->ld;
    DT *hb = hout + done;
    int s;
    switch (cpw) {
      case 1: s = launch_dots_inst<DT, 1, 8>(ws, Vb, v, nc, hb, nrm, g_rsq, g_w1sq, gate_step); break;
      case 2: s = launch_dots_inst<DT, 2, 8>(ws, Vb, v, nc, hb, nrm, g_rsq, g_w1sq, gate_step); break;
      case 3: s = launch_dots_inst<DT, 3, 4>(ws, Vb, v, nc, hb, nrm, g_rsq, g_w1sq, gate_step); break;
      case 4: s = launch_dots_inst<DT, 4, 4>(ws, Vb, v, nc, hb, nrm, g_rsq, g_w1sq, gate_step); break;
      case 5: s = launch_dots_inst<DT, 5, 2>(ws, Vb, v, nc, hb, nrm, g_rsq, g_w1sq, gate_step); break;
      case 6: s = launch_dots_inst<DT, 6, 2>(ws, Vb, v, nc, hb, nrm, g_rsq, g_w1sq, gate_step); break;
      case 7: s = launch_dots_inst<DT, 7, 2>(ws, Vb, v, nc, hb, nrm, g_rsq, g_w1sq, gate_step); break;
      default: s = launch_dots_inst<DT, 8, 2>(ws, Vb, v, nc, hb, nrm, g_rsq, g_w1sq, gate_step); break;
    }
    B2A_TRY(s);
    done += nc;
    first = false;
  } while (done < ncols);
  return B2A_OK;
}

template <class DT>
static int launch_update(b2a_ws *ws, int ncols, DT *v, const DT *h, double *nrm2, const double *g_rsq,
                         const double *g_w1sq, int gate_step = 0) {
  constexpr int PV = b2a::Scalar<DT>::per_vec;
  const int64_t nvec = cdiv(ws->n_local, PV);
  const int64_t grid = std::max<int64_t>(1, std::min<int64_t>(ws->upd_grid_max, cdiv(nvec, b2a::kCgsThreads)));
  prof_begin(ws->ctx, B2A_K_UPDATE, (double)(ncols + 2) * ws->n_local * sizeof(DT), gate_step);
  b2a::cgs_update_kernel<DT><<<(unsigned)grid, b2a::kCgsThreads, (ncols + 8) * sizeof(DT), ws->ctx->stream>>>(
      col<DT>(ws, 0), ws->ld, v, ws->n_local, ncols, h, ws->partials2, nrm2, &ws->state->ticket[1],
      &ws->state->poison, g_rsq, g_w1sq);
  prof_end(ws->ctx);
  ws->ctx->launches++;
  CUDA_TRY(cudaGetLastError());
  return B2A_OK;
}


// ---- TMA-pipelined Gram-Schmidt sweeps (kernels_cgs_tma.cuh) ---------------------------
static constexpr size_t kTmaSmemBudget = 220 * 1024;

// cuTensorMapEncodeTiled comes from the driver; resolve it through the runtime so that the
// library has no link-time dependency on libcuda (it must load on GPU-less hosts).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      (void)cudaGetLastError();
  }
  return fn;
}

// Tensor map over columns [0, ncols] of V (panel + the column being orthogonalised), box = one tile.
static bool make_panel_tmap(const b2a_ws *ws, int ncols, int RT, CUtensorMap *tm) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return false;
  const cuuint64_t inner = ws->esz / 8;  // ComplexF64 = 2 doubles per row
  if ((cuuint64_t)RT * inner > 256 || ncols + 1 > 256) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)ws->ld * inner, (cuuint64_t)(ncols + 1)};
  cuuint64_t gstride[1] = {(cuuint64_t)ws->ld * ws->esz};
  cuuint32_t box[2] = {(cuuint32_t)(RT * inner), (cuuint32_t)(ncols + 1)};
  cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, ws->dV, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)ws->tune_l2promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ==
         CUDA_SUCCESS;
}

// ring geometry for a sweep over (ncols + 1) columns of `elem`-byte elements
static bool tma_geometry(const b2a_ws *ws, int ncols, size_t elem, bool update, bool reverse, b2a::TmaGeom *g,
                         size_t *smem_bytes, int *grid) {
  if (ncols > b2a::kTmaMaxCols) return false;
  const size_t header = 256 + (update ? b2a::kTmaMaxCols * elem : 0);
  const size_t budget = kTmaSmemBudget / ws->tune_ctas - (ws->tune_ctas > 1 ? 2048 : 0);
  const size_t row_bytes = (size_t)(ncols + 1) * elem;
  const int min_rt = elem == 8 ? 64 : 32;
  const int max_rt = elem == 8 ? 256 : 128;  // tensor-map box: inner dimension <= 256 doubles
  // rows per tile: ~64 KB stages (update: >= 256 rows when two such stages fit, so that all eight
  // consumer warps own a 32-row slab in phase 1); never more rows than the vector has
  int RT = max_rt;
  while (RT > min_rt && (size_t)RT * row_bytes > 64 * 1024) RT >>= 1;
  if (update && RT < 256 && max_rt >= 256 && 2 * 256 * row_bytes + header <= budget) RT = 256;
  const int forced = update ? ws->tune_rt_upd : ws->tune_rt_dots;
  if (forced >= min_rt && forced <= max_rt && (forced & (forced - 1)) == 0) RT = forced;
  while (RT > min_rt && (int64_t)RT / 2 >= ws->n_local) RT >>= 1;
  const size_t stage_bytes = (size_t)RT * row_bytes;
  int stages = (int)std::min<size_t>(b2a::kTmaMaxStages, (budget - header) / stage_bytes);
  if (ws->tune_stages >= 2) stages = std::min(stages, ws->tune_stages);
  if (stages < 2) return false;
  const int64_t ntiles = cdiv(std::max<int64_t>(ws->n_local, 1), RT);
  if (ntiles > 2000000000LL) return false;
  int gr = (int)std::min<int64_t>((int64_t)ws->ctx->num_sms * ws->tune_ctas, ntiles);
  const int tpc = (int)cdiv(ntiles, gr);
  gr = (int)cdiv(ntiles, tpc);
  stages = std::min(stages, std::max(2, tpc));
  g->RT = RT;
  g->stages = stages;
  g->tiles_per_cta = tpc;
  g->ntiles = (int)ntiles;
  g->reverse = reverse ? 1 : 0;
  *smem_bytes = header + (size_t)stages * stage_bytes;
  *grid = gr;
  return true;
}

// Raise a kernel's dynamic shared-memory limit once per context (= per device: the attribute is per device and
// function, and one process may own contexts on several GPUs).
template <class K> static int ensure_smem_attr(b2a_ctx *ctx, K kern, size_t bytes) {
  const void *key = reinterpret_cast<const void *>(kern);
  for (const void *p : ctx->smem_attr_done)
    if (p == key) return B2A_OK;
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  ctx->smem_attr_done.push_back(key);
  return B2A_OK;
}

template <class DT, int CPW>
static int launch_dots_tma_inst(b2a_ws *ws, const DT *v, int ncols, const b2a::TmaGeom &g, size_t smem, int grid,
                                DT *hout, double *nrm2, const double *g_rsq, const double *g_w1sq, int gate_step) {
  auto kern = b2a::cgs_dots_tma_kernel<DT, CPW>;
  B2A_TRY(ensure_smem_attr(ws->ctx, kern, kTmaSmemBudget));
  CUtensorMap tm;
  if (!make_panel_tmap(ws, ncols, g.RT, &tm)) return fail(B2A_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  (void)v;  // v is column `ncols` of the tensor map
  prof_begin(ws->ctx, B2A_K_DOTS, (double)(ncols + 1) * ws->n_local * sizeof(DT), gate_step);
  CUDA_TRY(launch_pdl(kern, (unsigned)grid, (unsigned)b2a::kTmaThreads, smem, ws->ctx->stream, tm, ncols, g,
                      reinterpret_cast<DT *>(ws->partials), hout, nrm2, &ws->state->ticket[2],
                      (const int *)&ws->state->poison, g_rsq, g_w1sq, ws->peer));
  prof_end(ws->ctx);
  ws->ctx->launches++;
  CUDA_TRY(cudaGetLastError());
  return B2A_OK;
}

template <class DT>
static int launch_dots_tma(b2a_ws *ws, int ncols, const DT *v, DT *hout, double *nrm2, const double *g_rsq,
                           const double *g_w1sq, int gate_step, bool reverse) {
  b2a::TmaGeom g;
  size_t smem;
  int grid;
  if (!tma_geometry(ws, ncols, sizeof(DT), false, reverse, &g, &smem, &grid)) return 1;  // caller falls back
  switch (std::max(1, (ncols + 7) / 8)) {
    case 1: return launch_dots_tma_inst<DT, 1>(ws, v, ncols, g, smem, grid, hout, nrm2, g_rsq, g_w1sq, gate_step);
    case 2: return launch_dots_tma_inst<DT, 2>(ws, v, ncols, g, smem, grid, hout, nrm2, g_rsq, g_w1sq, gate_step);
    case 3: return launch_dots_tma_inst<DT, 3>(ws, v, ncols, g, smem, grid, hout, nrm2, g_rsq, g_w1sq, gate_step);
    case 4: return launch_dots_tma_inst<DT, 4>(ws, v, ncols, g, smem, grid, hout, nrm2, g_rsq, g_w1sq, gate_step);
    case 5: return launch_dots_tma_inst<DT, 5>(ws, v, ncols, g, smem, grid, hout, nrm2, g_rsq, g_w1sq, gate_step);
    case 6: return launch_dots_tma_inst<DT, 6>(ws, v, ncols, g, smem, grid, hout, nrm2, g_rsq, g_w1sq, gate_step);
    case 7: return launch_dots_tma_inst<DT, 7>(ws, v, ncols, g, smem, grid, hout, nrm2, g_rsq, g_w1sq, gate_step);
    default: return launch_dots_tma_inst<DT, 8>(ws, v, ncols, g, smem, grid, hout, nrm2, g_rsq, g_w1sq, gate_step);
  }
}

template <class DT, int CPW, bool SPEC>
static int launch_update_tma_inst(b2a_ws *ws, DT *v, int ncols, const b2a::TmaGeom &g, size_t smem, int grid,
                                  const DT *h, DT *cout, double *nrm2, const double *g_rsq, const double *g_w1sq,
                                  int gate_step) {
  auto kern = b2a::cgs_update_tma_kernel<DT, CPW, SPEC>;
  B2A_TRY(ensure_smem_attr(ws->ctx, kern, kTmaSmemBudget));
  CUtensorMap tm;
  if (!make_panel_tmap(ws, ncols, g.RT, &tm)) return fail(B2A_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  prof_begin(ws->ctx, B2A_K_UPDATE, (double)(ncols + 2) * ws->n_local * sizeof(DT), gate_step);
  CUDA_TRY(launch_pdl(kern, (unsigned)grid, (unsigned)b2a::kTmaThreads, smem, ws->ctx->stream, tm, v, ncols, g, h,
                      reinterpret_cast<DT *>(ws->partials), cout, nrm2, &ws->state->ticket[3],
                      (const int *)&ws->state->poison, g_rsq, g_w1sq, ws->peer));
  prof_end(ws->ctx);
  ws->ctx->launches++;
  CUDA_TRY(cudaGetLastError());
  return B2A_OK;
}

// v -= V h (+ ||v||^2); spec: also cout = V' v_new in the same pass
template <class DT>
static int launch_update_tma(b2a_ws *ws, int ncols, DT *v, const DT *h, DT *cout, double *nrm2, bool spec,
                             const double *g_rsq, const double *g_w1sq, int gate_step, bool reverse) {
  b2a::TmaGeom g;
  size_t smem;
  int grid;
  if (!tma_geometry(ws, ncols, sizeof(DT), true, reverse, &g, &smem, &grid)) return 1;
  if (!spec)
    return launch_update_tma_inst<DT, 1, false>(ws, v, ncols, g, smem, grid, h, cout, nrm2, g_rsq, g_w1sq, gate_step);
  switch (std::max(1, (ncols + 7) / 8)) {
    case 1: return launch_update_tma_inst<DT, 1, true>(ws, v, ncols, g, smem, grid, h, cout, nrm2, g_rsq, g_w1sq, gate_step);
    case 2: return launch_update_tma_inst<DT, 2, true>(ws, v, ncols, g, smem, grid, h, cout, nrm2, g_rsq, g_w1sq, gate_step);
    case 3: return launch_update_tma_inst<DT, 3, true>(ws, v, ncols, g, smem, grid, h, cout, nrm2, g_rsq, g_w1sq, gate_step);
    case 4: return launch_update_tma_inst<DT, 4, true>(ws, v, ncols, g, smem, grid, h, cout, nrm2, g_rsq, g_w1sq, gate_step);
    case 5: return launch_update_tma_inst<DT, 5, true>(ws, v, ncols, g, smem, grid, h, cout, nrm2, g_rsq, g_w1sq, gate_step);
    case 6: return launch_update_tma_inst<DT, 6, true>(ws, v, ncols, g, smem, grid, h, cout, nrm2, g_rsq, g_w1sq, gate_step);
    case 7: return launch_update_tma_inst<DT, 7, true>(ws, v, ncols, g, smem, grid, h, cout, nrm2, g_rsq, g_w1sq, gate_step);
    default: return launch_update_tma_inst<DT, 8, true>(ws, v, ncols, g, smem, grid, h, cout, nrm2, g_rsq, g_w1sq, gate_step);
  }
}

// ---- fused orthogonalisation: S1 -> S2 -> [S3] -> finish in ONE persistent kernel (kernels_cgs_sweep.cuh) ----
static bool fused_sweep_on(const b2a_ws *ws) {
  if (!ws->use_tma || ws->tune_ctas != 1 || ws->ctx->num_sms > b2a::kSweepPartStride) return false;
  if (ws->ctx->world == 1) return ws->fused_sweep >= 1;
  // row-sharded: the kernel all-reduces inside its grid barriers, which needs the NVLink peer block (a workspace
  // that fell back to host-launched NCCL collectives keeps the four-kernel path)
  // default (fused_sweep == 1): fused whenever the exchange is staged, i.e. the kernel has nothing to push
  return (ws->fused_sweep >= 2 || (ws->fused_sweep == 1 && ws->xchg_staged)) && ws->peer.P == ws->ctx->world;
}

template <class DT, int CPW>
static int launch_sweep_inst(b2a_ws *ws, int j, int step, const b2a::TmaGeom &g, size_t smem, int grid, int push) {
  auto kern = b2a::cgs_sweep_tma_kernel<DT, CPW>;
  B2A_TRY(ensure_smem_attr(ws->ctx, kern, kTmaSmemBudget + 2048));
  PdlScope pdl_scope(!(ws->sweep_pdl & 1));
  CUtensorMap tm;
  if (!make_panel_tmap(ws, j, g.RT, &tm)) return fail(B2A_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  b2a_ctx *ctx = ws->ctx;
  DT *v = col<DT>(ws, j);
  DT *h1 = reinterpret_cast<DT *>(ws->hb1);
  DT *h2 = reinterpret_cast<DT *>(ws->hb2);
  double *rsq = reinterpret_cast<double *>(h1 + j);
  double *w1sq = reinterpret_cast<double *>(h2 + j);
  DT *Hcol = reinterpret_cast<DT *>(ws->dH) + (int64_t)(j - 1) * (ws->maxdim + 1);
  const double ns = (double)ws->n_local * sizeof(DT);
  // P1 (j+1) n s + P2 (j+2) n s + P4 2 n s; the gated P3 adds (j+2) n s when it runs
  prof_begin(ctx, B2A_K_SWEEP, (2.0 * j + 5.0) * ns);
  if (ctx->prof_on) {
    ctx->prof_pending.back().extra_bytes = (j + 2.0) * ns;
    ctx->prof_pending.back().extra_step = j;
  }
  const unsigned long long epoch = ws->sweep_epoch;
  ws->sweep_epoch += 4;
  CUDA_TRY(launch_pdl(kern, (unsigned)grid, (unsigned)b2a::kTmaThreads, smem, ctx->stream, tm, v, ws->n_local, j, g,
                      reinterpret_cast<DT *>(ws->partials), h1, h2, rsq, w1sq, ws->w2sq, Hcol, ws->dinfo + j, ws->state,
                      ws->sweep_flag, epoch, step, ws->peer, ws->row_offset, push, ws->sweep_early_trigger,
                      ws->sweep_trace));
  prof_end(ctx);
  ctx->launches++;
  CUDA_TRY(cudaGetLastError());
  return B2A_OK;
}

// returns 1 when the geometry does not fit (caller falls back to the four-kernel path)
template <class DT> static int launch_sweep(b2a_ws *ws, int j, int step, int push) {
  b2a::TmaGeom g;
  size_t smem;
  int grid;
  if (!tma_geometry(ws, j, sizeof(DT), true, false, &g, &smem, &grid)) return 1;
  if (grid > ws->ctx->num_sms) return 1;  // the in-kernel grid barrier needs every CTA resident
  // same ring as the update kernel, larger header (two coefficient buffers instead of one)
  smem = smem - (256 + b2a::kTmaMaxCols * sizeof(DT)) + b2a::sweep_header_bytes<DT>();
  switch (std::max(1, (j + 7) / 8)) {
    case 1: return launch_sweep_inst<DT, 1>(ws, j, step, g, smem, grid, push);
    case 2: return launch_sweep_inst<DT, 2>(ws, j, step, g, smem, grid, push);
    case 3: return launch_sweep_inst<DT, 3>(ws, j, step, g, smem, grid, push);
    case 4: return launch_sweep_inst<DT, 4>(ws, j, step, g, smem, grid, push);
    case 5: return launch_sweep_inst<DT, 5>(ws, j, step, g, smem, grid, push);
    case 6: return launch_sweep_inst<DT, 6>(ws, j, step, g, smem, grid, push);
    case 7: return launch_sweep_inst<DT, 7>(ws, j, step, g, smem, grid, push);
    default: return launch_sweep_inst<DT, 8>(ws, j, step, g, smem, grid, push);
  }
}

static bool tma_path_ok(const b2a_ws *ws, int j, size_t elem) {
  b2a::TmaGeom g;
  size_t smem;
  int grid;
  return ws->use_tma && get_encode_tiled() != nullptr && j + 1 <= 256 &&
         tma_geometry(ws, j, elem, false, false, &g, &smem, &grid) &&
         tma_geometry(ws, j, elem, true, false, &g, &smem, &grid);
}

// One orthogonalisation of column index j (0-based; panel = columns 0..j-1).
// mode: 0 Arnoldi step, 1 re-seed, 2 normalise only (j == 0).
template <class DT> static int enqueue_cgs(b2a_ws *ws, int j, int mode, int step) {
  b2a_ctx *ctx = ws->ctx;
  DT *v = col<DT>(ws, j);
  DT *h1 = reinterpret_cast<DT *>(ws->hb1);
  DT *h2 = reinterpret_cast<DT *>(ws->hb2);
  double *rsq = reinterpret_cast<double *>(h1 + j);
  double *w1sq = reinterpret_cast<double *>(h2 + j);
  const size_t hd = (size_t)j * sizeof(DT) / sizeof(double);  // doubles in a j-vector of DT

  const bool tma = tma_path_ok(ws, j, sizeof(DT));
  // multi-GPU: the TMA kernels finish their own all-reduce over NVLink peer memory (peer_comm.cuh);
  // otherwise a host-launched NCCL all-reduce follows each reduction kernel
  const bool fused = tma && ws->peer.P > 1;
  if (mode == 0 && j >= 1 && tma && fused_sweep_on(ws)) {
    // the whole orthogonalisation as one persistent kernel with in-kernel grid barriers
    const int push = (ws->peer.P > 1 && ws->peer_x && !ws->push_separate && !ws->xchg_staged) ? 1 : 0;
    const int s = launch_sweep<DT>(ws, j, step, push);
    if (s == B2A_OK) {
      ws->x_pushed_col = push ? j : -1;
      return B2A_OK;
    }
    if (s != 1) return s;
  }
  if (j == 0 || mode == 2) {
    if (tma)
      B2A_TRY(launch_dots_tma<DT>(ws, 0, v, h1, rsq, nullptr, nullptr, 0, false));
    else
      B2A_TRY(launch_dots<DT>(ws, 0, v, h1, rsq, nullptr, nullptr));
    if (!fused) B2A_TRY(allreduce_f64(ctx, rsq, 1));
    mode = 2;
  } else if (tma) {
    // S1: h = V' v, rnorm^2                                                (expansion.jl:81-84)
    B2A_TRY(launch_dots_tma<DT>(ws, j, v, h1, rsq, nullptr, nullptr, 0, false));
    if (!fused) B2A_TRY(allreduce_f64(ctx, h1, hd + 1));
    // S2: v -= V h, wnorm^2, and speculatively c = V' v_new in the same pass   (expansion.jl:85-88,93)
    B2A_TRY(launch_update_tma<DT>(ws, j, v, h1, h2, w1sq, true, nullptr, nullptr, 0, true));
    if (!fused) B2A_TRY(allreduce_f64(ctx, h2, hd + 1));
    // S3, gated on the device by wnorm < eta * rnorm: v -= V c, wnorm^2       (expansion.jl:91-96)
    B2A_TRY(launch_update_tma<DT>(ws, j, v, h2, h2, ws->w2sq, false, rsq, w1sq, j, false));
    if (!fused) B2A_TRY(allreduce_f64(ctx, ws->w2sq, 1));
  } else {
    // pass 1: h = V' v, rnorm^2 ; v -= V h, wnorm^2          (expansion.jl:81-88)
    B2A_TRY(launch_dots<DT>(ws, j, v, h1, rsq, nullptr, nullptr));
    B2A_TRY(allreduce_f64(ctx, h1, hd + 1));
    B2A_TRY(launch_update<DT>(ws, j, v, h1, w1sq, nullptr, nullptr));
    B2A_TRY(allreduce_f64(ctx, w1sq, 1));
    // pass 2, gated on the device by wnorm < eta * rnorm      (expansion.jl:91-96)
    B2A_TRY(launch_dots<DT>(ws, j, v, h2, nullptr, rsq, w1sq, j));
    B2A_TRY(allreduce_f64(ctx, h2, hd));
    B2A_TRY(launch_update<DT>(ws, j, v, h2, ws->w2sq, rsq, w1sq, j));
    B2A_TRY(allreduce_f64(ctx, ws->w2sq, 1));
  }
  constexpr int PV = b2a::Scalar<DT>::per_vec;
  const int64_t nvec = cdiv(ws->n_local, PV);
  // 4 vectors per thread and pass; at most 4 CTAs per SM (one system-scope fence per CTA when pushing)
  const int64_t grid = std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->num_sms * ws->finish_grid_mult, cdiv(nvec, 256 * 4)));
  DT *Hcol = reinterpret_cast<DT *>(ws->dH) + (int64_t)(std::max(j, 1) - 1) * (ws->maxdim + 1);
  prof_begin(ctx, B2A_K_FINISH, 2.0 * ws->n_local * sizeof(DT));
  const int push = (mode == 0 && ws->peer.P > 1 && ws->peer_x && !ws->push_separate && !ws->xchg_staged) ? 1 : 0;  // Arnoldi step: the new column is the next mat-vec input
  CUDA_TRY(launch_pdl(b2a::cgs_finish_kernel<DT>, (unsigned)grid, 256u, 0, ctx->stream, v, ws->n_local, j,
                      (const DT *)h1, (const DT *)h2, (const double *)rsq, (const double *)w1sq,
                      (const double *)ws->w2sq, Hcol, ws->dinfo + j, ws->state, step, mode, ws->peer, ws->row_offset,
                      push));
  prof_end(ctx);
  ctx->launches++;
  ws->x_pushed_col = push ? j : -1;
  CUDA_TRY(cudaGetLastError());
  return B2A_OK;
}

// ---- operator -------------------------------------------------------------------
static double op_bytes(const b2a_op *A);

using b2a::XWait;

// vector kernel, one launch: rowptr / accumulate select the column block, xw what the launch waits for
template <class DT, int LPR, int U, bool HINT, bool COH, int E = 4>
static cudaError_t launch_spmv_vec_one(b2a_op *A, const int64_t *rowptr, const DT *x, DT *y, const int *poison,
                                       cudaStream_t st, int sms, const XWait &xw, int accumulate) {
  const int64_t threads = cdiv(A->n_local, U) * LPR;
  const int64_t grid = std::max<int64_t>(1, std::min<int64_t>((int64_t)sms * A->grid_mult, cdiv(threads, 256)));
  return launch_pdl(b2a::spmv_csr_vector_kernel<DT, LPR, U, HINT, COH, E>, (unsigned)grid, 256u, 0, st, A->n_local,
                    rowptr, (const int32_t *)A->d_idx, reinterpret_cast<const DT *>(A->d_vals), x, y, poison, xw,
                    accumulate);
}
template <class DT, int LPR, int U>
static cudaError_t launch_spmv_vec_u(b2a_op *A, const DT *x, DT *y, const int *poison, cudaStream_t st, int sms,
                                     const XWait &xw) {
  if (xw.mode) return launch_spmv_vec_one<DT, LPR, 2, false, true>(A, A->d_ptr, x, y, poison, st, sms, xw, 0);
  switch (A->entries_in_flight) {  // B2A_SPMV_E (experiments); results are bit-identical for every E
    case 1: return launch_spmv_vec_one<DT, LPR, U, false, false, 1>(A, A->d_ptr, x, y, poison, st, sms, xw, 0);
    case 2: return launch_spmv_vec_one<DT, LPR, U, false, false, 2>(A, A->d_ptr, x, y, poison, st, sms, xw, 0);
    default: return launch_spmv_vec_one<DT, LPR, U, false, false, 4>(A, A->d_ptr, x, y, poison, st, sms, xw, 0);
  }
}
// column-blocked operator (x larger than L2): one pass per block, L2-hinted loads, the first pass writes y, the
// others accumulate; every pass gathers from the same x, so only the first one waits
template <class DT, int LPR>
static cudaError_t launch_spmv_blocked(b2a_op *A, const DT *x, DT *y, const int *poison, cudaStream_t st, int sms,
                                       const XWait &xw, int64_t *launches) {
  for (int b = 0; b < A->nblocks; ++b) {
    const int64_t *rp = A->d_ptr + (size_t)b * (A->n_local + 1);
    cudaError_t e;
    if (xw.mode) {
      XWait w = xw;
      if (b > 0) w.mode = 0;
      e = launch_spmv_vec_one<DT, LPR, 2, true, true>(A, rp, x, y, poison, st, sms, w, b > 0 ? 1 : 0);
    } else {
      e = launch_spmv_vec_one<DT, LPR, 2, true, false>(A, rp, x, y, poison, st, sms, xw, b > 0 ? 1 : 0);
    }
    if (e != cudaSuccess) return e;
    if (b > 0) ++*launches;
  }
  return cudaSuccess;
}
// operator stored by OWNER GROUP (row-sharded, staged exchange): block b holds the columns of the ranks
// rank + b G .. rank + b G + G - 1 (mod P), i.e. the slices in their order of arrival; one L2-hinted pass per block, each
// waiting only for the slices of its own group, the first writing y and the others accumulating.
template <class DT, int LPR>
static cudaError_t launch_spmv_groups(b2a_op *A, const DT *x_buf, DT *y, const int *poison, cudaStream_t st, int sms,
                                      const XWait &xw, int64_t *launches) {
  const int P = xw.pv.P, G = A->owner_G;
  for (int b = 0; b < A->nblocks; ++b) {
    const int64_t *rp = A->d_ptr + (size_t)b * (A->n_local + 1);
    XWait w = xw;
    w.mode = 4;
    w.owner = b * G;
    w.count = std::min(G, P - b * G);
    const cudaError_t e = launch_spmv_vec_one<DT, LPR, 2, true, true>(A, rp, x_buf, y, poison, st, sms, w, b > 0 ? 1 : 0);
    if (e != cudaSuccess) return e;
    if (b > 0) ++*launches;
  }
  return cudaSuccess;
}

template <class DT, int LPR>
static cudaError_t launch_spmv_vec(b2a_op *A, const DT *x, DT *y, const int *poison, cudaStream_t st, int sms,
                                   const XWait &xw) {
  if (A->nblocks > 1) return launch_spmv_blocked<DT, LPR>(A, x, y, poison, st, sms, xw, &A->ctx->launches);
  switch (A->rows_in_flight) {
    case 1: return launch_spmv_vec_u<DT, LPR, 1>(A, x, y, poison, st, sms, xw);
    case 4: return launch_spmv_vec_u<DT, LPR, 4>(A, x, y, poison, st, sms, xw);
    default: return launch_spmv_vec_u<DT, LPR, 2>(A, x, y, poison, st, sms, xw);
  }
}

template <class DT, int LPC>
static void launch_spmv_csc(b2a_op *A, const DT *x, DT *y, const int *poison, cudaStream_t st, int sms) {
  const int64_t threads = A->n_global * LPC;
  const int64_t grid = std::max<int64_t>(1, std::min<int64_t>((int64_t)sms * 8, cdiv(threads, 256)));
  b2a::spmv_csc_scatter_kernel<DT, LPC><<<(unsigned)grid, 256, 0, st>>>(
      A->n_global, A->d_ptr, A->d_idx, reinterpret_cast<const DT *>(A->d_vals), x, y, poison);
}

template <class DT, int LPR>
static int launch_spmv_tma_inst(b2a_op *A, const DT *x, DT *y, const int *poison, b2a_ctx *ctx) {
  auto kern = b2a::spmv_csr_tma_kernel<DT, LPR>;
  const size_t stage = b2a::SpmvStage<DT>::bytes;
  // two CTAs per SM (two-stage rings): while one CTA reduces rows the other has its gathers in flight
  const int grid = std::min(2 * ctx->num_sms, A->ntiles);
  const int tpc = (int)cdiv(A->ntiles, grid);
  const int g2 = (int)cdiv(A->ntiles, tpc);
  int stages = (int)std::min<size_t>(b2a::kSpmvMaxStages, (kTmaSmemBudget / 2 - 128) / stage);
  if (A->tma_stages >= 2) stages = A->tma_stages;
  stages = std::max(2, std::min(stages, std::max(2, tpc)));
  const size_t smem = 128 + (size_t)stages * stage;
  B2A_TRY(ensure_smem_attr(ctx, kern, kTmaSmemBudget));
  kern<<<g2, b2a::kTmaThreads, smem, ctx->stream>>>(A->n_local, A->d_ptr, A->d_idx,
                                                      reinterpret_cast<const DT *>(A->d_vals), A->d_tile_row, A->ntiles,
                                                      tpc, stages, x, y, poison);
  return B2A_OK;
}
template <class DT> static int launch_spmv_tma(b2a_op *A, const DT *x, DT *y, const int *poison, b2a_ctx *ctx) {
  switch (A->lpr) {
    case 1:
    case 2: return launch_spmv_tma_inst<DT, 2>(A, x, y, poison, ctx);
    case 4: return launch_spmv_tma_inst<DT, 4>(A, x, y, poison, ctx);
    case 8: return launch_spmv_tma_inst<DT, 8>(A, x, y, poison, ctx);
    case 16: return launch_spmv_tma_inst<DT, 16>(A, x, y, poison, ctx);
    default: return launch_spmv_tma_inst<DT, 32>(A, x, y, poison, ctx);
  }
}

// ---- staged x exchange ------------------------------------------------------------------------------------
// Row-sharded mat-vec input (SURVEY 8(e) "x-exchange"): rank s owns rows [off_s, off_s + cnt_s) of the new basis
// vector and every rank needs (in general) all of it.  Stage k = 1 .. P-1 sends this rank's slice to rank
// (rank - k) mod P - a permutation per stage, so every NVLink port carries one slice in and one out at a time -
// as ONE copy-engine transfer (no SM store-issue limit, no SMs taken from the mat-vec) followed by an 8-byte
// copy-engine transfer that publishes the exchange number in the receiver's flag for this sender (stream order
// puts it behind the data).  No kernel is involved on the sending side: the mat-vec launches that spin on the
// flags can occupy every SM slot of a GPU, and a flag-publishing kernel queued behind them would never start
// while its peer waits for that very flag.  The receiver runs its mat-vec owner group by owner group in the same order
// (own block, rank+1, rank+2, ...): the gathers on the blocks that have arrived hide the transfer of the others.
// Two x buffers (exchange-number parity): a rank that is one mat-vec ahead never overwrites what a peer still reads.
static int xchg_ensure_streams(b2a_ctx *ctx, int want) {
  while ((int)ctx->xchg_streams.size() < want) {
    cudaStream_t s;
    int lo = 0, hi = 0;
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CUDA_TRY(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, hi));  // transfers go first
    cudaEvent_t e;
    CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->xchg_streams.push_back(s);
    ctx->xchg_done.push_back(e);
  }
  if (!ctx->xchg_ready) CUDA_TRY(cudaEventCreateWithFlags(&ctx->xchg_ready, cudaEventDisableTiming));
  return B2A_OK;
}

constexpr unsigned long long kXchgTableSize = 1ull << 16;
// (Re)fill the table so that it covers exchange number `seq`; a refill (every 65536 exchanges) drains the streams.
static int xchg_ensure_table(b2a_ctx *ctx, unsigned long long seq) {
  if (ctx->xchg_seq_table && seq >= ctx->xchg_table_base && seq - ctx->xchg_table_base < kXchgTableSize) return B2A_OK;
  if (!ctx->xchg_seq_table)
    CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&ctx->xchg_seq_table), kXchgTableSize * sizeof(unsigned long long)));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  for (auto st : ctx->xchg_streams) CUDA_TRY(cudaStreamSynchronize(st));
  std::vector<unsigned long long> host(kXchgTableSize);
  for (unsigned long long i = 0; i < kXchgTableSize; ++i) host[i] = seq + i;
  CUDA_TRY(cudaMemcpy(ctx->xchg_seq_table, host.data(), kXchgTableSize * sizeof(unsigned long long), cudaMemcpyHostToDevice));
  ctx->xchg_table_base = seq;
  return B2A_OK;
}

// Side streams of the staged exchange.  Measured at N = 8 (profiles/r2_variants_n8_streams.txt): the outgoing 56 MB
// take 145-180 us whatever the stream count (copy-engine rate), but with ONE stream the stages arrive in the order
// the mat-vec consumes them - 224 us per mat-vec and 36.2 ms per solve against 255 us / 40.1 ms with three streams
// (concurrent stages all arrive late) and 271 us / 41.9 ms with seven.
static int g_xchg_nstreams = getenv("B2A_XCHG_STREAMS") ? std::max(1, atoi(getenv("B2A_XCHG_STREAMS"))) : 1;
// B2A_XCHG_CHAIN=1 (experiment): two streams, the data transfer of stage k+1 starts right behind the data transfer of
// stage k (event), so the flag transfers and launch latencies leave the critical path while the stages stay ordered
static bool g_xchg_chain = getenv("B2A_XCHG_CHAIN") && getenv("B2A_XCHG_CHAIN")[0] == '1';

// Enqueue the exchange of workspace column jsrc0; returns the exchange number the consumers wait for.
template <class DT> static int enqueue_xchg(b2a_ws *ws, const DT *xl, unsigned long long *seq_out) {
  b2a_ctx *ctx = ws->ctx;
  const b2a::PeerView &pv = ws->peer;
  const int P = pv.P, me = pv.rank;
  const bool chain = g_xchg_chain && P > 2;
  const int ns = chain ? 2 : std::min(g_xchg_nstreams, P - 1);
  B2A_TRY(xchg_ensure_streams(ctx, ns));
  if (chain && !ctx->xchg_data_ev[0])
    for (int i = 0; i < 2; ++i) CUDA_TRY(cudaEventCreateWithFlags(&ctx->xchg_data_ev[i], cudaEventDisableTiming));
  const unsigned long long seq = ++ctx->x_seq;
  B2A_TRY(xchg_ensure_table(ctx, seq));
  const unsigned long long *seq_src = ctx->xchg_seq_table + (seq - ctx->xchg_table_base);
  const size_t bytes = (size_t)ws->n_local * sizeof(DT);
  const size_t xoff = pv.off_x + (seq & 1ull) * pv.x_stride + (size_t)ws->row_offset * sizeof(DT);
  CUDA_TRY(cudaEventRecord(ctx->xchg_ready, ctx->stream));
  for (int i = 0; i < ns; ++i) CUDA_TRY(cudaStreamWaitEvent(ctx->xchg_streams[i], ctx->xchg_ready, 0));
  // timed from "column final" to "every outgoing stage done": bytes = what this rank sends
  prof_begin(ctx, B2A_K_XCHG, (double)bytes * (P - 1), 0, ctx->xchg_streams[0]);
  for (int k = 1; k < P; ++k) {
    const int dst = (me - k + P) % P;
    cudaStream_t st = ctx->xchg_streams[(k - 1) % ns];
    unsigned long long *flag = reinterpret_cast<unsigned long long *>(pv.peer[dst] + pv.off_flag_xs) + me;
    // slice, then its flag: both copy-engine transfers, ordered by the stream
    if (chain && k > 1) CUDA_TRY(cudaStreamWaitEvent(st, ctx->xchg_data_ev[(k - 2) & 1], 0));
    if (bytes) CUDA_TRY(cudaMemcpyAsync(pv.peer[dst] + xoff, xl, bytes, cudaMemcpyDeviceToDevice, st));
    if (chain) CUDA_TRY(cudaEventRecord(ctx->xchg_data_ev[(k - 1) & 1], st));
    CUDA_TRY(cudaMemcpyAsync(flag, seq_src, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
  }
  for (int i = 0; i < ns; ++i) CUDA_TRY(cudaEventRecord(ctx->xchg_done[i], ctx->xchg_streams[i]));
  if (ctx->prof_on) {
    // e1 of the exchange record: when ALL side streams have finished their stages (the first stream waits for the
    // others - it has nothing else to do), so the record is the duration of this rank's whole outgoing exchange
    for (int i = 1; i < ns; ++i) CUDA_TRY(cudaStreamWaitEvent(ctx->xchg_streams[0], ctx->xchg_done[i], 0));
    for (auto it = ctx->prof_pending.rbegin(); it != ctx->prof_pending.rend(); ++it)
      if (it->kind == B2A_K_XCHG) {
        cudaEventRecord(it->e1, ctx->xchg_streams[0]);
        break;
      }
    CUDA_TRY(cudaEventRecord(ctx->xchg_done[0], ctx->xchg_streams[0]));
  }
  *seq_out = seq;
  return B2A_OK;
}

// plain CSR mat-vec on raw device vectors (inner operator of a shift-and-invert map; single GPU)
template <class DT> static int spmv_plain(b2a_ctx *ctx, b2a_op *A, const DT *x, DT *y, const int *poison) {
  XWait xw;
  cudaError_t le = cudaSuccess;
  switch (A->lpr) {
    case 1: {
      if (A->nblocks > 1) {
        le = launch_spmv_blocked<DT, 2>(A, x, y, poison, ctx->stream, ctx->num_sms, xw, &ctx->launches);
        break;
      }
      const int64_t grid = std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->num_sms * 8, cdiv(A->n_local, 256)));
      b2a::spmv_csr_scalar_kernel<DT, false><<<(unsigned)grid, 256, 0, ctx->stream>>>(
          A->n_local, A->d_ptr, A->d_idx, reinterpret_cast<const DT *>(A->d_vals), x, y, poison, xw, 0);
      break;
    }
    case 2: le = launch_spmv_vec<DT, 2>(A, x, y, poison, ctx->stream, ctx->num_sms, xw); break;
    case 4: le = launch_spmv_vec<DT, 4>(A, x, y, poison, ctx->stream, ctx->num_sms, xw); break;
    case 8: le = launch_spmv_vec<DT, 8>(A, x, y, poison, ctx->stream, ctx->num_sms, xw); break;
    case 16: le = launch_spmv_vec<DT, 16>(A, x, y, poison, ctx->stream, ctx->num_sms, xw); break;
    default: le = launch_spmv_vec<DT, 32>(A, x, y, poison, ctx->stream, ctx->num_sms, xw); break;
  }
  ctx->launches++;
  CUDA_TRY(le);
  CUDA_TRY(cudaGetLastError());
  return B2A_OK;
}

// y = (A - sigma I)^{-1} b by Jacobi-preconditioned CG (kernels_solve.cuh).  Iterations are enqueued in chunks; the
// device raises `done` itself, the host looks once per chunk.
template <class DT> static inline DT host_scalar(double re, double im);
template <> inline double host_scalar<double>(double re, double) { return re; }
template <> inline cdouble host_scalar<cdouble>(double re, double im) { return make_double2(re, im); }

static const bool g_trace = getenv("B2A_TRACE") != nullptr;
#define B2A_TRACE(...)                 \
  do {                                 \
    if (g_trace) {                     \
      fprintf(stderr, "[b2a] " __VA_ARGS__); \
      fputc('\n', stderr);             \
      fflush(stderr);                  \
    }                                  \
  } while (0)

template <class DT> static int enqueue_shift_invert(b2a_ws *ws, b2a_op *S, const DT *b, DT *y) {
  b2a_ctx *ctx = ws->ctx;
  B2A_TRACE("shift-invert solve: n=%lld inner=%p work=%p cg=%p host=%p", (long long)S->n_local, (void *)S->inner,
            S->solve_work, (void *)S->cg, (void *)S->cg_host);
  b2a_op *A = S->inner;
  const int64_t n = S->n_local;
  const int *poison = &ws->state->poison;
  DT *d = reinterpret_cast<DT *>(S->solve_work);
  DT *r = d + n, *z = r + n, *p = z + n, *q = p + n;
  const DT sigma = host_scalar<DT>(S->sigma_re, S->sigma_im);  // (the device helper make_scalar must not be called here)
  const int has_sigma = (S->sigma_re != 0.0 || S->sigma_im != 0.0) ? 1 : 0;
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->num_sms * 4, cdiv(n, b2a::kSolveThreads)));
  prof_begin(ctx, B2A_K_SPMV, 0.0);
  b2a::cg_init_kernel<DT><<<grid, b2a::kSolveThreads, 0, ctx->stream>>>(n, b, d, y, r, z, p, S->cg_partials, S->cg,
                                                                        S->solve_rtol * S->solve_rtol, poison);
  ctx->launches++;
  CUDA_TRY(cudaGetLastError());
  int it = 0;
  bool finished = false;
  while (!finished) {
    const int chunk = std::min(32, S->solve_maxit - it);
    for (int c = 0; c < chunk; ++c) {
      B2A_TRY(spmv_plain<DT>(ctx, A, p, q, poison));
      b2a::cg_pq_kernel<DT><<<grid, b2a::kSolveThreads, 0, ctx->stream>>>(n, p, q, sigma, has_sigma, S->cg_partials,
                                                                          S->cg, poison);
      b2a::cg_update_kernel<DT><<<grid, b2a::kSolveThreads, 0, ctx->stream>>>(n, p, q, d, y, r, z, S->cg_partials,
                                                                              S->cg, poison);
      b2a::cg_p_kernel<DT><<<grid, b2a::kSolveThreads, 0, ctx->stream>>>(n, z, p, S->cg, poison);
      ctx->launches += 3;
    }
    it += chunk;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(S->cg_host, S->cg, sizeof(b2a::CgState), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    B2A_TRACE("  chunk done: it=%d done=%d iters=%d rr=%g bb=%g", it, S->cg_host->done, S->cg_host->iters, S->cg_host->rr,
              S->cg_host->bb);
    finished = S->cg_host->done != 0 || it >= S->solve_maxit;
  }
  prof_end(ctx);
  const b2a::CgState &st = *S->cg_host;
  S->solves += 1;
  S->solve_iters += st.iters;
  const double rel = st.bb > 0.0 ? std::sqrt(st.rr / st.bb) : 0.0;
  S->solve_worst = std::max(S->solve_worst, rel);
  if (st.done == 2) return fail(B2A_ERR_SOLVE, "shift-and-invert: conjugate gradients broke down (p' (A - sigma I) p == 0; is the shifted operator definite?)");
  if (st.done == 0)
    return fail(B2A_ERR_SOLVE, "shift-and-invert: inner solve did not converge in " + std::to_string(S->solve_maxit) +
                                   " iterations (relative residual " + std::to_string(rel) + ")");
  return B2A_OK;
}

// y = A x for the local rows; x_local / y are workspace columns.
template <class DT> static int enqueue_matvec(b2a_ws *ws, b2a_op *A, int jsrc0, int jdst0) {
  b2a_ctx *ctx = ws->ctx;
  const DT *xl = col<DT>(ws, jsrc0);
  DT *y = col<DT>(ws, jdst0);
  const int *poison = &ws->state->poison;
  if (A->kind == OP_CALLBACK) {
    const int rc = A->fn(A->user, xl, y, ws->n_local, ctx->stream);
    if (rc != 0) return fail(B2A_ERR_CALLBACK, "matvec callback returned " + std::to_string(rc));
    return B2A_OK;
  }
  if (A->kind == OP_SHIFT_INVERT) return enqueue_shift_invert<DT>(ws, A, xl, y);
  const DT *x = xl;
  XWait xw;
  bool owner_passes = false;
  int xchg_streams_used = 0;
  if (ctx->world > 1 && ws->peer.P > 1 && ws->xchg_staged) {
    unsigned long long seq = 0;
    B2A_TRY(enqueue_xchg<DT>(ws, xl, &seq));
    xchg_streams_used = (g_xchg_chain && ws->peer.P > 2) ? 2 : std::min(g_xchg_nstreams, ws->peer.P - 1);
    const DT *xbuf = reinterpret_cast<const DT *>(ws->peer.peer[ws->peer.rank] + ws->peer.off_x + (seq & 1ull) * ws->peer.x_stride);
    xw.pv = ws->peer;
    xw.want = seq;
    x = xbuf;
    // own slice into the buffer (stream-ordered copy-engine transfer): every pass gathers from the one buffer
    CUDA_TRY(cudaMemcpyAsync(const_cast<DT *>(xbuf) + ws->row_offset, xl, (size_t)ws->n_local * sizeof(DT),
                             cudaMemcpyDeviceToDevice, ctx->stream));
    if (A->owner_blocks && A->kind == OP_CSR && ws->uniform_partition && ws->all_counts[0] == A->owner_W)
      owner_passes = true;  // one pass per owner group, each waiting for its own slices
    else
      xw.mode = 3;  // single pass (or generic column blocks): wait for all the others first
  } else if (ctx->world > 1 && ws->peer.P > 1 && ws->peer_x) {
    // push model over NVLink peer memory: normally the normalising cgs_finish kernel of the previous step
    // has already pushed this column into every rank's x buffer; otherwise push it explicitly
    if (ws->x_pushed_col != jsrc0) {
      const int64_t grid = std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->num_sms * 4, cdiv(ws->n_local, 256)));
      prof_begin(ctx, B2A_K_FILL, (double)ws->n_local * sizeof(DT) * ws->peer.P);
      b2a::xpush_kernel<DT><<<(unsigned)grid, 256, 0, ctx->stream>>>(xl, ws->n_local, ws->state, ws->peer, ws->row_offset);
      prof_end(ctx);
      ctx->launches++;
      CUDA_TRY(cudaGetLastError());
      ws->x_pushed_col = jsrc0;
    }
    xw.pv = ws->peer;
    xw.mode = 1;
    x = reinterpret_cast<const DT *>(ws->peer.peer[ws->peer.rank] + ws->peer.off_x);
  } else if (ctx->world > 1) {
    // x-exchange: every shard needs (in general) all of x.  Uniform partitions use one
    // all-gather; ragged ones a group of broadcasts.
    DT *xf = reinterpret_cast<DT *>(ws->xfull);
    const size_t dpe = sizeof(DT) / sizeof(double);
    if (ws->uniform_partition) {
      NCCL_TRY(g_nccl.AllGather(xl, xf, (size_t)ws->all_counts[0] * dpe, ncclFloat64, ctx->comm, ctx->stream));
    } else {
      NCCL_TRY(g_nccl.GroupStart());
      for (int r = 0; r < ctx->world; ++r)
        NCCL_TRY(g_nccl.Broadcast(xl, xf + ws->all_offsets[r], (size_t)ws->all_counts[r] * dpe, ncclFloat64, r,
                                  ctx->comm, ctx->stream));
      NCCL_TRY(g_nccl.GroupEnd());
    }
    x = xf;
  }
  PdlScope pdl_scope(fused_sweep_on(ws) && !(ws->sweep_pdl & 2));
  prof_begin(ctx, B2A_K_SPMV, op_bytes(A));
  cudaError_t le = cudaSuccess;
  if (A->kind == OP_CSC_SCATTER) {
    b2a::zero_vector_kernel<DT><<<(unsigned)std::min<int64_t>(ctx->num_sms * 8, std::max<int64_t>(1, cdiv(A->n_local, 256))), 256, 0, ctx->stream>>>(y, A->n_local, poison);
    ctx->launches++;
    switch (A->lpr) {
      case 1:
      case 2: launch_spmv_csc<DT, 2>(A, x, y, poison, ctx->stream, ctx->num_sms); break;
      case 4: launch_spmv_csc<DT, 4>(A, x, y, poison, ctx->stream, ctx->num_sms); break;
      case 8: launch_spmv_csc<DT, 8>(A, x, y, poison, ctx->stream, ctx->num_sms); break;
      case 16: launch_spmv_csc<DT, 16>(A, x, y, poison, ctx->stream, ctx->num_sms); break;
      default: launch_spmv_csc<DT, 32>(A, x, y, poison, ctx->stream, ctx->num_sms); break;
    }
  } else if (owner_passes) {
    switch (A->lpr) {
      case 1:
      case 2: le = launch_spmv_groups<DT, 2>(A, x, y, poison, ctx->stream, ctx->num_sms, xw, &ctx->launches); break;
      case 4: le = launch_spmv_groups<DT, 4>(A, x, y, poison, ctx->stream, ctx->num_sms, xw, &ctx->launches); break;
      case 8: le = launch_spmv_groups<DT, 8>(A, x, y, poison, ctx->stream, ctx->num_sms, xw, &ctx->launches); break;
      case 16: le = launch_spmv_groups<DT, 16>(A, x, y, poison, ctx->stream, ctx->num_sms, xw, &ctx->launches); break;
      default: le = launch_spmv_groups<DT, 32>(A, x, y, poison, ctx->stream, ctx->num_sms, xw, &ctx->launches); break;
    }
  } else if (A->use_tma && ctx->world == 1 && A->nblocks == 1) {
    B2A_TRY(launch_spmv_tma<DT>(A, x, y, poison, ctx));
  } else {
    switch (A->lpr) {
      case 1: {
        const int64_t grid = std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->num_sms * 8, cdiv(A->n_local, 256)));
        if (A->nblocks > 1) {
          le = launch_spmv_blocked<DT, 2>(A, x, y, poison, ctx->stream, ctx->num_sms, xw, &ctx->launches);
          break;
        }
        if (xw.mode)
          b2a::spmv_csr_scalar_kernel<DT, true><<<(unsigned)grid, 256, 0, ctx->stream>>>(
              A->n_local, A->d_ptr, A->d_idx, reinterpret_cast<const DT *>(A->d_vals), x, y, poison, xw, 0);
        else
          b2a::spmv_csr_scalar_kernel<DT, false><<<(unsigned)grid, 256, 0, ctx->stream>>>(
              A->n_local, A->d_ptr, A->d_idx, reinterpret_cast<const DT *>(A->d_vals), x, y, poison, xw, 0);
        break;
      }
      case 2: le = launch_spmv_vec<DT, 2>(A, x, y, poison, ctx->stream, ctx->num_sms, xw); break;
      case 4: le = launch_spmv_vec<DT, 4>(A, x, y, poison, ctx->stream, ctx->num_sms, xw); break;
      case 8: le = launch_spmv_vec<DT, 8>(A, x, y, poison, ctx->stream, ctx->num_sms, xw); break;
      case 16: le = launch_spmv_vec<DT, 16>(A, x, y, poison, ctx->stream, ctx->num_sms, xw); break;
      default: le = launch_spmv_vec<DT, 32>(A, x, y, poison, ctx->stream, ctx->num_sms, xw); break;
    }
  }
  prof_end(ctx);
  ctx->launches++;
  CUDA_TRY(le);
  CUDA_TRY(cudaGetLastError());
  // whatever follows on the main stream may overwrite the column that is still being sent: order it behind the
  // outgoing transfers (they finish while the gathers above wait for the incoming ones)
  for (int i = 0; i < xchg_streams_used; ++i) CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->xchg_done[i], 0));
  return B2A_OK;
}


// ---- basis rotation ---------------------------------------------------------------
template <class DT, int R>
static bool try_rotate(b2a_ws *ws, int col0, int K, int N, int move_src, int move_dst) {
  const int Npad = (N + 3) & ~3;
  const size_t smem = ((size_t)(K + 1) * R + (size_t)K * Npad) * sizeof(DT);
  if (smem > 200 * 1024) return false;
  auto kern = b2a::rotate_basis_kernel<DT, R>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    (void)cudaGetLastError();
    return false;
  }
  const int64_t grid = std::max<int64_t>(1, cdiv(ws->n_local, R));
  prof_begin(ws->ctx, B2A_K_ROTATE, (double)ws->n_local * sizeof(DT) * (K + N + (move_dst >= 0 ? 2.0 : 0.0)));
  kern<<<(unsigned)grid, 256, smem, ws->ctx->stream>>>(reinterpret_cast<DT *>(ws->dV), ws->ld, ws->n_local, col0, K,
                                                         N, reinterpret_cast<const DT *>(ws->dQ), move_src, move_dst);
  prof_end(ws->ctx);
  ws->ctx->launches++;
  return true;
}

template <class DT, int NB>
static bool try_rotate2_inst(b2a_ws *ws, int col0, int K, int N, int move_src, int move_dst, int per, int nchunks) {
  const size_t smem = ((size_t)(K + 1) * b2a::kRot2Rows + (size_t)4 * nchunks * K * 8) * sizeof(DT);
  if (smem > 200 * 1024) return false;
  auto kern = b2a::rotate_basis2_kernel<DT, NB>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    (void)cudaGetLastError();
    return false;
  }
  const int64_t grid = std::max<int64_t>(1, cdiv(ws->n_local, b2a::kRot2Rows));
  prof_begin(ws->ctx, B2A_K_ROTATE, (double)ws->n_local * sizeof(DT) * (K + N + (move_dst >= 0 ? 2.0 : 0.0)));
  kern<<<(unsigned)grid, 256, smem, ws->ctx->stream>>>(reinterpret_cast<DT *>(ws->dV), ws->ld, col0, K, N,
                                                         reinterpret_cast<const DT *>(ws->dQ), move_src, move_dst, per,
                                                         nchunks);
  prof_end(ws->ctx);
  ws->ctx->launches++;
  return true;
}
// register-blocked rotation: outputs split over 4 thread groups, NB (<= 8) outputs per chunk
template <class DT> static bool try_rotate2(b2a_ws *ws, int col0, int K, int N, int move_src, int move_dst) {
  if (getenv("B2A_ROTATE_V1")) return false;
  // ComplexF64 is FP64-FMA bound either way (23 TFLOP/s at K = 60, N = 45) and measured slightly faster with
  // the 1-row x 4-output kernel (1874 vs 2009 us at n = 2e6); Float64 gains from the halved smem traffic
  // (152 -> 127 us at cfg 2, 1325 -> 1204 us at 256^3)
  if (b2a::Scalar<DT>::is_complex) return false;
  const int per = (N + 3) / 4;
  const int nchunks = (per + 7) / 8;
  const int nb = (per + nchunks - 1) / nchunks;
  switch (nb) {
    case 1: return try_rotate2_inst<DT, 1>(ws, col0, K, N, move_src, move_dst, per, nchunks);
    case 2: return try_rotate2_inst<DT, 2>(ws, col0, K, N, move_src, move_dst, per, nchunks);
    case 3: return try_rotate2_inst<DT, 3>(ws, col0, K, N, move_src, move_dst, per, nchunks);
    case 4: return try_rotate2_inst<DT, 4>(ws, col0, K, N, move_src, move_dst, per, nchunks);
    case 5: return try_rotate2_inst<DT, 5>(ws, col0, K, N, move_src, move_dst, per, nchunks);
    case 6: return try_rotate2_inst<DT, 6>(ws, col0, K, N, move_src, move_dst, per, nchunks);
    case 7: return try_rotate2_inst<DT, 7>(ws, col0, K, N, move_src, move_dst, per, nchunks);
    default: return try_rotate2_inst<DT, 8>(ws, col0, K, N, move_src, move_dst, per, nchunks);
  }
}

// ---- TMA + DMMA rotation (kernels_rotate_mma.cuh) ------------------------------------------------------------
struct RotPlan {
  b2a::RotGeom g;
  int NT = 4, NCH = 1, KS = 1, b_in_smem = 1, grid = 1, ctas = 1;
  size_t smem = 0, b_elems = 0;
};
// Largest row tile (16 rows per consumer warp) that leaves at least two ring stages, Q fragments in shared memory
// when they fit beside them.
static bool rot_plan(const b2a_ws *ws, int K, int N, bool cplx, RotPlan *p, bool cplx_out = false) {
  const int inner = cplx ? 2 : 1, parts = cplx ? 2 : 1;
  const int ntiles_n = (N * ((cplx || cplx_out) ? 2 : 1) + 7) / 8;  // 8-wide output tiles of the real view
  int best_nt = 4, best_pad = 1 << 30;
  for (int nt = 4; nt >= 1; --nt) {
    const int pad = (int)cdiv(ntiles_n, nt) * nt;
    if (pad < best_pad) {
      best_pad = pad;
      best_nt = nt;
    }
  }
  p->NT = best_nt;
  p->NCH = (int)cdiv(ntiles_n, best_nt);
  p->KS = (K + 3) / 4;
  p->b_elems = (size_t)p->KS * p->NCH * p->NT * parts * 32;
  const int nbox = (int)cdiv(K, 256);
  const int box_cols = (int)round_up(cdiv(K, nbox), 4);
  const int kpad = nbox * box_cols;
  const size_t b_bytes = (p->b_elems * 8 + 127) / 128 * 128;
  // two CTAs per SM (16 consumer warps to cover tile-boundary bubbles) when two full-size tiles with two stages each
  // and Q fit twice; else one CTA with a deeper ring.  B2A_ROT_CTAS = 1 / 2 forces.
  int want_ctas = ws->rot_ctas;
  const size_t full_tile = (size_t)kpad * 128 * inner * 8;
  if (want_ctas == 0) want_ctas = (2 * full_tile + b_bytes + 256 + 1024 <= kTmaSmemBudget / 2) ? 2 : 1;
  p->ctas = want_ctas;
  const size_t budget = kTmaSmemBudget / want_ctas - 256 - (want_ctas > 1 ? 1024 : 0);
  const int order[8][2] = {{1, 8}, {1, 4}, {0, 8}, {0, 4}, {1, 2}, {0, 2}, {1, 1}, {0, 1}};
  for (const auto &o : order) {
    const int bsm = o[0], warps = o[1], R = 16 * warps;
    if ((int64_t)R / 2 >= std::max<int64_t>(ws->n_local, 1) && warps > 1) continue;  // tiny vectors: smaller tiles
    const size_t stage = (size_t)kpad * R * inner * 8;
    const size_t avail = budget - (bsm ? std::min(b_bytes, budget) : 0);
    int stages = (int)std::min<size_t>(b2a::kTmaMaxStages, avail / stage);
    if (bsm && b_bytes >= budget) continue;
    if (stages < 2) continue;
    const int64_t ntiles = cdiv(std::max<int64_t>(ws->n_local, 1), R);
    if (ntiles > 2000000000LL) return false;
    int gr = (int)std::min<int64_t>((int64_t)ws->ctx->num_sms * want_ctas, ntiles);
    const int tpc = (int)cdiv(ntiles, gr);
    gr = (int)cdiv(ntiles, tpc);
    stages = std::min(stages, std::max(2, tpc));
    p->g.R = R;
    p->g.warps = warps;
    p->g.stages = stages;
    p->g.tiles_per_cta = tpc;
    p->g.ntiles = (int)ntiles;
    p->g.nbox = nbox;
    p->g.box_cols = box_cols;
    p->g.kpad = kpad;
    p->b_in_smem = bsm;
    p->grid = gr;
    p->smem = 256 + (bsm ? b_bytes : 0) + (size_t)stages * stage;
    return true;
  }
  return false;
}

// Tensor map over the K input columns [col0, col0 + K) of V, box = R rows x box_cols columns.
static bool make_rot_tmap(const b2a_ws *ws, int col0, int K, const b2a::RotGeom &g, CUtensorMap *tm) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return false;
  const cuuint64_t inner = ws->esz / 8;
  cuuint64_t gdim[2] = {(cuuint64_t)ws->ld * inner, (cuuint64_t)K};
  cuuint64_t gstride[1] = {(cuuint64_t)ws->ld * ws->esz};
  cuuint32_t box[2] = {(cuuint32_t)(g.R * inner), (cuuint32_t)g.box_cols};
  cuuint32_t estr[2] = {1, 1};
  void *base = reinterpret_cast<char *>(ws->dV) + (size_t)col0 * ws->ld * ws->esz;
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ==
         CUDA_SUCCESS;
}

// Q (K x N packed, column-major) in MMA B-fragment order, see kernels_rotate_mma.cuh
template <class HT> static void pack_b_fragments(const HT *Qp, int K, int N, const RotPlan &p, double *out);
template <> void pack_b_fragments<double>(const double *Qp, int K, int N, const RotPlan &p, double *out) {
  const int NTT = p.NCH * p.NT;
  for (int ks = 0; ks < p.KS; ++ks)
    for (int nt = 0; nt < NTT; ++nt)
      for (int lane = 0; lane < 32; ++lane) {
        const int k = 4 * ks + (lane & 3), n = 8 * nt + (lane >> 2);
        out[((size_t)ks * NTT + nt) * 32 + lane] = (k < K && n < N) ? Qp[(size_t)n * K + k] : 0.0;
      }
}
template <> void pack_b_fragments<cplx>(const cplx *Qp, int K, int N, const RotPlan &p, double *out) {
  const int NTT = p.NCH * p.NT;
  for (int ks = 0; ks < p.KS; ++ks)
    for (int nt = 0; nt < NTT; ++nt)
      for (int lane = 0; lane < 32; ++lane) {
        const int c = 4 * ks + (lane & 3), np = 8 * nt + (lane >> 2);
        const int o = np >> 1, comp = np & 1;
        const cplx q = (c < K && o < N) ? Qp[(size_t)o * K + c] : cplx(0.0, 0.0);
        double *dst = out + ((size_t)ks * NTT + nt) * 64 + lane;
        dst[0] = comp == 0 ? q.real() : q.imag();    // multiplies Re V
        dst[32] = comp == 0 ? -q.imag() : q.real();  // multiplies Im V
      }
}

// out == nullptr: in place (the rotation); else the product goes to out[:, 0 : N) with leading dimension ldo
template <bool CPLX, bool CPLX_OUT, int NT, int MINB>
static int launch_rotate_mma_inst2(b2a_ws *ws, const CUtensorMap &tm, int col0, int N, const RotPlan &p, int move_src,
                                   int move_dst, double *out, int64_t ldo) {
  auto kern = b2a::rotate_mma_kernel<CPLX, CPLX_OUT, NT, MINB>;
  B2A_TRY(ensure_smem_attr(ws->ctx, kern, kTmaSmemBudget / MINB));
  double *V = reinterpret_cast<double *>(ws->dV);
  kern<<<(unsigned)p.grid, b2a::kRotThreads, p.smem, ws->ctx->stream>>>(
      tm, V, ws->ld, out ? out : V, out ? ldo : ws->ld, out ? 0 : col0, N, reinterpret_cast<const double *>(ws->dQ),
      p.KS, p.NCH, p.g, move_src, move_dst, p.b_in_smem);
  CUDA_TRY(cudaGetLastError());
  return B2A_OK;
}
template <bool CPLX, bool CPLX_OUT, int NT>
static int launch_rotate_mma_inst(b2a_ws *ws, const CUtensorMap &tm, int col0, int N, const RotPlan &p, int move_src,
                                  int move_dst, double *out = nullptr, int64_t ldo = 0) {
  if (p.ctas > 1) return launch_rotate_mma_inst2<CPLX, CPLX_OUT, NT, 2>(ws, tm, col0, N, p, move_src, move_dst, out, ldo);
  return launch_rotate_mma_inst2<CPLX, CPLX_OUT, NT, 1>(ws, tm, col0, N, p, move_src, move_dst, out, ldo);
}
template <bool CPLX, bool CPLX_OUT>
static int launch_rotate_mma_nt(b2a_ws *ws, const CUtensorMap &tm, int col0, int N, const RotPlan &p, int move_src,
                                int move_dst, double *out = nullptr, int64_t ldo = 0) {
  switch (p.NT) {
    case 1: return launch_rotate_mma_inst<CPLX, CPLX_OUT, 1>(ws, tm, col0, N, p, move_src, move_dst, out, ldo);
    case 2: return launch_rotate_mma_inst<CPLX, CPLX_OUT, 2>(ws, tm, col0, N, p, move_src, move_dst, out, ldo);
    case 3: return launch_rotate_mma_inst<CPLX, CPLX_OUT, 3>(ws, tm, col0, N, p, move_src, move_dst, out, ldo);
    default: return launch_rotate_mma_inst<CPLX, CPLX_OUT, 4>(ws, tm, col0, N, p, move_src, move_dst, out, ldo);
  }
}

// host -> device copy of a small array through one of the two pinned Q slots, no stream synchronisation
static int stage_q(b2a_ws *ws, size_t bytes, double **host_slot) {
  if (bytes > ws->q_bytes) return fail(B2A_ERR_INTERNAL, "rotation: Q staging slot too small");
  const int slot = ws->q_slot ^= 1;
  if (ws->q_used[slot]) CUDA_TRY(cudaEventSynchronize(ws->qev[slot]));
  *host_slot = reinterpret_cast<double *>(ws->pinned + ws->qpin_off + (size_t)slot * ws->q_bytes);
  return B2A_OK;
}
static int stage_q_send(b2a_ws *ws, const double *host_slot, size_t bytes) {
  CUDA_TRY(cudaMemcpyAsync(ws->dQ, host_slot, bytes, cudaMemcpyHostToDevice, ws->ctx->stream));
  CUDA_TRY(cudaEventRecord(ws->qev[ws->q_slot], ws->ctx->stream));
  ws->q_used[ws->q_slot] = true;
  return B2A_OK;
}

// returns 1 when this path cannot take the shape (caller falls back to the shared-memory DFMA kernels)
template <class HT>
static int rotate_mma(b2a_ws *ws, int col0, int K, int N, const HT *Qp, int move_src, int move_dst) {
  using DT = typename Dev<HT>::type;
  constexpr bool CPLX = b2a::Scalar<DT>::is_complex;
  if (!ws->rotate_mode || !ws->use_tma || get_encode_tiled() == nullptr) return 1;
  RotPlan p;
  if (!rot_plan(ws, K, N, CPLX, &p)) return 1;
  CUtensorMap tm;
  if (!make_rot_tmap(ws, col0, K, p.g, &tm)) return 1;
  double *slot = nullptr;
  B2A_TRY(stage_q(ws, p.b_elems * 8, &slot));
  pack_b_fragments<HT>(Qp, K, N, p, slot);
  B2A_TRY(stage_q_send(ws, slot, p.b_elems * 8));
  prof_begin(ws->ctx, B2A_K_ROTATE, (double)ws->n_local * sizeof(DT) * (K + N + (move_dst >= 0 ? 2.0 : 0.0)));
  const int s = launch_rotate_mma_nt<CPLX, CPLX>(ws, tm, col0, N, p, move_src, move_dst);
  prof_end(ws->ctx);
  ws->ctx->launches++;
  return s;
}

// X[:, 0 : N) = V[:, 0 : K) * Y with a COMPLEX K x N coefficient matrix Y (packed, column-major, host) into the
// device buffer dX (complex, leading dimension ldx >= ws->ld rows: whole tiles are written) - partialeigen's Q * Y
// (src/eigvals.jl:94) on the same TMA + DMMA kernel.  Returns 1 when the shape cannot be planned.
template <class HT> static int basis_times_mma(b2a_ws *ws, int K, int N, const cplx *Yp, cdouble *dX, int64_t ldx) {
  using DT = typename Dev<HT>::type;
  constexpr bool CPLX = b2a::Scalar<DT>::is_complex;
  if (!ws->rotate_mode || !ws->use_tma || get_encode_tiled() == nullptr) return 1;
  RotPlan p;
  if (!rot_plan(ws, K, N, CPLX, &p, true)) return 1;
  CUtensorMap tm;
  if (!make_rot_tmap(ws, 0, K, p.g, &tm)) return 1;
  double *slot = nullptr;
  if (p.b_elems * 8 > ws->q_bytes) return 1;
  B2A_TRY(stage_q(ws, p.b_elems * 8, &slot));
  if (CPLX) {
    pack_b_fragments<cplx>(Yp, K, N, p, slot);
  } else {
    // real basis: B = Y as interleaved (re, im) real columns, K x 2N
    std::vector<double> Yr((size_t)K * 2 * N);
    for (int o = 0; o < N; ++o)
      for (int c = 0; c < K; ++c) {
        Yr[(size_t)(2 * o) * K + c] = Yp[(size_t)o * K + c].real();
        Yr[(size_t)(2 * o + 1) * K + c] = Yp[(size_t)o * K + c].imag();
      }
    pack_b_fragments<double>(Yr.data(), K, 2 * N, p, slot);
  }
  B2A_TRY(stage_q_send(ws, slot, p.b_elems * 8));
  const int s = launch_rotate_mma_nt<CPLX, true>(ws, tm, 0, N, p, -1, -1, reinterpret_cast<double *>(dX), ldx);
  ws->ctx->launches++;
  return s;
}

// V[:, col0 : col0+N) <- V[:, col0 : col0+K) * Qp  (Qp = K x N packed, host), optional column move.
// Asynchronous: the kernel is enqueued, Qp may be freed on return (it has been copied into a pinned slot).
template <class HT>
static int rotate(b2a_ws *ws, int col0, int K, int N, const HT *Qp, int move_src, int move_dst) {
  using DT = typename Dev<HT>::type;
  if (N <= 0 || K <= 0) return B2A_OK;
  ws->x_pushed_col = -1;
  if (move_src == move_dst) move_src = move_dst = -1;
  {
    const int s = rotate_mma<HT>(ws, col0, K, N, Qp, move_src, move_dst);
    if (s != 1) return s;
  }
  // fallback: Q as it is (K x N) through a pinned slot, shared-memory DFMA kernels
  double *slot = nullptr;
  const size_t qb = (size_t)K * N * sizeof(HT);
  B2A_TRY(stage_q(ws, qb, &slot));
  std::memcpy(slot, Qp, qb);
  B2A_TRY(stage_q_send(ws, slot, qb));
  bool ok = try_rotate2<DT>(ws, col0, K, N, move_src, move_dst);
  if (!ok) ok = try_rotate<DT, 128>(ws, col0, K, N, move_src, move_dst);
  if (!ok) ok = try_rotate<DT, 64>(ws, col0, K, N, move_src, move_dst);
  if (!ok) ok = try_rotate<DT, 32>(ws, col0, K, N, move_src, move_dst);
  if (!ok) return fail(B2A_ERR_ARGUMENT, "basis rotation: Krylov dimension too large for one shared-memory tile");
  CUDA_TRY(cudaGetLastError());
  return B2A_OK;
}

static double cgs_pass_bytes(const b2a_ws *ws, int j) {
  return (2.0 * j + 3.0) * (double)ws->n_local * (double)ws->esz;  // B_cgs(j)
}
static double scal_bytes(const b2a_ws *ws) { return 2.0 * (double)ws->n_local * (double)ws->esz; }

static double op_bytes(const b2a_op *A) {
  if (A->kind == OP_CALLBACK || A->kind == OP_SHIFT_INVERT) return 0.0;
  const double s = (double)dtype_size(A->dtype);
  const double nptr = (A->kind == OP_CSC_SCATTER ? A->n_global : A->n_local) + 1.0;
  return (double)A->nnz * (s + 4.0) + 8.0 * nptr + 2.0 * (double)A->n_local * s;
}

static uint64_t reseed_key(uint64_t seed, uint64_t counter) {
  return b2a::splitmix64(seed ^ (0x632BE59BD9B4E019ull * (counter + 1)));
}

template <class DT> static int enqueue_fill(b2a_ws *ws, int j0, uint64_t key) {
  const int64_t grid = std::max<int64_t>(1, std::min<int64_t>((int64_t)ws->ctx->num_sms * 8, cdiv(ws->n_local, 256)));
  prof_begin(ws->ctx, B2A_K_FILL, (double)ws->n_local * sizeof(DT));
  b2a::fill_uniform_kernel<DT><<<(unsigned)grid, 256, 0, ws->ctx->stream>>>(col<DT>(ws, j0), ws->n_local,
                                                                            ws->row_offset, key, ws->ctx->d_zero);
  prof_end(ws->ctx);
  ws->ctx->launches++;
  CUDA_TRY(cudaGetLastError());
  return B2A_OK;
}

struct Readback {
  int poison;
  unsigned long long second_passes;
};

// reinitialize!(arnoldi, j, populate!) for 0-based target column j.
template <class HT> static int reinitialize(b2a_ws *ws, int j, int mode, uint64_t seed, int *ok, b2a_stats *st) {
  using DT = typename Dev<HT>::type;
  b2a_ctx *ctx = ws->ctx;
  const int64_t l0 = ctx->launches;
  if (mode == B2A_INIT_RAND) B2A_TRY(enqueue_fill<DT>(ws, j, reseed_key(seed, ws->reseed_counter++)));
  B2A_TRY(enqueue_cgs<DT>(ws, j, j == 0 ? 2 : 1, 0));
  int info = 0;
  CUDA_TRY(cudaMemcpyAsync(ws->pinned, ws->dinfo + j, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  std::memcpy(&info, ws->pinned, sizeof(int));
  prof_collect(ctx, &info, j, 1);
  if (ok) *ok = (j == 0) ? 1 : ((info & 2) ? 0 : 1);
  if (st) {
    const int passes = (j == 0) ? 0 : ((info & 1) ? 2 : 1);
    st->passes += passes;
    st->bytes += passes * cgs_pass_bytes(ws, j) + scal_bytes(ws);
    st->launches += ctx->launches - l0;
  }
  return B2A_OK;
}

// iterate_arnoldi!(A, arnoldi, from:to) - 1-based steps as in the reference.
template <class HT>
static int iterate_arnoldi(b2a_ws *ws, b2a_op *A, int from, int to, uint64_t seed, b2a_stats *st) {
  using DT = typename Dev<HT>::type;
  b2a_ctx *ctx = ws->ctx;
  const int m1 = ws->maxdim + 1;
  HT *H = reinterpret_cast<HT *>(ws->H.data());
  int j = from;
  while (j <= to) {
    const int64_t l0 = ctx->launches;
    for (int s = j; s <= to; ++s) {
      B2A_TRY(enqueue_matvec<DT>(ws, A, s - 1, s));  // V[:, s+1] = A V[:, s]   (expansion.jl:121)
      B2A_TRY(enqueue_cgs<DT>(ws, s, 0, s));         // orthogonalize!(arnoldi, s)  (expansion.jl:127)
    }
    // one read-back per sweep: new H columns, per-column info, sweep state
    const int ncols = to - j + 1;
    char *p = ws->pinned;
    const size_t hbytes = (size_t)ncols * m1 * sizeof(HT);
    CUDA_TRY(cudaMemcpyAsync(p, reinterpret_cast<char *>(ws->dH) + (size_t)(j - 1) * m1 * sizeof(HT), hbytes,
                             cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(p + hbytes, ws->dinfo + j, ncols * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(p + hbytes + ncols * sizeof(int), ws->state, sizeof(b2a::SweepState),
                             cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    b2a::SweepState state;
    std::memcpy(&state, p + hbytes + ncols * sizeof(int), sizeof(state));
    const int *info = reinterpret_cast<const int *>(p + hbytes);
    prof_collect(ctx, info, j, ncols);
    if (state.error) return fail(B2A_ERR_CUDA, "fused orthogonalisation: grid barrier timed out");
    const int last = state.poison ? state.poison : to;  // last step that really executed
    const HT *Hn = reinterpret_cast<const HT *>(p);
    for (int s = j; s <= last; ++s) {
      for (int i = 0; i <= s; ++i) H[(size_t)(s - 1) * m1 + i] = Hn[(size_t)(s - j) * m1 + i];
      if (st) {
        const int passes = (info[s - j] & 1) ? 2 : 1;
        st->matvecs += 1;
        st->passes += passes;
        st->second_passes += (info[s - j] & 1);
        st->bytes += op_bytes(A) + passes * cgs_pass_bytes(ws, s) + ((info[s - j] & 2) ? 0.0 : scal_bytes(ws));
      }
    }
    if (st) st->launches += ctx->launches - l0;
    if (!state.poison) break;
    // breakdown at step `last`: H[last+1, last] = 0 already; re-seed unless j == size(V,1)
    if (st) st->breakdowns += 1;
    ws->x_pushed_col = -1;  // the broken step pushed nothing
    CUDA_TRY(cudaMemsetAsync(&ws->state->poison, 0, sizeof(int), ctx->stream));
    if ((int64_t)last != ws->n_global) {
      int ok_unused;
      B2A_TRY((reinitialize<HT>(ws, last, B2A_INIT_RAND, seed, &ok_unused, st)));  // expansion.jl:128
    }
    j = last + 1;
  }
  return B2A_OK;
}

}  // namespace eng

// ================================================================= restart driver (host)
namespace drv {

using namespace b2a::host;

template <class HT>
static int rotate_basis(b2a_ws *ws, int purge, int k, int maxdim, const HT *Q, int ldq, b2a_stats *st) {
  const int K = maxdim - purge + 1, N = k - purge + 1;
  if (N <= 0) return B2A_OK;
  std::vector<HT> Qp((size_t)K * N);
  for (int o = 0; o < N; ++o)
    for (int c = 0; c < K; ++c) Qp[(size_t)o * K + c] = Q[(size_t)(purge - 1 + o) * ldq + (purge - 1 + c)];
  const int64_t l0 = ws->ctx->launches;
  B2A_TRY((eng::rotate<HT>(ws, purge - 1, K, N, Qp.data(), maxdim, k)));
  if (st) {
    st->launches += ws->ctx->launches - l0;
    st->bytes += (double)ws->n_local * ws->esz * (K + N + 2.0);  // B_rot
  }
  return B2A_OK;
}

template <class HT> static int rotate_final(b2a_ws *ws, int nconv, const HT *Q, int ldq, b2a_stats *st) {
  if (nconv <= 0) return B2A_OK;
  std::vector<HT> Qp((size_t)nconv * nconv);
  for (int o = 0; o < nconv; ++o)
    for (int c = 0; c < nconv; ++c) Qp[(size_t)o * nconv + c] = Q[(size_t)o * ldq + c];
  const int64_t l0 = ws->ctx->launches;
  B2A_TRY((eng::rotate<HT>(ws, 0, nconv, nconv, Qp.data(), -1, -1)));
  if (st) {
    st->launches += ws->ctx->launches - l0;
    st->bytes += (double)ws->n_local * ws->esz * (2.0 * nconv);
  }
  return B2A_OK;
}

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// _partialschur (src/run.jl:224-392)
template <class HT>
static int partialschur(b2a_ws *ws, b2a_op *A, int mindim, int maxdim, int nev, double tol, int restarts,
                        int which, int active, uint64_t seed, b2a_history *hist, double *eig_out) {
  const int m1 = ws->maxdim + 1;  // leading dimension of the workspace H
  Mat<HT> H{reinterpret_cast<HT *>(ws->H.data()), maxdim + 1, maxdim, m1};
  Mat<HT> Q{reinterpret_cast<HT *>(ws->Q.data()), maxdim, maxdim, ws->maxdim};
  RestartScratch<HT> S(maxdim);
  Ordering ordering{which};
  b2a_stats *st = &hist->stats;

  int k = mindim;
  int64_t prods = std::max(0, mindim - active + 1);  // run.jl:264
  double t0 = now_ms();
  B2A_TRY((eng::iterate_arnoldi<HT>(ws, A, active, mindim, seed, st)));  // run.jl:267
  hist->ms_expand += now_ms() - t0;

  int iters = 0;
  for (int iter = 1; iter <= restarts; ++iter) {
    t0 = now_ms();
    B2A_TRY((eng::iterate_arnoldi<HT>(ws, A, k + 1, maxdim, seed, st)));  // run.jl:272
    prods += std::max(0, maxdim - k);                                        // run.jl:275
    double t1 = now_ms();
    hist->ms_expand += t1 - t0;

    RestartPlan plan;
    try {
      plan = restart_decision(H, Q, maxdim, mindim, nev, tol, ordering, active, S);  // run.jl:278-360
    } catch (const QRNoConvergence &e) {
      return fail(B2A_ERR_QR, e.what());
    }
    double t2 = now_ms();
    hist->ms_small += t2 - t1;

    k = plan.k;
    B2A_TRY((rotate_basis<HT>(ws, plan.purge, k, maxdim, Q.p, Q.ld, st)));  // run.jl:363-365
    hist->ms_rotate += now_ms() - t2;

    ++iters;
    active = plan.nlock + 1;  // run.jl:368
    if (active > nev) break;  // run.jl:370
  }

  const int nconverged = active - 1;
  t0 = now_ms();
  for (int j = 1; j <= maxdim; ++j)
    for (int i = 1; i <= maxdim; ++i) Q(i, j) = (i == j) ? HT(1) : HT(0);
  sortschur(H, Q, nconverged, ordering);  // run.jl:379
  double t1 = now_ms();
  hist->ms_small += t1 - t0;
  B2A_TRY((rotate_final<HT>(ws, nconverged, Q.p, Q.ld, st)));  // run.jl:382-383
  CUDA_TRY(cudaStreamSynchronize(ws->ctx->stream));  // rotations are asynchronous; the call is synchronous on return
  prof_collect(ws->ctx, nullptr, 0, 0);
  hist->ms_rotate += now_ms() - t1;

  if (eig_out) {
    std::vector<cplx> lams(maxdim);
    copy_eigenvalues(lams.data(), H, 1, nconverged);  // run.jl:386
    for (int i = 0; i < nconverged; ++i) {
      eig_out[2 * i] = lams[i].real();
      eig_out[2 * i + 1] = lams[i].imag();
    }
  }
  hist->mvproducts = prods;
  hist->nconverged = nconverged;
  hist->converged = nconverged >= nev;
  hist->nev = nev;
  hist->restarts = iters;
  return B2A_OK;
}

}  // namespace drv

// ================================================================================= ABI
extern "C" {

int b2a_version(void) { return B2A_VERSION; }
const char *b2a_last_error(void) { return g_err.c_str(); }

static int ctx_common(int device, b2a_ctx *c) {
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(B2A_ERR_ARGUMENT, "no such CUDA device");
  CUDA_TRY(cudaSetDevice(device));
  c->device = device;
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(B2A_ERR_CUDA, std::string("libb200arnoldi is built for sm_100a only; device is ") + prop.name);
  c->num_sms = prop.multiProcessorCount;
  CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  {
    cudaMemPool_t pool;
    CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
    unsigned long long keep = ~0ull;
    CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  CUDA_TRY(cudaMalloc(&c->d_zero, sizeof(int)));
  CUDA_TRY(cudaMemset(c->d_zero, 0, sizeof(int)));
  return B2A_OK;
}

int b2a_ctx_create(int device, b2a_ctx **out) {
  if (!out) return fail(B2A_ERR_ARGUMENT, "out is NULL");
  b2a_ctx *c = new b2a_ctx();
  int s = ctx_common(device, c);
  if (s != B2A_OK) {
    delete c;
    return s;
  }
  *out = c;
  return B2A_OK;
}

int b2a_nccl_unique_id(void *out128) {
  B2A_TRY(load_nccl());
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  NCCL_TRY(g_nccl.GetUniqueId(&id));
  std::memcpy(out128, &id, sizeof(id));
  return B2A_OK;
}

int b2a_ctx_create_dist(int device, int rank, int world, const void *nccl_unique_id, b2a_ctx **out) {
  if (!out) return fail(B2A_ERR_ARGUMENT, "out is NULL");
  if (world < 1 || rank < 0 || rank >= world) return fail(B2A_ERR_ARGUMENT, "bad rank / world");
  b2a_ctx *c = new b2a_ctx();
  int s = ctx_common(device, c);
  if (s != B2A_OK) {
    delete c;
    return s;
  }
  c->rank = rank;
  c->world = world;
  if (world > 1) {
    s = load_nccl();
    if (s != B2A_OK) {
      delete c;
      return s;
    }
    ncclUniqueId id;
    std::memcpy(&id, nccl_unique_id, sizeof(id));
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
      delete c;
      return fail(B2A_ERR_NCCL, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r));
    }
  }
  *out = c;
  return B2A_OK;
}

static void peer_release(b2a_ctx *ctx);

int b2a_ctx_destroy(b2a_ctx *ctx) {
  if (!ctx) return B2A_OK;
  cudaSetDevice(ctx->device);
  for (auto &r : ctx->prof_pending) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  for (auto e : ctx->prof_pool) cudaEventDestroy(e);
  peer_release(ctx);
  for (auto st : ctx->xchg_streams) cudaStreamDestroy(st);
  for (auto e : ctx->xchg_done) cudaEventDestroy(e);
  if (ctx->xchg_ready) cudaEventDestroy(ctx->xchg_ready);
  if (ctx->xchg_seq_table) cudaFree(ctx->xchg_seq_table);
  for (int i = 0; i < 2; ++i)
    if (ctx->xchg_data_ev[i]) cudaEventDestroy(ctx->xchg_data_ev[i]);
  if (ctx->comm) g_nccl.CommDestroy(ctx->comm);
  if (ctx->d_zero) cudaFree(ctx->d_zero);
  for (auto &pc : ctx->pinned_cache) cudaFreeHost(pc.second);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return B2A_OK;
}

int b2a_ctx_stream(b2a_ctx *ctx, void **stream) {
  if (!ctx || !stream) return fail(B2A_ERR_ARGUMENT, "NULL argument");
  *stream = ctx->stream;
  return B2A_OK;
}
int b2a_ctx_sync(b2a_ctx *ctx) {
  if (!ctx) return fail(B2A_ERR_ARGUMENT, "NULL ctx");
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return B2A_OK;
}
int b2a_ctx_rank(b2a_ctx *ctx, int *rank, int *world) {
  if (!ctx) return fail(B2A_ERR_ARGUMENT, "NULL ctx");
  if (rank) *rank = ctx->rank;
  if (world) *world = ctx->world;
  return B2A_OK;
}
int b2a_ctx_launch_count(b2a_ctx *ctx, int64_t *launches) {
  if (!ctx || !launches) return fail(B2A_ERR_ARGUMENT, "NULL argument");
  *launches = ctx->launches;
  return B2A_OK;
}

int b2a_ctx_profile_enable(b2a_ctx *ctx, int on) {
  if (!ctx) return fail(B2A_ERR_ARGUMENT, "NULL ctx");
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  prof_collect(ctx, nullptr, 0, 0);
  ctx->prof_on = on != 0;
  for (int k = 0; k < B2A_K_COUNT; ++k) {
    ctx->prof_n[k] = 0;
    ctx->prof_ms[k] = ctx->prof_bytes[k] = 0.0;
  }
  return B2A_OK;
}
int b2a_ctx_profile_get(b2a_ctx *ctx, int kind, int64_t *launches, double *ms, double *bytes) {
  if (!ctx || kind < 0 || kind >= B2A_K_COUNT) return fail(B2A_ERR_ARGUMENT, "bad kernel kind");
  if (launches) *launches = ctx->prof_n[kind];
  if (ms) *ms = ctx->prof_ms[kind];
  if (bytes) *bytes = ctx->prof_bytes[kind];
  return B2A_OK;
}

// ------------------------------------------------------------------------- operators
extern "C++" {
template <class Src>
__global__ void convert_index_kernel(const Src *__restrict__ src, int64_t count, int64_t base, int64_t *dst64,
                                     int32_t *dst32) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    const int64_t v = (int64_t)src[i] - base;
    if (dst64) dst64[i] = v;
    if (dst32) dst32[i] = (int32_t)v;
  }
}

}  // extern "C++"

// upload a host index array (32/64 bit, base 0/1) to a device int64 or int32 array
static int upload_index(b2a_ctx *ctx, const void *host, int64_t count, int idx_width, int idx_base, int64_t *d64,
                        int32_t *d32) {
  if (count == 0) return B2A_OK;
  if (idx_width == 32 && idx_base == 0 && d32) {
    CUDA_TRY(cudaMemcpyAsync(d32, host, count * 4, cudaMemcpyHostToDevice, ctx->stream));
    return B2A_OK;
  }
  if (idx_width == 64 && idx_base == 0 && d64) {
    CUDA_TRY(cudaMemcpyAsync(d64, host, count * 8, cudaMemcpyHostToDevice, ctx->stream));
    return B2A_OK;
  }
  void *tmp = nullptr;
  const size_t bytes = (size_t)count * (idx_width / 8);
  CUDA_TRY(dev_alloc(ctx, &tmp, bytes));
  cudaError_t e = cudaMemcpyAsync(tmp, host, bytes, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    const unsigned grid = (unsigned)std::min<int64_t>(ctx->num_sms * 8, std::max<int64_t>(1, cdiv(count, 256)));
    if (idx_width == 32)
      convert_index_kernel<int32_t><<<grid, 256, 0, ctx->stream>>>((const int32_t *)tmp, count, idx_base, d64, d32);
    else
      convert_index_kernel<int64_t><<<grid, 256, 0, ctx->stream>>>((const int64_t *)tmp, count, idx_base, d64, d32);
    ctx->launches++;
    e = cudaStreamSynchronize(ctx->stream);
  }
  dev_free(ctx, tmp);
  CUDA_TRY(e);
  return B2A_OK;
}

static int pick_entries_in_flight(int64_t nnz, int64_t nrows, int lpr);
static void op_tuning(b2a_op *op) {
  if (const char *e = getenv("B2A_SPMV_U")) op->rows_in_flight = atoi(e);
  if (const char *e = getenv("B2A_SPMV_GRID")) op->grid_mult = std::max(1, atoi(e));

  if (const char *e = getenv("B2A_SPMV_LPR")) op->lpr = atoi(e);
  if (const char *e = getenv("B2A_SPMV_STAGES")) op->tma_stages = atoi(e);
  op->entries_in_flight = pick_entries_in_flight(op->nnz, op->n_local, op->lpr);
  if (const char *e = getenv("B2A_SPMV_E")) op->entries_in_flight = atoi(e);
}

extern "C++" {
// Cut the rows into tiles for the TMA-stream kernel: even row boundaries, at most kSpmvMaxRows rows
// and kSpmvMaxNnz non-zeros (plus alignment slack) per tile.  Returns false if some row is too long.
template <class RP> static bool build_row_tiles(RP rp, int64_t n_rows, int max_nnz, std::vector<int32_t> &tiles) {
  tiles.clear();
  if (n_rows >= 2147483000LL) return false;
  int64_t r = 0;
  while (r < n_rows) {
    tiles.push_back((int32_t)r);
    const int64_t nz0 = rp(r);
    int64_t r1 = r;
    while (r1 < n_rows && r1 - r < b2a::kSpmvMaxRows) {
      const int64_t nxt = std::min<int64_t>(r1 + 2, n_rows);
      if (rp(nxt) - nz0 > max_nnz) break;
      r1 = nxt;
    }
    if (r1 == r) return false;  // a row pair with more than kSpmvMaxNnz entries
    r = r1;
  }
  tiles.push_back((int32_t)n_rows);
  return true;
}

}  // extern "C++"

static int upload_tiles(b2a_ctx *ctx, b2a_op *op, const std::vector<int32_t> &tiles) {
  op->ntiles = (int)tiles.size() - 1;
  CUDA_TRY(dev_alloc(ctx, reinterpret_cast<void **>(&op->d_tile_row), tiles.size() * sizeof(int32_t)));
  CUDA_TRY(cudaMemcpyAsync(op->d_tile_row, tiles.data(), tiles.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  // opt-in (B2A_SPMV_TMA=1): measured on B200 the staged stream is gather-latency bound and loses to
  // the occupancy-driven LDG kernel (191 vs 98 us at cfg 2, 220 vs 146 us on a 160^3 Laplacian)
  op->use_tma = op->ntiles > 0 && getenv("B2A_SPMV_TMA") && getenv("B2A_SPMV_TMA")[0] == '1';
  return B2A_OK;
}

// Lanes per row of the CSR vector kernel.  Measured on B200 (tools/spmvbench.py sweeps): about four
// entries per lane is best - 7-point Laplacian: LPR 2 = 94 us vs LPR 8 = 146 us (n = 4.1e6);
// 16 nnz/row random: LPR 4 = 91 us vs LPR 16 = 106 us; ComplexF64 20 nnz/row: LPR 4 = 120 us vs 193 us.
static int pick_lanes(int64_t nnz, int64_t nrows) {
  const double avg = nrows > 0 ? (double)nnz / (double)nrows : 0.0;
  if (avg <= 1.5) return 1;
  int lpr = 2;
  while (lpr < 32 && lpr * 2 <= avg / 4.0) lpr *= 2;
  return lpr;
}

// Entries per lane and row issued together (template parameter E of the vector kernel): the largest power of two
// <= (entries per row) / LPR, at most 4.  Measured on B200 (tools/spmvbench.py --esweep, profiles/r2_spmv_esweep_*):
// 16 random entries per row, LPR 4: E = 4 88.6 us vs E = 1 93.2; 7-point Laplacian 160^3, LPR 2: E = 2 85.5 us
// (5.15 TB/s) vs E = 1 91.4 and E = 4 95.1 (masked slots cost registers); ComplexF64 20 per row, LPR 4: E = 4
// 238 us vs E = 1 279.
static int pick_entries_in_flight(int64_t nnz, int64_t nrows, int lpr) {
  const double per_lane = nrows > 0 ? (double)nnz / (double)nrows / (double)std::max(lpr, 1) : 0.0;
  int e = 1;
  while (e < 4 && e * 2 <= per_lane) e *= 2;
  return e;
}

static int op_alloc(b2a_ctx *ctx, b2a_op *op, int64_t nptr) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  CUDA_TRY(dev_alloc(ctx, reinterpret_cast<void **>(&op->d_ptr), (size_t)(nptr + 1) * 8 + 64));
  CUDA_TRY(dev_alloc(ctx, reinterpret_cast<void **>(&op->d_idx), (size_t)std::max<int64_t>(op->nnz, 1) * 4 + 64));
  CUDA_TRY(dev_alloc(ctx, &op->d_vals, (size_t)std::max<int64_t>(op->nnz, 1) * dtype_size(op->dtype) + 64));
  return B2A_OK;
}

int b2a_op_destroy(b2a_op *op) {
  if (!op) return B2A_OK;
  if (op->owns) {
    cudaSetDevice(op->ctx->device);
    dev_free(op->ctx, op->d_ptr);
    dev_free(op->ctx, op->d_idx);
    dev_free(op->ctx, op->d_vals);
  }
  if (op->d_tile_row) {
    cudaSetDevice(op->ctx->device);
    dev_free(op->ctx, op->d_tile_row);
  }
  if (op->kind == OP_SHIFT_INVERT) {
    cudaSetDevice(op->ctx->device);
    dev_free(op->ctx, op->solve_work);
    dev_free(op->ctx, op->cg);
    dev_free(op->ctx, op->cg_partials);
    if (op->cg_host) cudaFreeHost(op->cg_host);
  }
  delete op;
  return B2A_OK;
}

static int check_op_args(b2a_ctx *ctx, int dtype, int64_t n_local, int64_t n_global, int64_t row_offset,
                         int64_t nnz, int idx_width, int idx_base) {
  if (!ctx) return fail(B2A_ERR_ARGUMENT, "NULL ctx");
  ARG_CHECK(dtype == B2A_F64 || dtype == B2A_C64, "dtype must be B2A_F64 or B2A_C64");
  ARG_CHECK(n_local >= 0 && n_global >= 1 && row_offset >= 0 && nnz >= 0, "negative size");
  if (row_offset + n_local > n_global)
    return fail(B2A_ERR_DIMENSION, "row block exceeds the matrix order (matrix must be square n_global x n_global)");
  ARG_CHECK(n_global < (int64_t)2147483647, "matrix order must fit 32-bit column indices");
  ARG_CHECK(idx_width == 32 || idx_width == 64, "idx_width must be 32 or 64");
  ARG_CHECK(idx_base == 0 || idx_base == 1, "idx_base must be 0 or 1");
  return B2A_OK;
}

// ---- column blocking -------------------------------------------------------------------------
// Decide the number of column blocks.  Blocking pays only when
//   (i)   x does not fit a comfortable share of L2 (> 48 MB),
//   (ii)  the columns are scattered (mean |col - row| far beyond an L2-sized window; banded / stencil operators
//         are left alone), and
//   (iii) the blocks are not too many for the row density: every block re-reads the row pointers and
//         read-modify-writes y, 8 + 2 s bytes per row and block, while the unblocked kernel over-fetches
//         32 - s bytes per nonzero (a 32-byte sector per s-byte gather that misses L2).  Measured on a cfg-5
//         shard (n = 1.25e7 square, 15 nnz/row, 3 blocks): 1.7x for blocking, in line with the model (72 vs 360
//         bytes per row); the TRUE cfg-5 shard (1.25e7 rows x 1e8 columns, x = 763 MiB, 24 blocks of 32 MiB)
//         would pay 576 bytes per row for blocking against 360 without, so it stays unblocked.
// B2A_SPMV_BLOCK_MB: unset/-1 = automatic, 0 = never, > 0 = force blocks of that many MB of x.
static int col_block_plan(double x_bytes, size_t es, double nnz_per_row, double mean_col_distance_bytes) {
  if (x_bytes <= 48.0 * 1048576.0) return 1;
  if (mean_col_distance_bytes < 8.0 * 1048576.0) return 1;
  const double per_row_block = 8.0 + 2.0 * (double)es;
  const double overfetch_per_row = nnz_per_row * (32.0 - (double)es);
  for (double mb : {32.0, 48.0}) {  // 32 MB measured best (16: 1734, 24: 1358, 32: 1227, 48: 1250 us)
    const double nb = std::ceil(x_bytes / (mb * 1048576.0));
    if (nb * per_row_block < 0.8 * overfetch_per_row) return (int)std::min(4096.0, nb);
  }
  return 1;
}

extern "C++" {
template <class RP, class CI>
static int decide_col_blocks(RP rp, CI ci, int64_t n_rows, int64_t n_global, int64_t row_offset, size_t es) {
  double block_mb = -1.0;
  if (const char *e = getenv("B2A_SPMV_BLOCK_MB")) block_mb = atof(e);
  if (block_mb == 0.0 || n_rows == 0) return 1;
  const double x_bytes = (double)n_global * (double)es;
  if (block_mb > 0.0) return (int)std::min<double>(4096.0, std::max(1.0, std::ceil(x_bytes / (block_mb * 1048576.0))));
  if (x_bytes <= 48.0 * 1048576.0) return 1;
  // scatter estimate on a sample of rows
  const int64_t step = std::max<int64_t>(1, n_rows / 4096);
  double sum = 0.0;
  int64_t cnt = 0;
  for (int64_t r = 0; r < n_rows; r += step)
    for (int64_t i = rp(r); i < rp(r + 1); ++i) {
      sum += std::fabs((double)(ci(i) - (row_offset + r)));
      ++cnt;
    }
  if (cnt == 0) return 1;
  return col_block_plan(x_bytes, es, (double)rp(n_rows) / (double)n_rows, sum / (double)cnt * (double)es);
}

// Reorder the CSR entries block-major (stable counting sort by (column block, row)): bptr gets nblocks row-pointer
// arrays, bcol / bval the permuted entries.
template <class RP, class CI>
static void build_col_blocks(RP rp, CI ci, const char *vals, size_t es, int64_t n_rows, int64_t n_global, int nblocks,
                             std::vector<int64_t> &bptr, std::vector<int32_t> &bcol, std::vector<char> &bval) {
  const int64_t W = cdiv(n_global, nblocks);
  const int64_t nnz = rp(n_rows);
  const size_t stride = (size_t)n_rows + 1;
  bptr.assign(stride * nblocks, 0);
  for (int64_t r = 0; r < n_rows; ++r)
    for (int64_t i = rp(r); i < rp(r + 1); ++i) bptr[(size_t)(ci(i) / W) * stride + r + 1]++;
  int64_t run = 0;
  for (int b = 0; b < nblocks; ++b) {
    int64_t *p = bptr.data() + (size_t)b * stride;
    p[0] = run;
    for (int64_t r = 0; r < n_rows; ++r) {
      run += p[r + 1];
      p[r + 1] = run;
    }
  }
  bcol.resize((size_t)std::max<int64_t>(nnz, 1));
  bval.resize((size_t)std::max<int64_t>(nnz, 1) * es);
  std::vector<int64_t> cur((size_t)nblocks);
  for (int64_t r = 0; r < n_rows; ++r) {
    for (int b = 0; b < nblocks; ++b) cur[b] = bptr[(size_t)b * stride + r];
    for (int64_t i = rp(r); i < rp(r + 1); ++i) {
      const int64_t c = ci(i);
      const int64_t dst = cur[(size_t)(c / W)]++;
      bcol[(size_t)dst] = (int32_t)c;
      std::memcpy(&bval[(size_t)dst * es], vals + (size_t)i * es, es);
    }
  }
}
}  // extern "C++"

// ---- structure validation (device pass over the uploaded arrays) -------------------------------------------
// SparseMatrixCSC's constructor rejects malformed structure; here a 1-based array passed with idx_base = 0, a
// decreasing pointer array or a column >= n would otherwise become out-of-bounds gathers (or red.add scatters).
static int validate_structure(b2a_ctx *ctx, const int64_t *d_ptr, int64_t n_ptr, int64_t nnz, const int32_t *d_idx,
                              int64_t idx_bound, const char *what) {
  int *d_err = nullptr;
  CUDA_TRY(dev_alloc(ctx, reinterpret_cast<void **>(&d_err), sizeof(int)));
  cudaError_t e = cudaMemsetAsync(d_err, 0, sizeof(int), ctx->stream);
  int err = 0;
  if (e == cudaSuccess) {
    const int64_t work = std::max<int64_t>(n_ptr, nnz);
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->num_sms * 8, cdiv(work, 256)));
    b2a::validate_csr_kernel<<<grid, 256, 0, ctx->stream>>>(n_ptr, d_ptr, nnz, d_idx, idx_bound, d_err);
    ctx->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  dev_free(ctx, d_err);
  CUDA_TRY(e);
  if (err & 1) return fail(B2A_ERR_ARGUMENT, std::string(what) + ": pointer array must start at 0, be non-decreasing and end at nnz");
  if (err & 2) return fail(B2A_ERR_ARGUMENT, std::string(what) + ": index out of range (check idx_base)");
  return B2A_OK;
}

// ---- owner groups (row-sharded operators) --------------------------------------------------------------------
// Reorder the uploaded shard block-major by the GROUP of ranks that owns the column, on the device
// (kernels_blocks.cuh).  W = rows per rank of the uniform partition, deduced from this rank's own block; if the block
// does not look like a uniform partition the operator stays as it is (the mat-vec then waits for the whole exchange
// before its single pass).
// ranks per owner group: as many slices of W rows as fit 32 MB of x, at least one, at most all
static int owner_group_size(int64_t W, size_t es, int P) {
  const int64_t slice = std::max<int64_t>(1, W * (int64_t)es);
  return (int)std::min<int64_t>(P, std::max<int64_t>(1, (int64_t)(32.0 * 1048576.0) / slice));
}

static int build_owner_blocks(b2a_ctx *ctx, b2a_op *op) {
  const int P = ctx->world;
  if (P < 2 || P > b2a::kMaxOwnerBlocks || op->n_local <= 0 || op->nnz <= 0) return B2A_OK;
  // Measured on 2 x B200 (bench.py, profiles/r2_bench_n2_variants.txt): with ONE remote slice the 8 MB transfer takes
  // 23 us and a single pass that waits for it (105 us) beats two passes (114 us: every pass re-reads the row
  // pointers, read-modify-writes y and pays its own ramp-up) and an all-blocks-in-one-launch kernel (147 us).
  // B2A_OWNER_BLOCKS=0 switches the reordering off, B2A_OWNER_GROUP=<ranks per block> overrides the size rule below.
  bool want = true;
  if (const char *e = getenv("B2A_OWNER_BLOCKS")) want = e[0] != '0';
  if (!want) return B2A_OK;
  const int64_t W = cdiv(op->n_global, P);
  // ranks per block: as many as fit ~32 MB of x (the block size that measured best for the L2-blocked mat-vec,
  // DESIGN 5); one block = no reordering (the mat-vec waits for the whole exchange, best up to 4 x 8 MB).
  // Measured at N = 8 (profiles/r2_variants_n8.txt): blocks of ONE owner (8 MB, 8 passes or one fused launch)
  // 316-319 us per mat-vec - every pass re-reads the row pointers and read-modify-writes y for 2 entries per row;
  // no blocks 273 us, of which ~80 us wait for the exchange and ~190 us are gathers from a 64 MB x that no longer
  // sits in L2 (89 us on one GPU with 8 MB).
  int G = owner_group_size(W, dtype_size(op->dtype), P);
  if (const char *e = getenv("B2A_OWNER_GROUP")) G = std::min(P, std::max(1, atoi(e)));
  const int nb = (int)cdiv(P, G);
  if (nb < 2) return B2A_OK;
  if (op->row_offset != (int64_t)ctx->rank * W) return B2A_OK;
  if (op->n_local != std::min<int64_t>(W, op->n_global - op->row_offset)) return B2A_OK;
  const int64_t n = op->n_local, ld = n + 1, L = (int64_t)nb * ld;
  const b2a::OwnerGroups og{W, ctx->rank, P, G};
  const size_t es = dtype_size(op->dtype);
  int64_t *bptr = nullptr, *sums = nullptr;
  int32_t *bcol = nullptr;
  void *bval = nullptr;
  const int64_t nchunks = cdiv(L, b2a::kScanChunk);
  cudaError_t e = dev_alloc(ctx, reinterpret_cast<void **>(&bptr), (size_t)L * 8 + 64);
  if (e == cudaSuccess) e = dev_alloc(ctx, reinterpret_cast<void **>(&sums), (size_t)nchunks * 8);
  if (e == cudaSuccess) e = dev_alloc(ctx, reinterpret_cast<void **>(&bcol), (size_t)op->nnz * 4 + 64);
  if (e == cudaSuccess) e = dev_alloc(ctx, &bval, (size_t)op->nnz * es + 64);
  if (e == cudaSuccess) e = cudaMemsetAsync(bptr, 0, (size_t)L * 8, ctx->stream);
  if (e == cudaSuccess) {
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->num_sms * 16, cdiv(n, 256)));
    b2a::blk_count_kernel<<<grid, 256, 0, ctx->stream>>>(n, op->d_ptr, op->d_idx, og, bptr);
    b2a::scan_partial_kernel<<<(unsigned)nchunks, b2a::kScanThreads, 0, ctx->stream>>>(bptr, L, sums);
    b2a::scan_sums_kernel<<<1, b2a::kScanThreads, 0, ctx->stream>>>(sums, nchunks);
    b2a::scan_apply_kernel<<<(unsigned)nchunks, b2a::kScanThreads, 0, ctx->stream>>>(bptr, L, sums);
    if (op->dtype == B2A_F64)
      b2a::blk_scatter_kernel<double><<<grid, 256, 0, ctx->stream>>>(n, op->d_ptr, op->d_idx,
                                                                     reinterpret_cast<const double *>(op->d_vals), og, nb,
                                                                     bptr, bcol, reinterpret_cast<double *>(bval));
    else
      b2a::blk_scatter_kernel<cdouble><<<grid, 256, 0, ctx->stream>>>(n, op->d_ptr, op->d_idx,
                                                                      reinterpret_cast<const cdouble *>(op->d_vals), og, nb,
                                                                      bptr, bcol, reinterpret_cast<cdouble *>(bval));
    ctx->launches += 5;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  dev_free(ctx, sums);
  if (e != cudaSuccess) {
    dev_free(ctx, bptr);
    dev_free(ctx, bcol);
    dev_free(ctx, bval);
    CUDA_TRY(e);
  }
  dev_free(ctx, op->d_ptr);
  dev_free(ctx, op->d_idx);
  dev_free(ctx, op->d_vals);
  op->d_ptr = bptr;
  op->d_idx = bcol;
  op->d_vals = bval;
  op->nblocks = nb;
  op->owner_blocks = true;
  op->owner_W = W;
  op->owner_G = G;
  op->lpr = pick_lanes(op->nnz, n * nb);
  if (const char *env = getenv("B2A_SPMV_LPR")) op->lpr = atoi(env);
  op->use_tma = false;
  return B2A_OK;
}

int b2a_csr_create(b2a_ctx *ctx, int dtype, int64_t n_rows_local, int64_t n_global, int64_t row_offset, int64_t nnz,
                   const void *rowptr, const void *colind, const void *vals, int idx_width, int idx_base,
                   b2a_op **out) {
  B2A_TRY(check_op_args(ctx, dtype, n_rows_local, n_global, row_offset, nnz, idx_width, idx_base));
  if (!rowptr || (nnz > 0 && (!colind || !vals)) || !out) return fail(B2A_ERR_ARGUMENT, "NULL array");
  b2a_op *op = new b2a_op();
  op->ctx = ctx;
  op->dtype = dtype;
  op->kind = OP_CSR;
  op->n_local = n_rows_local;
  op->n_global = n_global;
  op->row_offset = row_offset;
  op->nnz = nnz;
  op->lpr = pick_lanes(nnz, n_rows_local);
  op_tuning(op);
  // column blocking decision (host pass over a sample of rows) and, if taken, the block-major reordering
  {
    auto rp32 = [&](int64_t r) { return (int64_t) reinterpret_cast<const int32_t *>(rowptr)[r] - idx_base; };
    auto rp64 = [&](int64_t r) { return reinterpret_cast<const int64_t *>(rowptr)[r] - idx_base; };
    auto ci32 = [&](int64_t i) { return (int64_t) reinterpret_cast<const int32_t *>(colind)[i] - idx_base; };
    auto ci64 = [&](int64_t i) { return reinterpret_cast<const int64_t *>(colind)[i] - idx_base; };
    const size_t es = dtype_size(dtype);
    // row-sharded operators are cut by OWNER block on the device instead (build_owner_blocks)
    if (ctx->world == 1)
      op->nblocks = idx_width == 32 ? decide_col_blocks(rp32, ci32, n_rows_local, n_global, row_offset, es)
                                    : decide_col_blocks(rp64, ci64, n_rows_local, n_global, row_offset, es);
    if (op->nblocks > 1) {
      // the host reordering indexes by column: check the structure first (the device check comes too late here)
      bool sane = (idx_width == 32 ? rp32(0) : rp64(0)) == 0 && (idx_width == 32 ? rp32(n_rows_local) : rp64(n_rows_local)) == nnz;
      for (int64_t r = 0; sane && r < n_rows_local; ++r)
        sane = idx_width == 32 ? rp32(r) <= rp32(r + 1) : rp64(r) <= rp64(r + 1);
      for (int64_t i = 0; sane && i < nnz; ++i) {
        const int64_t c = idx_width == 32 ? ci32(i) : ci64(i);
        sane = c >= 0 && c < n_global;
      }
      if (!sane) {
        delete op;
        return fail(B2A_ERR_ARGUMENT, "CSR: malformed structure (pointer array / index range; check idx_base)");
      }
      std::vector<int64_t> bptr;
      std::vector<int32_t> bcol;
      std::vector<char> bval;
      if (idx_width == 32)
        build_col_blocks(rp32, ci32, reinterpret_cast<const char *>(vals), es, n_rows_local, n_global, op->nblocks, bptr, bcol, bval);
      else
        build_col_blocks(rp64, ci64, reinterpret_cast<const char *>(vals), es, n_rows_local, n_global, op->nblocks, bptr, bcol, bval);
      op->lpr = pick_lanes(nnz, n_rows_local * op->nblocks);
      if (const char *e = getenv("B2A_SPMV_LPR")) op->lpr = atoi(e);
      int sb = op_alloc(ctx, op, (int64_t)bptr.size() - 1);
      if (sb == B2A_OK) sb = upload_index(ctx, bptr.data(), (int64_t)bptr.size(), 64, 0, op->d_ptr, nullptr);
      if (sb == B2A_OK) sb = upload_index(ctx, bcol.data(), nnz, 32, 0, nullptr, op->d_idx);
      if (sb == B2A_OK && nnz > 0) {
        cudaError_t e = cudaMemcpyAsync(op->d_vals, bval.data(), (size_t)nnz * es, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) sb = fail(B2A_ERR_CUDA, cudaGetErrorString(e));
      }
      if (sb == B2A_OK) sb = cudaStreamSynchronize(ctx->stream) == cudaSuccess ? B2A_OK : fail(B2A_ERR_CUDA, "sync");
      if (sb == B2A_OK) sb = validate_structure(ctx, op->d_ptr, (int64_t)bptr.size() - 1, nnz, op->d_idx, n_global, "CSR");
      if (sb != B2A_OK) {
        b2a_op_destroy(op);
        return sb;
      }
      *out = op;
      return B2A_OK;
    }
  }
  int s = op_alloc(ctx, op, n_rows_local);
  if (s == B2A_OK) s = upload_index(ctx, rowptr, n_rows_local + 1, idx_width, idx_base, op->d_ptr, nullptr);
  if (s == B2A_OK) s = upload_index(ctx, colind, nnz, idx_width, idx_base, nullptr, op->d_idx);
  if (s == B2A_OK && nnz > 0) {
    cudaError_t e = cudaMemcpyAsync(op->d_vals, vals, (size_t)nnz * dtype_size(dtype), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) s = fail(B2A_ERR_CUDA, cudaGetErrorString(e));
  }
  if (s == B2A_OK) s = validate_structure(ctx, op->d_ptr, n_rows_local, nnz, op->d_idx, n_global, "CSR");
  if (s == B2A_OK) s = build_owner_blocks(ctx, op);
  if (s == B2A_OK && n_rows_local > 0 && !op->owner_blocks) {
    std::vector<int32_t> tiles;
    bool ok;
    const int max_nnz = dtype == B2A_C64 ? b2a::SpmvTile<cdouble>::max_nnz : b2a::SpmvTile<double>::max_nnz;
    if (idx_width == 32)
      ok = build_row_tiles([&](int64_t r) { return (int64_t) reinterpret_cast<const int32_t *>(rowptr)[r] - idx_base; }, n_rows_local, max_nnz, tiles);
    else
      ok = build_row_tiles([&](int64_t r) { return reinterpret_cast<const int64_t *>(rowptr)[r] - idx_base; }, n_rows_local, max_nnz, tiles);
    if (ok) s = upload_tiles(ctx, op, tiles);
  }
  if (s != B2A_OK) {
    b2a_op_destroy(op);
    return s;
  }
  *out = op;
  return B2A_OK;
}

int b2a_csr_create_device(b2a_ctx *ctx, int dtype, int64_t n_rows_local, int64_t n_global, int64_t row_offset,
                          int64_t nnz, const int64_t *d_rowptr, const int32_t *d_colind, const void *d_vals,
                          b2a_op **out) {
  B2A_TRY(check_op_args(ctx, dtype, n_rows_local, n_global, row_offset, nnz, 32, 0));
  if (!d_rowptr || !out) return fail(B2A_ERR_ARGUMENT, "NULL array");
  b2a_op *op = new b2a_op();
  op->ctx = ctx;
  op->dtype = dtype;
  op->kind = OP_CSR;
  op->n_local = n_rows_local;
  op->n_global = n_global;
  op->row_offset = row_offset;
  op->nnz = nnz;
  op->lpr = pick_lanes(nnz, n_rows_local);
  op->d_ptr = const_cast<int64_t *>(d_rowptr);
  op->d_idx = const_cast<int32_t *>(d_colind);
  op->d_vals = const_cast<void *>(d_vals);
  op->owns = false;
  op_tuning(op);
  {
    const int sv = validate_structure(ctx, d_rowptr, n_rows_local, nnz, d_colind, n_global, "CSR");
    if (sv != B2A_OK) {
      delete op;
      return sv;
    }
  }
  if (n_rows_local > 0) {
    // NOTE: the TMA-stream kernel reads 16-byte aligned supersets of the colind / vals slices, i.e. up
    // to 3 entries past nnz: borrowed arrays must be padded accordingly, else set B2A_SPMV_TMA=0.
    std::vector<int64_t> rp((size_t)n_rows_local + 1);
    cudaError_t e = cudaMemcpyAsync(rp.data(), d_rowptr, rp.size() * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      delete op;
      CUDA_TRY(e);
    }
    std::vector<int32_t> tiles;
    const int max_nnz = dtype == B2A_C64 ? b2a::SpmvTile<cdouble>::max_nnz : b2a::SpmvTile<double>::max_nnz;
    if (getenv("B2A_SPMV_TMA_BORROWED") && build_row_tiles([&](int64_t r) { return rp[(size_t)r]; }, n_rows_local, max_nnz, tiles)) {
      int s2 = upload_tiles(ctx, op, tiles);
      if (s2 != B2A_OK) {
        delete op;
        return s2;
      }
    }
  }
  *out = op;
  return B2A_OK;
}

extern "C++" {
template <class Idx> static inline int64_t host_idx(const void *p, int64_t i, int base) {
  return (int64_t) reinterpret_cast<const Idx *>(p)[i] - base;
}

}  // extern "C++"

int b2a_csc_create(b2a_ctx *ctx, int dtype, int64_t n_global, int64_t nnz, const void *colptr, const void *rowval,
                   const void *nzval, int idx_width, int idx_base, int mode, b2a_op **out) {
  B2A_TRY(check_op_args(ctx, dtype, n_global, n_global, 0, nnz, idx_width, idx_base));
  if (!colptr || (nnz > 0 && (!rowval || !nzval)) || !out) return fail(B2A_ERR_ARGUMENT, "NULL array");
  ARG_CHECK(mode == 0 || mode == 1, "mode must be 0 (transpose at upload) or 1 (scatter kernel)");
  // row-sharded job: every rank passes the WHOLE matrix in Julia's layout (as `A.colptr / A.rowval / A.nzval` would be
  // on every process) and keeps the rows of its block of the uniform partition, transposed to CSR at upload
  if (ctx->world > 1 && mode == 1)
    return fail(B2A_ERR_ARGUMENT, "the CSC scatter kernel (mode 1) is single-GPU; row-sharded CSC operators use mode 0");
  if (mode == 1) {
    b2a_op *op = new b2a_op();
    op->ctx = ctx;
    op->dtype = dtype;
    op->kind = OP_CSC_SCATTER;
    op->n_local = op->n_global = n_global;
    op->nnz = nnz;
    op->lpr = std::max(2, pick_lanes(nnz, n_global));
    int s = op_alloc(ctx, op, n_global);
    if (s == B2A_OK) s = upload_index(ctx, colptr, n_global + 1, idx_width, idx_base, op->d_ptr, nullptr);
    if (s == B2A_OK) s = upload_index(ctx, rowval, nnz, idx_width, idx_base, nullptr, op->d_idx);
    if (s == B2A_OK && nnz > 0) {
      cudaError_t e = cudaMemcpyAsync(op->d_vals, nzval, (size_t)nnz * dtype_size(dtype), cudaMemcpyHostToDevice, ctx->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
      if (e != cudaSuccess) s = fail(B2A_ERR_CUDA, cudaGetErrorString(e));
    }
    if (s == B2A_OK) s = validate_structure(ctx, op->d_ptr, n_global, nnz, op->d_idx, n_global, "CSC");
    if (s != B2A_OK) {
      b2a_op_destroy(op);
      return s;
    }
    *out = op;
    return B2A_OK;
  }
  // mode 0: one-time stable transpose to CSR on the host (setup, not the hot path), then upload.  Only the rows of
  // this rank's block [r0, r0 + nloc) of the uniform partition are kept (the whole matrix on a single GPU).
  const size_t es = dtype_size(dtype);
  const int64_t W = cdiv(n_global, ctx->world);
  const int64_t r0 = std::min<int64_t>((int64_t)ctx->rank * W, n_global);
  const int64_t nloc = std::min<int64_t>(W, n_global - r0);
  std::vector<int64_t> rowptr((size_t)nloc + 1, 0);
  auto ridx = [&](int64_t i) {
    return idx_width == 32 ? host_idx<int32_t>(rowval, i, idx_base) : host_idx<int64_t>(rowval, i, idx_base);
  };
  auto cptr = [&](int64_t c) {
    return idx_width == 32 ? host_idx<int32_t>(colptr, c, idx_base) : host_idx<int64_t>(colptr, c, idx_base);
  };
  if (cptr(0) != 0 || cptr(n_global) != nnz)
    return fail(B2A_ERR_ARGUMENT, "CSC: pointer array must start at 0 and end at nnz (check idx_base)");
  for (int64_t c = 0; c < n_global; ++c)
    if (cptr(c) > cptr(c + 1)) return fail(B2A_ERR_ARGUMENT, "CSC: pointer array must be non-decreasing");
  int64_t nnz_loc = 0;
  for (int64_t i = 0; i < nnz; ++i) {
    const int64_t r = ridx(i);
    if (r < 0 || r >= n_global) return fail(B2A_ERR_ARGUMENT, "row index out of range in CSC input");
    if (r >= r0 && r < r0 + nloc) {
      rowptr[(size_t)(r - r0) + 1]++;
      ++nnz_loc;
    }
  }
  for (int64_t r = 0; r < nloc; ++r) rowptr[(size_t)r + 1] += rowptr[(size_t)r];
  std::vector<int64_t> fill(rowptr.begin(), rowptr.end() - 1);
  std::vector<int64_t> colind((size_t)std::max<int64_t>(nnz_loc, 1));
  std::vector<char> vals((size_t)std::max<int64_t>(nnz_loc, 1) * es);
  for (int64_t c = 0; c < n_global; ++c) {
    const int64_t s = cptr(c), e = cptr(c + 1);
    for (int64_t i = s; i < e; ++i) {
      const int64_t r = ridx(i);
      if (r < r0 || r >= r0 + nloc) continue;
      const int64_t dst = fill[(size_t)(r - r0)]++;
      colind[(size_t)dst] = c;
      std::memcpy(&vals[(size_t)dst * es], reinterpret_cast<const char *>(nzval) + (size_t)i * es, es);
    }
  }
  return b2a_csr_create(ctx, dtype, nloc, n_global, r0, nnz_loc, rowptr.data(), colind.data(), vals.data(), 64, 0, out);
}

int b2a_op_from_callback(b2a_ctx *ctx, int dtype, int64_t n_rows_local, int64_t n_global, b2a_matvec_fn matvec,
                         void *user, b2a_op **out) {
  B2A_TRY(check_op_args(ctx, dtype, n_rows_local, n_global, 0, 0, 32, 0));
  if (!matvec || !out) return fail(B2A_ERR_ARGUMENT, "NULL callback");
  b2a_op *op = new b2a_op();
  op->ctx = ctx;
  op->dtype = dtype;
  op->kind = OP_CALLBACK;
  op->n_local = n_rows_local;
  op->n_global = n_global;
  op->fn = matvec;
  op->user = user;
  op->owns = false;
  *out = op;
  return B2A_OK;
}

int b2a_op_shift_invert(b2a_ctx *ctx, b2a_op *A, double sigma_re, double sigma_im, int method, double rtol, int maxit,
                        b2a_op **out) {
  if (!ctx || !A || !out) return fail(B2A_ERR_ARGUMENT, "NULL argument");
  ARG_CHECK(A->ctx == ctx, "operator belongs to a different context");
  ARG_CHECK(method == B2A_SOLVE_CG, "unknown inner solver");
  ARG_CHECK(ctx->world == 1 && A->kind == OP_CSR && !A->owner_blocks,
            "shift-and-invert needs a single-GPU CSR operator (CSC inputs: upload with mode 0)");
  ARG_CHECK(A->dtype == B2A_C64 || sigma_im == 0.0, "a complex shift needs a ComplexF64 operator");
  // the diagonal scan of the preconditioner walks ONE row-pointer array
  ARG_CHECK(A->nblocks == 1, "shift-and-invert: column-blocked operators are not supported (set B2A_SPMV_BLOCK_MB=0)");
  CUDA_TRY(cudaSetDevice(ctx->device));
  b2a_op *op = new b2a_op();
  op->ctx = ctx;
  op->dtype = A->dtype;
  op->kind = OP_SHIFT_INVERT;
  op->n_local = A->n_local;
  op->n_global = A->n_global;
  op->row_offset = A->row_offset;
  op->owns = false;
  op->inner = A;
  op->sigma_re = sigma_re;
  op->sigma_im = sigma_im;
  op->solve_rtol = rtol > 0.0 ? rtol : 1e-13;
  op->solve_maxit = maxit > 0 ? maxit : 10000;
  const size_t es = dtype_size(A->dtype);
  const int64_t n = std::max<int64_t>(A->n_local, 1);
  const unsigned pgrid = (unsigned)ctx->num_sms * 4;
  cudaError_t e = dev_alloc(ctx, &op->solve_work, (size_t)5 * n * es);
  if (e == cudaSuccess) e = dev_alloc(ctx, reinterpret_cast<void **>(&op->cg), sizeof(b2a::CgState));
  if (e == cudaSuccess) e = dev_alloc(ctx, reinterpret_cast<void **>(&op->cg_partials), sizeof(double2) * 2 * pgrid);
  if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void **>(&op->cg_host), sizeof(b2a::CgState));
  if (e == cudaSuccess) e = cudaMemsetAsync(op->cg, 0, sizeof(b2a::CgState), ctx->stream);
  if (e == cudaSuccess) {
    // Jacobi preconditioner of the shifted operator
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->num_sms * 8, cdiv(n, 256)));
    if (A->dtype == B2A_F64)
      b2a::shifted_diag_kernel<double><<<grid, 256, 0, ctx->stream>>>(A->n_local, A->d_ptr, A->d_idx,
                                                                      reinterpret_cast<const double *>(A->d_vals),
                                                                      A->row_offset, sigma_re,
                                                                      reinterpret_cast<double *>(op->solve_work));
    else
      b2a::shifted_diag_kernel<cdouble><<<grid, 256, 0, ctx->stream>>>(A->n_local, A->d_ptr, A->d_idx,
                                                                       reinterpret_cast<const cdouble *>(A->d_vals),
                                                                       A->row_offset, make_double2(sigma_re, sigma_im),
                                                                       reinterpret_cast<cdouble *>(op->solve_work));
    ctx->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    b2a_op_destroy(op);
    CUDA_TRY(e);
  }
  *out = op;
  return B2A_OK;
}

int b2a_op_solve_stats(b2a_op *op, int64_t *solves, int64_t *iterations, double *worst_relres) {
  if (!op || op->kind != OP_SHIFT_INVERT) return fail(B2A_ERR_ARGUMENT, "not a shift-and-invert operator");
  if (solves) *solves = op->solves;
  if (iterations) *iterations = op->solve_iters;
  if (worst_relres) *worst_relres = op->solve_worst;
  return B2A_OK;
}

int b2a_op_bytes(b2a_op *op, double *bytes) {
  if (!op || !bytes) return fail(B2A_ERR_ARGUMENT, "NULL argument");
  *bytes = eng::op_bytes(op);
  return B2A_OK;
}

// ------------------------------------------------------------------------- workspace
static void peer_teardown(b2a_ws *ws);

int b2a_ws_destroy(b2a_ws *ws) {
  if (!ws) return B2A_OK;
  cudaSetDevice(ws->ctx->device);
  peer_teardown(ws);
  for (int i = 0; i < 2; ++i)
    if (ws->qev[i]) cudaEventDestroy(ws->qev[i]);
  dev_free(ws->ctx, ws->dV);
  dev_free(ws->ctx, ws->arena);
  dev_free(ws->ctx, ws->xfull);
  pinned_put(ws->ctx, ws->pinned, ws->pinned_bytes);
  delete ws;
  return B2A_OK;
}

// Release the context's peer block (collective: all ranks call it at the same point).
static void peer_release(b2a_ctx *ctx) {
  if (!ctx->peer_local && ctx->peer_opened.empty()) return;
  cudaStreamSynchronize(ctx->stream);
  for (void *p : ctx->peer_opened) cudaIpcCloseMemHandle(p);
  ctx->peer_opened.clear();
  if (ctx->world > 1 && ctx->comm) {
    // exporters may only free after every importer has closed its mapping: a tiny all-reduce as barrier
    double *dd = nullptr;
    if (dev_alloc(ctx, reinterpret_cast<void **>(&dd), sizeof(double)) == cudaSuccess) {
      cudaMemsetAsync(dd, 0, sizeof(double), ctx->stream);
      g_nccl.AllReduce(dd, dd, 1, ncclFloat64, ncclSum, ctx->comm, ctx->stream);
      cudaStreamSynchronize(ctx->stream);
      dev_free(ctx, dd);
    }
  }
  if (ctx->peer_local) cudaFree(ctx->peer_local);
  ctx->peer_local = nullptr;
  ctx->peer_view = b2a::PeerView();
  ctx->peer_slot = 0;
  ctx->peer_x_bytes = 0;
}

// Give workspace `ws` the context's NVLink peer communication block, (re)creating it when it is too
// small: allocate this rank's block, exchange CUDA IPC handles over the NCCL communicator and map every
// peer's block.  Every decision is taken from values all ranks share, so the ranks stay in lockstep;
// any failure leaves ws->peer.P == 1 (host-launched NCCL collectives are used instead).
static int peer_setup(b2a_ctx *ctx, b2a_ws *ws) {
  ws->peer = b2a::PeerView();
  ws->peer_borrowed = false;
  if (ctx->world < 2 || ctx->world > b2a::kPeerMaxRanks) return B2A_OK;
  if (getenv("B2A_NO_PEER") && getenv("B2A_NO_PEER")[0] == '1') return B2A_OK;
  if (ctx->peer_busy) return B2A_OK;  // one workspace at a time owns the block; others use NCCL
  const int P = ctx->world;
  const int slot = (ws->maxdim + 2) * 2 + 2;  // doubles: [h (complex worst case) | norm]
  const size_t x_bytes = ((size_t)ws->n_global * ws->esz + 64 + 255) / 256 * 256;
  if (ctx->peer_view.P == P && ctx->peer_slot >= slot && ctx->peer_x_bytes >= x_bytes) {
    ws->peer = ctx->peer_view;
    ws->peer_borrowed = true;
    ctx->peer_busy = true;
    return B2A_OK;
  }
  peer_release(ctx);
  size_t off = 0;
  auto carve = [&](size_t bytes) {
    const size_t at = off;
    off += (bytes + 255) / 256 * 256;
    return at;
  };
  const size_t o_hdr = carve(256);
  const size_t o_flag_ar = carve(sizeof(unsigned long long) * b2a::kPeerBufs * P);
  const size_t o_flag_x = carve(sizeof(unsigned long long) * P);
  const size_t o_flag_xs = carve(sizeof(unsigned long long) * P);
  const size_t o_data = carve(sizeof(double) * b2a::kPeerBufs * P * slot);
  const size_t o_ll = carve((size_t)16 * b2a::kPeerBufs * P * slot);
  const size_t o_x = carve(2 * x_bytes);  // two buffers: exchange-number parity (staged exchange)
  const size_t total = off;
  int okflag = 1;
  cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&ctx->peer_local), total);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    ctx->peer_local = nullptr;
    okflag = 0;
  } else {
    cudaMemset(ctx->peer_local, 0, total);
  }
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof(mine));
  if (okflag && cudaIpcGetMemHandle(&mine, ctx->peer_local) != cudaSuccess) {
    (void)cudaGetLastError();
    okflag = 0;
  }
  // exchange [okflag | handle] (padded to 128 bytes) with an NCCL all-gather
  const size_t rec = 128;
  std::vector<char> send(rec, 0), recv(rec * P, 0);
  std::memcpy(send.data(), &okflag, sizeof(int));
  std::memcpy(send.data() + 8, &mine, sizeof(mine));
  char *d = nullptr;
  CUDA_TRY(dev_alloc(ctx, reinterpret_cast<void **>(&d), rec * (P + 1)));
  CUDA_TRY(cudaMemcpyAsync(d, send.data(), rec, cudaMemcpyHostToDevice, ctx->stream));
  NCCL_TRY(g_nccl.AllGather(d, d + rec, rec, ncclChar, ctx->comm, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(recv.data(), d + rec, rec * P, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  dev_free(ctx, d);
  bool all_ok = true;
  for (int r = 0; r < P; ++r) {
    int f;
    std::memcpy(&f, recv.data() + rec * r, sizeof(int));
    all_ok = all_ok && f == 1;
  }
  int opened_ok = all_ok ? 1 : 0;
  b2a::PeerView pv;
  if (all_ok) {
    for (int r = 0; r < P; ++r) {
      if (r == ctx->rank) {
        pv.peer[r] = ctx->peer_local;
        continue;
      }
      cudaIpcMemHandle_t h;
      std::memcpy(&h, recv.data() + rec * r + 8, sizeof(h));
      void *ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        (void)cudaGetLastError();
        opened_ok = 0;
        break;
      }
      ctx->peer_opened.push_back(ptr);
      pv.peer[r] = reinterpret_cast<char *>(ptr);
    }
  }
  {  // second agreement round: did everybody manage to map everybody?
    double *dd = nullptr;
    CUDA_TRY(dev_alloc(ctx, reinterpret_cast<void **>(&dd), sizeof(double)));
    const double v = opened_ok ? 0.0 : 1.0;
    CUDA_TRY(cudaMemcpyAsync(dd, &v, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    NCCL_TRY(g_nccl.AllReduce(dd, dd, 1, ncclFloat64, ncclSum, ctx->comm, ctx->stream));
    double tot = 1.0;
    CUDA_TRY(cudaMemcpyAsync(&tot, dd, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    dev_free(ctx, dd);
    if (tot != 0.0) {
      peer_release(ctx);
      return B2A_OK;  // somebody failed: everyone keeps NCCL
    }
  }
  pv.P = P;
  pv.rank = ctx->rank;
  pv.slot = slot;
  pv.off_flag_ar = o_flag_ar;
  pv.off_data_ar = o_data;
  pv.off_flag_x = o_flag_x;
  pv.off_flag_xs = o_flag_xs;
  pv.off_x = o_x;
  pv.x_stride = x_bytes;
  pv.off_ll = o_ll;
  pv.ll = !(getenv("B2A_AR_LL") && getenv("B2A_AR_LL")[0] == '0');
  pv.seq_ar = reinterpret_cast<unsigned long long *>(ctx->peer_local + o_hdr);
  pv.seq_x = pv.seq_ar + 1;
  pv.err = reinterpret_cast<int *>(pv.seq_ar + 2);
  ctx->peer_view = pv;
  ctx->peer_slot = slot;
  ctx->peer_x_bytes = x_bytes;
  ws->peer = pv;
  ws->peer_borrowed = true;
  ctx->peer_busy = true;
  return B2A_OK;
}

static void peer_teardown(b2a_ws *ws) {
  if (ws->peer_borrowed) {
    cudaStreamSynchronize(ws->ctx->stream);
    ws->ctx->peer_busy = false;  // the block stays mapped in the context for the next workspace
  }
  ws->peer_borrowed = false;
  ws->peer = b2a::PeerView();
}

static int ws_create_impl(b2a_ctx *ctx, int dtype, int64_t n_local, int64_t n_global, int64_t row_offset, int maxdim,
                          b2a_ws *ws) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  const size_t es = dtype_size(dtype);
  ws->ctx = ctx;
  ws->dtype = dtype;
  ws->esz = es;
  ws->n_local = n_local;
  ws->n_global = n_global;
  ws->row_offset = row_offset;
  ws->maxdim = maxdim;
  const int m1 = maxdim + 1;
  ws->all_offsets.assign(ctx->world, 0);
  ws->all_counts.assign(ctx->world, n_local);
  int64_t ld_rows = n_local;
  if (ctx->world > 1) {
    // share the row partition; uniform blocks (all but the last rank equal) use all-gather
    int64_t *d = nullptr;
    CUDA_TRY(dev_alloc(ctx, reinterpret_cast<void **>(&d), sizeof(int64_t) * 2 * (ctx->world + 1)));
    int64_t mine[2] = {row_offset, n_local};
    CUDA_TRY(cudaMemcpyAsync(d, mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream));
    NCCL_TRY(g_nccl.AllGather(d, d + 2, 2, ncclInt64, ctx->comm, ctx->stream));
    std::vector<int64_t> all(2 * ctx->world);
    CUDA_TRY(cudaMemcpyAsync(all.data(), d + 2, sizeof(int64_t) * 2 * ctx->world, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    dev_free(ctx, d);
    int64_t expect = 0;
    const int64_t blk = all[1];
    ws->uniform_partition = true;
    for (int r = 0; r < ctx->world; ++r) {
      ws->all_offsets[r] = all[2 * r];
      ws->all_counts[r] = all[2 * r + 1];
      if (all[2 * r] != expect) return fail(B2A_ERR_ARGUMENT, "row blocks of the ranks must be contiguous and ordered");
      expect += all[2 * r + 1];
      if (r + 1 < ctx->world && all[2 * r + 1] != blk) ws->uniform_partition = false;
      if (r + 1 == ctx->world && all[2 * r + 1] > blk) ws->uniform_partition = false;
    }
    if (expect != n_global) return fail(B2A_ERR_DIMENSION, "row blocks do not add up to n_global");
    if (ws->uniform_partition) ld_rows = std::max(ld_rows, blk);  // all-gather reads blk rows from every rank
  }
  // multiple of 1024 rows: every TMA tile (RT | 1024) can be loaded full-size; padding rows stay zero
  ws->ld = round_up(std::max<int64_t>(ld_rows, 1), 1024);
  ws->use_tma = !(getenv("B2A_NO_TMA") && getenv("B2A_NO_TMA")[0] == '1');
  if (const char *e = getenv("B2A_TMA_RT_DOTS")) ws->tune_rt_dots = atoi(e);
  if (const char *e = getenv("B2A_TMA_RT_UPD")) ws->tune_rt_upd = atoi(e);
  if (const char *e = getenv("B2A_TMA_STAGES")) ws->tune_stages = atoi(e);
  if (const char *e = getenv("B2A_FINISH_GRID")) ws->finish_grid_mult = std::max(1, atoi(e));
  if (const char *e = getenv("B2A_PEER_X")) ws->peer_x = e[0] != '0';
  if (const char *e = getenv("B2A_PUSH_SEPARATE")) ws->push_separate = e[0] == '1';
  if (const char *e = getenv("B2A_FUSED_SWEEP")) ws->fused_sweep = std::max(0, std::min(2, atoi(e)));
  if (const char *e = getenv("B2A_SWEEP_TRIGGER")) ws->sweep_early_trigger = e[0] == '1';
  if (const char *e = getenv("B2A_SWEEP_PDL")) ws->sweep_pdl = atoi(e) & 3;
  if (const char *e = getenv("B2A_TMA_L2PROMO")) ws->tune_l2promo = std::max(0, std::min(3, atoi(e)));
  CUDA_TRY(dev_alloc(ctx, &ws->dV, (size_t)ws->ld * m1 * es));
  CUDA_TRY(cudaMemsetAsync(ws->dV, 0, (size_t)ws->ld * m1 * es, ctx->stream));  // padding rows stay zero forever
  ws->H.assign((size_t)m1 * maxdim * es, 0);
  ws->Q.assign((size_t)maxdim * maxdim * es, 0);
  ws->dots_grid_max = 2 * ctx->num_sms;
  ws->upd_grid_max = 8 * ctx->num_sms;
  // all small scratch arrays live in one zero-initialised arena (256-byte aligned slices)
  size_t off = 0;
  auto carve = [&](size_t bytes) {
    const size_t at = off;
    off += (bytes + 255) / 256 * 256;
    return at;
  };
  const size_t o_dH = carve((size_t)m1 * maxdim * es);
  const size_t o_info = carve(sizeof(int) * (m1 + 1));
  const size_t o_hb1 = carve((size_t)(m1 + 1) * es + 16);
  const size_t o_hb2 = carve((size_t)(m1 + 1) * es + 16);
  const size_t o_w2 = carve(16);
  // per-CTA partial sums: four-kernel chain 66 x grid; fused sweep: one region per in-kernel grid barrier
  const size_t o_part = carve(std::max<size_t>((size_t)66 * ws->dots_grid_max,
                                               (size_t)b2a::kSweepBarriers * b2a::kSweepPartRegion) * es);
  const size_t o_part2 = carve(sizeof(double) * ws->upd_grid_max);
  const size_t o_state = carve(sizeof(b2a::SweepState));
  // Q of a rotation: packed K x N for the DFMA kernels, MMA B-fragment order (padded) for the DMMA kernel
  {
    const size_t inner = es / 8, parts = inner;
    // n tiles of the real view: 2 (maxdim + 1) outputs cover a complex product of a real basis too (partialeigen)
    const size_t frag = (size_t)((maxdim + 4) / 4) * ((2 * (maxdim + 1) + 7) / 8 + 3) * parts * 32;
    ws->q_bytes = (std::max<size_t>(frag, (size_t)maxdim * maxdim * inner) * 8 + 255) / 256 * 256;
  }
  const size_t o_dQ = carve(ws->q_bytes);
  const size_t o_flag = carve(sizeof(unsigned long long));
  const bool want_trace = getenv("B2A_SWEEP_TRACE") && getenv("B2A_SWEEP_TRACE")[0] == '1';
  const size_t o_trace = carve(want_trace ? sizeof(unsigned long long) * b2a::kSweepTraceSlots * ctx->num_sms : 8);
  CUDA_TRY(dev_alloc(ctx, &ws->arena, off));
  CUDA_TRY(cudaMemsetAsync(ws->arena, 0, off, ctx->stream));
  char *base = reinterpret_cast<char *>(ws->arena);
  ws->dH = base + o_dH;
  ws->dinfo = reinterpret_cast<int *>(base + o_info);
  ws->hb1 = base + o_hb1;
  ws->hb2 = base + o_hb2;
  ws->w2sq = reinterpret_cast<double *>(base + o_w2);
  ws->partials = base + o_part;
  ws->partials2 = reinterpret_cast<double *>(base + o_part2);
  ws->state = reinterpret_cast<b2a::SweepState *>(base + o_state);
  ws->dQ = base + o_dQ;
  ws->sweep_flag = reinterpret_cast<unsigned long long *>(base + o_flag);
  ws->sweep_trace = want_trace ? reinterpret_cast<unsigned long long *>(base + o_trace) : nullptr;
  ws->qpin_off = ((size_t)m1 * maxdim * es + sizeof(int) * (m1 + 1) + sizeof(b2a::SweepState) + 64 + 255) / 256 * 256;
  const size_t want = ws->qpin_off + 2 * ws->q_bytes;
  CUDA_TRY(pinned_get(ctx, want, &ws->pinned, &ws->pinned_bytes));
  for (int i = 0; i < 2; ++i) CUDA_TRY(cudaEventCreateWithFlags(&ws->qev[i], cudaEventDisableTiming));
  if (const char *e = getenv("B2A_ROTATE")) ws->rotate_mode = e[0] != '0';
  if (const char *e = getenv("B2A_ROT_CTAS")) ws->rot_ctas = std::max(0, std::min(2, atoi(e)));
  B2A_TRY(peer_setup(ctx, ws));
  // staged exchange: default whenever the peer block is available (B2A_XCHG=0: push from the normalising kernel)
  ws->xchg_staged = ws->peer.P > 1 && ws->peer_x && !(getenv("B2A_XCHG") && getenv("B2A_XCHG")[0] == '0');
  if (ctx->world > 1 && (ws->peer.P == 1 || !ws->peer_x)) {  // NCCL all-gather needs its own gather buffer
    const int64_t nx = ws->uniform_partition ? ws->all_counts[0] * ctx->world : n_global;
    CUDA_TRY(dev_alloc(ctx, &ws->xfull, (size_t)nx * es));
  }
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return B2A_OK;
}

int b2a_ws_create(b2a_ctx *ctx, int dtype, int64_t n_rows_local, int64_t n_global, int64_t row_offset, int maxdim,
                  b2a_ws **out) {
  if (!ctx || !out) return fail(B2A_ERR_ARGUMENT, "NULL argument");
  ARG_CHECK(dtype == B2A_F64 || dtype == B2A_C64, "dtype must be B2A_F64 or B2A_C64");
  ARG_CHECK(n_rows_local >= 0 && n_global >= 1 && row_offset >= 0 && row_offset + n_rows_local <= n_global,
            "bad row block");
  ARG_CHECK(maxdim >= 1, "Krylov dimension must be positive");
  // ArnoldiMethod.jl:62-63
  ARG_CHECK((int64_t)maxdim <= n_global, "Krylov dimension should be less than matrix order.");
  // the in-place basis rotation stages 16 rows x maxdim columns per pipeline stage, two stages at least
  // (kernels_rotate_mma.cuh): maxdim <= 879 (Float64) / 439 (ComplexF64)
  ARG_CHECK((size_t)round_up(maxdim, 4) * 16 * dtype_size(dtype) * 2 + 256 <= eng::kTmaSmemBudget,
            "Krylov dimension too large for the device basis rotation (limit 879 for Float64, 439 for ComplexF64)");
  b2a_ws *ws = new b2a_ws();
  int s = ws_create_impl(ctx, dtype, n_rows_local, n_global, row_offset, maxdim, ws);
  if (s != B2A_OK) {
    b2a_ws_destroy(ws);
    return s;
  }
  *out = ws;
  return B2A_OK;
}

#define WS_COL_CHECK(ws, j) \
  ARG_CHECK((ws) && (j) >= 1 && (j) <= (ws)->maxdim + 1, "column index out of range")

int b2a_ws_set_col(b2a_ws *ws, int j, const void *host) {
  WS_COL_CHECK(ws, j);
  if (!host) return fail(B2A_ERR_ARGUMENT, "NULL host pointer");
  char *dst = reinterpret_cast<char *>(ws->dV) + (size_t)(j - 1) * ws->ld * ws->esz;
  ws->x_pushed_col = -1;
  CUDA_TRY(cudaMemcpyAsync(dst, host, (size_t)ws->n_local * ws->esz, cudaMemcpyHostToDevice, ws->ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ws->ctx->stream));
  return B2A_OK;
}
int b2a_ws_set_col_device(b2a_ws *ws, int j, const void *dev) {
  WS_COL_CHECK(ws, j);
  if (!dev) return fail(B2A_ERR_ARGUMENT, "NULL device pointer");
  char *dst = reinterpret_cast<char *>(ws->dV) + (size_t)(j - 1) * ws->ld * ws->esz;
  ws->x_pushed_col = -1;
  CUDA_TRY(cudaMemcpyAsync(dst, dev, (size_t)ws->n_local * ws->esz, cudaMemcpyDeviceToDevice, ws->ctx->stream));
  return B2A_OK;
}
int b2a_ws_comm_mode(b2a_ws *ws, int *mode) {
  if (!ws || !mode) return fail(B2A_ERR_ARGUMENT, "NULL argument");
  *mode = ws->ctx->world == 1 ? 0 : (ws->peer.P == ws->ctx->world ? 2 : 1);
  return B2A_OK;
}

int b2a_ws_debug_sweep_trace(b2a_ws *ws, unsigned long long *out, int max_ctas, int *slots) {
  if (!ws || !out || !slots) return fail(B2A_ERR_ARGUMENT, "NULL argument");
  *slots = b2a::kSweepTraceSlots;
  if (!ws->sweep_trace) return fail(B2A_ERR_ARGUMENT, "workspace was created without B2A_SWEEP_TRACE=1");
  const int nc = std::min(max_ctas, ws->ctx->num_sms);
  CUDA_TRY(cudaStreamSynchronize(ws->ctx->stream));
  CUDA_TRY(cudaMemcpy(out, ws->sweep_trace, sizeof(unsigned long long) * b2a::kSweepTraceSlots * nc,
                      cudaMemcpyDeviceToHost));
  return B2A_OK;
}

int b2a_ws_get_cols(b2a_ws *ws, int j0, int ncols, void *host, int64_t ld) {
  if (ncols == 0) return B2A_OK;
  WS_COL_CHECK(ws, j0);
  ARG_CHECK(ncols > 0 && j0 + ncols - 1 <= ws->maxdim + 1, "column range out of bounds");
  ARG_CHECK(host && ld >= ws->n_local, "bad host matrix");
  const char *src = reinterpret_cast<const char *>(ws->dV) + (size_t)(j0 - 1) * ws->ld * ws->esz;
  CUDA_TRY(cudaMemcpy2DAsync(host, (size_t)ld * ws->esz, src, (size_t)ws->ld * ws->esz, (size_t)ws->n_local * ws->esz,
                             ncols, cudaMemcpyDeviceToHost, ws->ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ws->ctx->stream));
  return B2A_OK;
}
int b2a_ws_col_ptr(b2a_ws *ws, int j, void **dev, int64_t *ld) {
  WS_COL_CHECK(ws, j);
  if (dev) *dev = reinterpret_cast<char *>(ws->dV) + (size_t)(j - 1) * ws->ld * ws->esz;
  if (ld) *ld = ws->ld;
  return B2A_OK;
}
int b2a_ws_host_arrays(b2a_ws *ws, void **H, int *ldh, void **Q, int *ldq) {
  if (!ws) return fail(B2A_ERR_ARGUMENT, "NULL ws");
  if (H) *H = ws->H.data();
  if (ldh) *ldh = ws->maxdim + 1;
  if (Q) *Q = ws->Q.data();
  if (ldq) *ldq = ws->maxdim;
  return B2A_OK;
}

// -------------------------------------------------------------------------- hot path
int b2a_reinitialize(b2a_ws *ws, int j, int mode, uint64_t seed, int *ok) {
  ARG_CHECK(ws && j >= 0 && j <= ws->maxdim, "column index out of range");
  ARG_CHECK(mode == B2A_INIT_RAND || mode == B2A_INIT_KEEP, "mode must be B2A_INIT_RAND or B2A_INIT_KEEP");
  CUDA_TRY(cudaSetDevice(ws->ctx->device));
  return ws->dtype == B2A_F64 ? eng::reinitialize<double>(ws, j, mode, seed, ok, nullptr)
                              : eng::reinitialize<cplx>(ws, j, mode, seed, ok, nullptr);
}

extern "C++" {
template <class HT> static int orthogonalize_impl(b2a_ws *ws, int j, void *h_host, int *ok) {
  using DT = typename Dev<HT>::type;
  const int m1 = ws->maxdim + 1;
  B2A_TRY((eng::enqueue_cgs<DT>(ws, j, 0, j)));
  CUDA_TRY(cudaMemcpyAsync(ws->pinned, reinterpret_cast<char *>(ws->dH) + (size_t)(j - 1) * m1 * sizeof(HT),
                           (size_t)(j + 1) * sizeof(HT), cudaMemcpyDeviceToHost, ws->ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(ws->pinned + (size_t)m1 * sizeof(HT), ws->dinfo + j, sizeof(int), cudaMemcpyDeviceToHost,
                           ws->ctx->stream));
  CUDA_TRY(cudaMemsetAsync(&ws->state->poison, 0, sizeof(int), ws->ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ws->ctx->stream));
  HT *H = reinterpret_cast<HT *>(ws->H.data());
  std::memcpy(H + (size_t)(j - 1) * m1, ws->pinned, (size_t)(j + 1) * sizeof(HT));
  if (h_host) std::memcpy(h_host, ws->pinned, (size_t)(j + 1) * sizeof(HT));
  int info;
  std::memcpy(&info, ws->pinned + (size_t)m1 * sizeof(HT), sizeof(int));
  prof_collect(ws->ctx, &info, j, 1);
  if (ok) *ok = (info & 2) ? 0 : 1;
  int err = 0;
  CUDA_TRY(cudaMemcpy(&err, &ws->state->error, sizeof(int), cudaMemcpyDeviceToHost));
  if (err) return fail(B2A_ERR_CUDA, "fused orthogonalisation: grid barrier timed out");
  return B2A_OK;
}

}  // extern "C++"

int b2a_orthogonalize(b2a_ws *ws, int j, void *h_host, int *ok) {
  ARG_CHECK(ws && j >= 1 && j <= ws->maxdim, "column index out of range");
  CUDA_TRY(cudaSetDevice(ws->ctx->device));
  return ws->dtype == B2A_F64 ? orthogonalize_impl<double>(ws, j, h_host, ok)
                              : orthogonalize_impl<cplx>(ws, j, h_host, ok);
}

static int check_ws_op(b2a_ws *ws, b2a_op *A) {
  if (!ws || !A) return fail(B2A_ERR_ARGUMENT, "NULL handle");
  ARG_CHECK(ws->ctx == A->ctx, "workspace and operator belong to different contexts");
  ARG_CHECK(ws->dtype == A->dtype, "workspace and operator have different element types");
  if (ws->n_global != A->n_global || ws->n_local != A->n_local)
    return fail(B2A_ERR_DIMENSION, "workspace and operator dimensions differ");
  return B2A_OK;
}

int b2a_ws_matvec(b2a_ws *ws, b2a_op *A, int jsrc, int jdst) {
  B2A_TRY(check_ws_op(ws, A));
  WS_COL_CHECK(ws, jsrc);
  WS_COL_CHECK(ws, jdst);
  ARG_CHECK(jsrc != jdst, "source and destination columns must differ");
  CUDA_TRY(cudaSetDevice(ws->ctx->device));
  if (ws->dtype == B2A_F64) return eng::enqueue_matvec<double>(ws, A, jsrc - 1, jdst - 1);
  return eng::enqueue_matvec<cdouble>(ws, A, jsrc - 1, jdst - 1);
}

int b2a_iterate_arnoldi(b2a_ws *ws, b2a_op *A, int from, int to, uint64_t seed, void *H_host, int ldh,
                        b2a_stats *stats) {
  B2A_TRY(check_ws_op(ws, A));
  ARG_CHECK(from >= 1 && to <= ws->maxdim, "step range out of bounds");
  CUDA_TRY(cudaSetDevice(ws->ctx->device));
  b2a_stats local{};
  b2a_stats *st = stats ? stats : &local;
  if (from <= to) {
    B2A_TRY(ws->dtype == B2A_F64 ? eng::iterate_arnoldi<double>(ws, A, from, to, seed, st)
                                 : eng::iterate_arnoldi<cplx>(ws, A, from, to, seed, st));
  }
  if (H_host && from <= to) {
    ARG_CHECK(ldh >= ws->maxdim + 1, "ldh too small");
    const int m1 = ws->maxdim + 1;
    for (int s = from; s <= to; ++s)
      std::memcpy(reinterpret_cast<char *>(H_host) + (size_t)(s - 1) * ldh * ws->esz,
                  ws->H.data() + (size_t)(s - 1) * m1 * ws->esz, (size_t)(s + 1) * ws->esz);
  }
  return B2A_OK;
}

int b2a_rotate_basis(b2a_ws *ws, int purge, int k, int maxdim, const void *Q_host, int ldq, b2a_stats *stats) {
  if (!ws) return fail(B2A_ERR_ARGUMENT, "NULL ws");
  ARG_CHECK(maxdim >= 1 && maxdim <= ws->maxdim && purge >= 1 && purge <= k && k <= maxdim, "bad (purge, k, maxdim)");
  if (!Q_host) {
    Q_host = ws->Q.data();
    ldq = ws->maxdim;
  }
  ARG_CHECK(ldq >= maxdim, "ldq too small");
  CUDA_TRY(cudaSetDevice(ws->ctx->device));
  B2A_TRY(ws->dtype == B2A_F64
              ? drv::rotate_basis<double>(ws, purge, k, maxdim, reinterpret_cast<const double *>(Q_host), ldq, stats)
              : drv::rotate_basis<cplx>(ws, purge, k, maxdim, reinterpret_cast<const cplx *>(Q_host), ldq, stats));
  CUDA_TRY(cudaStreamSynchronize(ws->ctx->stream));  // synchronous on return (inside partialschur it is not)
  prof_collect(ws->ctx, nullptr, 0, 0);
  return B2A_OK;
}

int b2a_rotate_final(b2a_ws *ws, int nconv, const void *Q_host, int ldq, b2a_stats *stats) {
  if (!ws) return fail(B2A_ERR_ARGUMENT, "NULL ws");
  ARG_CHECK(nconv >= 0 && nconv <= ws->maxdim, "bad nconv");
  if (!Q_host) {
    Q_host = ws->Q.data();
    ldq = ws->maxdim;
  }
  ARG_CHECK(ldq >= nconv, "ldq too small");
  CUDA_TRY(cudaSetDevice(ws->ctx->device));
  B2A_TRY(ws->dtype == B2A_F64
              ? drv::rotate_final<double>(ws, nconv, reinterpret_cast<const double *>(Q_host), ldq, stats)
              : drv::rotate_final<cplx>(ws, nconv, reinterpret_cast<const cplx *>(Q_host), ldq, stats));
  CUDA_TRY(cudaStreamSynchronize(ws->ctx->stream));
  prof_collect(ws->ctx, nullptr, 0, 0);
  return B2A_OK;
}

int b2a_basis_times(b2a_ws *ws, int nconv, const double *Y, int ldy, double *X, int64_t ldx) {
  if (!ws) return fail(B2A_ERR_ARGUMENT, "NULL ws");
  if (nconv == 0) return B2A_OK;
  ARG_CHECK(nconv >= 1 && nconv <= ws->maxdim + 1 && Y && X && ldy >= nconv && ldx >= ws->n_local, "bad arguments");
  b2a_ctx *ctx = ws->ctx;
  CUDA_TRY(cudaSetDevice(ctx->device));
  std::vector<cplx> Yp((size_t)nconv * nconv);
  for (int o = 0; o < nconv; ++o)
    for (int c = 0; c < nconv; ++c) Yp[(size_t)o * nconv + c] = cplx(Y[2 * ((size_t)o * ldy + c)], Y[2 * ((size_t)o * ldy + c) + 1]);
  cdouble *dY = nullptr, *dX = nullptr;
  const int64_t n = ws->n_local;
  // TMA + DMMA kernel (whole tiles are written: the device buffer has the workspace's padded leading dimension)
  {
    cudaError_t e = dev_alloc(ctx, reinterpret_cast<void **>(&dX), sizeof(cdouble) * (size_t)ws->ld * nconv);
    CUDA_TRY(e);
    const int s = ws->dtype == B2A_F64 ? eng::basis_times_mma<double>(ws, nconv, nconv, Yp.data(), dX, ws->ld)
                                       : eng::basis_times_mma<cplx>(ws, nconv, nconv, Yp.data(), dX, ws->ld);
    if (s == B2A_OK) {
      e = cudaMemcpy2DAsync(X, (size_t)ldx * 16, dX, (size_t)ws->ld * 16, (size_t)n * 16, nconv, cudaMemcpyDeviceToHost,
                            ctx->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
      dev_free(ctx, dX);
      CUDA_TRY(e);
      return B2A_OK;
    }
    dev_free(ctx, dX);
    dX = nullptr;
    if (s != 1) return s;
  }
  CUDA_TRY(dev_alloc(ctx, reinterpret_cast<void **>(&dY), sizeof(cdouble) * nconv * nconv));
  cudaError_t e = dev_alloc(ctx, reinterpret_cast<void **>(&dX), sizeof(cdouble) * (size_t)std::max<int64_t>(n, 1) * nconv);
  if (e != cudaSuccess) {
    dev_free(ctx, dY);
    CUDA_TRY(e);
  }
  e = cudaMemcpyAsync(dY, Yp.data(), sizeof(cdouble) * nconv * nconv, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ctx->num_sms * 8, cdiv(n, 256)));
    if (ws->dtype == B2A_F64)
      b2a::basis_times_kernel<double><<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<const double *>(ws->dV), ws->ld, n, nconv, nconv, dY, dX, n);
    else
      b2a::basis_times_kernel<cdouble><<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<const cdouble *>(ws->dV), ws->ld, n, nconv, nconv, dY, dX, n);
    ctx->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess)
    e = cudaMemcpy2DAsync(X, (size_t)ldx * 16, dX, (size_t)n * 16, (size_t)n * 16, nconv, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  dev_free(ctx, dY);
  dev_free(ctx, dX);
  CUDA_TRY(e);
  return B2A_OK;
}

// --------------------------------------------------------- fine-grained vector operations
extern "C++" {
// dots of column j0 (0-based) against columns [0, ncols): result in ws->hb1 = [h | ||v||^2], all-reduced
template <class DT> static int fine_dots(b2a_ws *ws, int ncols, int j0) {
  b2a_ctx *ctx = ws->ctx;
  DT *v = eng::col<DT>(ws, j0);
  DT *h1 = reinterpret_cast<DT *>(ws->hb1);
  double *rsq = reinterpret_cast<double *>(h1 + ncols);
  // the TMA kernels expect v to be column `ncols` of the panel; use them only in that layout
  const bool tma = (j0 == ncols) && eng::tma_path_ok(ws, ncols, sizeof(DT));
  if (tma) {
    B2A_TRY(eng::launch_dots_tma<DT>(ws, ncols, v, h1, rsq, nullptr, nullptr, 0, false));
    if (ws->peer.P == 1) B2A_TRY(eng::allreduce_f64(ctx, h1, (size_t)ncols * sizeof(DT) / 8 + 1));
  } else {
    B2A_TRY(eng::launch_dots<DT>(ws, ncols, v, h1, rsq, nullptr, nullptr));
    B2A_TRY(eng::allreduce_f64(ctx, h1, (size_t)ncols * sizeof(DT) / 8 + 1));
  }
  return B2A_OK;
}
template <class DT> static int fine_nrm2(b2a_ws *ws, int j0, double *result) {
  B2A_TRY(fine_dots<DT>(ws, 0, j0));
  CUDA_TRY(cudaMemcpyAsync(ws->pinned, ws->hb1, sizeof(double), cudaMemcpyDeviceToHost, ws->ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ws->ctx->stream));
  prof_collect(ws->ctx, nullptr, 0, 0);
  double sq;
  std::memcpy(&sq, ws->pinned, sizeof(double));
  *result = std::sqrt(sq);
  return B2A_OK;
}
template <class DT> static int fine_gemv_c(b2a_ws *ws, int ncols, int j0, void *h_host) {
  B2A_TRY(fine_dots<DT>(ws, ncols, j0));
  CUDA_TRY(cudaMemcpyAsync(ws->pinned, ws->hb1, (size_t)ncols * sizeof(DT), cudaMemcpyDeviceToHost, ws->ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ws->ctx->stream));
  prof_collect(ws->ctx, nullptr, 0, 0);
  std::memcpy(h_host, ws->pinned, (size_t)ncols * sizeof(DT));
  return B2A_OK;
}
template <class DT> static int fine_gemv_n_sub(b2a_ws *ws, int ncols, int j0, const void *h_host) {
  b2a_ctx *ctx = ws->ctx;
  DT *h2 = reinterpret_cast<DT *>(ws->hb2);
  std::memcpy(ws->pinned, h_host, (size_t)ncols * sizeof(DT));
  CUDA_TRY(cudaMemcpyAsync(h2, ws->pinned, (size_t)ncols * sizeof(DT), cudaMemcpyHostToDevice, ctx->stream));
  // LDG kernel: its norm output is not needed here (no collective)
  B2A_TRY(eng::launch_update<DT>(ws, ncols, eng::col<DT>(ws, j0), h2, ws->w2sq, nullptr, nullptr));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  prof_collect(ctx, nullptr, 0, 0);
  ws->x_pushed_col = -1;
  return B2A_OK;
}
}  // extern "C++"

#define FINE_CHECK(ws, j) \
  ARG_CHECK((ws) && (j) >= 1 && (j) <= (ws)->maxdim + 1, "column index out of range")

int b2a_ws_nrm2(b2a_ws *ws, int j, double *result) {
  FINE_CHECK(ws, j);
  if (!result) return fail(B2A_ERR_ARGUMENT, "NULL result");
  CUDA_TRY(cudaSetDevice(ws->ctx->device));
  return ws->dtype == B2A_F64 ? fine_nrm2<double>(ws, j - 1, result) : fine_nrm2<cdouble>(ws, j - 1, result);
}
int b2a_ws_gemv_c(b2a_ws *ws, int ncols, int j, void *h_host) {
  FINE_CHECK(ws, j);
  ARG_CHECK(ncols >= 0 && ncols <= ws->maxdim + 1 && h_host, "bad panel width");
  if (ncols == 0) return B2A_OK;
  CUDA_TRY(cudaSetDevice(ws->ctx->device));
  return ws->dtype == B2A_F64 ? fine_gemv_c<double>(ws, ncols, j - 1, h_host) : fine_gemv_c<cdouble>(ws, ncols, j - 1, h_host);
}
int b2a_ws_gemv_n_sub(b2a_ws *ws, int ncols, int j, const void *h_host) {
  FINE_CHECK(ws, j);
  ARG_CHECK(ncols >= 0 && ncols <= ws->maxdim + 1 && h_host, "bad panel width");
  ARG_CHECK(j > ncols, "the updated column must lie outside the panel");
  if (ncols == 0) return B2A_OK;
  CUDA_TRY(cudaSetDevice(ws->ctx->device));
  return ws->dtype == B2A_F64 ? fine_gemv_n_sub<double>(ws, ncols, j - 1, h_host) : fine_gemv_n_sub<cdouble>(ws, ncols, j - 1, h_host);
}
int b2a_ws_scal_div(b2a_ws *ws, int j, double alpha) {
  FINE_CHECK(ws, j);
  b2a_ctx *ctx = ws->ctx;
  CUDA_TRY(cudaSetDevice(ctx->device));
  const int64_t grid = std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->num_sms * 8, cdiv(ws->n_local, 256)));
  if (ws->dtype == B2A_F64)
    b2a::scal_div_kernel<double><<<(unsigned)grid, 256, 0, ctx->stream>>>(eng::col<double>(ws, j - 1), ws->n_local, alpha);
  else
    b2a::scal_div_kernel<cdouble><<<(unsigned)grid, 256, 0, ctx->stream>>>(eng::col<cdouble>(ws, j - 1), ws->n_local, alpha);
  ctx->launches++;
  CUDA_TRY(cudaGetLastError());
  ws->x_pushed_col = -1;
  return B2A_OK;
}
int b2a_ws_copy_col(b2a_ws *ws, int jsrc, int jdst) {
  FINE_CHECK(ws, jsrc);
  FINE_CHECK(ws, jdst);
  if (jsrc == jdst) return B2A_OK;
  CUDA_TRY(cudaSetDevice(ws->ctx->device));
  const char *src = reinterpret_cast<const char *>(ws->dV) + (size_t)(jsrc - 1) * ws->ld * ws->esz;
  char *dst = reinterpret_cast<char *>(ws->dV) + (size_t)(jdst - 1) * ws->ld * ws->esz;
  CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)ws->n_local * ws->esz, cudaMemcpyDeviceToDevice, ws->ctx->stream));
  ws->x_pushed_col = -1;
  return B2A_OK;
}

// ------------------------------------------------------------------- whole restart loop
int b2a_partialschur(b2a_ws *ws, b2a_op *A, const b2a_params *p, b2a_history *history, double *eigenvalues_c64) {
  B2A_TRY(check_ws_op(ws, A));
  if (!p || !history) return fail(B2A_ERR_ARGUMENT, "NULL params / history");
  const int64_t n = ws->n_global;
  const int vcols = ws->maxdim + 1;  // size(arnoldi.V, 2)
  // defaults of src/run.jl:152-163
  const int nev = p->nev != 0 ? p->nev : (int)std::min<int64_t>(6, n);
  const int mindim = p->mindim != 0 ? p->mindim : (int)std::min<int64_t>(std::min<int64_t>(std::max(10, nev), n), vcols - 1);
  const int maxdim = p->maxdim != 0 ? p->maxdim : (int)std::min<int64_t>(std::min<int64_t>(std::max(20, 2 * nev), n), vcols - 1);
  const double tol = p->tol >= 0 ? p->tol : std::sqrt(b2a::host::kEps);
  const int restarts = p->restarts >= 0 ? p->restarts : 200;
  const int start_from = p->start_from > 0 ? p->start_from : 1;
  const int initialize = p->initialize >= 0 ? p->initialize : (start_from == 1 ? B2A_INIT_RAND : B2A_INIT_NONE);
  // src/run.jl:165-174
  ARG_CHECK(nev >= 1, "nev cannot be less than 1");
  if (!(nev <= mindim && mindim <= maxdim && (int64_t)maxdim <= n))
    return fail(B2A_ERR_ARGUMENT, "nev <= mindim <= maxdim <= size(A, 1) does not hold, got " + std::to_string(nev) +
                                      " <= " + std::to_string(mindim) + " <= " + std::to_string(maxdim) + " <= " +
                                      std::to_string(n));
  ARG_CHECK(maxdim < vcols, "maxdim should be strictly less than size(arnoldi.V, 2)");
  ARG_CHECK(1 <= start_from && start_from <= maxdim, "start_from should be between 1 and maxdim");
  ARG_CHECK(p->which >= B2A_LM && p->which <= B2A_SI, "Unknown target");
  ARG_CHECK(initialize >= B2A_INIT_NONE && initialize <= B2A_INIT_KEEP, "bad initialize mode");
  CUDA_TRY(cudaSetDevice(ws->ctx->device));

  std::memset(history, 0, sizeof(*history));
  const int m1 = ws->maxdim + 1;
  // fill!(view(H, :, start_from:end), 0)   (run.jl:176)
  std::memset(ws->H.data() + (size_t)(start_from - 1) * m1 * ws->esz, 0, (size_t)(ws->maxdim - start_from + 1) * m1 * ws->esz);
  if (initialize != B2A_INIT_NONE) {
    int ok;
    B2A_TRY(ws->dtype == B2A_F64 ? eng::reinitialize<double>(ws, start_from - 1, initialize, p->seed, &ok, &history->stats)
                                 : eng::reinitialize<cplx>(ws, start_from - 1, initialize, p->seed, &ok, &history->stats));
  }
  try {
    if (ws->dtype == B2A_F64)
      return drv::partialschur<double>(ws, A, mindim, maxdim, nev, tol, restarts, p->which, start_from, p->seed, history, eigenvalues_c64);
    return drv::partialschur<cplx>(ws, A, mindim, maxdim, nev, tol, restarts, p->which, start_from, p->seed, history, eigenvalues_c64);
  } catch (const std::exception &e) {
    return fail(B2A_ERR_INTERNAL, e.what());
  }
}

// ------------------------------------------------------------------ host-only algebra
int b2a_host_local_schurfact(int dtype, void *H, int ldh, int rows, int cols, int from, int to, void *Q, int ldq, int qrows) {
  using namespace b2a::host;
  ARG_CHECK(H && rows >= 1 && cols >= 1 && ldh >= rows && from >= 1 && to <= cols && to <= rows, "bad arguments");
  try {
    if (dtype == B2A_F64) {
      Mat<double> Hm{(double *)H, rows, cols, ldh}, Qm{(double *)Q, qrows, cols, ldq};
      return local_schurfact(Hm, from, to, Qm) ? B2A_OK : B2A_ERR_QR;
    }
    Mat<cplx> Hm{(cplx *)H, rows, cols, ldh}, Qm{(cplx *)Q, qrows, cols, ldq};
    return local_schurfact(Hm, from, to, Qm) ? B2A_OK : fail(B2A_ERR_QR, "QR algorithm did not converge");
  } catch (const QRNoConvergence &e) {
    return fail(B2A_ERR_QR, e.what());
  }
}

extern "C++" {
template <class HT>
static int host_restart_impl(void *H, int ldh, void *Q, int ldq, int maxdim, int mindim, int nev, double tol, int which,
                             int active, int *k, int *purge, int *nlock, double *eig, double *res) {
  using namespace b2a::host;
  Mat<HT> Hm{(HT *)H, maxdim + 1, maxdim, ldh}, Qm{(HT *)Q, maxdim, maxdim, ldq};
  RestartScratch<HT> S(maxdim);
  RestartPlan plan;
  try {
    plan = restart_decision(Hm, Qm, maxdim, mindim, nev, tol, Ordering{which}, active, S);
  } catch (const QRNoConvergence &e) {
    return fail(B2A_ERR_QR, e.what());
  }
  if (k) *k = plan.k;
  if (purge) *purge = plan.purge;
  if (nlock) *nlock = plan.nlock;
  for (int i = 0; i < maxdim; ++i) {
    if (eig) {
      eig[2 * i] = S.lams[i].real();
      eig[2 * i + 1] = S.lams[i].imag();
    }
    if (res) res[i] = S.rs[i];
  }
  return B2A_OK;
}

}  // extern "C++"

int b2a_host_restart(int dtype, void *H, int ldh, void *Q, int ldq, int maxdim, int mindim, int nev, double tol,
                     int which, int active, int *k, int *purge, int *nlock, double *eigenvalues_c64, double *residuals) {
  ARG_CHECK(H && Q && maxdim >= 1 && ldh >= maxdim + 1 && ldq >= maxdim, "bad arguments");
  ARG_CHECK(nev >= 1 && nev <= mindim && mindim <= maxdim && active >= 1 && active <= maxdim, "bad dimensions");
  ARG_CHECK(which >= B2A_LM && which <= B2A_SI, "Unknown target");
  return dtype == B2A_F64 ? host_restart_impl<double>(H, ldh, Q, ldq, maxdim, mindim, nev, tol, which, active, k, purge, nlock, eigenvalues_c64, residuals)
                          : host_restart_impl<cplx>(H, ldh, Q, ldq, maxdim, mindim, nev, tol, which, active, k, purge, nlock, eigenvalues_c64, residuals);
}

int b2a_host_sortschur(int dtype, void *H, int ldh, void *Q, int ldq, int maxdim, int nconv, int which) {
  using namespace b2a::host;
  ARG_CHECK(H && Q && maxdim >= 1 && ldh >= maxdim + 1 && ldq >= maxdim && nconv >= 0 && nconv <= maxdim, "bad arguments");
  ARG_CHECK(which >= B2A_LM && which <= B2A_SI, "Unknown target");
  if (dtype == B2A_F64) {
    Mat<double> Hm{(double *)H, maxdim + 1, maxdim, ldh}, Qm{(double *)Q, maxdim, maxdim, ldq};
    for (int j = 1; j <= maxdim; ++j)
      for (int i = 1; i <= maxdim; ++i) Qm(i, j) = (i == j) ? 1.0 : 0.0;
    sortschur(Hm, Qm, nconv, Ordering{which});
  } else {
    Mat<cplx> Hm{(cplx *)H, maxdim + 1, maxdim, ldh}, Qm{(cplx *)Q, maxdim, maxdim, ldq};
    for (int j = 1; j <= maxdim; ++j)
      for (int i = 1; i <= maxdim; ++i) Qm(i, j) = (i == j) ? cplx(1.0) : cplx(0.0);
    sortschur(Hm, Qm, nconv, Ordering{which});
  }
  return B2A_OK;
}

int b2a_host_col_block_plan(int dtype, int64_t n_global, double nnz_per_row, double mean_col_distance, int *nblocks) {
  if (!nblocks || n_global < 1) return fail(B2A_ERR_ARGUMENT, "bad argument");
  const size_t es = dtype_size(dtype);
  *nblocks = col_block_plan((double)n_global * (double)es, es, nnz_per_row, mean_col_distance * (double)es);
  return B2A_OK;
}

int b2a_host_owner_group_plan(int dtype, int64_t n_global, int world, int rank, int *ranks_per_block, int *nblocks,
                              int *block_of_owner) {
  if (!ranks_per_block || !nblocks || n_global < 1 || world < 1 || rank < 0 || rank >= world)
    return fail(B2A_ERR_ARGUMENT, "bad argument");
  const int64_t W = cdiv(n_global, world);
  const int G = owner_group_size(W, dtype_size(dtype), world);
  *ranks_per_block = G;
  *nblocks = (int)cdiv(world, G);
  if (block_of_owner) {
    const b2a::OwnerGroups og{W, rank, world, G};
    for (int o = 0; o < world; ++o) block_of_owner[o] = ((o - og.me + og.P) % og.P) / og.G;
  }
  return B2A_OK;
}

int b2a_host_givens(int dtype, const double *f, const double *g, double *c, double *s, double *r) {
  using namespace b2a::host;
  if (!f || !g || !c || !s || !r) return fail(B2A_ERR_ARGUMENT, "NULL argument");
  if (dtype == B2A_F64) {
    auto G = givens(f[0], g[0]);
    *c = G.c;
    s[0] = G.s;
    r[0] = G.r;
  } else {
    auto G = givens(cplx(f[0], f[1]), cplx(g[0], g[1]));
    *c = G.c;
    s[0] = G.s.real();
    s[1] = G.s.imag();
    r[0] = G.r.real();
    r[1] = G.r.imag();
  }
  return B2A_OK;
}

}  // extern "C"
