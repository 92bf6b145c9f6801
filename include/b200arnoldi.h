/* b200arnoldi.h - C ABI of libb200arnoldi.so
 *
 * B200-native (sm_100a) implementation of the ONE hot path of ArnoldiMethod.jl:
 * the Arnoldi expansion (operator mat-vec + iterated classical Gram-Schmidt) and
 * the Krylov-Schur restart's basis rotation V <- V*Q.  Every entry point below
 * names the reference interface (path:line under the reference tree) it replaces.
 *
 * Conventions
 *  - every function returns an int status: 0 = B2A_OK, < 0 = error (see enum);
 *    b2a_last_error() returns a thread-local message for the last failure.
 *    No C++ exception crosses this boundary.
 *  - argument errors mirror the reference: B2A_ERR_ARGUMENT  <-> ArgumentError
 *    (src/run.jl:111-116,123-124,165-174,185; src/ArnoldiMethod.jl:62-63,87-90),
 *    B2A_ERR_DIMENSION <-> DimensionMismatch (src/run.jl:110 checksquare).
 *    Non-convergence is NOT an error (src/run.jl:388): it is reported in b2a_history.
 *  - matrices are column-major; all indices in this API are 1-based where the
 *    reference's are (Krylov column numbers j, start_from, purge, k), so a Julia
 *    caller passes its own integers unchanged.
 *  - host pointers are borrowed for the duration of the call only; device memory
 *    is owned by the opaque handles.
 *  - dtype: B2A_F64 = Float64, B2A_C64 = ComplexF64 (interleaved re,im).
 *  - thread-compatible, not thread-safe: one host thread per context.
 *  - one context per process and GPU.  Multi-GPU = one process per GPU, rows of A,
 *    V and v sharded in contiguous blocks (b2a_ctx_create_dist).
 */
#ifndef B200ARNOLDI_H
#define B200ARNOLDI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2A_VERSION 100 /* 0.1.0 */

typedef struct b2a_ctx b2a_ctx; /* device context: stream, scratch, communicator          */
typedef struct b2a_op b2a_op;   /* linear operator A (CSR / CSC / user callback)          */
typedef struct b2a_ws b2a_ws;   /* ArnoldiWorkspace: V on device, H and Q on the host     */

enum b2a_status {
  B2A_OK = 0,
  B2A_ERR_ARGUMENT = -1,  /* Julia ArgumentError                                       */
  B2A_ERR_DIMENSION = -2, /* Julia DimensionMismatch                                   */
  B2A_ERR_CUDA = -3,
  B2A_ERR_NCCL = -4,
  B2A_ERR_OOM = -5,
  B2A_ERR_QR = -6, /* "QR algorithm did not converge" (src/schurfact.jl:406)           */
  B2A_ERR_INTERNAL = -7,
  B2A_ERR_CALLBACK = -8, /* user mat-vec callback returned non-zero                     */
  B2A_ERR_SOLVE = -9     /* inner solve of a shift-and-invert operator did not converge     */
};

enum b2a_dtype { B2A_F64 = 0, B2A_C64 = 1 };

/* src/targets.jl:12-32 and _symbol_to_target, src/run.jl:181-185 */
enum b2a_which { B2A_LM = 0, B2A_LR = 1, B2A_SR = 2, B2A_LI = 3, B2A_SI = 4 };

/* how column start_from is populated before the run (src/run.jl:119-127,177) */
enum b2a_init {
  B2A_INIT_NONE = 0, /* partialschur!(...; initialize=false): column used as is          */
  B2A_INIT_RAND = 1, /* reinitialize!(arnoldi, start_from-1, rand!)                      */
  B2A_INIT_KEEP = 2  /* reinitialize!(arnoldi, 0, v -> copyto!(v, v1)): normalise what
                        b2a_ws_set_col() put there (the v1 keyword of partialschur)      */
};

/* ------------------------------------------------------------------ lifecycle */

int b2a_version(void);
const char *b2a_last_error(void);

/* Single-GPU context on CUDA device `device`. */
int b2a_ctx_create(int device, b2a_ctx **out);

/* Row-sharded multi-GPU context: this process is `rank` of `world`, one GPU each.
 * `nccl_unique_id` is the 128-byte ncclUniqueId produced by b2a_nccl_unique_id() on
 * rank 0 and distributed by the host program (torch.distributed / MPI / sockets). */
int b2a_ctx_create_dist(int device, int rank, int world, const void *nccl_unique_id, b2a_ctx **out);
int b2a_nccl_unique_id(void *out128);
int b2a_ctx_destroy(b2a_ctx *ctx);

/* The CUDA stream (cudaStream_t) all kernels of this context are launched on, so a
 * host can bracket calls with its own CUDA events. */
int b2a_ctx_stream(b2a_ctx *ctx, void **stream);
int b2a_ctx_sync(b2a_ctx *ctx);
int b2a_ctx_rank(b2a_ctx *ctx, int *rank, int *world);
/* number of this library's kernels launched on the context so far */
int b2a_ctx_launch_count(b2a_ctx *ctx, int64_t *launches);

/* Per-kernel-kind device timing (CUDA events on the context's stream around every launch of
 * this library's kernels).  Off by default; meant for benchmarks: kernel durations are
 * accumulated at the library's own synchronisation points (end of a sweep / rotation). */
enum b2a_kernel_kind {
  B2A_K_SPMV = 0,   /* operator mat-vec                                                  */
  B2A_K_DOTS = 1,   /* cgs_dots   (h = V' v, ||v||^2)        bytes/launch = (j+1) n s    */
  B2A_K_UPDATE = 2, /* cgs_update (v -= V h, ||v||^2)        bytes/launch = (j+2) n s    */
  B2A_K_FINISH = 3, /* cgs_finish (H column, v ./= wnorm)    bytes/launch = 2 n s        */
  B2A_K_ROTATE = 4, /* in-place V <- V Q                                                 */
  B2A_K_FILL = 5,   /* counter-based rand!                                               */
  B2A_K_SWEEP = 6,  /* fused orthogonalisation (dots + update [+ update] + finish in one
                       persistent kernel)     bytes/launch = (2j+5) n s [+ (j+2) n s]   */
  B2A_K_XCHG = 7,   /* staged x exchange of a row-sharded mat-vec, from "column final" to "all outgoing
                       stages done": bytes/launch = NVLink bytes this rank SENDS, (P-1) n_loc s          */
  B2A_K_COUNT = 8
};
int b2a_ctx_profile_enable(b2a_ctx *ctx, int on); /* also resets the accumulators */
int b2a_ctx_profile_get(b2a_ctx *ctx, int kind, int64_t *launches, double *ms, double *bytes);

/* --------------------------------------------------------------------- operator
 * Replaces the user operator of `mul!(y, A, x)`, `eltype(A)`, `size(A)`
 * (contract: src/run.jl:21-25; call site: src/expansion.jl:121).               */

/* CSR rows [row_offset, row_offset + n_rows_local) of a square n_global x n_global matrix.
 * rowptr has n_rows_local+1 entries, relative to this block (rowptr[0] == idx_base);
 * colind holds GLOBAL column numbers.  idx_width = 32 or 64 (bits of rowptr/colind
 * entries as passed); idx_base = 0 or 1 (Julia arrays are 1-based Int64).  All arrays are
 * host pointers and are copied to the device (colind narrowed to 32 bit). */
int b2a_csr_create(b2a_ctx *ctx, int dtype, int64_t n_rows_local, int64_t n_global, int64_t row_offset,
                   int64_t nnz, const void *rowptr, const void *colind, const void *vals, int idx_width,
                   int idx_base, b2a_op **out);

/* Same, from arrays already resident on this context's device (0-based, int64 rowptr,
 * int32 colind); the operator BORROWS them (no copy) - they must outlive the operator. */
int b2a_csr_create_device(b2a_ctx *ctx, int dtype, int64_t n_rows_local, int64_t n_global,
                          int64_t row_offset, int64_t nnz, const int64_t *d_rowptr,
                          const int32_t *d_colind, const void *d_vals, b2a_op **out);

/* Julia's native SparseMatrixCSC: colptr (n_cols+1), rowval, nzval of the WHOLE
 * n_global x n_global matrix - `mode` selects the kernel:
 *   0 = transpose once at upload, then the CSR kernel (deterministic).  In a row-sharded job every rank passes
 *       the whole matrix (as every Julia process would hold `A.colptr / A.rowval / A.nzval`) and keeps the rows
 *       of its block of the uniform partition (ceil(n / world) rows per rank);
 *   1 = native column-scatter kernel (red.global.add.f64; sums in arrival order; single GPU only). */
int b2a_csc_create(b2a_ctx *ctx, int dtype, int64_t n_global, int64_t nnz, const void *colptr,
                   const void *rowval, const void *nzval, int idx_width, int idx_base, int mode,
                   b2a_op **out);

/* Matrix-free operator: `matvec(user, x_dev, y_dev, n_local, stream)` must enqueue
 * y <- A*x on `stream` for the local rows (device pointers, dtype as given) and return 0.
 * This is the `mul!(y, A, x)` contract for LinearMaps-style operators (@cfunction). */
typedef int (*b2a_matvec_fn)(void *user, const void *x_dev, void *y_dev, int64_t n_local, void *stream);
int b2a_op_from_callback(b2a_ctx *ctx, int dtype, int64_t n_rows_local, int64_t n_global,
                         b2a_matvec_fn matvec, void *user, b2a_op **out);

/* Shift-and-invert operator  y = (A - sigma I)^{-1} x  on the device.
 * The reference leaves spectral transformations to the user: docs/src/index.md:234-262 ("Shift-and-invert with
 * LinearMaps.jl") and bench/partial_schur.jl:11-35 wrap a factorisation in a linear map and hand that to
 * partialschur with which = :LM; eigenvalues of A nearest sigma are then sigma + 1 / theta.  Here the map is an
 * iterative solve built on the CSR mat-vec of `A` (single GPU, CSR operator; `A` must outlive the result):
 *   method B2A_SOLVE_CG = Jacobi-preconditioned conjugate gradients, for A - sigma I Hermitian positive definite
 *   (e.g. the smallest modes of a Laplacian-type operator).  Every mat-vec with the result runs the inner solve to
 *   ||r|| <= rtol ||x|| (rtol <= 0: 1e-13) in at most maxit iterations (maxit <= 0: 10000); a solve that does not
 *   get there makes the mat-vec - and the partialschur call around it - fail with B2A_ERR_SOLVE. */
enum { B2A_SOLVE_CG = 0 };
int b2a_op_shift_invert(b2a_ctx *ctx, b2a_op *A, double sigma_re, double sigma_im, int method, double rtol,
                        int maxit, b2a_op **out);
/* inner-solver statistics of a shift-and-invert operator since its creation */
int b2a_op_solve_stats(b2a_op *op, int64_t *solves, int64_t *iterations, double *worst_relres);

int b2a_op_destroy(b2a_op *op);
/* algorithmic bytes one mat-vec moves (nnz*(s+4) + 8*(n+1) + 2*n*s) */
int b2a_op_bytes(b2a_op *op, double *bytes);

/* -------------------------------------------------------------------- workspace
 * Replaces ArnoldiWorkspace (src/ArnoldiMethod.jl:41-93).  V is n_local x (maxdim+1)
 * on the device; H ((maxdim+1) x maxdim) and Q (maxdim x maxdim) live on the host
 * inside the handle (they are indexed scalar-by-scalar by the m x m algebra).  V_tmp
 * does not exist: the basis rotation is done in place by row tiles.
 * Errors: maxdim > n_global -> B2A_ERR_ARGUMENT (src/ArnoldiMethod.jl:62-63).       */
int b2a_ws_create(b2a_ctx *ctx, int dtype, int64_t n_rows_local, int64_t n_global, int64_t row_offset,
                  int maxdim, b2a_ws **out);
int b2a_ws_destroy(b2a_ws *ws);

/* copy a host vector (n_rows_local entries) into column j (1-based) of V */
int b2a_ws_set_col(b2a_ws *ws, int j, const void *host);
/* same from a device pointer */
int b2a_ws_set_col_device(b2a_ws *ws, int j, const void *dev);
/* copy columns j0 .. j0+ncols-1 (1-based) of V to a host matrix with leading dim ld */
int b2a_ws_get_cols(b2a_ws *ws, int j0, int ncols, void *host, int64_t ld);
/* how the workspace's collectives run: 0 = single GPU, 1 = host-launched NCCL (all-reduce of h, all-gather of x),
 * 2 = fused into the kernels over NVLink peer memory (csrc/peer_comm.cuh).  One workspace per context owns the
 * peer block at a time; a second live workspace on the same context gets mode 1. */
int b2a_ws_comm_mode(b2a_ws *ws, int *mode);
/* diagnostics: per-CTA phase timestamps (ns, %globaltimer) of the last fused orthogonalisation kernel
 * (csrc/kernels_cgs_sweep.cuh); only for workspaces created with B2A_SWEEP_TRACE=1 in the environment.
 * out receives min(max_ctas, #SMs) x *slots values. */
int b2a_ws_debug_sweep_trace(b2a_ws *ws, unsigned long long *out, int max_ctas, int *slots);
/* device pointer of column j (1-based) and the leading dimension (elements) */
int b2a_ws_col_ptr(b2a_ws *ws, int j, void **dev, int64_t *ld);
/* host H / Q of the workspace (column-major, leading dims returned) */
int b2a_ws_host_arrays(b2a_ws *ws, void **H, int *ldh, void **Q, int *ldq);

/* ------------------------------------------------------------ hot path, per call
 * These let the reference's own `_partialschur` (src/run.jl:224-392) drive the device:
 * a Julia wrapper adds methods for ArnoldiWorkspace{T,<:B200Matrix} that ccall them. */

typedef struct b2a_stats {
  int64_t matvecs;       /* operator applications                                        */
  int64_t passes;        /* Gram-Schmidt passes executed (1 or 2 per step)               */
  int64_t second_passes; /* steps whose DGKS test fired (src/expansion.jl:91)            */
  int64_t breakdowns;    /* steps that returned false (src/expansion.jl:99-102)          */
  int64_t launches;      /* kernels launched by this call                                */
  double bytes;          /* algorithmic bytes moved by this call           */
} b2a_stats;

/* reinitialize!(arnoldi, j, populate!) - src/expansion.jl:12-59.
 * mode = B2A_INIT_RAND fills column j+1 with counter-based uniform [0,1) numbers keyed by
 * (seed, global row) - independent of the GPU count; B2A_INIT_KEEP keeps its contents.
 * *ok = 1 if the column is a valid new basis vector (the reference's Bool). */
int b2a_reinitialize(b2a_ws *ws, int j, int mode, uint64_t seed, int *ok);

/* orthogonalize!(arnoldi, j) - src/expansion.jl:69-109.  Orthogonalises column j+1 against
 * columns 1..j; writes H[1:j+1, j] into the workspace's host H (and into h_host if not
 * NULL: j+1 entries); *ok = 0 on breakdown (H[j+1,j] = 0, column left un-normalised). */
int b2a_orthogonalize(b2a_ws *ws, int j, void *h_host, int *ok);

/* iterate_arnoldi!(A, arnoldi, from:to) - src/expansion.jl:116-133.  All steps of the range
 * are enqueued asynchronously; H columns from..to of the workspace's host H are filled.
 * If H_host != NULL they are also copied there (column-major, leading dim ldh). */
int b2a_iterate_arnoldi(b2a_ws *ws, b2a_op *A, int from, int to, uint64_t seed, void *H_host, int ldh,
                        b2a_stats *stats);

/* The restart's change of basis - src/run.jl:363-365:
 *   V[:, purge:k] <- V[:, purge:maxdim] * Q[purge:maxdim, purge:k];  V[:, k+1] <- V[:, maxdim+1]
 * Q_host: maxdim x maxdim column-major with leading dim ldq (NULL = the workspace's Q). */
int b2a_rotate_basis(b2a_ws *ws, int purge, int k, int maxdim, const void *Q_host, int ldq,
                     b2a_stats *stats);

/* The final change of basis - src/run.jl:382-383:  V[:, 1:nconv] <- V[:, 1:nconv] * Q[1:nconv, 1:nconv] */
int b2a_rotate_final(b2a_ws *ws, int nconv, const void *Q_host, int ldq, b2a_stats *stats);

/* partialeigen's n-sized product - src/eigvals.jl:94:  X <- V[:, 1:nconv] * Y, Y complex
 * nconv x nconv (interleaved), X complex n_rows_local x nconv on the host (ld = ldx). */
int b2a_basis_times(b2a_ws *ws, int nconv, const double *Y_host_c64, int ldy, double *X_host_c64,
                    int64_t ldx);

/* single operator application y = A x on workspace columns: V[:, jdst] <- A V[:, jsrc] */
int b2a_ws_matvec(b2a_ws *ws, b2a_op *A, int jsrc, int jdst);

/* ---------------------------------------------------------- fine-grained vector operations
 * The BLAS-level operations the reference's GENERIC code performs on V-typed objects
 * (src/expansion.jl:17-57,70-107: norm, Vprev' * v, mul!(v, Vprev, h, -1, 1), v ./= a, copyto!), one
 * entry point each, so that the unmodified generic methods can run on a device array type (slow path,
 * one call + host round trip per operation; the fused sweeps above are the fast path).  Columns 1-based. */

/* norm(view(V, :, j)) - global 2-norm (all-reduced when sharded) */
int b2a_ws_nrm2(b2a_ws *ws, int j, double *result);
/* h = view(V, :, 1:ncols)' * view(V, :, j)   (h_host: ncols entries of the workspace dtype) */
int b2a_ws_gemv_c(b2a_ws *ws, int ncols, int j, void *h_host);
/* mul!(view(V, :, j), view(V, :, 1:ncols), h, -1, 1):  V[:, j] -= V[:, 1:ncols] * h */
int b2a_ws_gemv_n_sub(b2a_ws *ws, int ncols, int j, const void *h_host);
/* view(V, :, j) ./= alpha (alpha real) */
int b2a_ws_scal_div(b2a_ws *ws, int j, double alpha);
/* copyto!(view(V, :, jdst), view(V, :, jsrc)) */
int b2a_ws_copy_col(b2a_ws *ws, int jsrc, int jdst);

/* ------------------------------------------------------------- whole restart loop
 * partialschur / partialschur! - src/run.jl:100-179 + _partialschur :224-392, with the
 * m x m algebra (src/schurfact.jl, src/schursort.jl, src/restore_hessenberg.jl,
 * src/eigvals.jl, src/eigenvector_uppertriangular.jl, src/targets.jl) on the host in C++. */

typedef struct b2a_params {
  int32_t nev;        /* 0: min(6, n)                                     run.jl:103   */
  int32_t which;      /* enum b2a_which                                   run.jl:104   */
  double tol;         /* < 0: sqrt(eps)                                   run.jl:105   */
  int32_t mindim;     /* 0: min(max(10, nev), n[, size(V,2)-1])           run.jl:106   */
  int32_t maxdim;     /* 0: min(max(20, 2nev), n[, size(V,2)-1])          run.jl:107   */
  int32_t restarts;   /* < 0: 200                                         run.jl:108   */
  int32_t start_from; /* <= 0: 1                                          run.jl:155   */
  int32_t initialize; /* enum b2a_init; < 0: RAND if start_from == 1 else NONE  run.jl:156 */
  uint64_t seed;      /* seed of the counter-based start / re-seed vectors             */
} b2a_params;

typedef struct b2a_history {
  int64_t mvproducts; /* History.mvproducts                               run.jl:218   */
  int32_t nconverged; /* History.nconverged                                            */
  int32_t converged;  /* History.converged                                             */
  int32_t nev;        /* History.nev                                                   */
  int32_t restarts;   /* restart iterations executed                                   */
  b2a_stats stats;    /* totals over the run                                           */
  double ms_expand;   /* host wall clock spent waiting on expansion sweeps             */
  double ms_rotate;   /* ... on basis rotations                                        */
  double ms_small;    /* ... in the m x m host algebra                                 */
} b2a_history;

/* Runs the restart loop on workspace `ws`.  On return V[:, 1:nconverged] holds the Schur
 * vectors, host H[1:nconverged, 1:nconverged] the (quasi) upper triangular R, and
 * eigenvalues_c64 (if not NULL, 2*maxdim doubles) the eigenvalues (re,im interleaved). */
int b2a_partialschur(b2a_ws *ws, b2a_op *A, const b2a_params *params, b2a_history *history,
                     double *eigenvalues_c64);

/* ------------------------------------------------- host m x m algebra (no GPU needed)
 * Exposed for non-Julia hosts and for CPU-side testing of the C++ driver against the
 * oracle; a Julia host keeps using the reference's own functions instead. */

/* local_schurfact!(view(H, 1:m, :), from, to, Q) - src/schurfact.jl:393/492.
 * H: rows x cols column-major (ldh), Q: qrows x cols (ldq) or NULL. */
int b2a_host_local_schurfact(int dtype, void *H, int ldh, int rows, int cols, int from, int to, void *Q,
                             int ldq, int qrows);

/* One restart's host work, src/run.jl:278-360: Q <- I, Schur form, Ritz values and residuals,
 * lock / retain / purge partition, restore_arnoldi!.  H is (maxdim+1) x maxdim, Q maxdim x maxdim.
 * Outputs: k, purge, nlock; eigenvalues_c64 / residuals (maxdim each) if not NULL. */
int b2a_host_restart(int dtype, void *H, int ldh, void *Q, int ldq, int maxdim, int mindim, int nev,
                     double tol, int which, int active, int *k, int *purge, int *nlock,
                     double *eigenvalues_c64, double *residuals);

/* sortschur!(H, Q <- I, nconv, ordering) - src/run.jl:379,465-502 */
int b2a_host_sortschur(int dtype, void *H, int ldh, void *Q, int ldq, int maxdim, int nconv, int which);

/* Upload-time plan of the CSR mat-vec (no GPU needed): number of column blocks b2a_csr_create would use for an
 * operator with n_global columns, `nnz_per_row` entries per row and a mean |column - row| of
 * `mean_col_distance` (in elements); 1 = plain CSR.  See the cost model at col_block_plan() in csrc/b2a.cu. */
int b2a_host_col_block_plan(int dtype, int64_t n_global, double nnz_per_row, double mean_col_distance, int *nblocks);
/* Owner groups of a row-sharded operator (host logic, no GPU): how many ranks' slices of x form one column block of
 * the mat-vec (about 32 MB of x), how many blocks that gives, and - optionally - the block of every owner rank as
 * seen from `rank` (block 0 = own slice and the ones that arrive first in the staged exchange). */
int b2a_host_owner_group_plan(int dtype, int64_t n_global, int world, int rank, int *ranks_per_block, int *nblocks,
                              int *block_of_owner);

/* givensAlgorithm(f, g) -> (c, s, r); f, g, s, r are 1 (F64) or 2 (C64) doubles */
int b2a_host_givens(int dtype, const double *f, const double *g, double *c, double *s, double *r);

#ifdef __cplusplus
}
#endif
#endif /* B200ARNOLDI_H */
