"""Host-side mirror of the reference's public API for the hot path.

Same names, keyword arguments, defaults and error behaviour as
``src/run.jl:100-179`` / ``src/ArnoldiMethod.jl:41-137`` / ``src/eigvals.jl:92-95``;
Julia's ``ArgumentError`` maps to ``ValueError``, ``DimensionMismatch`` to
``DimensionMismatch`` (a ``ValueError`` subclass).  Everything n-sized happens
inside ``libb200arnoldi.so``; this file only marshals arguments.
"""

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from . import _lib as L
from . import sharding

_MASK64 = (1 << 64) - 1


def _dtype_code(dt):
    dt = np.dtype(dt)
    if np.issubdtype(dt, np.complexfloating):
        return L.C64, np.dtype(np.complex128)
    return L.F64, np.dtype(np.float64)  # vtype(A), src/run.jl:9-12: ints/bools/float32 operate in Float64


def _which_code(which):
    if not isinstance(which, str):
        which = getattr(which, "name", None) or type(which).__name__
    key = str(which).lstrip(":").upper()
    if key not in L.WHICH:
        raise ValueError(f"Unknown target: {which}")  # src/run.jl:185
    return L.WHICH[key]


# ----------------------------------------------------------------------------- context
class Context:
    """One per process and GPU (``b2a_ctx``).  ``world > 1``: one rank of a row-sharded job."""

    def __init__(self, device=0, rank=0, world=1, nccl_unique_id=None):
        self._h = C.c_void_p()
        lib = L.lib()
        if world == 1:
            L.check(lib.b2a_ctx_create(int(device), C.byref(self._h)))
        else:
            if nccl_unique_id is None or len(nccl_unique_id) != 128:
                raise ValueError("nccl_unique_id must be the 128 bytes of Context.nccl_unique_id()")
            buf = (C.c_char * 128).from_buffer_copy(bytes(nccl_unique_id))
            L.check(lib.b2a_ctx_create_dist(int(device), int(rank), int(world), buf, C.byref(self._h)))
        self.device, self.rank, self.world = int(device), int(rank), int(world)

    @staticmethod
    def nccl_unique_id():
        buf = (C.c_char * 128)()
        L.check(L.lib().b2a_nccl_unique_id(buf))
        return bytes(buf)

    @classmethod
    def from_torch_distributed(cls, device=None):
        """Build the context of this rank from an initialised ``torch.distributed`` group
        (the unique id travels over the group's own backend)."""
        import torch
        import torch.distributed as dist

        rank, world = dist.get_rank(), dist.get_world_size()
        if device is None:
            device = torch.cuda.current_device()
        if world == 1:
            return cls(device)
        payload = [cls.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(payload, src=0)
        return cls(device, rank, world, payload[0])

    @property
    def stream(self):
        p = C.c_void_p()
        L.check(L.lib().b2a_ctx_stream(self._h, C.byref(p)))
        return p.value or 0

    def synchronize(self):
        L.check(L.lib().b2a_ctx_sync(self._h))

    @property
    def launches(self):
        v = C.c_int64()
        L.check(L.lib().b2a_ctx_launch_count(self._h, C.byref(v)))
        return v.value

    def profile(self, on=True):
        """Enable / disable (and reset) per-kernel-kind CUDA-event timing."""
        L.check(L.lib().b2a_ctx_profile_enable(self._h, 1 if on else 0))

    def profile_report(self):
        """-> {kind: dict(launches, ms, bytes)} accumulated since ``profile(True)``."""
        out = {}
        for k, name in enumerate(L.KERNEL_KINDS):
            n, ms, by = C.c_int64(), C.c_double(), C.c_double()
            L.check(L.lib().b2a_ctx_profile_get(self._h, k, C.byref(n), C.byref(ms), C.byref(by)))
            out[name] = dict(launches=n.value, ms=ms.value, bytes=by.value)
        return out

    def close(self):
        if self._h:
            L.lib().b2a_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


# ---------------------------------------------------------------------------- operator
def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Operator:
    """The linear map ``A`` of ``mul!(y, A, x)`` (src/run.jl:21-25, src/expansion.jl:121)."""

    def __init__(self, ctx, handle, dtype, n_local, n_global, row_offset, keep=()):
        self.ctx, self._h = ctx, handle
        self.dtype = np.dtype(dtype)
        self.n_local, self.n_global, self.row_offset = int(n_local), int(n_global), int(row_offset)
        self.shape = (self.n_global, self.n_global)
        self._keep = keep

    # -- constructors
    @classmethod
    def from_csr_arrays(cls, ctx, indptr, indices, data, n_global, row_offset=0, idx_base=0):
        """Rows ``[row_offset, row_offset + len(indptr) - 1)`` of A; global column numbers."""
        code, dt = _dtype_code(data.dtype)
        data = np.ascontiguousarray(data, dtype=dt)
        indptr = np.ascontiguousarray(indptr)
        indices = np.ascontiguousarray(indices)
        if indices.dtype not in (np.int32, np.int64):
            indices = indices.astype(np.int64 if indices.dtype.itemsize > 4 else np.int32)
        if indptr.dtype != indices.dtype:
            # one index width crosses the ABI: convert the SMALL array (rowptr, n+1 entries), never the
            # nnz-sized one - unless the row pointers do not fit 32 bits
            if indices.dtype == np.int32 and int(indptr[-1]) + int(idx_base) >= 2**31 - 1:
                indices = indices.astype(np.int64)
            indptr = indptr.astype(indices.dtype)
        n_local = indptr.shape[0] - 1
        h = C.c_void_p()
        L.check(
            L.lib().b2a_csr_create(
                ctx._h, code, n_local, int(n_global), int(row_offset), int(data.shape[0]), _ptr(indptr), _ptr(indices),
                _ptr(data), indptr.dtype.itemsize * 8, int(idx_base), C.byref(h),
            )
        )
        return cls(ctx, h, dt, n_local, n_global, row_offset)

    @classmethod
    def from_csc_arrays(cls, ctx, colptr, rowval, nzval, n_global, idx_base=0, mode=0):
        """Julia's ``SparseMatrixCSC`` fields as they are (``idx_base=1`` for Julia arrays).
        mode 0: transpose once at upload + CSR kernel; mode 1: native scatter kernel."""
        code, dt = _dtype_code(nzval.dtype)
        nzval = np.ascontiguousarray(nzval, dtype=dt)
        colptr = np.ascontiguousarray(colptr)
        rowval = np.ascontiguousarray(rowval)
        if colptr.dtype != rowval.dtype or colptr.dtype not in (np.int32, np.int64):
            colptr, rowval = colptr.astype(np.int64), rowval.astype(np.int64)
        h = C.c_void_p()
        L.check(
            L.lib().b2a_csc_create(
                ctx._h, code, int(n_global), int(nzval.shape[0]), _ptr(colptr), _ptr(rowval), _ptr(nzval),
                colptr.dtype.itemsize * 8, int(idx_base), int(mode), C.byref(h),
            )
        )
        off, cnt = sharding.local_rows(int(n_global), ctx.rank, ctx.world)  # world == 1: the whole matrix
        return cls(ctx, h, dt, cnt, n_global, off)

    @classmethod
    def from_matrix(cls, ctx, A, layout="auto", csc_mode=0):
        """From a SciPy sparse matrix or a dense array.  ``ctx.world > 1``: this rank keeps
        its row block of the (replicated) global matrix."""
        import scipy.sparse as sp

        if A.ndim != 2 or A.shape[0] != A.shape[1]:
            # checksquare, src/run.jl:110
            raise L.DimensionMismatch(f"matrix is not square: dimensions are {tuple(A.shape)}")
        n = A.shape[0]
        _, dt = _dtype_code(A.dtype)
        if sp.issparse(A) and A.format == "csc" and layout in ("auto", "csc"):
            # Julia's native layout as it is; row-sharded jobs keep this rank's row block (transpose at upload)
            return cls.from_csc_arrays(ctx, A.indptr, A.indices, A.data.astype(dt, copy=False), n, 0,
                                       csc_mode if ctx.world == 1 else 0)
        M = A.tocsr() if sp.issparse(A) else sp.csr_matrix(np.asarray(A))
        M.sort_indices()
        off, cnt, ip, idx, dat = sharding.shard_csr(M.indptr, M.indices, M.data, n, ctx.rank, ctx.world)
        return cls.from_csr_arrays(ctx, ip, idx, dat.astype(dt, copy=False), n, off)

    @classmethod
    def from_callback(cls, ctx, dtype, n_local, n_global, fn):
        """Matrix-free operator.  ``fn(x_ptr, y_ptr, n_local, stream_ptr) -> int`` must enqueue
        ``y <- A x`` on the given CUDA stream (device pointers as ints) and return 0."""
        code, dt = _dtype_code(dtype)

        def tramp(_user, x, y, n, stream):
            try:
                return int(fn(x or 0, y or 0, int(n), stream or 0) or 0)
            except Exception:  # an exception must not cross the C boundary
                import traceback

                traceback.print_exc()
                return 1

        cfn = L.MATVEC_FN(tramp)
        h = C.c_void_p()
        L.check(L.lib().b2a_op_from_callback(ctx._h, code, int(n_local), int(n_global), cfn, None, C.byref(h)))
        return cls(ctx, h, dt, n_local, n_global, 0, keep=(cfn, fn))

    @classmethod
    def from_torch_function(cls, ctx, dtype, n, fn):
        """Matrix-free operator from ``fn(x: torch.Tensor) -> torch.Tensor`` (both on the GPU)."""
        import torch

        _, dt = _dtype_code(dtype)
        typestr = "<c16" if dt == np.complex128 else "<f8"

        class _Dev:
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {
                    "shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2, "strides": None,
                }

        def cb(x_ptr, y_ptr, n_local, stream_ptr):
            dev = torch.device("cuda", ctx.device)
            with torch.cuda.stream(torch.cuda.ExternalStream(stream_ptr, device=dev)):
                x = torch.as_tensor(_Dev(x_ptr, n_local), device=dev)
                y = torch.as_tensor(_Dev(y_ptr, n_local), device=dev)
                y.copy_(fn(x))
            return 0

        return cls.from_callback(ctx, dt, n, n, cb)

    @classmethod
    def shift_invert(cls, A, sigma=0.0, rtol=1e-13, maxit=10000):
        """The linear map ``x -> (A - sigma I) \\ x`` on the device (docs/src/index.md:234-262 builds it from a
        factorisation and LinearMaps.jl): Jacobi-preconditioned CG on the CSR mat-vec of ``A``, for a Hermitian positive
        definite ``A - sigma I``.  ``partialschur(op, which="LM")`` then finds the eigenvalues of ``A`` nearest
        ``sigma`` as ``sigma + 1 / theta``.  ``A`` (an ``Operator``) must stay alive."""
        sigma = complex(sigma)
        h = C.c_void_p()
        L.check(L.lib().b2a_op_shift_invert(A.ctx._h, A._h, sigma.real, sigma.imag, L.SOLVE_CG, float(rtol), int(maxit),
                                            C.byref(h)))
        return cls(A.ctx, h, A.dtype, A.n_local, A.n_global, A.row_offset, keep=(A,))

    @property
    def solve_stats(self):
        """(solves, inner iterations, worst relative residual) of a shift-and-invert operator."""
        a, b, w = C.c_int64(), C.c_int64(), C.c_double()
        L.check(L.lib().b2a_op_solve_stats(self._h, C.byref(a), C.byref(b), C.byref(w)))
        return a.value, b.value, w.value

    @property
    def bytes_per_matvec(self):
        v = C.c_double()
        L.check(L.lib().b2a_op_bytes(self._h, C.byref(v)))
        return v.value

    def close(self):
        if self._h:
            L.lib().b2a_op_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# --------------------------------------------------------------------------- workspace
class ArnoldiWorkspace:
    """``ArnoldiWorkspace`` (src/ArnoldiMethod.jl:41-93) with V resident in HBM.

    ``ArnoldiWorkspace(n, k)`` / ``ArnoldiWorkspace(v1, k)`` as in the reference (dtype from
    ``dtype=`` / from ``v1``).  ``H`` ((k+1) x k) and ``Q`` (k x k) are host arrays (NumPy
    views of the library's memory); ``V_tmp`` does not exist (in-place rotation).
    """

    def __init__(self, n_or_v1, krylov_dimension, dtype=np.float64, ctx=None, n_global=None, row_offset=None):
        ctx = ctx or default_context()
        v1 = None
        if np.ndim(n_or_v1) == 1:
            v1 = np.asarray(n_or_v1)
            dtype = v1.dtype
            n = v1.shape[0]
        else:
            n = int(n_or_v1)
        code, dt = _dtype_code(dtype)
        if n_global is None:  # n is the global order: take this rank's block
            n_global = n
            row_offset, n_local = sharding.local_rows(n, ctx.rank, ctx.world)
            if v1 is not None:
                v1 = v1[row_offset : row_offset + n_local]
        else:
            n_local = n
            row_offset = int(row_offset or 0)
        self.ctx, self.dtype = ctx, dt
        self.n_local, self.n_global, self.row_offset = int(n_local), int(n_global), int(row_offset)
        self.maxdim = int(krylov_dimension)
        self._h = C.c_void_p()
        L.check(
            L.lib().b2a_ws_create(ctx._h, code, self.n_local, self.n_global, self.row_offset, self.maxdim, C.byref(self._h))
        )
        Hp, Qp, ldh, ldq = C.c_void_p(), C.c_void_p(), C.c_int(), C.c_int()
        L.check(L.lib().b2a_ws_host_arrays(self._h, C.byref(Hp), C.byref(ldh), C.byref(Qp), C.byref(ldq)))
        m = self.maxdim
        ctype = C.c_double * ((m + 1) * m * (2 if dt == np.complex128 else 1))
        self.H = np.frombuffer(ctype.from_address(Hp.value), dtype=dt).reshape((m, m + 1)).T
        ctype = C.c_double * (m * m * (2 if dt == np.complex128 else 1))
        self.Q = np.frombuffer(ctype.from_address(Qp.value), dtype=dt).reshape((m, m)).T
        if v1 is not None:
            self.set_col(1, v1)

    def set_col(self, j, vec):
        """``copyto!(view(V, :, j), vec)`` - j is 1-based like the reference."""
        vec = np.ascontiguousarray(vec, dtype=self.dtype)
        if vec.shape != (self.n_local,):
            raise ValueError("v1 should have the same dimension as A")  # src/run.jl:123-124
        L.check(L.lib().b2a_ws_set_col(self._h, int(j), _ptr(vec)))

    def get_cols(self, j0, ncols, out=None):
        """Host copy of ``V[:, j0:j0+ncols-1]`` (1-based), column-major.  ``out``: optional
        (e.g. pinned) Fortran-ordered destination with at least ``ncols`` columns."""
        if out is None:
            out = np.zeros((self.n_local, ncols), dtype=self.dtype, order="F")
        else:
            assert out.dtype == self.dtype and out.flags.f_contiguous and out.shape[0] == self.n_local
            out = out[:, :ncols]
        if ncols > 0 and self.n_local > 0:
            L.check(L.lib().b2a_ws_get_cols(self._h, int(j0), int(ncols), _ptr(out), max(self.n_local, 1)))
        return out

    @property
    def V(self):
        return self.get_cols(1, self.maxdim + 1)

    def col_ptr(self, j):
        p, ld = C.c_void_p(), C.c_int64()
        L.check(L.lib().b2a_ws_col_ptr(self._h, int(j), C.byref(p), C.byref(ld)))
        return p.value, ld.value

    # -- fine-grained hot-path calls (what a Julia wrapper ccalls)
    def reinitialize(self, j=0, mode="rand", seed=0):
        ok = C.c_int()
        m = L.INIT_RAND if mode == "rand" else L.INIT_KEEP
        L.check(L.lib().b2a_reinitialize(self._h, int(j), m, int(seed) & _MASK64, C.byref(ok)))
        return bool(ok.value)

    def orthogonalize(self, j):
        ok = C.c_int()
        L.check(L.lib().b2a_orthogonalize(self._h, int(j), None, C.byref(ok)))
        return bool(ok.value)

    def matvec(self, A, jsrc, jdst):
        L.check(L.lib().b2a_ws_matvec(self._h, A._h, int(jsrc), int(jdst)))

    def iterate_arnoldi(self, A, frm, to, seed=0):
        st = L.Stats()
        L.check(L.lib().b2a_iterate_arnoldi(self._h, A._h, int(frm), int(to), int(seed) & _MASK64, None, 0, C.byref(st)))
        return st

    def rotate_basis(self, purge, k, maxdim, Q=None):
        st = L.Stats()
        if Q is None:
            L.check(L.lib().b2a_rotate_basis(self._h, int(purge), int(k), int(maxdim), None, 0, C.byref(st)))
        else:
            Qf = np.asfortranarray(Q, dtype=self.dtype)
            L.check(L.lib().b2a_rotate_basis(self._h, int(purge), int(k), int(maxdim), _ptr(Qf), Qf.shape[0], C.byref(st)))
        return st

    def rotate_final(self, nconv, Q=None):
        st = L.Stats()
        if Q is None:
            L.check(L.lib().b2a_rotate_final(self._h, int(nconv), None, 0, C.byref(st)))
        else:
            Qf = np.asfortranarray(Q, dtype=self.dtype)
            L.check(L.lib().b2a_rotate_final(self._h, int(nconv), _ptr(Qf), Qf.shape[0], C.byref(st)))
        return st

    @property
    def comm_mode(self):
        """'single' | 'nccl' (host-launched collectives) | 'peer' (collectives fused into the kernels over
        NVLink peer memory; one workspace per context owns the peer block at a time)."""
        m = C.c_int()
        L.check(L.lib().b2a_ws_comm_mode(self._h, C.byref(m)))
        return ("single", "nccl", "peer")[m.value]

    # -- BLAS-level operations of the reference's generic code (slow path: one call each)
    def norm(self, j):
        """``norm(view(V, :, j))``"""
        r = C.c_double()
        L.check(L.lib().b2a_ws_nrm2(self._h, int(j), C.byref(r)))
        return r.value

    def gemv_c(self, ncols, j):
        """``view(V, :, 1:ncols)' * view(V, :, j)``"""
        h = np.zeros(ncols, dtype=self.dtype)
        if ncols:
            L.check(L.lib().b2a_ws_gemv_c(self._h, int(ncols), int(j), _ptr(h)))
        return h

    def gemv_n_sub(self, ncols, j, h):
        """``mul!(view(V, :, j), view(V, :, 1:ncols), h, -1, 1)``"""
        h = np.ascontiguousarray(h, dtype=self.dtype)
        assert h.shape == (ncols,)
        if ncols:
            L.check(L.lib().b2a_ws_gemv_n_sub(self._h, int(ncols), int(j), _ptr(h)))

    def scal_div(self, j, alpha):
        """``view(V, :, j) ./= alpha``"""
        L.check(L.lib().b2a_ws_scal_div(self._h, int(j), float(alpha)))

    def copy_col(self, jsrc, jdst):
        """``copyto!(view(V, :, jdst), view(V, :, jsrc))``"""
        L.check(L.lib().b2a_ws_copy_col(self._h, int(jsrc), int(jdst)))

    def basis_times(self, Y):
        """``V[:, 1:nconv] * Y`` for a small complex Y (partialeigen)."""
        Y = np.asfortranarray(Y, dtype=np.complex128)
        nconv = Y.shape[0]
        X = np.zeros((self.n_local, nconv), dtype=np.complex128, order="F")
        if nconv and self.n_local:
            L.check(L.lib().b2a_basis_times(self._h, nconv, _ptr(Y), nconv, _ptr(X), self.n_local))
        return X

    def close(self):
        if self._h:
            self.H = self.Q = None
            L.lib().b2a_ws_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------ results
@dataclass
class History:
    """src/run.jl:217-222 plus device statistics."""

    mvproducts: int
    nconverged: int
    converged: bool
    nev: int
    restarts: int = 0
    stats: dict = field(default_factory=dict)
    timers_ms: dict = field(default_factory=dict)

    def __str__(self):  # src/show.jl:3-21
        s = "Converged" if self.converged else "Not converged"
        return f"{s}: {self.nconverged} of {self.nev} eigenvalues in {self.mvproducts} matrix-vector products"


class PartialSchur:
    """src/ArnoldiMethod.jl:130-137.  ``Q`` is fetched from the device on first use
    (this rank's row block when sharded); ``R`` and ``eigenvalues`` are host arrays."""

    def __init__(self, workspace, nconv, R, eigenvalues):
        self.workspace, self._nconv = workspace, nconv
        self.R, self.eigenvalues = R, eigenvalues
        self._Q = None

    @property
    def Q(self):
        if self._Q is None:
            self._Q = self.workspace.get_cols(1, self._nconv)
        return self._Q


def partialeigen(P):
    """``partialeigen`` (src/eigvals.jl:92-95): LAPACK ``eigen(R)`` on the host, ``Q * vecs``
    on the device."""
    vals, vecs = np.linalg.eig(P.R) if P.R.shape[0] else (np.zeros(0, complex), np.zeros((0, 0)))
    X = P.workspace.basis_times(vecs)
    if not np.iscomplexobj(vecs) and P.workspace.dtype == np.float64:
        X = np.ascontiguousarray(X.real)
    return vals, X


# -------------------------------------------------------------------------------- drivers
def _as_operator(A, ctx):
    if isinstance(A, Operator):
        return A, False
    return Operator.from_matrix(ctx, A), True


def _run(ws, op, nev, which, tol, mindim, maxdim, restarts, start_from, initialize, seed):
    p = L.Params(
        nev=int(nev), which=_which_code(which), tol=float(tol), mindim=int(mindim), maxdim=int(maxdim),
        restarts=int(restarts), start_from=int(start_from), initialize=int(initialize), seed=int(seed) & _MASK64,
    )
    hist = L.HistoryC()
    eig = np.zeros(2 * ws.maxdim, dtype=np.float64)
    L.check(L.lib().b2a_partialschur(ws._h, op._h, C.byref(p), C.byref(hist), _ptr(eig)))
    nconv = hist.nconverged
    lam = eig[: 2 * nconv].view(np.complex128).copy()
    R = np.array(ws.H[:nconv, :nconv], order="F")
    st = hist.stats
    history = History(
        hist.mvproducts, nconv, bool(hist.converged), hist.nev, hist.restarts,
        dict(matvecs=st.matvecs, passes=st.passes, second_passes=st.second_passes, breakdowns=st.breakdowns,
             launches=st.launches, bytes=st.bytes),
        dict(expand=hist.ms_expand, rotate=hist.ms_rotate, small=hist.ms_small),
    )
    return PartialSchur(ws, nconv, R, lam), history


def partialschur(A, v1=None, nev=None, which="LM", tol=None, mindim=None, maxdim=None, restarts=200, ctx=None, seed=0):
    """``partialschur(A; v1, nev, which, tol, mindim, maxdim, restarts)`` - src/run.jl:100-129.

    ``A``: SciPy sparse matrix (CSR, or CSC = Julia's native layout), dense array, or an
    ``Operator`` (e.g. a matrix-free callback).  Returns ``(PartialSchur, History)``.
    """
    ctx = ctx or (A.ctx if isinstance(A, Operator) else default_context())
    shape = A.shape
    if len(shape) != 2 or shape[0] != shape[1]:
        raise L.DimensionMismatch(f"matrix is not square: dimensions are {tuple(shape)}")  # run.jl:110
    n = shape[0]
    nev = min(6, n) if nev is None else nev
    tol = math.sqrt(np.finfo(np.float64).eps) if tol is None else tol
    mindim = min(max(10, nev), n) if mindim is None else mindim
    maxdim = min(max(20, 2 * nev), n) if maxdim is None else maxdim
    if nev < 1:
        raise ValueError("nev cannot be less than 1")
    if not (nev <= mindim <= maxdim <= n):
        raise ValueError(f"nev ≤ mindim ≤ maxdim ≤ size(A, 1) does not hold, got {nev} ≤ {mindim} ≤ {maxdim} ≤ {n}")
    _which_code(which)
    if v1 is not None and np.shape(v1)[0] != n:
        raise ValueError("v1 should have the same dimension as A")
    if v1 is not None and np.iscomplexobj(v1):
        # the reference takes the workspace type from v1 (`ArnoldiWorkspace(v1, maxdim)`, src/run.jl:125): a complex
        # start vector with a real A runs in ComplexF64.  Device kernels need one element type, so A is promoted.
        if isinstance(A, Operator):
            if A.dtype != np.complex128:
                raise ValueError("complex v1 with a Float64 device operator: build the operator from a complex matrix")
        elif not np.issubdtype(A.dtype, np.complexfloating):
            A = A.astype(np.complex128)
    op, owned = _as_operator(A, ctx)
    try:
        ws = ArnoldiWorkspace(n, maxdim, dtype=op.dtype, ctx=ctx)
        if v1 is not None:
            off, cnt = ws.row_offset, ws.n_local
            ws.set_col(1, np.asarray(v1)[off : off + cnt])
        init = L.INIT_RAND if v1 is None else L.INIT_KEEP
        return _run(ws, op, nev, which, tol, mindim, maxdim, restarts, 1, init, seed)
    finally:
        if owned:
            op.close()


def partialschur_(A, arnoldi, start_from=1, initialize=None, nev=None, which="LM", tol=None, mindim=None, maxdim=None,
                  restarts=200, seed=0):
    """``partialschur!(A, arnoldi; start_from, initialize, ...)`` - src/run.jl:152-179."""
    shape = A.shape
    if len(shape) != 2 or shape[0] != shape[1]:
        raise L.DimensionMismatch(f"matrix is not square: dimensions are {tuple(shape)}")
    n = shape[0]
    vcols = arnoldi.maxdim + 1
    nev = min(6, n) if nev is None else nev
    tol = math.sqrt(np.finfo(np.float64).eps) if tol is None else tol
    mindim = min(max(10, nev), n, vcols - 1) if mindim is None else mindim
    maxdim = min(max(20, 2 * nev), n, vcols - 1) if maxdim is None else maxdim
    initialize = (start_from == 1) if initialize is None else initialize
    if nev < 1:
        raise ValueError("nev cannot be less than 1")
    if not (nev <= mindim <= maxdim <= n):
        raise ValueError(f"nev ≤ mindim ≤ maxdim ≤ size(A, 1) does not hold, got {nev} ≤ {mindim} ≤ {maxdim} ≤ {n}")
    if not maxdim < vcols:
        raise ValueError("maxdim should be strictly less than size(arnoldi.V, 2)")
    if not (1 <= start_from <= maxdim):
        raise ValueError("start_from should be between 1 and maxdim")
    _which_code(which)
    op, owned = _as_operator(A, arnoldi.ctx)
    try:
        return _run(arnoldi, op, nev, which, tol, mindim, maxdim, restarts, start_from,
                    L.INIT_RAND if initialize else L.INIT_NONE, seed)
    finally:
        if owned:
            op.close()


# ------------------------------------------------------------------- RNG reference (host)
def _splitmix64(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(_MASK64)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def uniform_reference(seed, counter, n, dtype=np.float64, row_offset=0):
    """Host restatement of the device's counter-based ``rand!`` stand-in (kernels_rotate.cuh
    ``fill_uniform_kernel``): the vector the ``counter``-th re-seed of a run with ``seed``
    produces, for rows ``row_offset .. row_offset+n-1``."""
    with np.errstate(over="ignore"):
        key = _splitmix64(np.uint64(seed & _MASK64) ^ (np.uint64(0x632BE59BD9B4E019) * np.uint64(counter + 1)))
        g = np.arange(row_offset, row_offset + n, dtype=np.uint64)

        def u(idx):
            bits = _splitmix64(key + idx * np.uint64(0xD1342543DE82EF95))
            return (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)

        if np.issubdtype(np.dtype(dtype), np.complexfloating):
            return u(np.uint64(2) * g) + 1j * u(np.uint64(2) * g + np.uint64(1))
        return u(np.uint64(2) * g)
