"""Oracle: Arnoldi expansion + Krylov-Schur restart loop on the CPU.

Restates the reference's
  * ``src/ArnoldiMethod.jl:41-137`` - ``ArnoldiWorkspace``, ``RitzValues``, ``PartialSchur``
  * ``src/expansion.jl``            - ``reinitialize!`` / ``orthogonalize!`` / ``iterate_arnoldi!``
  * ``src/run.jl:100-392,510-517``  - ``partialschur`` / ``partialschur!`` / ``_partialschur``
  * ``src/eigvals.jl:92-95``        - ``partialeigen``

The n-sized arithmetic uses the same class of kernels the Julia path uses:
``A @ x`` (SciPy sparse mat-vec, single-threaded, stands in for
``SparseArrays.mul!``) and BLAS ``gemv``/``gemm``/``nrm2`` through NumPy
(stands in for ``LinearAlgebra.mul!``/``norm``).

Index variables are 1-based as in the cited lines; array accesses subtract 1.
Test infrastructure only - see ``oracle/__init__.py``.
"""

import math
import time
from dataclasses import dataclass, field

import numpy as np

from . import dense_small as ds

ETA = math.sqrt(2.0) / 2.0  # expansion.jl:33,74 - ARPACK's DGKS constant


def vtype(A):
    """run.jl:9-12: the floating type a matrix of ``eltype(A)`` operates on."""
    dt = np.dtype(getattr(A, "dtype", np.float64))
    return np.complex128 if np.issubdtype(dt, np.complexfloating) else np.float64


class ArnoldiWorkspace:
    """ArnoldiMethod.jl:41-93.  V is n x (k+1) column-major, H is (k+1) x k."""

    def __init__(self, dtype_or_V, n_or_H=None, krylov_dimension=None, V_tmp=None, Q=None):
        if isinstance(dtype_or_V, np.ndarray) and dtype_or_V.ndim == 2:
            V, H = dtype_or_V, n_or_H
            if V.shape[1] != H.shape[0]:
                raise ValueError("V should have the same number of columns as H has rows.")
            if H.shape[0] != H.shape[1] + 1:
                raise ValueError("H should have one more row than it has columns.")
            self.V, self.H = V, H
            self.V_tmp = np.empty_like(V) if V_tmp is None else V_tmp
            self.Q = np.empty((H.shape[1], H.shape[1]), dtype=H.dtype, order="F") if Q is None else Q
            return
        dtype, n, k = np.dtype(dtype_or_V), int(n_or_H), int(krylov_dimension)
        if not k <= n:
            raise ValueError("Krylov dimension should be less than matrix order.")
        self.V = np.zeros((n, k + 1), dtype=dtype, order="F")
        self.V_tmp = np.zeros((n, k + 1), dtype=dtype, order="F")
        self.H = np.zeros((k + 1, k), dtype=dtype, order="F")
        self.Q = np.zeros((k, k), dtype=dtype, order="F")


@dataclass
class History:
    """run.jl:217-222."""

    mvproducts: int
    nconverged: int
    converged: bool
    nev: int
    # extras (not in the reference): phase timers for the CPU-baseline report
    timers: dict = field(default_factory=dict)
    second_passes: int = 0
    restarts_done: int = 0


@dataclass
class PartialSchur:
    """ArnoldiMethod.jl:130-137."""

    Q: np.ndarray
    R: np.ndarray
    eigenvalues: np.ndarray


class _Stats:
    def __init__(self):
        self.t = {"matvec": 0.0, "orth": 0.0, "rotate": 0.0, "small": 0.0}
        self.second_passes = 0
        self.restarts = 0


def _mul(A, x, out):
    """``mul!(y, A, x)`` - the operator contract of run.jl:24-25."""
    if hasattr(A, "mul"):
        A.mul(out, x)
    else:
        out[:] = A @ x


def _default_rand(rng):
    def populate(v):
        if np.iscomplexobj(v):
            v[:] = rng.random(v.shape[0]) + 1j * rng.random(v.shape[0])
        else:
            v[:] = rng.random(v.shape[0])

    return populate


def reinitialize(arnoldi, j=0, populate=None, rng=None):
    """``reinitialize!`` (expansion.jl:12-59)."""
    V = arnoldi.V
    v = V[:, j]
    if populate is None:
        populate = _default_rand(rng if rng is not None else np.random.default_rng())
    populate(v)
    rnorm = np.linalg.norm(v)
    if j == 0:
        v /= rnorm
        return True
    Vprev = V[:, :j]
    h = Vprev.conj().T @ v
    v -= Vprev @ h
    wnorm = np.linalg.norm(v)
    if wnorm < ETA * rnorm:
        rnorm = wnorm
        h = Vprev.conj().T @ v
        v -= Vprev @ h
        wnorm = np.linalg.norm(v)
    if wnorm <= ETA * rnorm:
        return False
    v /= wnorm
    return True


def orthogonalize(arnoldi, j, stats=None):
    """``orthogonalize!`` (expansion.jl:69-109): CGS + one DGKS correction."""
    V, H = arnoldi.V, arnoldi.H
    Vprev = V[:, :j]
    v = V[:, j]
    rnorm = np.linalg.norm(v)
    h = Vprev.conj().T @ v
    v -= Vprev @ h
    wnorm = np.linalg.norm(v)
    if wnorm < ETA * rnorm:
        rnorm = wnorm
        correction = Vprev.conj().T @ v
        v -= Vprev @ correction
        h += correction
        wnorm = np.linalg.norm(v)
        if stats is not None:
            stats.second_passes += 1
    H[:j, j - 1] = h
    if wnorm <= ETA * rnorm:
        H[j, j - 1] = 0
        return False
    H[j, j - 1] = wnorm
    v /= wnorm
    return True


def iterate_arnoldi(A, arnoldi, frm, to, rng=None, stats=None):
    """``iterate_arnoldi!(A, arnoldi, from:to)`` (expansion.jl:116-133)."""
    V = arnoldi.V
    for j in range(frm, to + 1):
        t0 = time.perf_counter()
        _mul(A, V[:, j - 1], V[:, j])
        t1 = time.perf_counter()
        ok = orthogonalize(arnoldi, j, stats)
        if ok is False and j != V.shape[0]:
            reinitialize(arnoldi, j, rng=rng)
        if stats is not None:
            stats.t["matvec"] += t1 - t0
            stats.t["orth"] += time.perf_counter() - t1
    return arnoldi


def include_conjugate_pair(real_T, lams, ord_, i):
    """run.jl:510-517."""
    if not real_T:
        return i
    if i >= len(ord_):
        return i
    l1 = lams[ord_[i - 1] - 1]
    l2 = lams[ord_[i] - 1]
    return i + 1 if (l1.imag != 0 and l1.conjugate() == l2) else i


def _check_args(n_rows, n_cols, nev, mindim, maxdim):
    if n_rows != n_cols:
        raise IndexError(f"matrix is not square: dimensions are ({n_rows}, {n_cols})")  # DimensionMismatch
    if nev < 1:
        raise ValueError("nev cannot be less than 1")
    if not (nev <= mindim <= maxdim <= n_rows):
        raise ValueError(
            f"nev ≤ mindim ≤ maxdim ≤ size(A, 1) does not hold, got {nev} ≤ {mindim} ≤ {maxdim} ≤ {n_rows}"
        )


def partialschur(
    A,
    v1=None,
    nev=None,
    which="LM",
    tol=None,
    mindim=None,
    maxdim=None,
    restarts=200,
    rng=None,
):
    """``partialschur`` (run.jl:100-129).  Returns ``(PartialSchur, History)``."""
    n_rows, n_cols = A.shape
    T = vtype(A)
    if nev is None:
        nev = min(6, n_rows)
    if tol is None:
        tol = math.sqrt(ds.EPS)
    if mindim is None:
        mindim = min(max(10, nev), n_rows)
    if maxdim is None:
        maxdim = min(max(20, 2 * nev), n_rows)
    _check_args(n_rows, n_cols, nev, mindim, maxdim)
    ordering = ds.Ordering(which)
    rng = rng if rng is not None else np.random.default_rng()
    arnoldi = ArnoldiWorkspace(T, n_rows, maxdim)
    if v1 is None:
        reinitialize(arnoldi, 0, rng=rng)
    else:
        v1 = np.asarray(v1)
        if v1.shape[0] != n_rows:
            raise ValueError("v1 should have the same dimension as A")

        def _copy(v):
            v[:] = v1

        reinitialize(arnoldi, 0, populate=_copy)
    return _partialschur(A, arnoldi, mindim, maxdim, nev, tol, restarts, ordering, 1, rng)


def partialschur_inplace(
    A,
    arnoldi,
    start_from=1,
    initialize=None,
    nev=None,
    which="LM",
    tol=None,
    mindim=None,
    maxdim=None,
    restarts=200,
    rng=None,
):
    """``partialschur!`` (run.jl:152-179): resume / user-supplied workspace."""
    n_rows, n_cols = A.shape
    ncolsV = arnoldi.V.shape[1]
    if initialize is None:
        initialize = start_from == 1
    if nev is None:
        nev = min(6, n_rows)
    if tol is None:
        tol = math.sqrt(ds.EPS)
    if mindim is None:
        mindim = min(max(10, nev), n_rows, ncolsV - 1)
    if maxdim is None:
        maxdim = min(max(20, 2 * nev), n_rows, ncolsV - 1)
    _check_args(n_rows, n_cols, nev, mindim, maxdim)
    if not maxdim < ncolsV:
        raise ValueError("maxdim should be strictly less than size(arnoldi.V, 2)")
    if not (1 <= start_from <= maxdim):
        raise ValueError("start_from should be between 1 and maxdim")
    ordering = ds.Ordering(which)
    rng = rng if rng is not None else np.random.default_rng()
    arnoldi.H[:, start_from - 1 :] = 0
    if initialize:
        reinitialize(arnoldi, start_from - 1, rng=rng)
    return _partialschur(A, arnoldi, mindim, maxdim, nev, tol, restarts, ordering, start_from, rng)


def restart_decision(H, Q, maxdim, mindim, nev, tol, ordering, active, real_T, x=None, G=None):
    """One restart's host work: run.jl:278-360.

    Mutates H and Q; returns ``(k, purge, nlock, effective_nev, lams, rs)``.
    Split out so the tests can compare the C++ host driver step by step.
    """
    if x is None:
        x = np.zeros(maxdim, dtype=np.complex128)
    if G is None:
        G = ds.Reflector(maxdim, H.dtype)
    lams = np.zeros(maxdim, dtype=np.complex128)
    rs = np.zeros(maxdim, dtype=np.float64)
    groups = [0] * maxdim

    Q[:] = 0
    Q[np.arange(maxdim), np.arange(maxdim)] = 1  # run.jl:278
    ds.local_schurfact(H[:maxdim, :], active, maxdim, Q)  # run.jl:281

    ord_ = list(range(1, maxdim + 1))  # run.jl:284
    ds.copy_eigenvalues(lams, H)  # run.jl:285
    ds.copy_residuals(rs, H, Q, H[maxdim, maxdim - 1], x, active, maxdim)  # run.jl:286
    ds.sort_perm(ord_, lams, ordering)  # run.jl:289
    H_frob_norm = float(np.linalg.norm(H))  # run.jl:292

    def isconverged(i):  # run.jl:206-208
        return rs[i - 1] <= max(ds.EPS * H_frob_norm, tol * abs(lams[i - 1]))

    effective_nev = include_conjugate_pair(real_T, lams, ord_, nev)  # run.jl:298

    nlock = 0
    for i in range(1, effective_nev + 1):  # run.jl:301-308
        if isconverged(ord_[i - 1]):
            groups[ord_[i - 1] - 1] = 1
            nlock += 1
        else:
            groups[ord_[i - 1] - 1] = 2

    ideal_size = min(nlock + mindim, (mindim + maxdim) // 2)  # run.jl:316
    k = effective_nev
    i = effective_nev + 1
    while i <= maxdim:  # run.jl:320-339
        is_pair = include_conjugate_pair(real_T, lams, ord_, i) == i + 1
        num = 2 if is_pair else 1
        if k < ideal_size and not isconverged(ord_[i - 1]):
            group = 2
            k += num
        else:
            group = 3
        if is_pair:
            groups[ord_[i - 1] - 1] = group
            groups[ord_[i] - 1] = group
            i += 2
        else:
            groups[ord_[i - 1] - 1] = group
            i += 1

    purge = 1  # run.jl:350-353
    while purge < active and groups[purge - 1] == 1:
        purge += 1

    ds.partition_schur_three_way(H, Q, groups)  # run.jl:355
    ds.restore_arnoldi(H, nlock + 1, k, Q, G)  # run.jl:360
    return k, purge, nlock, effective_nev, lams, rs


def _partialschur(A, arnoldi, mindim, maxdim, nev, tol, restarts, ordering, active=1, rng=None):
    """``_partialschur`` (run.jl:224-392)."""
    H, V, V_tmp, Q = arnoldi.H, arnoldi.V, arnoldi.V_tmp, arnoldi.Q
    real_T = not np.iscomplexobj(H)
    stats = _Stats()

    x = np.zeros(maxdim, dtype=np.complex128)
    G = ds.Reflector(maxdim, H.dtype)

    k = mindim
    prods = len(range(active, mindim + 1))  # run.jl:264
    iterate_arnoldi(A, arnoldi, active, mindim, rng, stats)  # run.jl:267

    for _ in range(restarts):
        iterate_arnoldi(A, arnoldi, k + 1, maxdim, rng, stats)  # run.jl:272
        prods += len(range(k + 1, maxdim + 1))  # run.jl:275

        t0 = time.perf_counter()
        k, purge, nlock, _, _, _ = restart_decision(
            H, Q, maxdim, mindim, nev, tol, ordering, active, real_T, x, G
        )
        t1 = time.perf_counter()

        # run.jl:363-365 - the n-sized change of basis
        V_tmp[:, purge - 1 : k] = V[:, purge - 1 : maxdim] @ Q[purge - 1 : maxdim, purge - 1 : k]
        V[:, purge - 1 : k] = V_tmp[:, purge - 1 : k]
        V[:, k] = V[:, maxdim]
        t2 = time.perf_counter()
        stats.t["small"] += t1 - t0
        stats.t["rotate"] += t2 - t1
        stats.restarts += 1

        active = nlock + 1  # run.jl:368
        if active > nev:
            break

    nconverged = active - 1

    t0 = time.perf_counter()
    Q[:] = 0
    Q[np.arange(Q.shape[0]), np.arange(Q.shape[0])] = 1
    ds.sortschur(H, Q, nconverged, ordering)  # run.jl:379
    t1 = time.perf_counter()
    V_tmp[:, :nconverged] = V[:, :nconverged] @ Q[:nconverged, :nconverged]  # run.jl:382
    V[:, :nconverged] = V_tmp[:, :nconverged]  # run.jl:383
    stats.t["small"] += t1 - t0
    stats.t["rotate"] += time.perf_counter() - t1

    lams = np.zeros(maxdim, dtype=np.complex128)
    ds.copy_eigenvalues(lams, H, 1, nconverged)  # run.jl:386

    history = History(
        prods, nconverged, nconverged >= nev, nev, dict(stats.t), stats.second_passes, stats.restarts
    )
    schur = PartialSchur(V[:, :nconverged], H[:nconverged, :nconverged], lams[:nconverged].copy())
    return schur, history


def partialeigen(P):
    """``partialeigen`` (eigvals.jl:92-95): LAPACK eigen of R, then Q * vecs."""
    vals, vecs = np.linalg.eig(P.R)
    return vals, P.Q @ vecs
