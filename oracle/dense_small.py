"""Oracle: the m x m dense algebra of the Krylov-Schur restart (host side).

Restates, in NumPy, the reference's
  * ``src/schurfact.jl``      - rotations, implicit QR sweeps, ``local_schurfact!``
  * ``src/schursort.jl``      - Sylvester solves, ``swap*!``, ``rotate_right!``
  * ``src/restore_hessenberg.jl`` - ``reflector!``, ``restore_arnoldi!``
  * ``src/eigvals.jl:6-65``   - ``copy_eigenvalues!``, ``eigenvalue``
  * ``src/eigenvector_uppertriangular.jl`` - ``collect_eigen!``
  * ``src/targets.jl``        - orderings with the stable index tie-break
  * ``src/run.jl:394-545``    - partition / sort / residual helpers

Index variables are 1-based exactly as in the cited lines; every array access
subtracts 1.  Matrices are NumPy arrays (float64 or complex128) mutated in
place.  Test infrastructure only - see ``oracle/__init__.py``.
"""

import cmath
import functools
import math

import numpy as np

from .givens import givens_algorithm

EPS = float(np.finfo(np.float64).eps)


def _is_real(M):
    return not np.iscomplexobj(M)


def _conj(x):
    return x.conjugate() if isinstance(x, (complex, np.complexfloating)) else x


def is_offdiagonal_small(H, i, tol=EPS):
    """schurfact.jl:7-11."""
    return abs(H[i, i - 1]) <= tol * (abs(H[i - 1, i - 1]) + abs(H[i, i]))


# --------------------------------------------------------------------------
# Rotations (schurfact.jl:19-148).  lmul applies [c s; -conj(s) c] to rows
# i, i+1; rmul multiplies columns i, i+1 by the conjugate transpose.
# --------------------------------------------------------------------------
class Rotation2:
    __slots__ = ("c", "s", "i")

    def __init__(self, c, s, i):
        self.c, self.s, self.i = c, s, i


class Rotation3:
    __slots__ = ("c1", "s1", "c2", "s2", "i")

    def __init__(self, c1, s1, c2, s2, i):
        self.c1, self.s1, self.c2, self.s2, self.i = c1, s1, c2, s2, i


def get_rotation2(p1, p2, i):
    """schurfact.jl:57-60."""
    c, s, nrm = givens_algorithm(p1, p2)
    return Rotation2(c, s, i), nrm


def get_rotation3(p1, p2, p3, i):
    """schurfact.jl:65-69: rotate (p2,p3) first, then (p1, nrm1)."""
    c1, s1, nrm1 = givens_algorithm(p2, p3)
    c2, s2, nrm2 = givens_algorithm(p1, nrm1)
    return Rotation3(c1, s1, c2, s2, i), nrm2


def lmul(G, A, frm=None, to=None):
    """``lmul!(G, A, from, to)``: act on rows, columns from:to (schurfact.jl:82-134)."""
    if A is None:
        return
    if frm is None:
        frm, to = 1, A.shape[1]
    if to < frm:
        return
    cols = slice(frm - 1, to)
    i = G.i
    if isinstance(G, Rotation2):
        a1 = A[i - 1, cols].copy()
        a2 = A[i, cols].copy()
        A[i - 1, cols] = G.c * a1 + G.s * a2
        A[i, cols] = -_conj(G.s) * a1 + G.c * a2
    else:
        a1 = A[i - 1, cols].copy()
        a2 = A[i, cols].copy()
        a3 = A[i + 1, cols].copy()
        a2p = G.c1 * a2 + G.s1 * a3
        a3p = -_conj(G.s1) * a2 + G.c1 * a3
        A[i - 1, cols] = G.c2 * a1 + G.s2 * a2p
        A[i, cols] = -_conj(G.s2) * a1 + G.c2 * a2p
        A[i + 1, cols] = a3p


def rmul(A, G, frm=None, to=None):
    """``rmul!(A, G, from, to)``: act on columns, rows from:to (schurfact.jl:102-148)."""
    if A is None:
        return
    if frm is None:
        frm, to = 1, A.shape[0]
    if to < frm:
        return
    rows = slice(frm - 1, to)
    i = G.i
    if isinstance(G, Rotation2):
        a1 = A[rows, i - 1].copy()
        a2 = A[rows, i].copy()
        A[rows, i - 1] = a1 * G.c + a2 * _conj(G.s)
        A[rows, i] = a1 * -G.s + a2 * G.c
    else:
        a1 = A[rows, i - 1].copy()
        a2 = A[rows, i].copy()
        a3 = A[rows, i + 1].copy()
        a2p = a2 * G.c1 + a3 * _conj(G.s1)
        a3p = a2 * -G.s1 + a3 * G.c1
        A[rows, i - 1] = a1 * G.c2 + a2p * _conj(G.s2)
        A[rows, i] = a1 * -G.s2 + a2p * G.c2
        A[rows, i + 1] = a3p


# --------------------------------------------------------------------------
# Implicit QR sweeps (schurfact.jl:150-320)
# --------------------------------------------------------------------------
def double_shift_schur(H, frm, to, trace, determinant, Q=None):
    """Francis double-shift bulge chase, real H (schurfact.jl:150-249)."""
    m, n = H.shape
    H11 = H[frm - 1, frm - 1]
    H21 = H[frm, frm - 1]
    H12 = H[frm - 1, frm]
    H22 = H[frm, frm]
    H32 = H[frm + 1, frm]

    p1 = H11 * H11 + H12 * H21 - trace * H11 + determinant
    p2 = H21 * (H11 + H22 - trace)
    p3 = H32 * H21

    G1, _ = get_rotation3(p1, p2, p3, frm)
    lmul(G1, H, frm, n)
    rmul(H, G1, 1, min(frm + 3, m))
    rmul(Q, G1)

    for i in range(frm + 1, to - 1):  # i = from+1 : to-2
        p1 = H[i - 1, i - 2]
        p2 = H[i, i - 2]
        p3 = H[i + 1, i - 2]
        G, nrm = get_rotation3(p1, p2, p3, i)
        H[i - 1, i - 2] = nrm
        H[i, i - 2] = 0
        H[i + 1, i - 2] = 0
        lmul(G, H, i, n)
        rmul(H, G, 1, min(i + 3, m))
        rmul(Q, G)

    Gn, nrm = get_rotation2(H[to - 2, to - 3], H[to - 1, to - 3], to - 1)
    H[to - 2, to - 3] = nrm
    H[to - 1, to - 3] = 0
    lmul(Gn, H, to - 1, n)
    rmul(H, Gn, 1, to)
    rmul(Q, Gn)
    return H


def single_shift_schur(H, frm, to, mu, Q=None):
    """Single-shift bulge chase (schurfact.jl:251-320)."""
    m, n = H.shape
    H11 = H[frm - 1, frm - 1]
    H21 = H[frm, frm - 1]
    p1 = H11 - mu
    p2 = H21
    G1, _ = get_rotation2(p1, p2, frm)
    lmul(G1, H, frm, n)
    rmul(H, G1, 1, min(frm + 2, m))
    rmul(Q, G1)
    for i in range(frm + 1, to):  # i = from+1 : to-1
        p1 = H[i - 1, i - 2]
        p2 = H[i, i - 2]
        G, nrm = get_rotation2(p1, p2, i)
        H[i - 1, i - 2] = nrm
        H[i, i - 2] = 0
        lmul(G, H, i, n)
        rmul(H, G, 1, min(i + 2, m))
        rmul(Q, G)
    return H


def _sign(x):
    return int(x > 0) - int(x < 0)


def upper_triangular_2x2(H11, H12, H21, H22):
    """schurfact.jl:327-357 -> (is_real, c, s)."""
    if H21 == 0 or (H11 - H22 == 0 and _sign(H12) != _sign(H21)):
        return False, 1.0, 0.0
    if H12 == 0:
        return True, 0.0, 1.0
    p = (H11 - H22) / 2
    bcmax = max(abs(H12), abs(H21))
    bcmis = min(abs(H12), abs(H21)) * _sign(H12) * _sign(H21)
    scale = max(abs(p), bcmax)
    z = (p / scale) * p + (bcmax / scale) * bcmis
    if z < 0:
        return False, 1.0, 0.0
    H11_min_lam = p + math.copysign(math.sqrt(scale) * math.sqrt(z), p)
    nrm = math.hypot(H21, H11_min_lam)
    return True, H11_min_lam / nrm, H21 / nrm


def use_single_shift(H11, H12, H21, H22):
    """schurfact.jl:363-388 -> (is_single, shift)."""
    scale = abs(H11) + abs(H12) + abs(H21) + abs(H22)
    H11 /= scale
    H12 /= scale
    H21 /= scale
    H22 /= scale
    t = (H11 + H22) / 2
    d = (H11 - t) * (H22 - t) - H12 * H21
    if d > 0:
        return False, 0.0
    sqrt_discr = math.sqrt(abs(d))
    l1 = t + sqrt_discr
    l2 = t - sqrt_discr
    lam = l1 if abs(H22 - l1) < abs(H22 - l2) else l2
    return True, lam * scale


class QRDidNotConverge(RuntimeError):
    """schurfact.jl:406 ``throw("QR algorithm did not converge")``."""


def local_schurfact(H, start, to, Q=None, tol=EPS, maxiter=None):
    """``local_schurfact!`` (real: schurfact.jl:393-487, generic: :492-538)."""
    if maxiter is None:
        maxiter = 100 * H.shape[0]
    if _is_real(H):
        it = 0
        while to > start:
            it += 1
            if it > maxiter:
                raise QRDidNotConverge("QR algorithm did not converge")
            frm = to
            while frm > start:
                if is_offdiagonal_small(H, frm - 1, tol):
                    H[frm - 1, frm - 2] = 0.0
                    break
                frm -= 1
            if frm == to:
                to -= 1
                continue
            C11, C12 = H[to - 2, to - 2], H[to - 2, to - 1]
            C21, C22 = H[to - 1, to - 2], H[to - 1, to - 1]
            if frm + 1 == to:
                is_real, cs, sn = upper_triangular_2x2(C11, C12, C21, C22)
                if is_real:
                    G = Rotation2(cs, sn, frm)
                    lmul(G, H, frm, H.shape[1])
                    rmul(H, G, 1, to)
                    rmul(Q, G)
                    H[to - 1, to - 2] = 0.0
                to -= 2
                continue
            is_single, mu = use_single_shift(C11, C12, C21, C22)
            if is_single:
                single_shift_schur(H, frm, to, mu, Q)
            else:
                trace = C11 + C22
                determinant = C11 * C22 - C12 * C21
                double_shift_schur(H, frm, to, trace, determinant, Q)
        return True

    it = 0
    while True:
        it += 1
        if it > maxiter:
            return False
        frm = to
        while frm > start and not is_offdiagonal_small(H, frm - 1, tol):
            frm -= 1
        if frm == to:
            if frm >= 2:  # guard for the latent from == 1 edge
                H[frm - 1, frm - 2] = 0
            to -= 1
        else:
            H11, H12 = H[to - 2, to - 2], H[to - 2, to - 1]
            H21, H22 = H[to - 1, to - 2], H[to - 1, to - 1]
            d = H11 * H22 - H21 * H12
            t = H11 + H22
            sqr = cmath.sqrt(t * t - 4 * d)
            l1 = (t + sqr) / 2
            l2 = (t - sqr) / 2
            lam = l1 if abs(H22 - l1) < abs(H22 - l2) else l2
            single_shift_schur(H, frm, to, lam, Q)
        if to <= start:
            break
    return True


# --------------------------------------------------------------------------
# Eigenvalues of a quasi-triangular matrix (eigvals.jl:6-65)
# --------------------------------------------------------------------------
def copy_eigenvalues(lams, A, first=1, last=None, tol=EPS):
    """``copy_eigenvalues!(λs, A, range, tol)``."""
    if last is None:
        last = A.shape[1]
    i = first
    while i < last:
        if is_offdiagonal_small(A, i, tol):
            lams[i - 1] = A[i - 1, i - 1]
            i += 1
        else:
            d = A[i - 1, i - 1] * A[i, i] - A[i - 1, i] * A[i, i - 1]
            x = (A[i - 1, i - 1] + A[i, i]) / 2
            y = cmath.sqrt(complex(x * x - d))
            lams[i - 1] = x + y
            lams[i] = x - y
            i += 2
    if i == last:
        lams[i - 1] = A[i - 1, i - 1]
    return lams


def eigenvalues(A, tol=EPS):
    return copy_eigenvalues(np.empty(A.shape[1], dtype=np.complex128), A, 1, A.shape[1], tol)


def eigenvalue(R, i):
    """eigvals.jl:41-54: eigenvalue of the block starting at ``i``."""
    n = min(R.shape)
    if i == n or R[i, i - 1] == 0:
        return complex(R[i - 1, i - 1])
    d = R[i - 1, i - 1] * R[i, i] - R[i - 1, i] * R[i, i - 1]
    x = (R[i - 1, i - 1] + R[i, i]) / 2
    y = cmath.sqrt(complex(x * x - d))
    return x + y


def is_start_of_11_block(R, i):
    """schursort.jl:505."""
    return i == R.shape[1] or R[i, i - 1] == 0


def is_end_of_11_block(R, i):
    """schursort.jl:506."""
    return i == 1 or R[i - 1, i - 2] == 0


# --------------------------------------------------------------------------
# Tiny Sylvester equations with completely pivoted LU (schursort.jl:61-202)
# --------------------------------------------------------------------------
def lu_complete_pivoting(A):
    """schursort.jl:79-140 -> (LU, p, q, singular); p, q are 1-based."""
    A = np.array(A)
    N = A.shape[0]
    p = [N] * N
    q = [N] * N
    singular = False
    for k in range(1, N):
        m, n, maxval = 1, 1, 0.0
        for j in range(k, N + 1):
            for i in range(k, N + 1):
                if abs(A[i - 1, j - 1]) > maxval:
                    m, n, maxval = i, j, abs(A[i - 1, j - 1])
        p[k - 1] = m
        q[k - 1] = n
        for j in range(k, N + 1):
            A[k - 1, j - 1], A[m - 1, j - 1] = A[m - 1, j - 1], A[k - 1, j - 1]
        for j in range(k, N + 1):
            A[j - 1, k - 1], A[j - 1, n - 1] = A[j - 1, n - 1], A[j - 1, k - 1]
        Akk = A[k - 1, k - 1]
        if Akk == 0:
            singular = True
            break
        for i in range(k + 1, N + 1):
            A[i - 1, k - 1] /= Akk
        for j in range(k + 1, N + 1):
            Akj = A[k - 1, j - 1]
            for i in range(k + 1, N + 1):
                A[i - 1, j - 1] -= A[i - 1, k - 1] * Akj
    if A[N - 1, N - 1] == 0:
        singular = True
    return A, p, q, singular


def lu_solve(LU, p, q, b):
    """schursort.jl:142-168."""
    N = LU.shape[0]
    x = np.array(b, dtype=LU.dtype)
    for i in range(1, N + 1):
        x[i - 1], x[p[i - 1] - 1] = x[p[i - 1] - 1], x[i - 1]
        for j in range(i + 1, N + 1):
            x[j - 1] -= LU[j - 1, i - 1] * x[i - 1]
    for i in range(N, 0, -1):
        for j in range(N, i, -1):
            x[i - 1] -= LU[i - 1, j - 1] * x[j - 1]
        x[i - 1] /= LU[i - 1, i - 1]
        x[i - 1], x[q[i - 1] - 1] = x[q[i - 1] - 1], x[i - 1]
    return x


def sylvsystem(A, B):
    """schursort.jl:170-185."""
    na, nb = A.shape[0], B.shape[0]
    dt = np.result_type(A, B)
    if na == 1 and nb == 2:
        return np.array(
            [[A[0, 0] - B[0, 0], -B[1, 0]], [-B[0, 1], A[0, 0] - B[1, 1]]], dtype=dt
        )
    if na == 2 and nb == 1:
        return np.array(
            [[A[0, 0] - B[0, 0], A[0, 1]], [A[1, 0], A[1, 1] - B[0, 0]]], dtype=dt
        )
    return np.array(
        [
            [A[0, 0] - B[0, 0], A[0, 1], -B[1, 0], 0],
            [A[1, 0], A[1, 1] - B[0, 0], 0, -B[1, 0]],
            [-B[0, 1], 0, A[0, 0] - B[1, 1], A[0, 1]],
            [0, -B[0, 1], A[1, 0], A[1, 1] - B[1, 1]],
        ],
        dtype=dt,
    )


def sylv(A, B, C):
    """Solve A X - X B = C for 1x1 / 2x2 blocks (schursort.jl:198-202)."""
    A = np.atleast_2d(A)
    B = np.atleast_2d(B)
    C = np.atleast_2d(C)
    N, M = A.shape[0], B.shape[0]
    with np.errstate(all="ignore"):
        LU, p, q, singular = lu_complete_pivoting(sylvsystem(A, B))
        rhs = C.reshape(N * M, order="F")
        x = lu_solve(LU, p, q, rhs)
    return x.reshape((N, M), order="F"), singular


# --------------------------------------------------------------------------
# Swaps of adjacent diagonal blocks (schursort.jl:222-503)
# --------------------------------------------------------------------------
def _one(R):
    return 1.0 if _is_real(R) else complex(1.0)


def swap22(R, i, Q=None):
    m, n = R.shape
    A = R[i - 1 : i + 1, i - 1 : i + 1].copy()
    B = R[i + 1 : i + 3, i + 1 : i + 3].copy()
    C = R[i - 1 : i + 1, i + 1 : i + 3].copy()
    X, singular = sylv(A, B, C)
    if singular:
        return R
    one = _one(R)
    c1, s1, nrm1 = givens_algorithm(-X[1, 0], one)
    c2, s2, _ = givens_algorithm(-X[0, 0], nrm1)
    X22 = c1 * -X[1, 1]
    X32 = -_conj(s1) * -X[1, 1]
    X22 = -_conj(s2) * -X[0, 1] + c2 * X22
    c3, s3, nrm3 = givens_algorithm(X32, one)
    c4, s4, _ = givens_algorithm(X22, nrm3)
    G1 = Rotation3(c1, s1, c2, s2, i)
    G2 = Rotation3(c3, s3, c4, s4, i + 1)
    lmul(G1, R, i, n)
    rmul(R, G1, 1, i + 3)
    lmul(G2, R, i, n)
    rmul(R, G2, 1, i + 3)
    R[i + 1, i - 1] = 0
    R[i + 2, i - 1] = 0
    R[i + 1, i] = 0
    R[i + 2, i] = 0
    rmul(Q, G1)
    rmul(Q, G2)
    return R


def swap21(R, i, Q=None):
    m, n = R.shape
    A = R[i - 1 : i + 1, i - 1 : i + 1].copy()
    B = R[i + 1 : i + 2, i + 1 : i + 2].copy()
    C = R[i - 1 : i + 1, i + 1 : i + 2].copy()
    X, singular = sylv(A, B, C)
    if singular:
        return R
    one = _one(R)
    c1, s1, nrm1 = givens_algorithm(-X[1, 0], one)
    c2, s2, _ = givens_algorithm(-X[0, 0], nrm1)
    G1 = Rotation3(c1, s1, c2, s2, i)
    lmul(G1, R, i, n)
    rmul(R, G1, 1, i + 2)
    R[i, i - 1] = 0
    R[i + 1, i - 1] = 0
    rmul(Q, G1)
    return R


def swap12(R, i, Q=None):
    m, n = R.shape
    A = R[i - 1 : i, i - 1 : i].copy()
    B = R[i : i + 2, i : i + 2].copy()
    C = R[i - 1 : i, i : i + 2].copy()
    X, singular = sylv(A, B, C)
    if singular:
        return R
    one = _one(R)
    c1, s1, _ = givens_algorithm(-X[0, 0], one)
    X22 = -_conj(s1) * -X[0, 1]
    c2, s2, _ = givens_algorithm(X22, one)
    G1 = Rotation2(c1, s1, i)
    G2 = Rotation2(c2, s2, i + 1)
    lmul(G1, R, i, n)
    rmul(R, G1, 1, i + 2)
    lmul(G2, R, i, n)
    rmul(R, G2, 1, i + 2)
    R[i + 1, i - 1] = 0
    R[i + 1, i] = 0
    rmul(Q, G1)
    rmul(Q, G2)
    return R


def swap11(R, i, Q=None):
    m, n = R.shape
    R11 = R[i - 1, i - 1]
    R12 = R[i - 1, i]
    R22 = R[i, i]
    G, _ = get_rotation2(R12, R22 - R11, i)
    lmul(G, R, i + 2, n)
    rmul(R, G, 1, i - 1)
    R[i - 1, i - 1] = R22
    R[i, i] = R11
    rmul(Q, G)
    return R


def swap(R, i, curr_11, next_11, Q=None):
    """schursort.jl:489-503."""
    if curr_11:
        if next_11:
            swap11(R, i, Q)
        else:
            swap12(R, i, Q)
    else:
        if next_11:
            swap21(R, i, Q)
        else:
            swap22(R, i, Q)


def rotate_right(R, frm, to, Q=None):
    """schursort.jl:19-32."""
    i = to
    while i > frm:
        curr_11 = is_start_of_11_block(R, i)
        prev_11 = is_end_of_11_block(R, i - 1)
        j = i - 1 if prev_11 else i - 2
        swap(R, j, prev_11, curr_11, Q)
        i = j


def partition_schur_three_way(R, Q, groups):
    """run.jl:394-457.  ``groups`` is indexed by original Schur position."""
    hi = mi = lo = 1
    while hi <= len(groups):
        group = groups[hi - 1]
        blocksize = 1 if is_start_of_11_block(R, hi) else 2
        if group == 3:
            hi += blocksize
        elif group == 2:
            rotate_right(R, mi, hi, Q)
            hi += blocksize
            mi += blocksize
        else:
            rotate_right(R, lo, hi, Q)
            hi += blocksize
            mi += blocksize
            lo += blocksize


# --------------------------------------------------------------------------
# Orderings (targets.jl:34-75)
# --------------------------------------------------------------------------
def _isless(a, b):
    """Julia ``isless`` on floats: NaN is largest, -0.0 < 0.0."""
    if math.isnan(a):
        return False
    if math.isnan(b):
        return True
    if a == 0 and b == 0:
        return math.copysign(1.0, a) < math.copysign(1.0, b)
    return a < b


_KEYS = {
    "LM": (abs, True),
    "LR": (lambda z: z.real, True),
    "SR": (lambda z: z.real, False),
    "LI": (lambda z: z.imag, True),
    "SI": (lambda z: z.imag, False),
}


class Ordering:
    """``get_order(which)`` (targets.jl:71-75) as an ``lt`` functor."""

    def __init__(self, which):
        which = str(which).lstrip(":").upper()
        if which not in _KEYS:
            raise ValueError(f"Unknown target: {which}")  # run.jl:185 ArgumentError
        self.which = which
        self.f, self.reverse = _KEYS[which]

    def lt(self, a, b):
        fa, fb = self.f(complex(a)), self.f(complex(b))
        return _isless(fb, fa) if self.reverse else _isless(fa, fb)


def sort_perm(ord_, lams, ordering):
    """``sort!(ord, QuickSort, OrderPerm(λs, ordering))`` (run.jl:289, targets.jl:61-67)."""

    def cmp(i, j):
        fst, snd = lams[i - 1], lams[j - 1]
        if ordering.lt(fst, snd):
            return -1
        if ordering.lt(snd, fst):
            return 1
        return -1 if i < j else (1 if i > j else 0)

    ord_.sort(key=functools.cmp_to_key(cmp))
    return ord_


def sortschur(R, Q, to, ordering):
    """run.jl:465-502: insertion sort of the leading ``to`` eigenvalues."""
    if to <= 1:
        return
    next_idx = 1
    while next_idx <= to:
        curr_idx = next_idx
        curr_size = 1 if is_start_of_11_block(R, curr_idx) else 2
        curr_lam = eigenvalue(R, curr_idx)
        while curr_idx > 1:
            prev_size = 1 if is_end_of_11_block(R, curr_idx - 1) else 2
            prev_idx = curr_idx - prev_size
            prev_lam = eigenvalue(R, prev_idx)
            if not ordering.lt(curr_lam, prev_lam):
                break
            swap(R, prev_idx, prev_size == 1, curr_size == 1, Q)
            curr_idx -= prev_size
        next_idx += curr_size


# --------------------------------------------------------------------------
# Householder reflector + restore_arnoldi! (restore_hessenberg.jl)
# --------------------------------------------------------------------------
def reflector(y, k):
    """``reflector!(y, k) -> tau'`` (restore_hessenberg.jl:16-45)."""
    xnrm = 0.0
    for idx in range(k - 1):
        xnrm += abs(y[idx]) ** 2
    alpha = y[k - 1]
    if xnrm == 0 and alpha.imag == 0:
        return 0 * alpha
    xnrm = math.sqrt(xnrm)
    beta = -math.copysign(math.hypot(abs(alpha), xnrm), alpha.real)
    tau = (beta - alpha) / beta
    alpha = 1 / (alpha - beta)
    y[: k - 1] *= alpha
    y[k - 1] = beta
    return _conj(tau)


class Reflector:
    """restore_hessenberg.jl:47-59."""

    def __init__(self, max_len, dtype):
        self.vec = np.zeros(max_len, dtype=dtype)
        self.offset = 1
        self.len = 0
        self.tau = 0


def reflector_lmul(G, H, frm, to):
    """restore_hessenberg.jl:138-159."""
    ln, off, z, tau = G.len, G.offset, G.vec, G.tau
    if tau == 0:
        return
    rows = slice(off - 1, off - 1 + ln - 1)
    last = ln + off - 2
    zz = z[: ln - 1]
    for col in range(frm - 1, to):
        dot = np.dot(np.conj(zz), H[rows, col]) + H[last, col]
        dot *= tau
        H[rows, col] -= dot * zz
        H[last, col] -= dot


def reflector_rmul(H, G, frm, to):
    """restore_hessenberg.jl:161-182."""
    ln, off, z, tau = G.len, G.offset, G.vec, G.tau
    if tau == 0:
        return
    cols = slice(off - 1, off - 1 + ln - 1)
    last = off + ln - 2
    zz = z[: ln - 1]
    for row in range(frm - 1, to):
        dot = np.dot(H[row, cols], zz) + H[row, last]
        dot *= _conj(tau)
        H[row, cols] -= dot * np.conj(zz)
        H[row, last] -= dot


def restore_arnoldi(H, frm, to, Q, G):
    """``restore_arnoldi!`` (restore_hessenberg.jl:75-134)."""
    if not frm < to:
        return
    m, n = H.shape
    nrm = Q[n - 1, frm - 1]
    for i in range(frm, to):  # i = from : to-1
        c, s, nrm = givens_algorithm(Q[n - 1, i], nrm)
        g = Rotation2(c, -s, i)
        rmul(H, g, 1, min(i + 2, to))
        lmul(g, H, 1, to)
        rmul(Q, g, 1, n)
    H[to, to - 1] = Q[Q.shape[0] - 1, to - 1] * H[m - 1, n - 1]
    G.offset = frm
    for i in range(to - frm, 1, -1):  # i = to-from : -1 : 2
        G.len = i
        row = frm + i
        for j in range(1, i + 1):
            G.vec[j - 1] = _conj(H[row - 1, j + frm - 2])
        G.tau = reflector(G.vec, i)
        reflector_rmul(H, G, 1, row - 1)
        for j in range(1, i):
            H[row - 1, j + frm - 2] = 0
        H[row - 1, i - 2 + frm] = _conj(G.vec[i - 1])
        reflector_lmul(G, H, frm, to)
        reflector_rmul(Q, G, 1, n)


# --------------------------------------------------------------------------
# Eigenvectors of (quasi) upper triangular R (eigenvector_uppertriangular.jl)
# --------------------------------------------------------------------------
def shifted_backward_sub(x, R, lam, k):
    """eigenvector_uppertriangular.jl:6-68 (real quasi-triangular or generic)."""
    real_R = _is_real(R)
    while k > 0:
        if real_R and k > 1 and R[k - 1, k - 2] != 0:
            R11, R12 = R[k - 2, k - 2] - lam, R[k - 2, k - 1]
            R21, R22 = R[k - 1, k - 2], R[k - 1, k - 1] - lam
            det = R11 * R22 - R21 * R12
            a1 = (R22 * x[k - 2] - R12 * x[k - 1]) / det
            a2 = (-R21 * x[k - 2] + R11 * x[k - 1]) / det
            x[k - 2] = a1
            x[k - 1] = a2
            if k > 2:
                x[: k - 2] -= R[: k - 2, k - 2] * x[k - 2] + R[: k - 2, k - 1] * x[k - 1]
            k -= 2
        else:
            sigma = R[k - 1, k - 1] - lam
            if sigma == 0:
                x[k - 1] = sigma
            else:
                x[k - 1] /= sigma
                if k > 1:
                    x[: k - 1] -= R[: k - 1, k - 1] * x[k - 1]
            k -= 1
    return x


def collect_eigen(x, R, j):
    """``collect_eigen!(x, R, j) -> k`` (eigenvector_uppertriangular.jl:76-154)."""
    n = R.shape[1]
    if _is_real(R):
        if j < n and R[j, j - 1] != 0:
            j += 1
        if j > 1 and R[j - 1, j - 2] != 0:
            R11, R21 = R[j - 2, j - 2], R[j - 1, j - 2]
            R12, R22 = R[j - 2, j - 1], R[j - 1, j - 1]
            det = R11 * R22 - R21 * R12
            tr = R11 + R22
            lam = (tr + cmath.sqrt(complex(tr * tr - 4 * det))) / 2
            x[j - 2] = -R12 / (R11 - lam)
            x[j - 1] = 1
            for i in range(1, j - 1):
                x[i - 1] = -R[i - 1, j - 2] * x[j - 2] - R[i - 1, j - 1]
            shifted_backward_sub(x, R, lam, j - 2)
        else:
            lam = R[j - 1, j - 1]
            x[j - 1] = 1
            x[: j - 1] = -R[: j - 1, j - 1]
            shifted_backward_sub(x, R, lam, j - 1)
    else:
        lam = R[j - 1, j - 1]
        x[j - 1] = 1
        x[: j - 1] = -R[: j - 1, j - 1]
        shifted_backward_sub(x, R, lam, j - 1)
    nrm = 0.0
    for k in range(j):
        nrm += abs(x[k]) ** 2
    x[:j] *= 1.0 / math.sqrt(nrm)
    return j


def copy_residuals(rs, H, Q, h_last, x, first, last):
    """``copy_residuals!`` (run.jl:524-545); range = first:last (1-based)."""
    rs[:] = 0
    m = H.shape[1]
    for i in range(first, last + 1):
        x[:] = 0
        ln = collect_eigen(x, H, i)
        tmp = complex(np.dot(Q[m - 1, :ln], x[:ln]))
        rs[i - 1] = abs(tmp * h_last)
    return rs
