"""Plane rotations: restatement of ``LinearAlgebra.givensAlgorithm``.

The reference calls Julia's stdlib ``givensAlgorithm(f, g)`` (a port of LAPACK
``dlartg`` / ``zlartg``) at ``src/schurfact.jl:58,66-67``,
``src/schursort.jl:224-235,258-265,288-289`` and
``src/restore_hessenberg.jl:91``.  The stdlib is not part of the reference
tree, so this restates the published LAPACK 3.x algorithm:

    [ c        s ] [ f ]   [ r ]
    [ -conj(s) c ] [ g ] = [ 0 ],   c real.

Real case: ``g == 0 -> (1, 0, f)``; ``f == 0 -> (0, 1, g)``; otherwise
``r = sqrt(f^2 + g^2)`` with power-of-two rescaling against over/underflow and
the LAPACK sign rule ``|f| > |g| and c < 0 -> flip (c, s, r)``.
Complex case: ``zlartg`` (c >= 0 real, s complex).

Test infrastructure only - see ``oracle/__init__.py``.
"""

import math

import numpy as np

# floatmin2(Float64) in Julia == 2^-485 ; its inverse bounds the safe range.
_SAFMN2 = math.ldexp(1.0, -485)
_SAFMX2 = math.ldexp(1.0, 485)
_SAFMIN = 2.2250738585072014e-308


def _givens_real(f, g):
    if g == 0:
        return 1.0, 0.0, f
    if f == 0:
        return 0.0, 1.0, g
    f1, g1 = f, g
    scale = max(abs(f1), abs(g1))
    if scale >= _SAFMX2:
        count = 0
        while True:
            count += 1
            f1 *= _SAFMN2
            g1 *= _SAFMN2
            scale = max(abs(f1), abs(g1))
            if scale < _SAFMX2 or count >= 20:
                break
        r = math.sqrt(f1 * f1 + g1 * g1)
        c, s = f1 / r, g1 / r
        for _ in range(count):
            r *= _SAFMX2
    elif scale <= _SAFMN2:
        count = 0
        while True:
            count += 1
            f1 *= _SAFMX2
            g1 *= _SAFMX2
            scale = max(abs(f1), abs(g1))
            if scale > _SAFMN2:
                break
        r = math.sqrt(f1 * f1 + g1 * g1)
        c, s = f1 / r, g1 / r
        for _ in range(count):
            r *= _SAFMN2
    else:
        r = math.sqrt(f1 * f1 + g1 * g1)
        c, s = f1 / r, g1 / r
    if abs(f) > abs(g) and c < 0:
        c, s, r = -c, -s, -r
    return c, s, r


def _abs1(z):
    return max(abs(z.real), abs(z.imag))


def _abs2(z):
    return z.real * z.real + z.imag * z.imag


def _givens_complex(f, g):
    f = complex(f)
    g = complex(g)
    scale = max(_abs1(f), _abs1(g))
    fs, gs = f, g
    count = 0
    if scale >= _SAFMX2:
        while True:
            count += 1
            fs *= _SAFMN2
            gs *= _SAFMN2
            scale *= _SAFMN2
            if scale < _SAFMX2 or count >= 20:
                break
    elif scale <= _SAFMN2:
        if g == 0:
            return 1.0, 0j, f
        while True:
            count -= 1
            fs *= _SAFMX2
            gs *= _SAFMX2
            scale *= _SAFMX2
            if scale > _SAFMN2:
                break
    f2 = _abs2(fs)
    g2 = _abs2(gs)
    if f2 <= max(g2, 1.0) * _SAFMIN:
        # rare: f negligible against g
        if f == 0:
            d = abs(gs)
            return 0.0, complex(gs.real / d, -gs.imag / d), complex(abs(g))
        f2s = abs(fs)
        g2s = math.sqrt(g2)
        c = f2s / g2s
        if _abs1(f) > 1:
            d = abs(f)
            ff = complex(f.real / d, f.imag / d)
        else:
            dr = _SAFMX2 * f.real
            di = _SAFMX2 * f.imag
            d = math.hypot(dr, di)
            ff = complex(dr / d, di / d)
        s = ff * complex(gs.real / g2s, -gs.imag / g2s)
        r = c * f + s * g
        return c, s, r
    # common case
    f2s = math.sqrt(1.0 + g2 / f2)
    r = complex(f2s * fs.real, f2s * fs.imag)
    c = 1.0 / f2s
    d = f2 + g2
    s = complex(r.real / d, r.imag / d) * gs.conjugate()
    if count > 0:
        for _ in range(count):
            r *= _SAFMX2
    elif count < 0:
        for _ in range(-count):
            r *= _SAFMN2
    return c, s, r


def givens_algorithm(f, g):
    """Return ``(c, s, r)`` like Julia's ``LinearAlgebra.givensAlgorithm``."""
    cplx = (complex, np.complexfloating)
    if isinstance(f, cplx) or isinstance(g, cplx):
        return _givens_complex(complex(f), complex(g))
    return _givens_real(float(f), float(g))
