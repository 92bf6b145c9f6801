"""b200arnoldi - B200-native hot path of ArnoldiMethod.jl behind its own API.

Host-side mirror (Python over ctypes) of the reference's public interface for
the Arnoldi expansion + Krylov-Schur restart path:

    partialschur(A; v1, nev, which, tol, mindim, maxdim, restarts)   src/run.jl:100
    partialschur!(A, arnoldi; start_from, initialize, ...)           src/run.jl:152
    partialeigen(P)                                                  src/eigvals.jl:92
    ArnoldiWorkspace(n, k) / (v1, k)                                 src/ArnoldiMethod.jl:41

All n-sized arithmetic runs in hand-written sm_100a CUDA kernels inside
``libb200arnoldi.so`` (C ABI: ``include/b200arnoldi.h``).  There is no CPU
fallback; importing this package needs the built shared library.
"""

from ._lib import B200Error, DimensionMismatch, LIB_PATH, lib  # noqa: F401
from .api import (  # noqa: F401
    ArnoldiWorkspace,
    Context,
    History,
    Operator,
    PartialSchur,
    default_context,
    partialeigen,
    partialschur,
    partialschur_,
    uniform_reference,
)
from . import sharding  # noqa: F401

__all__ = [
    "partialschur",
    "partialschur_",
    "partialeigen",
    "ArnoldiWorkspace",
    "Operator",
    "Context",
    "History",
    "PartialSchur",
]
