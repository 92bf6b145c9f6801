"""CPU oracle: a NumPy/SciPy restatement of ArnoldiMethod.jl's Krylov-Schur path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``arnoldimethod.jl_b200/`` (the
product) imports this package.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it,
and only as the checker / the reported CPU baseline.

Parity status: the reference is pure Julia and Julia is not available in the
build image, so the reference itself cannot be executed here.  The oracle is
pinned against every known-answer value the reference publishes for this path
(README eigenvalues ``readme.md:40-49``, the matvec count ``readme.md:52``,
the exact ``mvproducts == 7 / 3 / 5`` counts of ``test/partial_schur.jl``,
hard-coded matrices of ``test/schurfact.jl`` / ``test/sort_schur.jl``), see
``tests/test_oracle_*.py``.  Bitwise parity with Julia is UNPINNED (no golden
H/V dumps exist upstream; RNG, BLAS summation order and ``givensAlgorithm``
rounding live outside the reference repo) - parity is therefore defined on the
reference's own invariants and tolerances.

Index convention: index *variables* keep the reference's 1-based values so
that every line can be checked against the cited ``file:line``; array accesses
subtract one explicitly.
"""

from .givens import givens_algorithm  # noqa: F401
from .krylov_schur import (  # noqa: F401
    ArnoldiWorkspace,
    History,
    PartialSchur,
    iterate_arnoldi,
    orthogonalize,
    partialeigen,
    partialschur,
    partialschur_inplace,
    reinitialize,
)
