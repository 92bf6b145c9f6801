"""Small run that touches every kernel once - meant to be run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp

import b200arnoldi as b2a

rng = np.random.default_rng(0)
for T in (np.float64, np.complex128):
    n = 3001
    A = sp.random(n, n, 9 / n, random_state=rng, format="csr").astype(T)
    if T is np.complex128:
        A = A + 1j * sp.random(n, n, 9 / n, random_state=rng, format="csr")
    d = np.zeros(n, dtype=T)
    d[:8] = 5 + 10 * 0.8 ** np.arange(8)
    A = (A + sp.diags(d)).tocsr()
    v1 = rng.random(n).astype(T)
    P, hist = b2a.partialschur(A, nev=4, tol=1e-8, v1=v1)
    assert hist.converged
    vals, X = b2a.partialeigen(P)
    P2, h2 = b2a.partialschur(A.tocsc(), nev=4, tol=1e-8, seed=3)  # CSC upload + rand! fill
    ctx = b2a.default_context()
    op = b2a.Operator.from_csc_arrays(ctx, A.tocsc().indptr, A.tocsc().indices, A.tocsc().data, n, mode=1)
    ws = b2a.ArnoldiWorkspace(v1, 70, ctx=ctx)  # 70 > 64 columns: exercises the LDG Gram-Schmidt kernels too
    ws.reinitialize(0, "keep")
    ws.iterate_arnoldi(op, 1, 70)
    ws.norm(3), ws.gemv_c(5, 6), ws.scal_div(71, 2.0), ws.copy_col(1, 2)
    ws.rotate_basis(3, 41, 70, np.asfortranarray(np.linalg.qr(rng.standard_normal((70, 70)))[0].astype(T)))  # DMMA, N > 32
    # shift-and-invert (Jacobi-CG kernels) on a Hermitian positive definite operator
    B = sp.random(n, n, 3 / n, random_state=rng, format="csr").astype(T)
    S_ = (B.conj().T @ B + sp.identity(n) * 2.0).tocsr().astype(T)
    S_.sort_indices()
    ops = b2a.Operator.from_matrix(ctx, S_)
    inv = b2a.Operator.shift_invert(ops, sigma=0.5)
    ws.matvec(inv, 1, 2)
    print(T.__name__, "ok", hist, inv.solve_stats)
print("SANITIZE_SMOKE_DONE")
