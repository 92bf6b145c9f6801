"""Independent second opinion on the oracle.  The reference cannot be executed here (no Julia), so besides
the reference's own known answers (test_oracle_partialschur.py, test_golden.py) the oracle's eigenvalues
are checked against two implementations that share no code with it: ARPACK's implicitly restarted Arnoldi
(`scipy.sparse.linalg.eigs`, the method docs/src/index.md:368-375 compares the reference with) and LAPACK's
dense `eig`.  Same matrix, same target, same number of wanted eigenvalues; tolerance 10 * tol * |lambda| as in
the GPU parity definition (SURVEY 8(c) i)."""

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import oracle

WHICH_ARPACK = {"LM": "LM", "LR": "LR", "SR": "SR", "LI": "LI", "SI": "SI"}


def spectrum_matrix(rng, T, n, which):
    """Sparse non-symmetric matrix with a designed, well separated outer spectrum for the target."""
    B = sp.random(n, n, 4.0 / n, random_state=rng, format="csr") * 0.05
    if T is np.complex128:
        B = B + 1j * sp.random(n, n, 4.0 / n, random_state=rng, format="csr") * 0.05
    d = rng.random(n) * 0.5 + 0.0j
    k = 12
    lead = 2.0 + 1.5 * 0.85 ** np.arange(k)
    if which == "LM":
        d[:k] = lead * np.exp(1j * np.linspace(0.0, 1.2, k)) if T is np.complex128 else lead
    elif which == "LR":
        d[:k] = lead
    elif which == "SR":
        d[:k] = -lead
    elif which == "LI":
        d[:k] = 0.3 + 1j * lead
    else:
        d[:k] = 0.3 - 1j * lead
    if T is np.float64:
        d = d.real
    return (B + sp.diags(d)).tocsr().astype(T)


def closest_match(a, b, tol):
    b = list(b)
    for x in a:
        j = int(np.argmin([abs(x - y) for y in b]))
        assert abs(x - b[j]) <= tol * max(1.0, abs(x)), (x, b[j])
        b.pop(j)


@pytest.mark.parametrize("T,which", [(np.float64, "LM"), (np.float64, "LR"), (np.float64, "SR"),
                                     (np.complex128, "LM"), (np.complex128, "LR"), (np.complex128, "SR"),
                                     (np.complex128, "LI"), (np.complex128, "SI")])
def test_oracle_agrees_with_arpack_and_lapack(T, which):
    rng = np.random.default_rng(sum(map(ord, T.__name__ + which)))
    n, nev, tol = 400, 6, 1e-9
    A = spectrum_matrix(rng, T, n, which)
    v1 = rng.random(n).astype(T)
    P, hist = oracle.partialschur(A, v1=v1, nev=nev, tol=tol, which=which)
    assert hist.converged
    lam = np.asarray(P.eigenvalues)[:nev]

    # LAPACK: the nev extreme eigenvalues of the dense matrix for this target
    ev = np.linalg.eigvals(A.toarray())
    key = {"LM": -np.abs(ev), "LR": -ev.real, "SR": ev.real, "LI": -ev.imag, "SI": ev.imag}[which]
    closest_match(lam, ev[np.argsort(key)][:nev], 10 * tol)

    # ARPACK (implicitly restarted Arnoldi), same start vector
    w = spla.eigs(A, k=nev, which=WHICH_ARPACK[which], v0=v1, tol=1e-12, ncv=40, return_eigenvectors=False)
    closest_match(lam, w, 10 * tol)

    # and the Schur relation itself, by the reference's own bounds (test/partial_schur.jl:23-25,38)
    Q, R = P.Q, P.R
    assert np.linalg.norm(A @ Q - Q @ R) < n * tol
    assert np.linalg.norm(Q.conj().T @ Q - np.eye(Q.shape[1])) < 1000 * np.finfo(float).eps


def test_real_matrix_with_complex_pairs_against_arpack():
    """Real non-symmetric operator whose dominant eigenvalues are conjugate pairs: the pair is never split
    (src/run.jl:298,510-517), so nev + 1 values may come back; ARPACK returns the same pairs."""
    rng = np.random.default_rng(11)
    n, nev, tol = 300, 5, 1e-9
    blocks = []
    for i in range(4):  # 2x2 rotation-scaling blocks: eigenvalues r e^{+-i t}
        r, t = 3.0 - 0.4 * i, 0.5 + 0.3 * i
        blocks.append(r * np.array([[np.cos(t), -np.sin(t)], [np.sin(t), np.cos(t)]]))
    D = sp.block_diag(blocks + [sp.diags(rng.random(n - 8) * 0.5)], format="csr")
    A = (D + sp.random(n, n, 3.0 / n, random_state=rng, format="csr") * 0.02).tocsr()
    v1 = rng.random(n)
    P, hist = oracle.partialschur(A, v1=v1, nev=nev, tol=tol, which="LM")
    assert hist.converged and hist.nconverged in (nev, nev + 1)
    lam = np.asarray(P.eigenvalues)
    assert hist.nconverged == nev + 1  # 5 wanted, the third pair is completed
    w = spla.eigs(A, k=nev + 1, which="LM", v0=v1, tol=1e-12, ncv=40, return_eigenvectors=False)
    closest_match(lam, w, 10 * tol)
    # conjugate pairs are exact conjugates (src/eigvals.jl:20-24)
    for x in lam:
        assert np.any(lam == np.conj(x))
