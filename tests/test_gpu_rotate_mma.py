"""The basis rotation on the FP64 tensor pipe (kernels_rotate_mma.cuh: TMA ring + DMMA.8x8x4) against NumPy, for the
shapes the shared-memory kernels could not take (wide bases: the reference handles any maxdim), against the DFMA
kernels on the same input, and a whole wide-basis solve against the oracle.  Reference: src/run.jl:363-365, 382-383."""

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu

import b200arnoldi as b2a  # noqa: E402
import oracle  # noqa: E402

TYPES = [np.float64, np.complex128]


def randn(rng, T, *shape):
    x = rng.standard_normal(shape)
    if np.issubdtype(T, np.complexfloating):
        x = x + 1j * rng.standard_normal(shape)
    return x.astype(T)


@pytest.fixture(scope="module")
def ctx():
    return b2a.default_context()


def rotate_once(ctx, T, V, Q, maxdim, purge, k):
    n = V.shape[0]
    ws = b2a.ArnoldiWorkspace(n, maxdim, dtype=T, ctx=ctx)
    for c in range(maxdim + 1):
        ws.set_col(c + 1, V[:, c])
    ws.rotate_basis(purge, k, maxdim, Q)
    got = ws.V
    ws.close()
    return got


@pytest.mark.parametrize("T", TYPES)
@pytest.mark.parametrize("n,maxdim,purge,k", [
    (9, 6, 1, 4),            # fewer rows than one MMA slab
    (5000, 40, 1, 26),       # cfg-2 shape: one chunk of four n tiles
    (5000, 40, 12, 30),      # odd K, N
    (3000, 70, 1, 67),       # N > 32 outputs: several n chunks
    (2000, 200, 1, 150),     # nev = 100 default basis: does not fit the DFMA kernels' shared-memory tile
    (1500, 300, 41, 222),    # K > 256: two TMA boxes per tile
    (700, 64, 1, 64),        # k == maxdim: no column move
])
def test_rotate_mma_matches_numpy(ctx, T, n, maxdim, purge, k, monkeypatch):
    monkeypatch.setenv("B2A_ROTATE", "1")
    rng = np.random.default_rng(n + maxdim)
    V = np.asfortranarray(randn(rng, T, n, maxdim + 1))
    Q = np.asfortranarray(randn(rng, T, maxdim, maxdim))
    got = rotate_once(ctx, T, V, Q, maxdim, purge, k)
    ref = V.copy()
    ref[:, purge - 1 : k] = V[:, purge - 1 : maxdim] @ Q[purge - 1 : maxdim, purge - 1 : k]
    if k < maxdim:
        ref[:, k] = V[:, maxdim]
    assert np.abs(got - ref).max() <= 1e-13 * maxdim * np.abs(V).max() * np.abs(Q).max()
    keep = [c for c in range(maxdim + 1) if not (purge - 1 <= c <= min(k, maxdim))]
    assert np.array_equal(got[:, keep], V[:, keep])  # untouched columns, bit for bit


@pytest.mark.parametrize("T", TYPES)
def test_rotate_mma_vs_dfma_kernels(ctx, T, monkeypatch):
    """Same input through both kernels: they differ only in summation order (4-wide MMA k groups vs a running FMA)."""
    rng = np.random.default_rng(3)
    n, maxdim, purge, k = 20001, 40, 5, 29
    V = np.asfortranarray(randn(rng, T, n, maxdim + 1))
    Q = np.asfortranarray(np.linalg.qr(randn(rng, T, maxdim, maxdim))[0])
    monkeypatch.setenv("B2A_ROTATE", "1")
    a = rotate_once(ctx, T, V, Q, maxdim, purge, k)
    monkeypatch.setenv("B2A_ROTATE", "0")
    b = rotate_once(ctx, T, V, Q, maxdim, purge, k)
    assert np.abs(a - b).max() <= 64 * np.finfo(np.float64).eps * np.abs(V).max() * np.sqrt(maxdim)


def test_wide_basis_solve_matches_oracle(ctx):
    """nev = 100 -> default maxdim = 200 (src/run.jl:108): every restart rotates a 200-column basis."""
    rng = np.random.default_rng(21)
    n, nev = 3000, 100
    d = np.concatenate([100 + 50 * np.linspace(0, 1, 400) ** 0.5, rng.random(n - 400) * 90])  # 13 restarts in the oracle
    A = (sp.diags(d) + sp.random(n, n, 2 / n, random_state=rng) * 0.01).tocsr()
    v1 = rng.random(n)
    P, hist = b2a.partialschur(A, nev=nev, tol=1e-8, which="LM", v1=v1, ctx=ctx)
    Po, ho = oracle.partialschur(A, v1=v1, nev=nev, tol=1e-8, which="LM")
    assert hist.converged and ho.converged
    assert abs(hist.mvproducts - ho.mvproducts) <= 100  # one restart's worth (maxdim - mindim)
    lam, lamo = np.sort_complex(P.eigenvalues)[-nev:], np.sort_complex(Po.eigenvalues)[-nev:]
    assert np.abs(lam - lamo).max() <= 10 * 1e-8 * np.abs(lamo).max()
    Q = P.Q
    assert np.linalg.norm(A @ Q - Q @ P.R) < n * 1e-8
    assert np.linalg.norm(Q.conj().T @ Q - np.eye(Q.shape[1])) < 1e-11
