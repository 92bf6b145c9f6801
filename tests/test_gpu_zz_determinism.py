"""Bit-reproducibility of the CUDA path ("Determinism"): no FP64 atomics on the CSR path, static tile
ranges per CTA, fixed-order two-stage reductions - so two runs on the same inputs give identical bits, for the
fused orthogonalisation kernel and for the four-kernel chain.  Restart decisions are threshold tests on
residuals (src/run.jl:206-208), so this is what makes `mvproducts` and the locked set repeatable.
(Collected last on purpose: it is a property test of the whole path.)"""

import numpy as np
import pytest
import scipy.sparse as sp

import b200arnoldi as b2a

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return b2a.default_context()


def problem(T, n, k, seed):
    rng = np.random.default_rng(seed)
    cols = rng.integers(0, n, size=(n, k))
    vals = rng.standard_normal((n, k)) * 0.3
    if T is np.complex128:
        vals = vals + 0.3j * rng.standard_normal((n, k))
    A = sp.csr_matrix((vals.ravel(), cols.ravel(), np.arange(0, n * k + 1, k)), shape=(n, n))
    d = np.zeros(n, dtype=T)
    d[:12] = 4.0 + 10.0 * 0.8 ** np.arange(12)
    A = (A + sp.diags(d)).tocsr().astype(T)
    A.sum_duplicates()
    A.sort_indices()
    v1 = rng.random(n).astype(T)
    return A, v1


@pytest.mark.parametrize("T", [np.float64, np.complex128])
@pytest.mark.parametrize("fused", ["1", "0"], ids=["fused", "four_kernels"])
def test_sweep_is_bit_reproducible(ctx, monkeypatch, T, fused):
    monkeypatch.setenv("B2A_FUSED_SWEEP", fused)
    n, mx = 300_000, 24
    A, v1 = problem(T, n, 8, 3)
    op = b2a.Operator.from_matrix(ctx, A)
    runs = []
    for rep in range(2):
        ws = b2a.ArnoldiWorkspace(v1, mx, ctx=ctx)
        ws.reinitialize(0, "keep")
        ws.iterate_arnoldi(op, 1, mx)
        runs.append((np.array(ws.H).copy(), ws.V.copy()))
        ws.close()
    op.close()
    assert np.array_equal(runs[0][0], runs[1][0])
    assert np.array_equal(runs[0][1], runs[1][1])


@pytest.mark.parametrize("T", [np.float64, np.complex128])
def test_partialschur_is_bit_reproducible(ctx, T):
    A, v1 = problem(T, 50_000, 6, 4)
    out = []
    for rep in range(2):
        P, hist = b2a.partialschur(A, nev=6, tol=1e-8, which="LM", v1=v1, ctx=ctx)
        out.append((hist.mvproducts, hist.nconverged, np.array(P.eigenvalues), np.array(P.R), np.array(P.Q)))
        P.workspace.close()
    assert out[0][0] == out[1][0] and out[0][1] == out[1][1]
    for a, b in zip(out[0][2:], out[1][2:]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("T", [np.float64, np.complex128])
def test_fused_sweep_barrier_regions_many_ctas_one_tile(ctx, monkeypatch, T):
    """Widest window of the round-1 write-after-read hazard on the per-CTA partial sums of the fused sweep: many CTAs,
    ONE resident tile each, so phase P2 takes ~1 us and a fast CTA stores its barrier-B partials while a slow one is
    still summing barrier A's.  Every barrier now owns its region (kernels_cgs_sweep.cuh kSweepPartRegion): the H
    column must equal the oracle's AND be bit-identical over many repetitions."""
    import oracle

    monkeypatch.setenv("B2A_FUSED_SWEEP", "1")
    rng = np.random.default_rng(17)
    j = 24
    n = 148 * (256 if T is np.float64 else 128)  # one TMA tile per CTA at this panel width
    Vp = np.linalg.qr(rng.standard_normal((n, j)) + (1j * rng.standard_normal((n, j)) if T is np.complex128 else 0))[0]
    x = (rng.standard_normal(n) + (1j * rng.standard_normal(n) if T is np.complex128 else 0)).astype(T)
    x = x + Vp @ rng.standard_normal(j) * 100  # wnorm / rnorm ~ 0.4 < eta: DGKS fires, all three barriers run
    ws = b2a.ArnoldiWorkspace(n, j + 1, dtype=T, ctx=ctx)
    for c in range(j):
        ws.set_col(c + 1, Vp[:, c].astype(T))
    arn = oracle.ArnoldiWorkspace(T, n, j + 1)
    arn.V[:, :j] = Vp
    arn.V[:, j] = x
    assert oracle.orthogonalize(arn, j) is True
    first = None
    for rep in range(300):
        ws.set_col(j + 1, x)
        assert ws.orthogonalize(j) is True
        h = np.array(ws.H[: j + 1, j - 1]).copy()
        if first is None:
            first = h
            assert np.abs(h - arn.H[: j + 1, j - 1]).max() <= 1e-12 * np.abs(arn.H[: j + 1, j - 1]).max()
        assert np.array_equal(h, first), rep
    ws.close()
