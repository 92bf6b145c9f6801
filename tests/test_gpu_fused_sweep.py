"""GPU parity of the two orthogonalisation paths of src/expansion.jl:69-109:

  B2A_FUSED_SWEEP=1 (default on one GPU): kernels_cgs_sweep.cuh, ONE persistent kernel with in-kernel
                     grid barriers;
  B2A_FUSED_SWEEP=0: the four kernels S1 dots / S2 update / gated S3 update / finish (still the path of
                     row-sharded workspaces, re-seeds and panels wider than 64 columns).

The kernel-level and sweep-level tests of test_gpu_kernels.py are run under BOTH settings (same oracle,
same tolerances), the launch counter proves which path really ran, and the two paths are compared
directly on identical inputs at sizes where every CTA streams many tiles (ring re-use across the
phase changes of the fused kernel).
"""

import numpy as np
import pytest
import scipy.sparse as sp

import b200arnoldi as b2a
import oracle
import test_gpu_kernels as K

pytestmark = pytest.mark.gpu

TYPES = K.TYPES


@pytest.fixture(scope="module")
def ctx():
    return b2a.default_context()


@pytest.fixture(autouse=True, params=["1", "0"], ids=["fused", "four_kernels"])
def fused(request, monkeypatch):
    monkeypatch.setenv("B2A_FUSED_SWEEP", request.param)
    return request.param


def _orth(ctx, T, Vp, v, fused_on, monkeypatch):
    monkeypatch.setenv("B2A_FUSED_SWEEP", "1" if fused_on else "0")
    n, j = Vp.shape
    ws = b2a.ArnoldiWorkspace(n, j + 1, dtype=T, ctx=ctx)
    for c in range(j):
        ws.set_col(c + 1, Vp[:, c])
    ws.set_col(j + 1, v)
    l0 = ctx.launches
    ok = ws.orthogonalize(j)
    nl = ctx.launches - l0
    h = np.array(ws.H[: j + 1, j - 1])
    vout = ws.get_cols(j + 1, 1)[:, 0].copy()
    ws.close()
    return ok, h, vout, nl


@pytest.mark.parametrize("T", TYPES)
@pytest.mark.parametrize("n,j", [(777, 3), (40_000, 9), (300_000, 24), (1_000_000, 40)])
def test_fused_equals_unfused(ctx, fused, monkeypatch, T, n, j):
    """Same inputs through both paths: one launch instead of four, identical decisions, h and v equal
    up to the summation order (the two paths tile the rows differently)."""
    if fused == "0":
        pytest.skip("direct comparison: runs both paths itself")
    if T is np.complex128 and n > 500_000:
        n = 500_000
    rng = np.random.default_rng(100 + j)
    Vp, _ = np.linalg.qr(K.randn(rng, T, n, j))
    Vp = np.asfortranarray(Vp)
    for kind in ("generic", "nearly_dependent"):
        v = K.randn(rng, T, n)
        if kind == "nearly_dependent":  # forces the second pass (expansion.jl:91): phase P3 runs
            v = Vp @ K.randn(rng, T, j) + 1e-6 * v
        ok_u, h_u, v_u, nl_u = _orth(ctx, T, Vp, v, False, monkeypatch)
        ok_f, h_f, v_f, nl_f = _orth(ctx, T, Vp, v, True, monkeypatch)
        assert nl_u in (4, 5) and nl_f == 1, (nl_u, nl_f)  # TMA path: 4 kernels (LDG fallback: 5)
        assert ok_u == ok_f is True
        tol = 1e-13 if kind == "generic" else 1e-9  # cancellation loses 6 digits in the dependent case
        assert K.relerr(h_f[:j], h_u[:j]) <= 1e-13
        assert abs(h_f[j] - h_u[j]) <= tol * abs(h_u[j])
        assert np.linalg.norm(v_f - v_u) <= tol
        assert np.abs(Vp.conj().T @ v_f).max() < 1e-13
        assert abs(np.linalg.norm(v_f) - 1) < 1e-14


@pytest.mark.parametrize("T", TYPES)
@pytest.mark.parametrize("n,j", [(1000, 1), (4097, 7), (30001, 20), (30000, 40), (12345, 61), (9000, 70)])
def test_orthogonalize_matches_oracle_fused(ctx, T, n, j):
    K.test_orthogonalize_matches_oracle(ctx, T, n, j)  # j = 70 exceeds the TMA panel width: falls back


@pytest.mark.parametrize("T", TYPES)
def test_orthogonalize_breakdown_fused(ctx, T):
    K.test_orthogonalize_breakdown(ctx, T)


@pytest.mark.parametrize("T", TYPES)
def test_iterate_arnoldi_relation_fused(ctx, T):
    K.test_iterate_arnoldi_relation(ctx, T)


@pytest.mark.parametrize("T", TYPES)
def test_iterate_arnoldi_medium_fused(ctx, T):
    K.test_iterate_arnoldi_medium(ctx, T)


def test_invariant_subspace_breakdown_fused(ctx):
    K.test_invariant_subspace_breakdown(ctx)


@pytest.mark.parametrize("T", TYPES)
def test_partialschur_fused_vs_oracle(ctx, T):
    """End to end with the fused kernel: same start vector as the oracle => same mvproducts (within one
    restart), eigenvalues to 10 tol |lambda|, ||AQ - QR|| and ||Q'Q - I|| within the reference's bounds."""
    rng = np.random.default_rng(21)
    n, nev, tol = 20_000, 6, 1e-8
    A = K.random_csr(rng, T, n, 6) * 0.1
    d = np.zeros(n, dtype=T)
    d[:12] = 4.0 + 10.0 * 0.8 ** np.arange(12)
    if T is np.complex128:
        d[:12] = d[:12] * np.exp(1j * np.linspace(0.0, 1.0, 12))
    A = (A + sp.diags(d)).tocsr()
    v1 = K.randn(rng, T, n)
    P, hist = b2a.partialschur(A, nev=nev, tol=tol, which="LM", v1=v1)
    Po, ho = oracle.partialschur(A, v1=v1, nev=nev, tol=tol, which="LM")
    assert hist.converged and ho.converged
    assert abs(hist.mvproducts - ho.mvproducts) <= 10
    lam, lam_o = np.asarray(P.eigenvalues), np.asarray(Po.eigenvalues)
    for x in lam[:nev]:
        assert np.min(np.abs(lam_o - x)) <= 10 * tol * abs(x)
    Q, R = P.Q, P.R
    assert np.linalg.norm(A @ Q - Q @ R) < n * tol
    assert np.linalg.norm(Q.conj().T @ Q - np.eye(Q.shape[1])) < 1000 * K.EPS


def test_fused_bench_shape_sweep(ctx, fused):
    """cfg-2 shape (n = 1e6, maxdim 40): a full sweep with the fused kernel keeps the Arnoldi relation
    (independent SciPy mat-vec) and orthonormality - size-independent properties at BASELINE's size."""
    rng = np.random.default_rng(7)
    n, mx = 1_000_000, 40
    cols = rng.integers(0, n, size=(n, 8))
    vals = rng.standard_normal((n, 8)) * 0.5
    A = sp.csr_matrix((vals.ravel(), cols.ravel(), np.arange(0, 8 * n + 1, 8)), shape=(n, n))
    A.sum_duplicates()
    v1 = rng.random(n)
    ws = b2a.ArnoldiWorkspace(v1, mx, ctx=ctx)
    ws.reinitialize(0, "keep")
    op = b2a.Operator.from_matrix(ctx, A)
    l0 = ctx.launches
    st = ws.iterate_arnoldi(op, 1, mx)
    # one mat-vec + one fused sweep per Arnoldi step (four-kernel path: mat-vec + S1 + S2 + S3 + finish)
    assert ctx.launches - l0 == (2 if fused == "1" else 5) * mx
    assert st.matvecs == mx and st.breakdowns == 0
    V, H = ws.V, np.array(ws.H)
    assert np.linalg.norm(A @ V[:, :mx] - V @ H) < 1e-12 * np.linalg.norm(H)
    assert np.linalg.norm(V.T @ V - np.eye(mx + 1)) < 1e-13
    ws.close()
    op.close()
