"""Shift-and-invert on the device (reference: docs/src/index.md:234-262 "Shift-and-invert with
LinearMaps.jl", bench/partial_schur.jl:11-35): the linear map x -> (A - sigma I) \\ x as a Jacobi-CG solve on the CSR
mat-vec, handed to partialschur with which = :LM exactly like the reference's LinearMap.  Checked against SciPy's sparse
LU (the factorisation the reference example uses) and against the oracle driven by that LU map from the same v1."""

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

pytestmark = pytest.mark.gpu

import b200arnoldi as b2a  # noqa: E402
import oracle  # noqa: E402


@pytest.fixture(scope="module")
def ctx():
    return b2a.default_context()


def stencil3d(nx, ny, nz, wx=1.0, wy=1.0, wz=1.0):
    def t(n):
        return sp.diags([-np.ones(n - 1), 2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1], format="csr")

    ix, iy, iz = sp.identity(nx), sp.identity(ny), sp.identity(nz)
    return (wx * sp.kron(sp.kron(t(nx), iy), iz) + wy * sp.kron(sp.kron(ix, t(ny)), iz)
            + wz * sp.kron(sp.kron(ix, iy), t(nz))).tocsr()


class LUMap:
    """The reference example's `LinearMap((y, x) -> ldiv!(y, F, x))` for the oracle."""

    def __init__(self, A, sigma=0.0):
        self.shape, self.dtype = A.shape, A.dtype
        self.lu = spla.splu((A - sigma * sp.identity(A.shape[0])).tocsc())

    def __matmul__(self, x):
        return self.lu.solve(np.asarray(x))


@pytest.mark.parametrize("sigma", [0.0, -0.7])
def test_inner_solve_matches_sparse_lu(ctx, sigma):
    A = stencil3d(20, 18, 16, 1.0, 1.3, 1.7)
    n = A.shape[0]
    b = np.random.default_rng(1).standard_normal(n)
    op = b2a.Operator.from_matrix(ctx, A)
    S = b2a.Operator.shift_invert(op, sigma=sigma, rtol=1e-13)
    ws = b2a.ArnoldiWorkspace(n, 2, ctx=ctx)
    ws.set_col(1, b)
    ws.matvec(S, 1, 2)
    x = ws.get_cols(2, 1)[:, 0]
    ref = LUMap(A, sigma) @ b
    assert np.linalg.norm(x - ref) <= 1e-10 * np.linalg.norm(ref)
    assert np.linalg.norm((A - sigma * sp.identity(n)) @ x - b) <= 1e-12 * np.linalg.norm(b) * 10
    solves, iters, worst = S.solve_stats
    assert solves == 1 and 0 < iters < 2000 and worst <= 1e-13
    ws.close()
    S.close()
    op.close()


def test_complex_hermitian_positive_definite_solve(ctx):
    rng = np.random.default_rng(2)
    n = 3000
    B = sp.random(n, n, 4 / n, random_state=rng) + 1j * sp.random(n, n, 4 / n, random_state=rng)
    A = (B.conj().T @ B + sp.diags(1.0 + rng.random(n))).tocsr().astype(np.complex128)
    A.sort_indices()
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    op = b2a.Operator.from_matrix(ctx, A)
    S = b2a.Operator.shift_invert(op, sigma=0.0)
    ws = b2a.ArnoldiWorkspace(n, 2, dtype=np.complex128, ctx=ctx)
    ws.set_col(1, b)
    ws.matvec(S, 1, 2)
    x = ws.get_cols(2, 1)[:, 0]
    ref = LUMap(A) @ b
    assert np.linalg.norm(x - ref) <= 1e-10 * np.linalg.norm(ref)
    ws.close()
    S.close()
    op.close()


def test_smallest_laplacian_modes_by_shift_invert(ctx):
    """The reference recipe end to end: eigenvalues of A nearest 0 are 1 / theta for the largest-magnitude theta of
    inv(A).  Same v1 for the device map (CG) and the oracle driven by SciPy's LU: same restart path, same answer, and
    an order of magnitude fewer outer steps than asking for :SR directly."""
    nx, ny, nz = 24, 22, 20
    w = (1.0, 1.37, 1.83)
    A = stencil3d(nx, ny, nz, *w)
    n = A.shape[0]
    lam = [wi * (2 - 2 * np.cos(np.arange(1, m + 1) * np.pi / (m + 1))) for wi, m in zip(w, (nx, ny, nz))]
    exact = np.sort((lam[0][:, None, None] + lam[1][None, :, None] + lam[2][None, None, :]).ravel())
    v1 = np.random.default_rng(5).random(n)
    op = b2a.Operator.from_matrix(ctx, A)
    S = b2a.Operator.shift_invert(op, sigma=0.0, rtol=1e-13)
    P, hist = b2a.partialschur(S, nev=10, which="LM", tol=1e-8, v1=v1, ctx=ctx)
    Po, ho = oracle.partialschur(LUMap(A), v1=v1, nev=10, which="LM", tol=1e-8)
    assert hist.converged and ho.converged and hist.nconverged == ho.nconverged
    assert abs(hist.mvproducts - ho.mvproducts) <= 10, (hist.mvproducts, ho.mvproducts)
    assert hist.mvproducts < 150  # the direct :SR run of test_stencil_smallest_real_matches_oracle needs ~450
    got = np.sort(1.0 / P.eigenvalues.real)
    assert np.allclose(got[:10], exact[:10], rtol=1e-7)
    assert np.allclose(np.sort(1.0 / Po.eigenvalues.real)[:10], got[:10], rtol=1e-7)
    # Schur vectors of inv(A) are Schur vectors of A: A Q = Q inv(R)
    Q, R = P.Q, P.R
    assert np.linalg.norm(A @ Q @ R - Q) < 1e-6
    vals, X = b2a.partialeigen(P)
    for i in range(len(vals)):
        x = X[:, i].real / np.linalg.norm(X[:, i].real)
        assert np.linalg.norm(A @ x - x / vals[i].real) <= 1e-6 / abs(vals[i])
    solves, iters, worst = S.solve_stats
    assert solves == hist.mvproducts and worst <= 1e-13
    P.workspace.close()
    S.close()
    op.close()


def test_singular_shift_is_an_error_not_a_wrong_answer(ctx):
    """A shift that hits an eigenvalue makes A - sigma I singular: the solve cannot converge and the mat-vec must fail
    loudly (B2A_ERR_SOLVE), not return garbage."""
    A = stencil3d(12, 11, 10)
    n = A.shape[0]
    sigma = sum(2 - 2 * np.cos(np.pi / (m + 1)) for m in (12, 11, 10))  # the smallest eigenvalue, exactly
    op = b2a.Operator.from_matrix(ctx, A)
    S = b2a.Operator.shift_invert(op, sigma=sigma, maxit=300)
    ws = b2a.ArnoldiWorkspace(n, 2, ctx=ctx)
    ws.set_col(1, np.random.default_rng(6).standard_normal(n))
    with pytest.raises(b2a.B200Error):
        ws.matvec(S, 1, 2)
    ws.close()
    S.close()
    op.close()


def test_shift_invert_argument_checks(ctx):
    A = stencil3d(6, 5, 4)
    op = b2a.Operator.from_matrix(ctx, A)
    with pytest.raises(ValueError):
        b2a.Operator.shift_invert(op, sigma=1j)  # complex shift on a Float64 operator
    cb = b2a.Operator.from_torch_function(ctx, np.float64, A.shape[0], lambda x: x)
    with pytest.raises(ValueError):
        b2a.Operator.shift_invert(cb)  # needs a CSR operator
    op.close()
