"""Upload-time argument checks and element-type rules of the drop-in boundary (round-2 advisor findings).

* A malformed sparse structure must be an ArgumentError (ValueError here), as `SparseMatrixCSC`'s constructor makes
  it in the reference, never an out-of-bounds device access.
* `partialschur(A; v1)` takes the workspace element type from v1 (src/run.jl:125 `ArnoldiWorkspace(v1, maxdim)`): a
  complex start vector with a real matrix runs in ComplexF64."""

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu

import b200arnoldi as b2a  # noqa: E402
import oracle  # noqa: E402


@pytest.fixture(scope="module")
def ctx():
    return b2a.default_context()


def small_csr(n=200, k=5, seed=0):
    rng = np.random.default_rng(seed)
    A = (sp.random(n, n, k / n, random_state=rng) + sp.identity(n)).tocsr()
    A.sort_indices()
    return A


def test_one_based_arrays_with_base_zero_are_rejected(ctx):
    A = small_csr()
    ip, idx = A.indptr.astype(np.int64) + 1, A.indices.astype(np.int64) + 1  # Julia-style arrays ...
    op = b2a.Operator.from_csr_arrays(ctx, ip, idx, A.data, A.shape[0], idx_base=1)  # ... are fine when declared
    op.close()
    with pytest.raises(ValueError):
        b2a.Operator.from_csr_arrays(ctx, ip, idx, A.data, A.shape[0], idx_base=0)


def test_column_out_of_range_is_rejected(ctx):
    A = small_csr()
    idx = A.indices.copy()
    idx[7] = A.shape[0]  # one past the end
    with pytest.raises(ValueError):
        b2a.Operator.from_csr_arrays(ctx, A.indptr, idx, A.data, A.shape[0])
    idx[7] = -1
    with pytest.raises(ValueError):
        b2a.Operator.from_csr_arrays(ctx, A.indptr, idx, A.data, A.shape[0])


def test_non_monotone_row_pointers_are_rejected(ctx):
    A = small_csr()
    ip = A.indptr.copy()
    ip[10], ip[11] = ip[11] + 1, ip[10]
    with pytest.raises(ValueError):
        b2a.Operator.from_csr_arrays(ctx, ip, A.indices, A.data, A.shape[0])
    ip = A.indptr.copy()
    ip[-1] -= 1  # does not end at nnz
    with pytest.raises(ValueError):
        b2a.Operator.from_csr_arrays(ctx, ip, A.indices, A.data, A.shape[0])


@pytest.mark.parametrize("mode", [0, 1])
def test_malformed_csc_is_rejected(ctx, mode):
    A = small_csr().tocsc()
    rv = A.indices.copy()
    rv[3] = A.shape[0] + 5
    with pytest.raises(ValueError):
        b2a.Operator.from_csc_arrays(ctx, A.indptr, rv, A.data, A.shape[0], mode=mode)
    cp = A.indptr.copy()
    cp[5] = cp[6] + 2
    with pytest.raises(ValueError):
        b2a.Operator.from_csc_arrays(ctx, cp, A.indices, A.data, A.shape[0], mode=mode)
    op = b2a.Operator.from_csc_arrays(ctx, A.indptr, A.indices, A.data, A.shape[0], mode=mode)  # the good one loads
    op.close()


def test_complex_start_vector_with_real_matrix(ctx):
    rng = np.random.default_rng(4)
    n = 400
    A = (sp.random(n, n, 6 / n, random_state=rng) * 0.2 + sp.diags(np.concatenate([10 - np.arange(8), np.zeros(n - 8)]))).tocsr()
    v1 = rng.random(n) + 1j * rng.random(n)
    P, hist = b2a.partialschur(A, nev=4, tol=1e-9, which="LM", v1=v1, ctx=ctx)
    assert P.Q.dtype == np.complex128 and P.R.dtype == np.complex128  # the run happened in ComplexF64
    Po, ho = oracle.partialschur(A.astype(np.complex128), v1=v1, nev=4, tol=1e-9, which="LM")
    assert hist.converged and ho.converged
    assert abs(hist.mvproducts - ho.mvproducts) <= 10
    assert np.allclose(np.sort_complex(P.eigenvalues)[-4:], np.sort_complex(Po.eigenvalues)[-4:], atol=1e-7)
    assert np.linalg.norm(A @ P.Q - P.Q @ P.R) < n * 1e-9
    op = b2a.Operator.from_matrix(ctx, A)  # a Float64 device operator cannot be promoted after the fact
    with pytest.raises(ValueError):
        b2a.partialschur(op, nev=4, v1=v1, ctx=ctx)
    op.close()
