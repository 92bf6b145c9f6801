"""GPU parity, kernel by kernel: the CUDA path (through the C ABI) against the oracle /
NumPy on identical seeded inputs.  Tolerances (SURVEY 8(c) v): h, wnorm relative error
<= 1e-13; ||v_gpu - v_ref|| <= 1e-13 ||v||.  RNG fill and breakdown zeros are bit-exact.
"""

import numpy as np
import pytest
import scipy.sparse as sp

import b200arnoldi as b2a
import oracle

pytestmark = pytest.mark.gpu

EPS = np.finfo(np.float64).eps
TYPES = [np.float64, np.complex128]


def randn(rng, T, *shape):
    if T is np.complex128:
        return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    return rng.standard_normal(shape)


def random_csr(rng, T, n, nnz_per_row, ragged=False):
    if ragged:
        counts = rng.integers(0, 2 * nnz_per_row + 1, size=n)
        counts[rng.integers(0, n, size=max(1, n // 50))] = 0  # empty rows
        counts[rng.integers(0, n, size=3)] = min(n, 40 * nnz_per_row + 7)  # a few long rows
    else:
        counts = np.full(n, nnz_per_row)
    indptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    indices = rng.integers(0, n, size=indptr[-1]).astype(np.int32)
    data = randn(rng, T, indptr[-1])
    A = sp.csr_matrix((data, indices, indptr), shape=(n, n))
    A.sum_duplicates()
    A.sort_indices()
    return A


def relerr(a, b):
    d = np.linalg.norm(np.asarray(a) - np.asarray(b))
    s = np.linalg.norm(np.asarray(b))
    return d / s if s > 0 else d


@pytest.fixture(scope="module")
def ctx():
    return b2a.default_context()


# ------------------------------------------------------------------------------ SpMV
@pytest.mark.parametrize("T", TYPES)
@pytest.mark.parametrize("nnz_per_row,ragged", [(1, False), (2, False), (5, False), (7, True), (16, False), (16, True), (40, False), (3, True)])
def test_spmv_csr(ctx, T, nnz_per_row, ragged):
    rng = np.random.default_rng(100 + nnz_per_row)
    n = 20011
    A = random_csr(rng, T, n, nnz_per_row, ragged)
    op = b2a.Operator.from_matrix(ctx, A)
    ws = b2a.ArnoldiWorkspace(n, 2, dtype=T, ctx=ctx)
    x = randn(rng, T, n)
    ws.set_col(1, x)
    ws.matvec(op, 1, 2)
    y = ws.get_cols(2, 1)[:, 0]
    ref = A @ x
    scale = (abs(A) @ abs(x)).max()
    assert np.abs(y - ref).max() <= 64 * EPS * scale
    assert np.array_equal(ws.get_cols(1, 1)[:, 0], x)  # x untouched


@pytest.mark.parametrize("T", TYPES)
@pytest.mark.parametrize("mode", [0, 1])
def test_spmv_csc_julia_layout(ctx, T, mode):
    """Julia's SparseMatrixCSC arrays as they are: 1-based Int64 colptr / rowval."""
    rng = np.random.default_rng(7)
    n = 5003
    A = random_csr(rng, T, n, 9, ragged=True).tocsc()
    A.sort_indices()
    colptr = A.indptr.astype(np.int64) + 1
    rowval = A.indices.astype(np.int64) + 1
    op = b2a.Operator.from_csc_arrays(ctx, colptr, rowval, A.data, n, idx_base=1, mode=mode)
    ws = b2a.ArnoldiWorkspace(n, 2, dtype=T, ctx=ctx)
    x = randn(rng, T, n)
    ws.set_col(1, x)
    ws.matvec(op, 1, 2)
    y = ws.get_cols(2, 1)[:, 0]
    scale = (abs(A) @ abs(x)).max()
    assert np.abs(y - A @ x).max() <= 64 * EPS * scale
    if mode == 0:  # transpose path is deterministic and equals the CSR kernel bit for bit
        op2 = b2a.Operator.from_matrix(ctx, A.tocsr())
        ws.matvec(op2, 1, 2)
        assert np.array_equal(ws.get_cols(2, 1)[:, 0], y)


def test_spmv_index_widths_and_bases(ctx):
    rng = np.random.default_rng(8)
    n = 3001
    A = random_csr(rng, np.float64, n, 6, ragged=True)
    x = rng.standard_normal(n)
    ref = None
    for width in (np.int32, np.int64):
        for base in (0, 1):
            op = b2a.Operator.from_csr_arrays(
                ctx, A.indptr.astype(width) + base, A.indices.astype(width) + base, A.data, n, idx_base=base)
            ws = b2a.ArnoldiWorkspace(n, 2, ctx=ctx)
            ws.set_col(1, x)
            ws.matvec(op, 1, 2)
            y = ws.get_cols(2, 1)[:, 0]
            if ref is None:
                ref = y
                assert np.abs(y - A @ x).max() < 1e-12
            assert np.array_equal(y, ref)


# --------------------------------------------------------------------- Gram-Schmidt
def orthonormal_panel(rng, T, n, j):
    Qm, _ = np.linalg.qr(randn(rng, T, n, j))
    return np.asfortranarray(Qm)


@pytest.mark.parametrize("T", TYPES)
@pytest.mark.parametrize("n,j", [(1000, 1), (4097, 7), (30001, 20), (30000, 40), (12345, 61), (9000, 70)])
def test_orthogonalize_matches_oracle(ctx, T, n, j):
    rng = np.random.default_rng(n + j)
    maxdim = j + 1
    Vp = orthonormal_panel(rng, T, n, j)
    for kind in ("generic", "nearly_dependent"):
        v = randn(rng, T, n)
        if kind == "nearly_dependent":  # forces the DGKS second pass (expansion.jl:91)
            v = Vp @ randn(rng, T, j) + 1e-6 * v
        arn = oracle.ArnoldiWorkspace(T, n, maxdim)
        arn.V[:, :j] = Vp
        arn.V[:, j] = v
        ok_ref = oracle.orthogonalize(arn, j)

        ws = b2a.ArnoldiWorkspace(n, maxdim, dtype=T, ctx=ctx)
        for c in range(j):
            ws.set_col(c + 1, Vp[:, c])
        ws.set_col(j + 1, v)
        ok = ws.orthogonalize(j)
        assert ok == ok_ref
        h_gpu, h_ref = ws.H[: j + 1, j - 1], arn.H[: j + 1, j - 1]
        assert relerr(h_gpu, h_ref) <= 1e-13
        # generic: 1e-13.  nearly dependent: v loses 6 digits to cancellation, so wnorm (and the
        # normalised v) are only determined to ~1e6 * eps; stated tolerance 1e-9.
        assert abs(h_gpu[j] - h_ref[j]) <= (1e-13 if kind == "generic" else 1e-9) * abs(h_ref[j])
        v_gpu = ws.get_cols(j + 1, 1)[:, 0]
        assert np.linalg.norm(v_gpu - arn.V[:, j]) <= (1e-13 if kind == "generic" else 1e-9)
        assert np.abs(Vp.conj().T @ v_gpu).max() < 1e-13
        assert abs(np.linalg.norm(v_gpu) - 1) < 1e-14
        ws.close()


@pytest.mark.parametrize("T", TYPES)
def test_orthogonalize_breakdown(ctx, T):
    """v in span(V): returns false, H[j+1,j] == 0 exactly, v not normalised (expansion.jl:99-102)."""
    rng = np.random.default_rng(5)
    n, j = 5000, 6
    Vp = orthonormal_panel(rng, T, n, j)
    v = Vp @ randn(rng, T, j)
    ws = b2a.ArnoldiWorkspace(n, j + 1, dtype=T, ctx=ctx)
    for c in range(j):
        ws.set_col(c + 1, Vp[:, c])
    ws.set_col(j + 1, v)
    assert ws.orthogonalize(j) is False
    assert ws.H[j, j - 1] == 0
    assert relerr(ws.H[:j, j - 1], Vp.conj().T @ v) < 1e-12
    # a zero vector: rnorm = 0 -> no second pass, 0 <= 0 -> breakdown
    ws.set_col(j + 1, np.zeros(n, dtype=T))
    assert ws.orthogonalize(j) is False
    assert ws.H[j, j - 1] == 0 and np.all(ws.H[:j, j - 1] == 0)


@pytest.mark.parametrize("T", TYPES)
def test_reinitialize(ctx, T):
    rng = np.random.default_rng(6)
    n, j = 7001, 5
    ws = b2a.ArnoldiWorkspace(n, 8, dtype=T, ctx=ctx)
    # j == 0: fill + normalise only (expansion.jl:27-30); the fill is bit-reproducible on the host
    assert ws.reinitialize(0, "rand", seed=42) is True
    u = b2a.uniform_reference(42, 0, n, T)
    v0 = ws.get_cols(1, 1)[:, 0]
    assert relerr(v0, u / np.linalg.norm(u)) < 1e-15
    assert (u.real >= 0).all() and (u.real < 1).all()
    # j > 0: random column orthonormal against V[:, 1:j]
    Vp = orthonormal_panel(rng, T, n, j)
    for c in range(j):
        ws.set_col(c + 1, Vp[:, c])
    assert ws.reinitialize(j, "rand", seed=42) is True
    w = ws.get_cols(j + 1, 1)[:, 0]
    u1 = b2a.uniform_reference(42, 1, n, T)  # second draw of this workspace
    ref = u1 - Vp @ (Vp.conj().T @ u1)
    ref -= Vp @ (Vp.conj().T @ ref)
    assert relerr(w, ref / np.linalg.norm(ref)) < 1e-12
    assert np.abs(Vp.conj().T @ w).max() < 1e-14 and abs(np.linalg.norm(w) - 1) < 1e-14
    # keep mode == the v1 path of partialschur: normalise what is there, input untouched on the host
    x = randn(rng, T, n)
    ws.set_col(1, x)
    assert ws.reinitialize(0, "keep") is True
    assert relerr(ws.get_cols(1, 1)[:, 0], x / np.linalg.norm(x)) < 1e-15


# ------------------------------------------------------------------ basis rotation
@pytest.mark.parametrize("T", TYPES)
@pytest.mark.parametrize("n,maxdim,purge,k", [(1000, 6, 1, 4), (20001, 20, 1, 12), (20001, 20, 3, 13),
                                              (4099, 40, 7, 31), (4099, 40, 12, 30), (3000, 60, 1, 45),
                                              (2000, 12, 1, 12), (777, 90, 5, 60)])
def test_rotate_basis(ctx, T, n, maxdim, purge, k):
    rng = np.random.default_rng(n + maxdim + purge)
    V = np.asfortranarray(randn(rng, T, n, maxdim + 1))
    Q = np.asfortranarray(randn(rng, T, maxdim, maxdim))
    ws = b2a.ArnoldiWorkspace(n, maxdim, dtype=T, ctx=ctx)
    for c in range(maxdim + 1):
        ws.set_col(c + 1, V[:, c])
    ws.rotate_basis(purge, k, maxdim, Q)
    ref = V.copy()
    ref[:, purge - 1 : k] = V[:, purge - 1 : maxdim] @ Q[purge - 1 : maxdim, purge - 1 : k]  # run.jl:363-364
    if k < maxdim:
        ref[:, k] = V[:, maxdim]  # run.jl:365
    got = ws.V
    assert np.abs(got - ref).max() <= 1e-13 * maxdim * np.abs(V).max() * np.abs(Q).max()
    # columns outside purge:k+1 are untouched, bit for bit
    keep = [c for c in range(maxdim + 1) if not (purge - 1 <= c <= min(k, maxdim))]
    assert np.array_equal(got[:, keep], V[:, keep])
    ws.close()


@pytest.mark.parametrize("T", TYPES)
def test_rotate_final_and_basis_times(ctx, T):
    rng = np.random.default_rng(11)
    n, maxdim, nconv = 5003, 20, 11
    V = np.asfortranarray(randn(rng, T, n, maxdim + 1))
    Q = np.asfortranarray(randn(rng, T, maxdim, maxdim))
    ws = b2a.ArnoldiWorkspace(n, maxdim, dtype=T, ctx=ctx)
    for c in range(maxdim + 1):
        ws.set_col(c + 1, V[:, c])
    ws.rotate_final(nconv, Q)
    ref = V.copy()
    ref[:, :nconv] = V[:, :nconv] @ Q[:nconv, :nconv]  # run.jl:382-383
    assert np.abs(ws.V - ref).max() < 1e-12
    Y = randn(rng, np.complex128, nconv, nconv)
    X = ws.basis_times(Y)
    assert np.abs(X - ref[:, :nconv] @ Y).max() < 1e-12


# ------------------------------------------------------------------- expansion sweep
@pytest.mark.parametrize("T", TYPES)
def test_iterate_arnoldi_relation(ctx, T):
    """test/expansion.jl:12-32 - A V[:,1:m] = V H and orthonormality, plus H against the oracle."""
    rng = np.random.default_rng(12)
    n, mx = 10, 6
    A = (sp.random(n, n, 0.1, random_state=rng) + sp.identity(n)).tocsr().astype(T)
    v1 = randn(rng, T, n)
    ws = b2a.ArnoldiWorkspace(v1, mx, ctx=ctx)
    ws.reinitialize(0, "keep")
    op = b2a.Operator.from_matrix(ctx, A)
    ws.iterate_arnoldi(op, 1, 3)
    V, H = ws.V, np.array(ws.H)
    assert np.allclose(A @ V[:, :3], V[:, :4] @ H[:4, :3])
    assert np.linalg.norm(V[:, :4].conj().T @ V[:, :4] - np.eye(4)) < np.sqrt(EPS) / 100
    ws.iterate_arnoldi(op, 4, mx)
    V, H = ws.V, np.array(ws.H)
    assert np.allclose(A @ V[:, :mx], V @ H)
    assert np.linalg.norm(V.conj().T @ V - np.eye(mx + 1)) < np.sqrt(EPS) / 100

    arn = oracle.ArnoldiWorkspace(T, n, mx)
    arn.V[:, 0] = v1 / np.linalg.norm(v1)
    oracle.iterate_arnoldi(A, arn, 1, mx)
    assert np.abs(H - arn.H).max() < 1e-12
    # compare basis vectors while the recurrence is well conditioned (|h_{j+1,j}| not tiny)
    good = 1 + int(np.argmax(np.append(np.abs(np.diag(arn.H, -1)) < 1e-6, True)))
    assert good >= 4
    assert np.abs(V[:, :good] - arn.V[:, :good]).max() < 1e-9


@pytest.mark.parametrize("T", TYPES)
def test_iterate_arnoldi_medium(ctx, T):
    rng = np.random.default_rng(13)
    n, mx = 50000, 30
    A = random_csr(rng, T, n, 8) + sp.identity(n) * 3
    A = A.tocsr()
    v1 = randn(rng, T, n)
    ws = b2a.ArnoldiWorkspace(v1, mx, ctx=ctx)
    ws.reinitialize(0, "keep")
    op = b2a.Operator.from_matrix(ctx, A)
    st = ws.iterate_arnoldi(op, 1, mx)
    assert st.matvecs == mx and st.breakdowns == 0
    V, H = ws.V, np.array(ws.H)
    assert np.linalg.norm(A @ V[:, :mx] - V @ H) < 1e-12 * np.linalg.norm(H)
    assert np.linalg.norm(V.conj().T @ V - np.eye(mx + 1)) < 1e-13
    arn = oracle.ArnoldiWorkspace(T, n, mx)
    arn.V[:, 0] = v1 / np.linalg.norm(v1)
    oracle.iterate_arnoldi(A, arn, 1, mx)
    assert np.abs(H - arn.H).max() < 1e-10 * np.abs(arn.H).max()


def test_invariant_subspace_breakdown(ctx):
    """test/expansion.jl:34-55: block diagonal A, v1 = e1 -> H[5,4] == 0 exactly, V orthonormal."""
    rng = np.random.default_rng(14)
    A = np.zeros((8, 8))
    A[:4, :4] = rng.random((4, 4))
    A[4:, 4:] = rng.random((4, 4))
    e1 = np.zeros(8)
    e1[0] = 1
    ws = b2a.ArnoldiWorkspace(e1, 5, ctx=ctx)
    op = b2a.Operator.from_matrix(ctx, A)
    st = ws.iterate_arnoldi(op, 1, 5, seed=3)
    V, H = ws.V, np.array(ws.H)
    assert H[4, 3] == 0
    assert st.breakdowns == 1
    assert np.linalg.norm(V.T @ V - np.eye(6)) < np.sqrt(EPS) / 100
    assert np.allclose(A @ V[:, :5], V @ H, atol=1e-13)


def test_callback_operator(ctx):
    """The matrix-free `mul!(y, A, x)` contract: a torch function as the operator."""
    import torch

    rng = np.random.default_rng(15)
    n = 4000
    d = np.linspace(1, 5, n)
    dt = torch.tensor(d, device="cuda")

    def fn(x):  # y = D x + shift-by-one coupling
        return dt * x + 0.1 * torch.roll(x, 1)

    op = b2a.Operator.from_torch_function(ctx, np.float64, n, fn)
    v1 = rng.random(n)
    ws = b2a.ArnoldiWorkspace(v1, 12, ctx=ctx)
    ws.reinitialize(0, "keep")
    ws.iterate_arnoldi(op, 1, 12)
    V, H = ws.V, np.array(ws.H)
    Ad = sp.diags(d) + 0.1 * sp.csr_matrix((np.ones(n), (np.arange(n), (np.arange(n) - 1) % n)), shape=(n, n))
    assert np.linalg.norm(Ad @ V[:, :12] - V @ H) < 1e-12


def test_spmv_tma_stream_kernel_opt_in(ctx, monkeypatch):
    """The TMA-staged CSR-stream kernel (opt-in, B2A_SPMV_TMA=1) gives the same result as SciPy."""
    monkeypatch.setenv("B2A_SPMV_TMA", "1")
    rng = np.random.default_rng(77)
    for T, k, ragged in [(np.float64, 16, False), (np.float64, 7, True), (np.complex128, 20, True), (np.float64, 3, True)]:
        n = 30011
        A = random_csr(rng, T, n, k, ragged)
        op = b2a.Operator.from_matrix(ctx, A)
        ws = b2a.ArnoldiWorkspace(n, 2, dtype=T, ctx=ctx)
        x = randn(rng, T, n)
        ws.set_col(1, x)
        ws.matvec(op, 1, 2)
        y = ws.get_cols(2, 1)[:, 0]
        assert np.abs(y - A @ x).max() <= 64 * EPS * (abs(A) @ abs(x)).max()


@pytest.mark.parametrize("T", TYPES)
def test_generic_method_on_device_ops(ctx, T):
    """The reference's GENERIC orthogonalize! (src/expansion.jl:69-109) executed operation by operation
    through the BLAS-level entry points (norm, V'v, v -= V h, v ./= a) equals the fused device sweep."""
    rng = np.random.default_rng(31)
    n, j = 20003, 9
    Vp = orthonormal_panel(rng, T, n, j)
    for kind in ("generic", "nearly_dependent"):
        v = randn(rng, T, n)
        if kind == "nearly_dependent":
            v = Vp @ randn(rng, T, j) + 1e-6 * v
        ws1 = b2a.ArnoldiWorkspace(n, j + 1, dtype=T, ctx=ctx)
        ws2 = b2a.ArnoldiWorkspace(n, j + 1, dtype=T, ctx=ctx)
        for w in (ws1, ws2):
            for c in range(j):
                w.set_col(c + 1, Vp[:, c])
            w.set_col(j + 1, v)
        # generic method, one device call per BLAS operation
        eta = np.sqrt(2) / 2
        rnorm = ws1.norm(j + 1)
        assert abs(rnorm - np.linalg.norm(v)) <= 1e-14 * rnorm
        h = ws1.gemv_c(j, j + 1)
        ws1.gemv_n_sub(j, j + 1, h)
        wnorm = ws1.norm(j + 1)
        second = wnorm < eta * rnorm
        if second:
            rnorm = wnorm
            c = ws1.gemv_c(j, j + 1)
            ws1.gemv_n_sub(j, j + 1, c)
            h = h + c
            wnorm = ws1.norm(j + 1)
        assert wnorm > eta * rnorm
        ws1.scal_div(j + 1, wnorm)
        assert second == (kind == "nearly_dependent")
        # fused sweep
        assert ws2.orthogonalize(j)
        tol = 1e-13 if kind == "generic" else 1e-9
        assert relerr(ws2.H[:j, j - 1], h) <= 1e-13
        assert abs(ws2.H[j, j - 1] - wnorm) <= tol * wnorm
        assert np.linalg.norm(ws1.get_cols(j + 1, 1) - ws2.get_cols(j + 1, 1)) <= tol
        ws1.copy_col(j + 1, 1)
        assert np.array_equal(ws1.get_cols(1, 1), ws1.get_cols(j + 1, 1))
        ws1.close()
        ws2.close()


def test_csr_from_device_arrays(ctx):
    """b2a_csr_create_device: operator over CSR arrays that already live in HBM (borrowed, not copied)."""
    import ctypes as C

    import torch
    from arnoldimethod_jl_b200 import _lib as L

    rng = np.random.default_rng(88)
    n = 10007
    A = random_csr(rng, np.float64, n, 6, ragged=True)
    d_ptr = torch.from_numpy(A.indptr.astype(np.int64)).cuda()
    d_idx = torch.from_numpy(A.indices.astype(np.int32)).cuda()
    d_val = torch.from_numpy(A.data).cuda()
    torch.cuda.synchronize()
    h = C.c_void_p()
    L.check(L.lib().b2a_csr_create_device(ctx._h, L.F64, n, n, 0, A.nnz, d_ptr.data_ptr(), d_idx.data_ptr(),
                                          d_val.data_ptr(), C.byref(h)))
    op = b2a.Operator(ctx, h, np.float64, n, n, 0, keep=(d_ptr, d_idx, d_val))
    ws = b2a.ArnoldiWorkspace(n, 2, ctx=ctx)
    x = rng.standard_normal(n)
    ws.set_col(1, x)
    ws.matvec(op, 1, 2)
    assert np.abs(ws.get_cols(2, 1)[:, 0] - A @ x).max() < 1e-12
    assert op.bytes_per_matvec == A.nnz * 12 + 8 * (n + 1) + 16 * n


@pytest.mark.parametrize("T", TYPES)
@pytest.mark.parametrize("block_mb", ["0.02", "0.05", "0.3"])
def test_spmv_column_blocked(ctx, monkeypatch, T, block_mb):
    """Column-blocked (block-major CSR, one pass per block, accumulate) SpMV equals SciPy; forced through
    B2A_SPMV_BLOCK_MB on a small matrix (automatic mode only triggers when x exceeds 48 MB)."""
    monkeypatch.setenv("B2A_SPMV_BLOCK_MB", block_mb)
    rng = np.random.default_rng(91)
    n = 20011
    for k, ragged in [(15, False), (6, True), (1, False)]:
        A = random_csr(rng, T, n, k, ragged)
        op = b2a.Operator.from_matrix(ctx, A)
        ws = b2a.ArnoldiWorkspace(n, 2, dtype=T, ctx=ctx)
        x = randn(rng, T, n)
        ws.set_col(1, x)
        ws.matvec(op, 1, 2)
        y = ws.get_cols(2, 1)[:, 0]
        assert np.abs(y - A @ x).max() <= 64 * EPS * (abs(A) @ abs(x)).max()
        ws.matvec(op, 1, 2)  # idempotent: the first block pass overwrites y
        assert np.array_equal(ws.get_cols(2, 1)[:, 0], y)
    # Julia-style 1-based Int64 input goes through the same builder
    A = random_csr(rng, T, n, 9, True)
    op = b2a.Operator.from_csr_arrays(ctx, A.indptr.astype(np.int64) + 1, A.indices.astype(np.int64) + 1, A.data, n, idx_base=1)
    ws = b2a.ArnoldiWorkspace(n, 2, dtype=T, ctx=ctx)
    x = randn(rng, T, n)
    ws.set_col(1, x)
    ws.matvec(op, 1, 2)
    assert np.abs(ws.get_cols(2, 1)[:, 0] - A @ x).max() <= 64 * EPS * (abs(A) @ abs(x)).max()


def test_partialschur_with_column_blocked_operator(ctx, monkeypatch):
    monkeypatch.setenv("B2A_SPMV_BLOCK_MB", "0.1")
    rng = np.random.default_rng(92)
    n = 40000
    A = sp.random(n, n, 12 / n, random_state=rng, format="csr") * 0.3
    d = np.zeros(n)
    d[:12] = 5 + 20 * 0.8 ** np.arange(12)
    A = (A + sp.diags(d)).tocsr()
    v1 = rng.random(n)
    P, hist = b2a.partialschur(A, nev=6, tol=1e-8, v1=v1)
    monkeypatch.setenv("B2A_SPMV_BLOCK_MB", "0")
    P0, hist0 = b2a.partialschur(A, nev=6, tol=1e-8, v1=v1)
    assert hist.converged and hist.mvproducts == hist0.mvproducts
    assert np.allclose(np.sort_complex(P.eigenvalues), np.sort_complex(P0.eigenvalues), atol=1e-10)
    assert np.linalg.norm(A @ P.Q - P.Q @ P.R) < n * 1e-8
