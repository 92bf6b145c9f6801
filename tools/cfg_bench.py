"""BASELINE configs 3 and 4 at (scaled) size: correctness properties + per-kernel roofline numbers.

    python tools/cfg_bench.py cfg3 [N]     7-point Laplacian N^3 (default 256), nev 10, maxdim 20, :SR, 3 restarts
    python tools/cfg_bench.py cfg4 [n]     ComplexF64 random CSR, 20 nnz/row (default n = 2e6), nev 30, maxdim 60, :LM
    python tools/cfg_bench.py cfg5shard    one GPU's share of cfg 5 run as a stand-alone problem: Float64 random CSR,
                                           n = 1.25e7, 15 nnz/row, nev 20, maxdim 40, :LM, 2 restarts
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp

import b200arnoldi as b2a


def laplacian_csr(N):
    """7-point Laplacian on an N^3 grid (diag 6, off -1, Dirichlet) built directly as CSR arrays."""
    n = N ** 3
    idx = np.arange(n, dtype=np.int64)
    i, j, k = idx // (N * N), (idx // N) % N, idx % N
    cols = [idx]
    vals = [np.full(n, 6.0)]
    for cond, off in ((i > 0, -N * N), (i < N - 1, N * N), (j > 0, -N), (j < N - 1, N), (k > 0, -1), (k < N - 1, 1)):
        c = np.where(cond, idx + off, -1)
        cols.append(c)
        vals.append(np.where(cond, -1.0, 0.0))
    C = np.stack(cols, axis=1)
    V = np.stack(vals, axis=1)
    order = np.argsort(np.where(C < 0, np.iinfo(np.int64).max, C), axis=1, kind="stable")
    C = np.take_along_axis(C, order, axis=1)
    V = np.take_along_axis(V, order, axis=1)
    mask = C >= 0
    counts = mask.sum(axis=1)
    indptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return indptr, C[mask].astype(np.int32), V[mask], n


def report(name, ctx, hist, wall, extra):
    prof = ctx.profile_report()
    kern = {k: dict(launches=r["launches"], avg_us=round(1e3 * r["ms"] / r["launches"], 1),
                    gbs=round(r["bytes"] / (r["ms"] * 1e-3) / 1e9)) for k, r in prof.items() if r["launches"]}
    print(json.dumps(dict(config=name, mvproducts=hist.mvproducts, restarts=hist.restarts, nconverged=hist.nconverged,
                          second_pass_rate=round(hist.stats["second_passes"] / max(1, hist.mvproducts), 3),
                          steps_per_s=round(hist.mvproducts / wall, 1), wall_ms=round(1e3 * wall, 2),
                          algorithmic_GBs=round(hist.stats["bytes"] / wall / 1e9), kernels=kern, **extra)), flush=True)


def main():
    which = sys.argv[1]
    ctx = b2a.default_context()
    rng = np.random.default_rng(0)
    if which == "cfg3":
        N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
        indptr, indices, data, n = laplacian_csr(N)
        op = b2a.Operator.from_csr_arrays(ctx, indptr, indices, data, n)
        kw = dict(nev=10, which="SR", tol=1e-6, mindim=10, maxdim=20, restarts=3)
        T = np.float64
    elif which == "cfg5shard":
        n, k = 12_500_000, 15
        indptr = np.arange(0, (n + 1) * k, k, dtype=np.int64)
        indices = rng.integers(0, n, size=n * k, dtype=np.int32)
        indices = np.sort(indices.reshape(n, k), axis=1).ravel()
        data = rng.standard_normal(n * k) * 0.5
        d = np.arange(40)
        data.reshape(n, k)[:40, 0] = 5 + 20 * 0.9 ** d
        indices.reshape(n, k)[:40, 0] = d
        op = b2a.Operator.from_csr_arrays(ctx, indptr, indices, data, n)
        kw = dict(nev=20, which="LM", tol=1e-6, mindim=20, maxdim=40, restarts=2)
        T = np.float64
    else:
        n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 2_000_000
        k = 20
        indptr = np.arange(0, (n + 1) * k, k, dtype=np.int64)
        indices = np.sort(rng.integers(0, n, size=(n, k)), axis=1).astype(np.int32).ravel()
        data = (rng.standard_normal(n * k) + 1j * rng.standard_normal(n * k)) * 0.35
        d = np.arange(60)
        data.reshape(n, k)[:60, 0] = (5 + 20 * 0.9 ** d) * np.exp(1j * np.linspace(0, 1, 60))
        indices.reshape(n, k)[:60, 0] = d  # (row order no longer sorted for these rows: allowed)
        op = b2a.Operator.from_csr_arrays(ctx, indptr, indices, data, n)
        kw = dict(nev=30, which="LM", tol=1e-6, mindim=30, maxdim=60, restarts=2)
        T = np.complex128
    v1 = rng.random(n).astype(T)
    P = None
    for rep in range(2):  # warm-up, then measured with per-kernel events
        if P is not None:
            P.workspace.close()  # give the pool its memory back before the measured repetition
        ctx.profile(rep == 1)
        ctx.synchronize()
        t0 = time.perf_counter()
        P, hist = b2a.partialschur(op, v1=v1, **kw)
        ctx.synchronize()
        wall = time.perf_counter() - t0
    ws = P.workspace
    # properties that do not need the oracle: orthonormal basis, Arnoldi/Krylov-Schur relation on a sample
    m = kw["maxdim"]
    V = ws.get_cols(1, 6)
    G = V.conj().T @ V
    A = sp.csr_matrix((data, indices, indptr), shape=(n, n))
    extra = dict(n=n, nnz=int(indptr[-1]), orth_err=float(np.abs(G - np.eye(6)).max()))
    if hist.nconverged:
        Q, R = P.Q, P.R
        extra["residual_AQ_QR"] = float(np.linalg.norm(A @ Q - Q @ R))
    else:
        # no locked vectors yet: check the relation A v_1 = V[:, 1:2] H[1:2, 1] of a fresh 1-step sweep
        ws2 = b2a.ArnoldiWorkspace(v1, 4, ctx=ctx)
        ws2.reinitialize(0, "keep")
        ws2.iterate_arnoldi(op, 1, 3)
        V2, H2 = ws2.V, np.array(ws2.H)
        extra["arnoldi_relation_err"] = float(np.linalg.norm(A @ V2[:, :3] - V2 @ H2[:, :3]) / np.linalg.norm(H2))
    report(which, ctx, hist, wall, extra)
    ctx.profile(False)


main()
