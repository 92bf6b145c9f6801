"""SpMV microbenchmark: CUDA-event time of b2a_ws_matvec on a synthetic CSR matrix.
    python tools/spmvbench.py [n] [nnz_per_row] [--laplace N] [--sweep] [--complex]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp

import b200arnoldi as b2a


def laplacian3d(N):
    I = sp.identity(N, format="csr")
    T1 = sp.diags([-np.ones(N - 1), 2 * np.ones(N), -np.ones(N - 1)], [-1, 0, 1], format="csr")
    return (sp.kron(sp.kron(T1, I), I) + sp.kron(sp.kron(I, T1), I) + sp.kron(sp.kron(I, I), T1)).tocsr()


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    T = np.complex128 if "--complex" in sys.argv else np.float64
    rng = np.random.default_rng(0)
    if "--laplace" in sys.argv:
        N = int(args[0]) if args else 128
        A = laplacian3d(N).astype(T)
        n = N ** 3
    else:
        n = int(float(args[0])) if args else 1_000_000
        k = int(args[1]) if len(args) > 1 else 16
        indptr = np.arange(0, (n + 1) * k, k, dtype=np.int64)
        cols = np.sort(rng.integers(0, n, size=(n, k), dtype=np.int32), axis=1).ravel()
        vals = rng.standard_normal(n * k).astype(T)
        A = sp.csr_matrix((vals, cols, indptr), shape=(n, n))
    x = rng.standard_normal(n).astype(T)
    ref = A @ x
    ctx = b2a.default_context()
    configs = [dict()]
    if "--blocksweep" in sys.argv:
        configs = [dict(B2A_SPMV_BLOCK_MB=str(mb)) for mb in (0, 16, 24, 32, 48, -1)]
    if "--esweep" in sys.argv:  # entries in flight per lane x lanes per row x rows in flight
        lprs = [int(v) for v in os.environ.get("SWEEP_LPR", "2,4,8").split(",")]
        configs = [dict(B2A_SPMV_E=str(e), B2A_SPMV_LPR=str(l), B2A_SPMV_U=str(u))
                   for e in (1, 2, 4) for l in lprs for u in (2, 4)]
    if "--sweep" in sys.argv:
        lprs = [int(v) for v in os.environ.get("SWEEP_LPR", "8,16").split(",")]
        configs = [dict(B2A_SPMV_U=str(u), B2A_SPMV_GRID=str(g), B2A_SPMV_LPR=str(l))
                   for l in lprs for u in (2, 4, 8) for g in (8, 16)]
    for cfg in configs:
        for k_ in ("B2A_SPMV_U", "B2A_SPMV_GRID", "B2A_SPMV_LPR", "B2A_SPMV_BLOCK_MB", "B2A_SPMV_E"):
            os.environ.pop(k_, None)
        os.environ.update(cfg)
        op = b2a.Operator.from_matrix(ctx, A)
        ws = b2a.ArnoldiWorkspace(n, 2, dtype=T, ctx=ctx)
        ws.set_col(1, x)
        for it in range(25):
            if it == 5:
                ctx.profile(True)
            ws.matvec(op, 1, 2)
        ctx.synchronize()
        ws.reinitialize(0, "keep")  # a library sync point that collects the profile records
        r = ctx.profile_report()["spmv"]
        ctx.profile(False)
        ws.set_col(1, x)
        ws.matvec(op, 1, 2)
        err = np.abs(ws.get_cols(2, 1)[:, 0] - ref).max() / np.abs(ref).max()
        us = 1e3 * r["ms"] / r["launches"]
        print(cfg, f"{us:.1f} us  {r['bytes'] / r['launches'] / us / 1e3:.0f} GB/s  relerr {err:.1e}", flush=True)
        ws.close()
        op.close()


main()
