"""Multi-GPU parity check, launched one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_gpu_check.py

Row-sharded partialschur over NCCL must give the same H / eigenvalues as the oracle run on the
whole matrix (shard-count invariance), for Float64 and ComplexF64.  Prints DIST_GPU_CHECK_OK.

    ... tests/dist_gpu_check.py --big [rows_per_gpu]

BASELINE-size variant (default 1e6 rows per GPU, bench.py's cfg-2 matrix, any world size): the first sweep's H
against the oracle (<= 1e-13 relative), then the complete solve: `mvproducts` within one restart of the oracle's,
eigenvalues <= 10 tol |lambda|, ||A Q - Q R|| <= n tol with an independent SciPy mat-vec of the gathered Schur
vectors, ||Q'Q - I||.  Rank 0 builds the whole matrix for the oracle; the GPUs only ever see their own shard.
Prints one JSON line and DIST_GPU_BIG_OK."""

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

import b200arnoldi as b2a
import oracle


def big(ctx, rank, world, rows_per_gpu):
    import json
    import time

    import bench
    from arnoldimethod_jl_b200 import _lib as L
    from arnoldimethod_jl_b200.api import _run

    bench.N_PER_GPU = rows_per_gpu  # make_shard keys its generator by row block
    n = rows_per_gpu * world
    off = rank * rows_per_gpu
    indptr, indices, data = bench.make_shard(n, off, rows_per_gpu)
    v1 = bench.make_v1(n, off, rows_per_gpu)
    op = b2a.Operator.from_csr_arrays(ctx, indptr, indices, data, n, row_offset=off)
    ws = b2a.ArnoldiWorkspace(rows_per_gpu, bench.MAXDIM, ctx=ctx, n_global=n, row_offset=off)
    assert os.environ.get("B2A_NO_PEER") == "1" or ws.comm_mode == "peer", ws.comm_mode
    out = dict(world=world, n=n, rows_per_gpu=rows_per_gpu, collectives=ws.comm_mode,
               env={k: v for k, v in os.environ.items() if k.startswith("B2A_")})

    # --- first sweep: H against the oracle
    steps = 20
    ws.set_col(1, v1)
    ws.reinitialize(0, "keep")
    ws.iterate_arnoldi(op, 1, steps)
    H = np.array(ws.H)[: steps + 1, :steps].copy()
    Vl = ws.get_cols(1, steps + 1)
    dump = sys.argv[sys.argv.index("--dump") + 1] if "--dump" in sys.argv else None
    t0 = time.time()
    if rank == 0 and not dump:
        blocks = [bench.make_shard(n, r * rows_per_gpu, rows_per_gpu) for r in range(world)]
        ip = np.concatenate([[0]] + [b[0][1:] + r * rows_per_gpu * bench.NNZ_PER_ROW for r, b in enumerate(blocks)])
        A = sp.csr_matrix((np.concatenate([b[2] for b in blocks]), np.concatenate([b[1] for b in blocks]), ip), shape=(n, n))
        v1g = np.concatenate([bench.make_v1(n, r * rows_per_gpu, rows_per_gpu) for r in range(world)])
        arn = oracle.ArnoldiWorkspace(np.float64, n, bench.MAXDIM)
        arn.V[:, 0] = v1g / np.linalg.norm(v1g)
        oracle.iterate_arnoldi(A, arn, 1, steps)
        Ho = arn.H[: steps + 1, :steps]
        out["H_relerr_first_sweep"] = float(np.abs(H - Ho).max() / np.abs(Ho).max())
        out["V_abserr_first_sweep_rank0_rows"] = float(np.abs(Vl - arn.V[:rows_per_gpu, : steps + 1]).max())
        assert out["H_relerr_first_sweep"] <= 1e-13, out
        assert out["V_abserr_first_sweep_rank0_rows"] <= 1e-12, out
        del arn

    # --- the complete solve
    ws.set_col(1, v1)
    P, hist = _run(ws, op, bench.NEV, bench.WHICH, bench.TOL, bench.MINDIM, bench.MAXDIM, 200, 1, L.INIT_KEEP, 0)
    Ql = np.ascontiguousarray(ws.get_cols(1, hist.nconverged))
    parts = [torch.empty((rows_per_gpu, Ql.shape[1]), dtype=torch.float64, device="cuda") for _ in range(world)]
    dist.all_gather(parts, torch.from_numpy(Ql).cuda())
    out.update(mvproducts=int(hist.mvproducts), restarts=int(hist.restarts), nconverged=int(hist.nconverged),
               converged=bool(hist.converged))
    if dump:
        # 8-GPU leases are dear: the residual is computed here rank-parallel (every rank multiplies its own row shard
        # on the host), the ORACLE comparison runs afterwards on a CPU box from the dumped small arrays
        # (tools/check_big_dump.py regenerates the matrix from the same seeds).
        Qfull = torch.cat(parts).cpu().numpy()
        A_loc = sp.csr_matrix((data, indices, indptr), shape=(rows_per_gpu, n))
        sq = torch.tensor([float(np.linalg.norm(A_loc @ Qfull - Ql @ P.R) ** 2)], dtype=torch.float64, device="cuda")
        dist.all_reduce(sq)
        if rank == 0:
            out["residual_AQ_QR"] = float(np.sqrt(sq.item()))
            out["residual_bound_n_tol"] = n * bench.TOL
            out["orthogonality"] = float(np.linalg.norm(Qfull.T @ Qfull - np.eye(Qfull.shape[1])))
            np.savez(dump, H_first_sweep=H, V_first_rows=Vl[:2000], eigenvalues=P.eigenvalues, R=P.R,
                     meta=json.dumps(out))
            print(json.dumps(out), flush=True)
            assert hist.converged and out["residual_AQ_QR"] <= n * bench.TOL and out["orthogonality"] < 1e-12, out
            print("DIST_GPU_BIG_DUMPED", flush=True)
        ws.close()
        op.close()
        dist.barrier()
        return
    if rank == 0:
        Q = torch.cat(parts).cpu().numpy()
        out["residual_AQ_QR"] = float(np.linalg.norm(A @ Q - Q @ P.R))
        out["residual_bound_n_tol"] = n * bench.TOL
        out["orthogonality"] = float(np.linalg.norm(Q.T @ Q - np.eye(Q.shape[1])))
        Po, ho = oracle.partialschur(A, v1=v1g, nev=bench.NEV, mindim=bench.MINDIM, maxdim=bench.MAXDIM,
                                     which=bench.WHICH, tol=bench.TOL)
        out["oracle_mvproducts"] = int(ho.mvproducts)
        lam = np.sort_complex(P.eigenvalues)[-bench.NEV:]
        lamo = np.sort_complex(Po.eigenvalues)[-bench.NEV:]
        out["eig_relerr_max"] = float((np.abs(lam - lamo) / np.abs(lamo)).max())
        out["oracle_seconds"] = round(time.time() - t0, 1)
        print(json.dumps(out), flush=True)
        assert hist.converged and ho.converged
        assert abs(hist.mvproducts - ho.mvproducts) <= bench.MAXDIM - bench.MINDIM, out
        assert out["eig_relerr_max"] <= 10 * bench.TOL, out
        assert out["residual_AQ_QR"] <= n * bench.TOL and out["orthogonality"] < 1e-12, out
    ws.close()
    op.close()
    dist.barrier()
    if rank == 0:
        print("DIST_GPU_BIG_OK", flush=True)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = b2a.Context.from_torch_distributed(local)
    if "--big" in sys.argv:
        i = sys.argv.index("--big")
        rows = int(float(sys.argv[i + 1])) if len(sys.argv) > i + 1 else 1_000_000
        big(ctx, rank, world, rows)
        dist.destroy_process_group()
        return

    for T in (np.float64, np.complex128):
        rng = np.random.default_rng(5)
        n, nev = 60001, 6
        vals = rng.standard_normal(n * 8) * 0.3
        if T is np.complex128:
            vals = vals + 0.3j * rng.standard_normal(n * 8)
        A = sp.csr_matrix((vals, rng.integers(0, n, n * 8), np.arange(0, n * 8 + 1, 8)), shape=(n, n))
        d = np.zeros(n, dtype=T)
        d[:12] = 5 + 20 * 0.8 ** np.arange(12)
        A = (A + sp.diags(d)).tocsr().astype(T)
        A.sort_indices()
        v1 = rng.random(n).astype(T)

        # sweep parity: H after 15 Arnoldi steps
        ws = b2a.ArnoldiWorkspace(v1, 20, ctx=ctx)
        if os.environ.get("B2A_NO_PEER") != "1":
            assert ws.comm_mode == "peer", ws.comm_mode  # collectives fused into the kernels over NVLink
        op = b2a.Operator.from_matrix(ctx, A)
        ws.reinitialize(0, "keep")
        ws.iterate_arnoldi(op, 1, 15)
        arn = oracle.ArnoldiWorkspace(T, n, 20)
        arn.V[:, 0] = v1 / np.linalg.norm(v1)
        oracle.iterate_arnoldi(A, arn, 1, 15)
        H = np.array(ws.H)
        err = np.abs(H[:16, :15] - arn.H[:16, :15]).max() / np.abs(arn.H).max()
        assert err < 1e-11, err
        Vl = ws.get_cols(1, 16)
        assert np.abs(Vl - arn.V[ws.row_offset : ws.row_offset + ws.n_local, :16]).max() < 1e-9
        ws.close()

        # whole solve
        P, hist = b2a.partialschur(A, nev=nev, tol=1e-8, which="LM", v1=v1, ctx=ctx)
        Po, ho = oracle.partialschur(A, v1=v1, nev=nev, tol=1e-8, which="LM")
        assert hist.converged and hist.nconverged == ho.nconverged
        assert abs(hist.mvproducts - ho.mvproducts) <= 10
        assert np.allclose(np.sort_complex(P.eigenvalues), np.sort_complex(Po.eigenvalues), atol=1e-7)
        # residual on the local rows: (A Q - Q R)[rows]
        Ql = P.Q
        parts = [torch.zeros((int(c), Ql.shape[1]), dtype=torch.complex128, device="cuda")
                 for c in b2a.sharding.row_partition(n, world)[1]]
        dist.all_gather(parts, torch.from_numpy(np.ascontiguousarray(Ql).astype(np.complex128)).cuda())
        Q = torch.cat(parts).cpu().numpy()
        if T is np.float64:
            Q = Q.real
        res = np.linalg.norm(A @ Q - Q @ P.R)
        assert res < n * 1e-8, res
        assert np.linalg.norm(Q.conj().T @ Q - np.eye(Q.shape[1])) < 1e-12
        if rank == 0:
            print(f"{T.__name__}: world={world} mvproducts={hist.mvproducts} (oracle {ho.mvproducts}) "
                  f"H err={err:.1e} ||AQ-QR||={res:.2e} collectives={P.workspace.comm_mode} "
                  f"fused_sweep={os.environ.get('B2A_FUSED_SWEEP', 'default')}", flush=True)
        # Julia's native CSC layout, row-sharded: every rank passes the whole matrix and keeps its row block
        # (b2a_csc_create mode 0 transposes at upload) - same rows, same order, hence the very same run
        P.workspace.close()
        Pc, hc = b2a.partialschur(A.tocsc(), nev=nev, tol=1e-8, which="LM", v1=v1, ctx=ctx)
        assert hc.mvproducts == hist.mvproducts and hc.nconverged == hist.nconverged
        assert np.array_equal(Pc.eigenvalues, P.eigenvalues) and np.array_equal(Pc.R, P.R)
        Pc.workspace.close()
        # the result keeps its workspace (Q is a view of V): release it so that the next workspace of this
        # context gets the NVLink peer block again (one owner at a time; others fall back to NCCL)
        P.workspace.close()
        op.close()
    # breakdown under sharding (test/expansion.jl:34-55 shape): block-diagonal A, v1 = e1 -> the Krylov space is
    # invariant after 4 steps: H[5,4] == 0 exactly on every rank, the re-seeded column (global-row keyed RNG,
    # identical for every GPU count) keeps V orthonormal, and the sweep resumes
    rng = np.random.default_rng(9)
    nb = 4
    n = 4096
    B = np.zeros((n, n))
    B[:nb, :nb] = rng.random((nb, nb))
    B[nb:, nb:] = np.diag(np.linspace(1, 2, n - nb)) + 0.01 * rng.random((n - nb, n - nb))
    e1 = np.zeros(n)
    e1[0] = 1
    ws = b2a.ArnoldiWorkspace(e1, 8, ctx=ctx)
    op = b2a.Operator.from_matrix(ctx, sp.csr_matrix(B))
    st = ws.iterate_arnoldi(op, 1, 8, seed=3)
    H = np.array(ws.H)
    assert H[4, 3] == 0 and st.breakdowns == 1, (H[4, 3], st.breakdowns)
    Vl = ws.get_cols(1, 9)
    parts = [torch.zeros((int(c), 9), dtype=torch.float64, device="cuda") for c in b2a.sharding.row_partition(n, world)[1]]
    dist.all_gather(parts, torch.from_numpy(np.ascontiguousarray(Vl)).cuda())
    V = torch.cat(parts).cpu().numpy()
    assert np.linalg.norm(V.T @ V - np.eye(9)) < 1e-13
    assert np.linalg.norm(B @ V[:, :8] - V @ H) < 1e-12
    # the re-seeded vector does not depend on the number of GPUs
    if rank == 0:
        np.save("/tmp/b2a_reseed_check.npy", V[:, 4])
        print(f"breakdown: world={world} H[5,4]={H[4, 3]} breakdowns={st.breakdowns} ||V'V-I||={np.linalg.norm(V.T @ V - np.eye(9)):.1e}", flush=True)
    ws.close()

    dist.barrier()
    if rank == 0:
        print("DIST_GPU_CHECK_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
