"""Multi-GPU parity check, launched one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_gpu_check.py

Row-sharded partialschur over NCCL must give the same H / eigenvalues as the oracle run on the
whole matrix (shard-count invariance), for Float64 and ComplexF64.  Prints DIST_GPU_CHECK_OK."""

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

import b200arnoldi as b2a
import oracle


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = b2a.Context.from_torch_distributed(local)

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
