"""BASELINE configs 4 and 5 at their TRUE size and GPU count, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
        tools/cfg_dist_bench.py cfg4            # ComplexF64, n = 5e6, 20 nnz/row, nev 30, maxdim 60, :LM
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
        tools/cfg_dist_bench.py cfg5            # Float64, n = 1e8, 15 nnz/row (1.5e9 nnz), nev 20, maxdim 40, :LM
    ... cfg5 --n 2e7 --restarts 3              # scaled / capped variants; also runs on 1 GPU without torchrun

Every rank generates only its own row block (uniformly random GLOBAL columns, seeded per rank; the designed
top spectrum d_i = 5 + 20 * 0.9^i on the first rows so that the solve converges), so no host ever holds the
whole matrix.  Reports (rank 0, one JSON line): Arnoldi steps/s of the whole job, the algorithmic HBM bytes of
SURVEY 8(d) per second summed over the GPUs and as a fraction of N x the measured copy peak, per-kernel
CUDA-event times of rank 0, how the collectives ran, and - with --residual - the Frobenius residual
||A Q - Q R|| of the converged Schur vectors, computed with one extra distributed mat-vec per vector
(`b2a_ws_matvec`, the x exchange included) and an all-reduce of the local row blocks' squared norms.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

CONFIGS = {
    "cfg4": dict(n=5_000_000, k=20, T=np.complex128, nev=30, mindim=30, maxdim=60, scale=0.35),
    "cfg5": dict(n=100_000_000, k=15, T=np.float64, nev=20, mindim=20, maxdim=40, scale=0.5),
    "cfg2": dict(n=1_000_000, k=16, T=np.float64, nev=20, mindim=20, maxdim=40, scale=0.5),
}


def make_block(cfg, n, off, cnt, rank):
    """Rows [off, off + cnt) as CSR arrays with global column numbers (sorted within each row)."""
    rng = np.random.default_rng([7, rank])
    k, T = cfg["k"], cfg["T"]
    cdt = np.int32 if n < 2 ** 31 else np.int64
    cols = rng.integers(0, n, size=(cnt, k), dtype=cdt)
    vals = rng.standard_normal((cnt, k)) * cfg["scale"]
    if T is np.complex128:
        vals = vals + 1j * rng.standard_normal((cnt, k)) * cfg["scale"]
    ntop = cfg["maxdim"]
    rows = np.arange(off, off + cnt)
    top = rows < ntop
    if top.any():
        d = 5.0 + 20.0 * 0.9 ** rows[top]
        if T is np.complex128:
            d = d * np.exp(1j * np.linspace(0.0, 1.0, ntop)[rows[top]])
        cols[top, 0] = rows[top]
        vals[top, 0] = d
    order = np.argsort(cols, axis=1, kind="stable")
    cols = np.take_along_axis(cols, order, axis=1)
    vals = np.take_along_axis(vals, order, axis=1)
    indptr = np.arange(0, (cnt + 1) * k, k, dtype=np.int64)
    return indptr, cols.ravel(), np.ascontiguousarray(vals.ravel().astype(T))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=sorted(CONFIGS))
    ap.add_argument("--n", type=float, default=None, help="override the matrix order")
    ap.add_argument("--restarts", type=int, default=4, help="restart cap of the measured solve (200 = to convergence)")
    ap.add_argument("--tol", type=float, default=1e-6)
    ap.add_argument("--residual", action="store_true")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    import b200arnoldi as b2a
    from arnoldimethod_jl_b200 import _lib as L
    from arnoldimethod_jl_b200.api import _run

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = b2a.Context.from_torch_distributed(local) if world > 1 else b2a.Context(local)
    cfg = dict(CONFIGS[args.config])
    n = int(args.n) if args.n else cfg["n"]
    T = cfg["T"]
    offs, cnts = b2a.sharding.row_partition(n, world)
    off, cnt = int(offs[rank]), int(cnts[rank])

    t0 = time.perf_counter()
    indptr, cols, vals = make_block(cfg, n, off, cnt, rank)
    t_gen = time.perf_counter() - t0
    op = b2a.Operator.from_csr_arrays(ctx, indptr, cols, vals, n, row_offset=off)
    ws = b2a.ArnoldiWorkspace(cnt, cfg["maxdim"], dtype=T, ctx=ctx, n_global=n, row_offset=off)
    v1 = np.random.default_rng([8, rank]).random(cnt).astype(T)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))

    def barrier():
        ctx.synchronize()
        if world > 1:
            dist.barrier()

    def solve(restarts):
        ws.set_col(1, v1)
        return _run(ws, op, cfg["nev"], "LM", args.tol, cfg["mindim"], cfg["maxdim"], restarts, 1, L.INIT_KEEP, 0)

    solve(1)  # warm-up: pools, tensor maps, peer block
    barrier()
    ctx.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    P, hist = solve(args.restarts)
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    prof = ctx.profile_report()
    ctx.profile(False)

    out = None
    if rank == 0:
        peak = 6543.7
        try:
            peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
        gbs = hist.stats["bytes"] * world / (ms * 1e-3) / 1e9  # every rank executes the same byte model
        out = dict(config=args.config, n=n, nnz=int(n) * cfg["k"], gpus=world, dtype=T.__name__, restarts=hist.restarts,
                   mvproducts=hist.mvproducts, nconverged=hist.nconverged, converged=bool(hist.converged),
                   second_pass_rate=round(hist.stats["second_passes"] / max(1, hist.mvproducts), 3),
                   ms=round(ms, 2), steps_per_s=round(hist.mvproducts / (ms * 1e-3), 1),
                   algorithmic_GBs_aggregate=round(gbs, 1), frac_of_aggregate_hbm_peak=round(gbs / (peak * world), 4),
                   collectives=ws.comm_mode, gen_s=round(t_gen, 1),
                   env={k: v for k, v in os.environ.items() if k.startswith("B2A_")},
                   nvlink_bytes_sent_per_step_per_gpu=(world - 1) * cnt * np.dtype(T).itemsize,
                   kernels_rank0={k: dict(launches=r["launches"], avg_us=round(1e3 * r["ms"] / r["launches"], 1),
                                          gbs=round(r["bytes"] / (r["ms"] * 1e-3) / 1e9))
                                  for k, r in prof.items() if r["launches"]})

    if args.residual and hist.nconverged:
        # E = A Q - Q R on this rank's rows (one extra distributed mat-vec per Schur vector, x exchange included).
        # Per eigenpair (lambda, y) of R with ||y|| = 1 the eigenvector is x = Q y and  A x - lambda x = E y  exactly,
        # so ||A x - lambda x|| / |lambda| (the reference's convergence measure, src/run.jl:192-208) needs no further
        # mat-vec: squared norms of the row blocks of E y are summed over the ranks.
        nc, spare = hist.nconverged, cfg["maxdim"] + 1
        Q = ws.get_cols(1, nc)  # local row block of the Schur vectors
        R = P.R
        E = np.empty((cnt, nc), dtype=Q.dtype, order="F")
        for i in range(nc):
            ws.matvec(op, i + 1, spare)
            E[:, i] = ws.get_cols(spare, 1)[:, 0] - Q @ R[:, i]
        lam, Y = np.linalg.eig(R)
        Y = Y / np.linalg.norm(Y, axis=0)
        EY = E @ Y
        sq = np.concatenate([[np.linalg.norm(E) ** 2], np.sum(np.abs(EY) ** 2, axis=0),
                             [np.linalg.norm(Q.conj().T @ Q) ** 2 if world == 1 else 0.0]])
        t = torch.from_numpy(sq).cuda()
        if world > 1:
            dist.all_reduce(t)
        t = t.cpu().numpy()
        if rank == 0:
            pair = np.sqrt(t[1 : 1 + nc]) / np.abs(lam)
            out["residual_AQ_QR"] = float(np.sqrt(t[0]))
            out["residual_bound_n_tol"] = n * args.tol
            out["eigenpair_residual_rel_max"] = float(pair.max())
            out["eigenpair_residual_rel"] = [float(f"{v:.3e}") for v in pair]
            out["eigenvalues_abs"] = [float(f"{abs(v):.6f}") for v in lam]
            out["tol"] = args.tol
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
