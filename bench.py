#!/usr/bin/env python
"""bench.py - Arnoldi matvec+orthog steps/sec on BASELINE.json's headline config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1], SURVEY 8(d) cfg 2): CSR Float64, n = 1e6 rows per GPU,
exactly 16 nnz/row at uniformly random columns, values N(0,1)*0.5, plus the designed top
spectrum d_i = 5 + 20*0.9^i on the first 40 diagonal entries so that
partialschur(A; nev=20, mindim=20, maxdim=40, which=:LM, tol=1e-6) CONVERGES; synthetic, seeded.

A bench "step" = one complete partialschur call (Arnoldi expansion + Krylov-Schur restarts to
convergence).  `value` = Arnoldi steps (History.mvproducts = matvec + orthogonalisation) per
second with A and v1 already resident in HBM; `e2e` = the same through the public API from
pinned HOST CSR arrays: upload of A and v1, solve, download of the Schur vectors Q, R and the
eigenvalues, all inside the timed region.  N > 1: one process per GPU (torchrun), rows sharded
in contiguous blocks of 1e6 (weak scaling: the matrix order grows with N), NCCL all-reduces
of h / norms and an all-gather of x per step; `value` is then (global steps/s) x N, i.e.
1e6-row shard-steps per second summed over GPUs.

--impl reference: the reference's own CPU path.  Julia is not in the image, so this is the
oracle port (NumPy/SciPy restatement, see oracle/__init__.py) on the host cores: SciPy CSR
mat-vec (single-threaded like Julia's SparseArrays.mul!) + OpenBLAS gemv/gemm/nrm2 on all
cores.  Each step is the SAME complete solve as the GPU arm's step (same matrix, start vector, tolerance; to
convergence: 81 Arnoldi steps, 5 restarts, about 6 s per step on 16 cores), so both arms share one `config`;
warm-up steps are capped at one restart (untimed).  At N > 1 the CPU arm still solves ONE shard-sized problem: `value`
counts shard-steps (a step on an N-shard matrix = N shard-steps, CPU time per step being linear in n).
"""

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_PER_GPU = 1_000_000
NNZ_PER_ROW = 16
NEV, MINDIM, MAXDIM, TOL, WHICH = 20, 20, 40, 1e-6, "LM"
NTOP = 40
METRIC = "Arnoldi matvec+orthog steps/sec (partialschur, random CSR Float64 n=1e6/GPU, nnz=16/row, nev=20, maxdim=40)"
UNIT = "steps/s"


# ------------------------------------------------------------------------------ workload
def make_shard(n_global, row_offset, n_local, seed=0):
    """Rows [row_offset, row_offset + n_local) of the synthetic matrix, as CSR arrays (0-based,
    int64 rowptr, int32 global colind) - generated per row block so that every rank builds
    only its own shard and the matrix does not depend on the GPU count per block."""
    rng = np.random.default_rng([seed, row_offset // N_PER_GPU])
    nnz_r = NNZ_PER_ROW
    cols = rng.integers(0, n_global, size=(n_local, nnz_r), dtype=np.int64)
    vals = rng.standard_normal((n_local, nnz_r)) * 0.5
    rows = np.arange(row_offset, row_offset + n_local, dtype=np.int64)
    # designed diagonal (first NTOP global rows): put it in slot 0 of the row
    top = rows < NTOP
    cols[top, 0] = rows[top]
    vals[top, 0] = 5.0 + 20.0 * 0.9 ** rows[top]
    order = np.argsort(cols, axis=1, kind="stable")
    cols = np.take_along_axis(cols, order, axis=1)
    vals = np.take_along_axis(vals, order, axis=1)
    indptr = np.arange(0, (n_local + 1) * nnz_r, nnz_r, dtype=np.int64)
    return indptr, cols.astype(np.int32).ravel(), vals.ravel()


def make_v1(n_global, row_offset, n_local, seed=1):
    rng = np.random.default_rng([seed, row_offset // N_PER_GPU])
    return rng.random(n_local)


NCU_TRAFFIC_FILE = os.path.join("profiles", "ncu_traffic.json")


def ncu_traffic_table():
    """{kernel kind: dram bytes / algorithmic bytes} from the committed summary of the `ncu --set full` captures
    (profiles/ncu_traffic.json names the .txt summaries it was read from).  bench.py cannot run ncu inside the timed
    process, so `roofline.traffic` = this measured ratio x the algorithmic bytes per launch measured live."""
    try:
        return json.load(open(os.path.join(ROOT, NCU_TRAFFIC_FILE)))
    except Exception:
        return {}


LIMITER_NOTES = {
    "spmv": "uniformly random columns: every 8-byte gather of x is its own 32-byte sector; ncu "
            "(profiles/r2_ncu_spmv_summary.txt): 24.6 M L1TEX sectors / 148 SMs / 169 K cycles = 0.98 per SM-cycle - the "
            "L1TEX tag stage (1 sector/clk) is saturated; L2 at 69 %, DRAM traffic == algorithmic bytes (30 %). "
            "HBM-bound only on structured matrices (7-pt stencil: 5.15 TB/s = 79 % of peak).",
    "cgs_sweep": "HBM stream of the Krylov panel through a TMA ring, 2-3 passes per launch (dots, update + "
                 "speculative dots, gated second update, normalisation fused in one persistent kernel); the passes "
                 "run at 6.0-6.5 TB/s, the rest is 2-3 in-kernel grid barriers (~4 us each) and the launch",
    "cgs_update": "HBM stream of the Krylov panel through a TMA ring (fused update + speculative dots)",
    "cgs_dots": "HBM stream of the Krylov panel through a TMA ring",
    "rotate": "in-place V <- V Q: TMA ring + FP64 tensor pipe (DMMA.8x8x4), HBM / FP64 ridge",
    "xchg": "NVLink: staged copy-engine exchange of x, hidden behind the owner-blocked mat-vec",
}


# --------------------------------------------------------------------------- clock sampler
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []
        self.marks = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.t = threading.Thread(target=self._pump, daemon=True)
        self.t.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        sm, smmax, reasons = [], [], set()
        for t, line in self.lines:
            if not (t0 <= t <= t1 + 0.15):
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smmax.append(float(parts[2]))
            except ValueError:
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smmax)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------ reference
def oracle_run(indptr, indices, data, n, v1, restarts):
    import scipy.sparse as sp

    import oracle

    A = sp.csr_matrix((data, indices, indptr), shape=(n, n))
    t0 = time.perf_counter()
    P, hist = oracle.partialschur(A, v1=v1, nev=NEV, mindim=MINDIM, maxdim=MAXDIM, which=WHICH, tol=TOL,
                                  restarts=restarts)
    dt = time.perf_counter() - t0
    return hist, dt


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = N_PER_GPU  # bounded sample: one GPU's shard-sized problem, restarts capped
    indptr, indices, data = make_shard(n, 0, n)
    v1 = make_v1(n, 0, n)
    restarts = 200  # to convergence, exactly the GPU arm's step (5 restarts, 81 Arnoldi steps)
    for _ in range(args.warmup):
        oracle_run(indptr, indices, data, n, v1, 1)  # untimed warm-up: one restart is enough to page everything in
    t_tot, mv_tot, timers = 0.0, 0, {}
    for _ in range(args.steps):
        hist, dt = oracle_run(indptr, indices, data, n, v1, restarts)
        t_tot += dt
        mv_tot += hist.mvproducts
        for k, v in hist.timers.items():
            timers[k] = timers.get(k, 0.0) + v
    value = mv_tot / t_tot
    cores = host_threads()
    sample = (f"same matrix (n=1e6, 16 nnz/row, designed top spectrum), start vector and tolerance as the GPU arm; each "
              f"step = the complete partialschur solve to convergence ({mv_tot // max(args.steps, 1)} Arnoldi steps, "
              f"{hist.restarts_done} restarts); SciPy CSR matvec 1 thread (as Julia SparseArrays.mul!), OpenBLAS "
              f"gemv/gemm on {cores} threads; at N > 1 still ONE shard-sized problem (value counts shard-steps)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(max(1, args.gpus), mv=mv_tot // max(args.steps, 1), restarts=hist.restarts_done),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "phase_seconds": {k: round(v, 3) for k, v in timers.items()}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(n_gpus, restarts_cap=None, mv=None, restarts=None):
    cfg = {
        "workload": "BASELINE cfg2: random CSR Float64, n=1e6 rows/GPU, 16 nnz/row, designed top spectrum "
                    "(d_i = 5+20*0.9^i, i<40); partialschur nev=20 mindim=20 maxdim=40 which=LM tol=1e-6",
        "mode": "designed-spectrum solve to CONVERGENCE (restarts cap 200; SURVEY 8(d)'s parity matrix), not the "
                "fixed-10-restart run on the pure-random matrix: the per-step work is the same, and the solve can be "
                "checked (converged, residual)",
        "restarts_cap": 200,
        "step_unit": "value = Arnoldi steps/s x n_gpus (1e6-row shard-steps/s); the CPU arm solves one shard-sized "
                     "problem at every N",
        "n_global": N_PER_GPU * n_gpus, "nnz_global": N_PER_GPU * n_gpus * NNZ_PER_ROW,
        "nev": NEV, "mindim": MINDIM, "maxdim": MAXDIM, "which": WHICH, "tol": TOL,
        "parallelism": f"row-sharded x{n_gpus}" if n_gpus > 1 else "single GPU",
        "l2": "inputs larger than L2: Krylov panel 328 MB + CSR 216 MB per GPU vs 126 MB L2",
    }
    if restarts_cap is not None:
        cfg["restarts_cap"] = restarts_cap
    if mv is not None:
        cfg["mvproducts_per_step"] = mv
        cfg["restarts_per_step"] = restarts
    return cfg


# ------------------------------------------------------------------------------- GPU arm
def pinned_like(a):
    import torch

    t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, pin_memory=True)
    out = t.numpy()
    out[...] = a
    return out, t


def run_gpu(args):
    import torch
    import torch.distributed as dist

    import b200arnoldi as b2a

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = b2a.Context.from_torch_distributed(local_rank) if world > 1 else b2a.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    n_global = N_PER_GPU * world
    off, n_local = rank * N_PER_GPU, N_PER_GPU
    indptr, indices, data = make_shard(n_global, off, n_local)
    v1 = make_v1(n_global, off, n_local)
    (indptr_p, _k1), (indices_p, _k2), (data_p, _k3), (v1_p, _k4) = map(pinned_like, (indptr, indices, data, v1))
    q_out, _k5 = pinned_like(np.zeros((n_local, NEV + 1), order="F").T)  # pinned, then viewed column-major
    q_out = q_out.T

    # ---------------- resident arm: A and v1 in HBM before the timed region
    op = b2a.Operator.from_csr_arrays(ctx, indptr_p, indices_p, data_p, n_global, row_offset=off)
    ws = b2a.ArnoldiWorkspace(n_local, MAXDIM, ctx=ctx, n_global=n_global, row_offset=off)

    d_v1 = torch.from_numpy(v1).to(torch.device("cuda", local_rank))  # resident copy of the start vector

    def resident_step():
        # partialschur!(A, arnoldi; start_from=1, initialize=false-like "keep"): column 1 <- v1 (device to
        # device, the run normalises it in place), then the whole restart loop inside the library
        from arnoldimethod_jl_b200 import _lib as L
        from arnoldimethod_jl_b200.api import _run

        L.check(L.lib().b2a_ws_set_col_device(ws._h, 1, d_v1.data_ptr()))
        return _run(ws, op, NEV, WHICH, TOL, MINDIM, MAXDIM, 200, 1, L.INIT_KEEP, 0)

    for _ in range(args.warmup):
        P, hist = resident_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    l0 = ctx.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.perf_counter()
    ev0.record(stream)
    mv = 0
    algo_bytes = 0.0
    for _ in range(args.steps):
        P, hist = resident_step()
        mv += hist.mvproducts
        algo_bytes += hist.stats["bytes"]
    ev1.record(stream)
    barrier()
    t_wall1 = time.perf_counter()
    ms_dev = ev0.elapsed_time(ev1)
    launches = ctx.launches - l0
    if rank == 0:
        time.sleep(0.2)
        sampler.stop()
    t = torch.tensor([ms_dev], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    steps_per_s_global = mv / (ms_total * 1e-3)
    value = steps_per_s_global * world
    hbm_gbs = algo_bytes * world / (ms_total * 1e-3) / 1e9  # algorithmic bytes (per-rank model x N) / time

    # correctness of what was timed: converged, residual ||A x - lambda x|| <= tol |lambda| on this shard's rows
    last_hist = hist
    converged = bool(hist.converged)

    # ---------------- per-kernel pass (CUDA events around every launch) -> roofline of the top kernel
    ctx.profile(True)
    resident_step()
    ctx.synchronize()
    prof = ctx.profile_report()
    ctx.profile(False)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "B200_PROFILING.md fallback 6650 GB/s (of fallback)"
    kern = {}
    for name, r in prof.items():
        if r["launches"]:
            kern[name] = {"launches": r["launches"], "ms": round(r["ms"], 4),
                          "avg_us": round(1e3 * r["ms"] / r["launches"], 2),
                          "algo_bytes_per_launch": r["bytes"] / r["launches"],
                          "gbs": round(r["bytes"] / (r["ms"] * 1e-3) / 1e9, 1) if r["ms"] > 0 else None}
    top = max(kern, key=lambda k: kern[k]["ms"]) if kern else None
    total_ms = sum(k["ms"] for k in kern.values()) or 1.0
    roofline = None
    if top:
        ach = kern[top]["gbs"]
        roofline = {"bound": "hbm", "kernel": top, "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 4), "traffic": None, "peak_source": peak_src,
                    "share_of_kernel_time": round(kern[top]["ms"] / total_ms, 3),
                    "algo_bytes_per_launch": kern[top]["algo_bytes_per_launch"],
                    "avg_launch_us": kern[top]["avg_us"], "kernels": kern}
        tt = ncu_traffic_table().get(top)
        if tt:
            roofline["traffic"] = round(tt["dram_over_algorithmic"] * kern[top]["algo_bytes_per_launch"], 1)
            roofline["traffic_source"] = (f"{NCU_TRAFFIC_FILE}: dram__bytes_read+write / algorithmic bytes = "
                                          f"{tt['dram_over_algorithmic']} ({tt['source']}) x the algorithmic bytes "
                                          f"per launch measured in this run")
        roofline["limiter"] = LIMITER_NOTES.get(top)
        if world > 1 and top == "spmv":
            roofline["limiter"] = (
                "row-sharded mat-vec: the launch(es) wait for the slices of x they gather from (staged copy-engine "
                "exchange over NVLink, 'xchg' in roofline.kernels = its true duration on this rank), so the time "
                "includes that wait; the gathers themselves run at the L1TEX sector rate as on one GPU. " + LIMITER_NOTES["spmv"])
        # context: the HBM-streaming Gram-Schmidt sweeps (dots + update) taken together, and all kernels
        gs = [prof[k] for k in ("cgs_dots", "cgs_update", "cgs_sweep") if k in prof and prof[k]["launches"]]
        if gs:
            gs_ms, gs_bytes = sum(r["ms"] for r in gs), sum(r["bytes"] for r in gs)
            roofline["gram_schmidt"] = {"achieved": round(gs_bytes / (gs_ms * 1e-3) / 1e9, 1), "unit": "GB/s",
                                        "frac": round(gs_bytes / (gs_ms * 1e-3) / 1e9 / peak, 4),
                                        "share_of_kernel_time": round(gs_ms / total_ms, 3)}
        roofline["all_kernels_gbs"] = round(sum(r["bytes"] for r in prof.values()) / (total_ms * 1e-3) / 1e9, 1)

    # the resident workspace / operator are done: release them (at N > 1 this also hands the context's NVLink
    # peer block to the e2e workspaces)
    ws.close()
    op.close()

    if os.environ.get("B2A_BENCH_QUICK") == "1":
        # variant sweeps on multi-GPU leases (tools/gpu_r2_multi8.sh): device-timed arm only, one compact line
        if rank == 0:
            print(json.dumps({"quick": True, "n_gpus": world, "ms_per_step": ms_total / max(args.steps, 1),
                              "value": value, "mvproducts": mv // max(args.steps, 1), "converged": converged,
                              "hbm_frac_aggregate": round(hbm_gbs / (peak * world), 4),
                              "kernels_us": {k: (v["launches"], v["avg_us"]) for k, v in kern.items()},
                              "env": {k: v for k, v in os.environ.items() if k.startswith("B2A_")}}), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---------------- e2e arm: host CSR in pinned memory -> upload -> solve -> download Q, R, eigenvalues
    def e2e_step():
        op2 = b2a.Operator.from_csr_arrays(ctx, indptr_p, indices_p, data_p, n_global, row_offset=off)
        ws2 = b2a.ArnoldiWorkspace(n_local, MAXDIM, ctx=ctx, n_global=n_global, row_offset=off)
        ws2.set_col(1, v1_p)
        from arnoldimethod_jl_b200 import _lib as L
        from arnoldimethod_jl_b200.api import _run

        P2, h2 = _run(ws2, op2, NEV, WHICH, TOL, MINDIM, MAXDIM, 200, 1, L.INIT_KEEP, 0)
        Q = ws2.get_cols(1, h2.nconverged, out=q_out)
        R, lam = P2.R, P2.eigenvalues
        ws2.close()
        op2.close()
        return h2, Q, R, lam

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    mv2 = 0
    for _ in range(e2e_steps):
        h2, Q, R, lam = e2e_step()
        mv2 += h2.mvproducts
    barrier()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = mv2 / float(te.item()) * world
    h2d = indptr_p.nbytes + indices_p.nbytes + data_p.nbytes + v1_p.nbytes
    d2h = int(Q.nbytes + R.nbytes + lam.nbytes)

    # residual ||A Q - Q R||_F of the e2e result with an INDEPENDENT SciPy mat-vec, off the clock.  N > 1: every rank
    # gathers the Schur vectors (NCCL all-gather of the row blocks), multiplies its own row shard of A on the host and
    # the squared norms of the row blocks are summed over the ranks.
    import scipy.sparse as sp

    Qc = np.ascontiguousarray(np.array(Q))
    if world > 1:
        parts = [torch.empty((n_local, Qc.shape[1]), dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(Qc).cuda())
        Qfull = torch.cat(parts).cpu().numpy()
        del parts
    else:
        Qfull = Qc
    A_loc = sp.csr_matrix((data, indices, indptr), shape=(n_local, n_global))
    sq = torch.tensor([float(np.linalg.norm(A_loc @ Qfull - Qc @ R) ** 2),
                       float(np.linalg.norm(Qfull.T @ Qfull - np.eye(Qfull.shape[1])) ** 2)],
                      dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(sq)
        sq[1] /= world  # the orthogonality defect was computed redundantly by every rank
    resid = float(np.sqrt(sq[0].item()))
    ortho = float(np.sqrt(sq[1].item()))
    del Qfull

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        restarts = 2
        hist_o, dt = oracle_run(indptr, indices, data, n_global, v1, restarts)
        cores = host_threads()
        cpu_baseline = {
            "value": hist_o.mvproducts / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle port (Julia absent): same matrix and v1, partialschur with restarts capped at "
                      f"{restarts} = {hist_o.mvproducts} Arnoldi steps in {dt:.1f} s; SciPy CSR matvec 1 thread "
                      f"(as Julia SparseArrays.mul!), OpenBLAS gemv/gemm on {cores} threads",
            "phase_seconds": {k: round(v, 3) for k, v in hist_o.timers.items()},
        }

    if rank == 0:
        clocks = sampler.summary(t_wall0, t_wall1)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world, mv=mv // max(args.steps, 1), restarts=last_hist.restarts),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": 1e3 * float(te.item()) / e2e_steps},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "global_steps_per_s": steps_per_s_global,
            "hbm_gbs_aggregate_algorithmic": round(hbm_gbs, 1),
            "hbm_frac_aggregate": round(hbm_gbs / (peak * world), 4),
            "converged": converged, "nconverged": int(last_hist.nconverged),
            "second_pass_rate": round(last_hist.stats["second_passes"] / max(last_hist.mvproducts, 1), 3),
            "residual_AQ_QR": resid, "residual_bound_n_tol": n_global * TOL, "orthogonality_QtQ_I": ortho,
            "host_ms_per_step": {k: round(v, 3) for k, v in last_hist.timers_ms.items()},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
