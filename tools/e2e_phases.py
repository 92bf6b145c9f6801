"""Where does the end-to-end time go?  Times the phases of bench.py's e2e step separately, on 1 GPU or - under
torchrun - on every rank of a row-sharded job (mean and max over ranks per phase, several repetitions: the N = 8
e2e number of round 1 swung 4x between runs).

    python tools/e2e_phases.py
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tools/e2e_phases.py
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import bench
import b200arnoldi as b2a
from arnoldimethod_jl_b200 import _lib as L
from arnoldimethod_jl_b200.api import _run

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = b2a.Context.from_torch_distributed(local) if world > 1 else b2a.Context(local)

n_loc = bench.N_PER_GPU
n = n_loc * world
off = rank * n_loc
indptr, indices, data = bench.make_shard(n, off, n_loc)
v1 = bench.make_v1(n, off, n_loc)
(indptr_p, _a), (indices_p, _b), (data_p, _c), (v1_p, _d) = map(bench.pinned_like, (indptr, indices, data, v1))
q_out, _e = bench.pinned_like(np.zeros((n_loc, bench.NEV + 1), order="F").T)
q_out = q_out.T
PHASES = ("upload_A", "ws_create", "upload_v1", "solve", "download_Q", "free")


def step(timers):
    t = time.perf_counter()
    op = b2a.Operator.from_csr_arrays(ctx, indptr_p, indices_p, data_p, n, row_offset=off)
    ctx.synchronize(); t1 = time.perf_counter(); timers["upload_A"] += t1 - t
    ws = b2a.ArnoldiWorkspace(n_loc, bench.MAXDIM, ctx=ctx, n_global=n, row_offset=off)
    ctx.synchronize(); t2 = time.perf_counter(); timers["ws_create"] += t2 - t1
    ws.set_col(1, v1_p)
    t3 = time.perf_counter(); timers["upload_v1"] += t3 - t2
    P, h = _run(ws, op, bench.NEV, bench.WHICH, bench.TOL, bench.MINDIM, bench.MAXDIM, 200, 1, L.INIT_KEEP, 0)
    t4 = time.perf_counter(); timers["solve"] += t4 - t3
    Q = ws.get_cols(1, h.nconverged, out=q_out)
    t5 = time.perf_counter(); timers["download_Q"] += t5 - t4
    ws.close(); op.close()
    ctx.synchronize(); t6 = time.perf_counter(); timers["free"] += t6 - t5


for rep in range(3):
    timers = {k: 0.0 for k in PHASES}
    N = 5
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(N):
        step(timers)
    total = time.perf_counter() - t0
    vec = torch.tensor([timers[k] / N * 1e3 for k in PHASES] + [total / N * 1e3], dtype=torch.float64, device="cuda")
    mx, mean = vec.clone(), vec.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(mean)
        mean /= world
    if rank == 0:
        names = PHASES + ("total",)
        print(json.dumps(dict(rep=rep, world=world,
                              mean_ms={k: round(float(v), 2) for k, v in zip(names, mean.cpu())},
                              max_ms={k: round(float(v), 2) for k, v in zip(names, mx.cpu())})), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
