"""Where does the end-to-end time go?  Times the phases of bench.py's e2e step separately."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
import b200arnoldi as b2a
from arnoldimethod_jl_b200 import _lib as L
from arnoldimethod_jl_b200.api import _run

n = bench.N_PER_GPU
indptr, indices, data = bench.make_shard(n, 0, n)
v1 = bench.make_v1(n, 0, n)
(indptr_p, _a), (indices_p, _b), (data_p, _c), (v1_p, _d) = map(bench.pinned_like, (indptr, indices, data, v1))
q_out, _e = bench.pinned_like(np.zeros((n, bench.NEV + 1), order="F").T)
q_out = q_out.T
ctx = b2a.Context(0)


def step(timers):
    t = time.perf_counter()
    op = b2a.Operator.from_csr_arrays(ctx, indptr_p, indices_p, data_p, n)
    ctx.synchronize(); t1 = time.perf_counter(); timers["upload_A"] += t1 - t
    ws = b2a.ArnoldiWorkspace(n, bench.MAXDIM, ctx=ctx, n_global=n, row_offset=0)
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
    timers = dict(upload_A=0.0, ws_create=0.0, upload_v1=0.0, solve=0.0, download_Q=0.0, free=0.0)
    N = 5
    for _ in range(N):
        step(timers)
    print({k: round(1e3 * v / N, 2) for k, v in timers.items()}, "total ms", round(1e3 * sum(timers.values()) / N, 2), flush=True)
