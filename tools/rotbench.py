"""Basis-rotation microbenchmark (run.jl:363-365 shapes), CUDA events through the library's own profiler.

    python tools/rotbench.py [--json out.json]

Shapes: BASELINE cfg 2 first restart (n=1e6, K=40 -> N=26) and a later one (K=29 -> N=19), cfg 3-like
(n=1.6e7, K=20 -> N=15), cfg 4 ComplexF64 (n=2e6, K=60 -> N=45), and a wide basis (maxdim 200).  Both kernels:
B2A_ROTATE=1 (TMA + DMMA) and B2A_ROTATE=0 (shared-memory DFMA).  HBM bound = n s (K + N + 2) / measured copy
peak; FLOP bound = 2 n K N (x4 complex) / measured DMMA peak (tools/fp64_peak.bin) when given with --fp64-peak.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import b200arnoldi as b2a

SHAPES = [
    ("cfg2 first restart", np.float64, 1_000_000, 40, 1, 26),
    ("cfg2 later restart", np.float64, 1_000_000, 40, 12, 30),
    ("cfg3-like 256^3", np.float64, 16_777_216, 20, 1, 15),
    ("cfg4 complex", np.complex128, 2_000_000, 60, 1, 45),
    ("wide basis maxdim 200", np.float64, 400_000, 200, 1, 150),
]


def run(T, n, maxdim, purge, k, mode, reps=10):
    os.environ["B2A_ROTATE"] = str(mode)
    ctx = b2a.default_context()
    rng = np.random.default_rng(0)
    ws = b2a.ArnoldiWorkspace(n, maxdim, dtype=T, ctx=ctx)
    x = rng.standard_normal(n).astype(T)
    for c in range(maxdim + 1):
        ws.set_col(c + 1, np.roll(x, c))
    Q = np.linalg.qr(rng.standard_normal((maxdim, maxdim)))[0].astype(T)
    for it in range(reps + 2):
        if it == 2:
            ctx.profile(True)
        ws.rotate_basis(purge, k, maxdim, Q)
    rep = ctx.profile_report()["rotate"]
    ctx.profile(False)
    ws.close()
    return 1e3 * rep["ms"] / rep["launches"], rep["bytes"] / rep["launches"]


if __name__ == "__main__":
    peak = 6543.7
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    fp64 = None
    if "--fp64-peak" in sys.argv:
        fp64 = float(sys.argv[sys.argv.index("--fp64-peak") + 1])
    rows = []
    if "--only" in sys.argv:
        SHAPES[:] = [SHAPES[int(sys.argv[sys.argv.index("--only") + 1])]]
    for name, T, n, maxdim, purge, k in SHAPES:
        K, N = maxdim - purge + 1, k - purge + 1
        flops = 2.0 * n * K * N * (4 if T is np.complex128 else 1)
        for mode in (1, 0):
            try:
                us, by = run(T, n, maxdim, purge, k, mode)
            except Exception as e:  # the DFMA kernels cannot take wide bases
                rows.append(dict(shape=name, kernel="dmma" if mode else "dfma", error=str(e)[:120]))
                continue
            r = dict(shape=name, dtype=np.dtype(T).name, n=n, K=K, N=N, kernel="dmma" if mode else "dfma",
                     us=round(us, 1), algo_GB=round(by / 1e9, 3), gbs=round(by / us / 1e3, 1),
                     frac_hbm=round(by / us / 1e3 / peak, 3), hbm_bound_us=round(by / peak / 1e3, 1),
                     tflops=round(flops / us / 1e6, 2))
            if fp64:
                r["flop_bound_us"] = round(flops / fp64 / 1e6, 1)
            rows.append(r)
            print(r, flush=True)
    if "--json" in sys.argv:
        json.dump(rows, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
