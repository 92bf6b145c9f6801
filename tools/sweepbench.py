"""Wall-clock (CUDA events) of whole Arnoldi sweeps on the bench matrix (cfg 2) under different
orthogonalisation paths, plus the per-phase trace of the fused kernel (B2A_SWEEP_TRACE=1).

    python tools/sweepbench.py            # runs every variant in a subprocess (env is read at load / ws creation)
    python tools/sweepbench.py --one      # one measurement with the current environment
"""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

VARIANTS = [
    ("unfused", {}),
    ("fused", {"B2A_FUSED_SWEEP": "1"}),
    ("fused_pdl1", {"B2A_FUSED_SWEEP": "1", "B2A_SWEEP_PDL": "1"}),
    ("fused_pdl2", {"B2A_FUSED_SWEEP": "1", "B2A_SWEEP_PDL": "2"}),
    ("fused_pdl3", {"B2A_FUSED_SWEEP": "1", "B2A_SWEEP_PDL": "3"}),
    ("fused_rt128", {"B2A_FUSED_SWEEP": "1", "B2A_TMA_RT_UPD": "128"}),
    ("fused_stages3", {"B2A_FUSED_SWEEP": "1", "B2A_TMA_STAGES": "3"}),
]


def one():
    import ctypes as C

    import numpy as np
    import torch

    import b200arnoldi as b2a
    import bench
    from arnoldimethod_jl_b200 import _lib as L

    n, mx = bench.N_PER_GPU, bench.MAXDIM
    indptr, indices, data = bench.make_shard(n, 0, n)
    v1 = bench.make_v1(n, 0, n)
    ctx = b2a.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream)
    op = b2a.Operator.from_csr_arrays(ctx, indptr, indices, data, n)
    ws = b2a.ArnoldiWorkspace(n, mx, ctx=ctx)
    out = {}
    for lo, hi, tag in ((1, mx, "sweep_1_40"), (21, mx, "sweep_21_40")):
        times = []
        for rep in range(8):
            ws.set_col(1, v1)
            ws.reinitialize(0, "keep")
            if lo > 1:
                ws.iterate_arnoldi(op, 1, lo - 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ctx.synchronize()
            e0.record(stream)
            st = ws.iterate_arnoldi(op, lo, hi)
            e1.record(stream)
            ctx.synchronize()
            times.append(e0.elapsed_time(e1) * 1e3 / (hi - lo + 1))
        out[tag + "_us_per_step"] = round(float(np.median(times[2:])), 2)
        out[tag + "_second_passes"] = int(st.second_passes)
    # per-kernel events (serialised by the events: no PDL overlap)
    ctx.profile(True)
    ws.set_col(1, v1)
    ws.reinitialize(0, "keep")
    ws.iterate_arnoldi(op, 1, mx)
    rep = ctx.profile_report()
    ctx.profile(False)
    out["events_us"] = {k: round(1e3 * r["ms"] / r["launches"], 1) for k, r in rep.items() if r["launches"]}
    if os.environ.get("B2A_SWEEP_TRACE") == "1" and os.environ.get("B2A_FUSED_SWEEP") == "1":
        # phase trace of the last fused launch (step 40: j = 40)
        slots = C.c_int()
        buf = (C.c_ulonglong * (148 * 8))()
        L.check(L.lib().b2a_ws_debug_sweep_trace(ws._h, buf, 148, C.byref(slots)))
        t = np.array(buf[:], dtype=np.float64).reshape(148, 8)
        t = t[t[:, 0] > 0]  # CTAs that exist (the grid can be smaller than the SM count)
        t0 = t[:, 0].min()
        names = ["start", "P1_end", "A_done", "P2_end", "B_done", "P3_end", "C_done", "end"]
        out["trace_last_step_us"] = {
            nm: [round(float((t[:, k].min() - t0) / 1e3), 1), round(float((t[:, k].max() - t0) / 1e3), 1)]
            for k, nm in enumerate(names)}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    if "--one" in sys.argv:
        one()
    else:
        for name, env in VARIANTS:
            e = dict(os.environ)
            e.update(env)
            e["B2A_SWEEP_TRACE"] = "1"
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=e, capture_output=True,
                               text=True, timeout=300)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ("ERR " + r.stderr[-400:])
            print(name, line, flush=True)
