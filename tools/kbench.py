"""Kernel microbenchmark: per-kind CUDA-event timings of one orthogonalisation at (n, j),
optionally sweeping the TMA ring geometry through the B2A_TMA_* environment overrides.

    python tools/kbench.py [n] [j] [--sweep] [--complex]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import b200arnoldi as b2a


def run(n, j, T, reps=20, nearly=False):
    ctx = b2a.default_context()
    rng = np.random.default_rng(0)
    ws = b2a.ArnoldiWorkspace(n, j + 1, dtype=T, ctx=ctx)
    x = rng.standard_normal(n).astype(T)
    for c in range(j):
        ws.set_col(c + 1, np.roll(x, c) / np.linalg.norm(x))
    out = {}
    for it in range(reps + 3):
        ws.set_col(j + 1, x)
        if it == 3:
            ctx.profile(True)
        ws.orthogonalize(j)
    rep = ctx.profile_report()
    ctx.profile(False)
    for k, r in rep.items():
        if r["launches"]:
            out[k] = (1e3 * r["ms"] / r["launches"], r["bytes"] / (r["ms"] * 1e-3) / 1e9, r["launches"])
    ws.close()
    return out


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n = int(float(args[0])) if args else 1_000_000
    j = int(args[1]) if len(args) > 1 else 40
    T = np.complex128 if "--complex" in sys.argv else np.float64
    configs = [dict()]
    if "--sweep" in sys.argv:
        configs = [dict(), dict(B2A_FUSED_SWEEP="0"), dict(B2A_NO_TMA="1", B2A_FUSED_SWEEP="0")]
        for l2 in (0, 1, 2, 3):
            configs.append(dict(B2A_TMA_L2PROMO=str(l2)))
        for rt in (128, 256):
            configs.append(dict(B2A_TMA_RT_DOTS=str(rt), B2A_TMA_RT_UPD=str(rt)))
    for cfg in configs:
        for k in ("B2A_NO_TMA", "B2A_FUSED_SWEEP", "B2A_TMA_RT_DOTS", "B2A_TMA_RT_UPD", "B2A_TMA_STAGES", "B2A_TMA_L2PROMO"):
            os.environ.pop(k, None)
        os.environ.update(cfg)
        r = run(n, j, T)
        print(cfg, {k: f"{v[0]:.1f}us {v[1]:.0f}GB/s x{v[2]}" for k, v in r.items()}, flush=True)
