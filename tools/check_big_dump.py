"""Oracle comparison of a BASELINE-size multi-GPU run, done OFFLINE on a CPU box (8-GPU leases are too dear to keep
idle while one host core runs the oracle on an 8e6-row matrix):

    (GPU box)  torchrun --nproc-per-node 8 tests/dist_gpu_check.py --big 1000000 --dump gpurun_out/r2_big_n8.npz
    (here)     python tools/check_big_dump.py gpurun_out/r2_big_n8.npz  > profiles/r2_dist_big_n8_vs_oracle.log

Regenerates bench.py's matrix and start vector from the same seeds, runs the oracle (first sweep + complete solve) and
checks: H of the first sweep <= 1e-13 relative, rows of V <= 1e-12, `mvproducts` within one restart, eigenvalues
<= 10 tol |lambda|.  (||A Q - Q R|| and ||Q'Q - I|| were computed on the GPU box with SciPy, rank-parallel.)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp

import bench
import oracle

d = np.load(sys.argv[1], allow_pickle=False)
meta = json.loads(str(d["meta"]))
world, rows = meta["world"], meta["rows_per_gpu"]
n = world * rows
bench.N_PER_GPU = rows
t0 = time.time()
blocks = [bench.make_shard(n, r * rows, rows) for r in range(world)]
ip = np.concatenate([[0]] + [b[0][1:] + r * rows * bench.NNZ_PER_ROW for r, b in enumerate(blocks)])
A = sp.csr_matrix((np.concatenate([b[2] for b in blocks]), np.concatenate([b[1] for b in blocks]), ip), shape=(n, n))
del blocks
v1 = np.concatenate([bench.make_v1(n, r * rows, rows) for r in range(world)])
H = d["H_first_sweep"]
steps = H.shape[1]
arn = oracle.ArnoldiWorkspace(np.float64, n, bench.MAXDIM)
arn.V[:, 0] = v1 / np.linalg.norm(v1)
oracle.iterate_arnoldi(A, arn, 1, steps)
Ho = arn.H[: steps + 1, :steps]
out = dict(meta)
out["H_relerr_first_sweep"] = float(np.abs(H - Ho).max() / np.abs(Ho).max())
Vr = d["V_first_rows"]
out["V_abserr_first_sweep_first_rows"] = float(np.abs(Vr - arn.V[: Vr.shape[0], : Vr.shape[1]]).max())
del arn
Po, ho = oracle.partialschur(A, v1=v1, nev=bench.NEV, mindim=bench.MINDIM, maxdim=bench.MAXDIM, which=bench.WHICH,
                             tol=bench.TOL)
lam = np.sort_complex(d["eigenvalues"])[-bench.NEV:]
lamo = np.sort_complex(Po.eigenvalues)[-bench.NEV:]
out["oracle_mvproducts"] = int(ho.mvproducts)
out["eig_relerr_max"] = float((np.abs(lam - lamo) / np.abs(lamo)).max())
out["oracle_seconds"] = round(time.time() - t0, 1)
print(json.dumps(out))
assert out["H_relerr_first_sweep"] <= 1e-13 and out["V_abserr_first_sweep_first_rows"] <= 1e-12, out
assert ho.converged and abs(meta["mvproducts"] - ho.mvproducts) <= bench.MAXDIM - bench.MINDIM, out
assert out["eig_relerr_max"] <= 10 * bench.TOL, out
print("BIG_DUMP_MATCHES_ORACLE")
