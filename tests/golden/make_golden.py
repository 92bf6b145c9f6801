"""Generate tests/golden/*.npz.

The reference (ArnoldiMethod.jl) is pure Julia and cannot be executed in the build image, so these
fixtures are produced by the ORACLE (oracle/, the NumPy restatement pinned to the reference's published
known answers) from fixed start vectors.  They freeze the oracle's behaviour: the CPU tests check the
oracle against them (regression) and the GPU tests check the CUDA path against them, so the two
implementations are compared through committed numbers and not only through a live run.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402


def tridiag(n):
    return sp.diags([-np.ones(n - 1), 2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1], format="csr")


def main():
    # 1. README example (readme.md:28-60): tridiag(-1,2,-1), n=100, nev=10, tol=1e-6, :SR
    A = tridiag(100)
    v1 = np.random.default_rng(2024).random(100)
    arn = oracle.ArnoldiWorkspace(np.float64, 100, 20)
    arn.V[:, 0] = v1 / np.linalg.norm(v1)
    oracle.iterate_arnoldi(A, arn, 1, 20)
    P, hist = oracle.partialschur(A, v1=v1, nev=10, tol=1e-6, which="SR")
    np.savez(os.path.join(HERE, "readme_tridiag_n100.npz"), v1=v1, H_first_sweep=arn.H, eigenvalues=P.eigenvalues,
             R=P.R, mvproducts=hist.mvproducts, nconverged=hist.nconverged,
             residual=np.linalg.norm(A @ P.Q - P.Q @ P.R))

    # 2. complex non-symmetric, all five targets (60 x 60)
    rng = np.random.default_rng(13)
    n = 60
    d = rng.standard_normal(n) * 10 + 10j * rng.standard_normal(n)
    Ac = np.diag(d) + 0.01 * (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    v1c = rng.random(n) + 1j * rng.random(n)
    out = dict(A=Ac, v1=v1c)
    for which in ("LM", "LR", "SR", "LI", "SI"):
        P, hist = oracle.partialschur(Ac, v1=v1c, nev=4, which=which, tol=1e-9, restarts=500)
        out[f"eig_{which}"] = P.eigenvalues
        out[f"mv_{which}"] = hist.mvproducts
    np.savez(os.path.join(HERE, "complex_targets_n60.npz"), **out)

    # 3. real non-symmetric with a conjugate pair at the cut (200 x 200)
    rng = np.random.default_rng(12)
    Ar = rng.standard_normal((200, 200))
    v1r = rng.random(200)
    P, hist = oracle.partialschur(Ar, v1=v1r, nev=8, tol=1e-8, which="LM", restarts=400)
    np.savez(os.path.join(HERE, "real_conjugate_pair_n200.npz"), A=Ar, v1=v1r, eigenvalues=P.eigenvalues,
             mvproducts=hist.mvproducts, nconverged=hist.nconverged)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
