"""N > 1 host logic on the CPU: row-block sharding + the collectives the sharded Arnoldi step
needs (all-gather of x, all-reduce of [h; ||v||^2], all-reduce of ||v||^2), with world_size 2
over gloo.  The per-rank arithmetic is the oracle's; the partitioning code is the product's
(arnoldimethod.jl_b200/sharding.py), i.e. exactly what the GPU ranks use to cut A, V and v."""

import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp

import b200arnoldi as b2a
import oracle

sharding = b2a.sharding


def test_row_partition_uniform_blocks():
    for n, w in [(10, 1), (10, 2), (10, 3), (1000001, 8), (7, 8), (16, 4)]:
        off, cnt = sharding.row_partition(n, w)
        assert off[0] == 0 and cnt.sum() == n
        assert np.all(off[1:] == off[:-1] + cnt[:-1])
        blk = -(-n // w)
        assert np.all(cnt[:-1][cnt[:-1] > 0] <= blk) and np.all(cnt <= blk)
        # all-gather friendly: every rank but the trailing ones owns exactly blk rows
        full = cnt == blk
        assert np.all(full[: max(0, int(np.argmin(full)) if not full.all() else w)])


def test_shard_csr_reassembles():
    rng = np.random.default_rng(0)
    n = 103
    A = sp.random(n, n, 0.07, random_state=rng, format="csr")
    A.sort_indices()
    x = rng.standard_normal(n)
    for w in (1, 2, 3, 8):
        y = np.zeros(n)
        for r in range(w):
            off, cnt, ip, idx, dat = sharding.shard_csr(A.indptr, A.indices, A.data, n, r, w)
            assert ip[0] == 0 and len(ip) == cnt + 1
            Al = sp.csr_matrix((dat, idx, ip), shape=(cnt, n))
            y[off : off + cnt] = Al @ x
        assert np.allclose(y, A @ x)


def test_halo_plan_stencil_vs_random():
    n, w = 64 * 64, 4
    T1 = sp.diags([-np.ones(63), 2 * np.ones(64), -np.ones(63)], [-1, 0, 1])
    L = (sp.kron(T1, sp.identity(64)) + sp.kron(sp.identity(64), T1)).tocsr()
    off, cnt, ip, idx, dat = sharding.shard_csr(L.indptr, L.indices, L.data, n, 1, w)
    plan = sharding.halo_plan(idx, n, w)
    assert len(plan[0]) == 64 and len(plan[2]) == 64 and len(plan[3]) == 0  # one grid line per neighbour
    assert len(plan[1]) == cnt
    rng = np.random.default_rng(1)
    R = sp.random(n, n, 16 / n, random_state=rng, format="csr")
    off, cnt, ip, idx, dat = sharding.shard_csr(R.indptr, R.indices, R.data, n, 1, w)
    plan = sharding.halo_plan(idx, n, w)
    assert all(len(p) > 0.9 * n / w for p in plan)  # unstructured: needs (almost) all of x


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _sharded_arnoldi_worker(rank, world, port, n, maxdim, seed, out_dir):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(seed)
    A = (sp.random(n, n, 8 / n, random_state=rng, format="csr") + 3 * sp.identity(n)).tocsr()
    A.sort_indices()
    v1 = rng.random(n)
    off, cnt, ip, idx, dat = sharding.shard_csr(A.indptr, A.indices, A.data, n, rank, world)
    Al = sp.csr_matrix((dat, idx, ip), shape=(cnt, n))
    offs, cnts = sharding.row_partition(n, world)
    blk = int(cnts[0])

    def allreduce(a):
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
        dist.all_reduce(t)
        return t.numpy()

    def allgather_x(xl):
        pad = np.zeros(blk)
        pad[:cnt] = xl
        outs = [torch.zeros(blk, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(outs, torch.from_numpy(pad))
        return np.concatenate([o.numpy()[: int(c)] for o, c in zip(outs, cnts)])

    eta = np.sqrt(2) / 2
    V = np.zeros((cnt, maxdim + 1), order="F")
    H = np.zeros((maxdim + 1, maxdim))
    v = v1[off : off + cnt].copy()
    V[:, 0] = v / np.sqrt(allreduce([v @ v])[0])
    for j in range(1, maxdim + 1):
        x = allgather_x(V[:, j - 1])  # x-exchange
        w_ = Al @ x  # local rows of A x
        red = allreduce(np.append(V[:, :j].T @ w_, w_ @ w_))  # [h; ||v||^2] in ONE all-reduce
        h, rnorm = red[:j], np.sqrt(red[j])
        w_ = w_ - V[:, :j] @ h
        wnorm = np.sqrt(allreduce([w_ @ w_])[0])
        if wnorm < eta * rnorm:  # every rank takes the same branch: the scalars are all-reduced
            c = allreduce(V[:, :j].T @ w_)
            w_ = w_ - V[:, :j] @ c
            h = h + c
            wnorm = np.sqrt(allreduce([w_ @ w_])[0])
        H[:j, j - 1] = h
        H[j, j - 1] = wnorm
        V[:, j] = w_ / wnorm
    np.save(os.path.join(out_dir, f"H{rank}.npy"), H)
    np.save(os.path.join(out_dir, f"V{rank}.npy"), V)
    dist.destroy_process_group()


def test_sharded_arnoldi_matches_unsharded_oracle(tmp_path):
    """Shard-count invariance: 2 ranks give the same H as 1 to reduction-order tolerance."""
    import torch.multiprocessing as mp

    n, maxdim, seed, world = 4001, 12, 7, 2
    mp.spawn(_sharded_arnoldi_worker, args=(world, _free_port(), n, maxdim, seed, str(tmp_path)), nprocs=world,
             join=True)
    rng = np.random.default_rng(seed)
    A = (sp.random(n, n, 8 / n, random_state=rng, format="csr") + 3 * sp.identity(n)).tocsr()
    v1 = rng.random(n)
    arn = oracle.ArnoldiWorkspace(np.float64, n, maxdim)
    arn.V[:, 0] = v1 / np.linalg.norm(v1)
    oracle.iterate_arnoldi(A, arn, 1, maxdim)
    H0 = np.load(tmp_path / "H0.npy")
    H1 = np.load(tmp_path / "H1.npy")
    assert np.array_equal(H0, H1)  # replicated H is bit-identical on every rank
    assert np.abs(H0 - arn.H).max() < 1e-12 * np.abs(arn.H).max()
    V = np.vstack([np.load(tmp_path / "V0.npy"), np.load(tmp_path / "V1.npy")])
    assert np.abs(V - arn.V).max() < 1e-10


def _owner_plan(dtype_code, n, world, rank):
    import ctypes as C

    from arnoldimethod_jl_b200 import _lib as L

    g, nb = C.c_int(), C.c_int()
    blk = (C.c_int * world)()
    L.check(L.lib().b2a_host_owner_group_plan(dtype_code, int(n), int(world), int(rank), C.byref(g), C.byref(nb), blk))
    return g.value, nb.value, list(blk)


@pytest.mark.parametrize("world,n", [(2, 2_000_000), (3, 9_000_000), (4, 4_000_000), (6, 9_000_000), (8, 8_000_000),
                                     (8, 100_000_000), (5, 50_000_000)])
def test_staged_exchange_schedule_matches_owner_groups(world, n):
    """The staged x exchange and the owner-group mat-vec agree, for the library's own plan
    (`b2a_host_owner_group_plan`, the function `b2a_csr_create` uses):
      * stage k = 1 .. P-1 sends the slice of rank s to rank (s - k) mod P: every stage is a permutation, so each
        NVLink port carries one slice in and one out, and after P - 1 stages every rank holds every slice;
      * rank r therefore receives owners r+1, r+2, ... in that order, and the column blocks of its operator are numbered
        in exactly that order (block 0 = own slice first): a pass never waits for a slice that arrives after a slice of
        a later block;
      * the blocked mat-vec (one pass per block, accumulating) reproduces the plain one."""
    P = world
    have = [{r} for r in range(P)]
    for k in range(1, P):
        receivers = [(s - k) % P for s in range(P)]
        assert sorted(receivers) == list(range(P))  # a permutation per stage
        for s, r in enumerate(receivers):
            have[r].add(s)
    assert all(h == set(range(P)) for h in have)
    for r in range(P):
        G, nb, blk = _owner_plan(0, n, P, r)
        assert nb == -(-P // G) and blk[r] == 0
        arrival = [r] + [(r + k) % P for k in range(1, P)]
        order = [blk[o] for o in arrival]
        assert order == sorted(order)  # blocks are consumed in arrival order
        assert all(order.count(b) == min(G, P - b * G) for b in range(nb))
    # numerics of the blocked mat-vec on a small shard with this very block map
    rng = np.random.default_rng(world)
    m = 40 * P
    W = -(-m // P)
    A = sp.random(W, m, 0.2, random_state=rng, format="csr")
    x = rng.standard_normal(m)
    G, nb, blk = _owner_plan(0, n, P, 1 % P)
    y = np.zeros(W)
    for b in range(nb):
        cols = np.array([blk[min(c // W, P - 1)] == b for c in range(m)])
        y += A[:, cols] @ x[cols]
    assert np.allclose(y, A @ x, rtol=1e-13, atol=1e-13)
