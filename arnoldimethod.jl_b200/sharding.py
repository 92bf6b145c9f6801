"""Row-block sharding of the n dimension: pure host logic, no GPU.

GPU p of P owns the contiguous rows [offset_p, offset_p + count_p) of A, of every
column of V and of v.  Blocks are uniform (``ceil(n / P)`` rows, the last rank
takes the remainder) so that the x-exchange before the mat-vec is a single
all-gather.  H and Q are replicated on every rank's host.
"""

import numpy as np


def row_partition(n, world):
    """-> (offsets, counts) of the uniform contiguous row blocks."""
    if world < 1:
        raise ValueError("world must be >= 1")
    blk = -(-n // world)
    offsets = np.minimum(np.arange(world, dtype=np.int64) * blk, n)
    counts = np.minimum(offsets + blk, n) - offsets
    return offsets, counts


def local_rows(n, rank, world):
    offsets, counts = row_partition(n, world)
    return int(offsets[rank]), int(counts[rank])


def shard_csr(indptr, indices, data, n, rank, world):
    """Rows of this rank from a global CSR (0-based).  Column numbers stay GLOBAL.

    Returns ``(row_offset, n_local, indptr_local, indices_local, data_local)`` with
    ``indptr_local[0] == 0``.
    """
    off, cnt = local_rows(n, rank, world)
    lo, hi = int(indptr[off]), int(indptr[off + cnt])
    ip = np.asarray(indptr[off : off + cnt + 1], dtype=np.int64) - lo
    return off, cnt, ip, np.asarray(indices[lo:hi]), np.asarray(data[lo:hi])


def needed_columns(indices_local):
    """Distinct x entries a shard references (n_x of SURVEY 8(d))."""
    return np.unique(np.asarray(indices_local))


def halo_plan(indices_local, n, world):
    """Per source rank: the sorted global x entries this shard needs from it.

    For unstructured matrices every list is (almost) the whole remote block, i.e. the
    exchange degenerates to an all-gather; for banded / stencil matrices only the
    neighbouring ranks contribute a thin halo (cfg 3: one 512^2 plane each side).
    """
    offsets, counts = row_partition(n, world)
    need = needed_columns(indices_local)
    owner = np.searchsorted(offsets + counts, need, side="right")
    return [need[owner == r] for r in range(world)]
