"""world_size-2 gloo test (CPU) of the multi-GPU host logic (hyperqueue_b200/sharded.py): block ranges,
the all-gather of per-group counts and the derived (counts_all, ranks_before), and that tasks ranked
locally + ranks_before reproduce the single-rank assignment.  The device phases are replaced by a numpy
stand-in that follows the same contract (count per group / emit the tasks whose global rank < k[g])."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import parity as P


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_total, q, seed, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hyperqueue_b200.sharded import block_range, shard_exchange
    rng = np.random.default_rng(seed)
    cls = rng.integers(0, q, n_total)
    lvl = rng.integers(0, 4, n_total)
    G = 4 * q
    group = lvl * q + cls
    k = np.minimum(np.bincount(group, minlength=G), rng.integers(0, 40, G))     # "solver": k[g] tasks per group
    lo, hi = block_range(n_total, rank, world)
    local_counts = torch.from_numpy(np.bincount(group[lo:hi], minlength=G).astype(np.int32))
    counts_all, before = shard_exchange(local_counts, rank, world)
    assert np.array_equal(counts_all.numpy(), np.bincount(group, minlength=G))
    assert np.array_equal(before.numpy(), np.bincount(group[:lo], minlength=G))
    # local stable rank inside the group + ranks_before = global rank; emit those below k[g]
    g_loc = group[lo:hi]
    order = np.argsort(g_loc, kind="stable")
    rank_in = np.empty(hi - lo, dtype=np.int64)
    starts = np.searchsorted(g_loc[order], np.arange(G))
    rank_in[order] = np.arange(hi - lo) - starts[g_loc[order]]
    chosen = np.nonzero(rank_in + before.numpy()[g_loc] < k[g_loc])[0] + lo
    ret[rank] = chosen.tolist()
    dist.barrier()
    dist.destroy_process_group()


def test_shard_exchange_reproduces_single_rank_selection():
    n_total, q, seed, world = 5000, 6, 3, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_total, q, seed, ret), nprocs=world, join=True)
    got = sorted(ret[0] + ret[1])
    rng = np.random.default_rng(seed)
    cls = rng.integers(0, q, n_total); lvl = rng.integers(0, 4, n_total)
    G = 4 * q
    group = lvl * q + cls
    k = np.minimum(np.bincount(group, minlength=G), rng.integers(0, 40, G))
    exp = []
    seen = np.zeros(G, dtype=np.int64)
    for t in range(n_total):
        if seen[group[t]] < k[group[t]]:
            exp.append(t)
        seen[group[t]] += 1
    assert got == exp


def test_block_ranges_cover_everything_in_order():
    from hyperqueue_b200.sharded import block_range
    for n, w in [(10, 3), (1_000_000, 8), (7, 8), (0, 2)]:
        rs = [block_range(n, r, w) for r in range(w)]
        assert rs[0][0] == 0 and rs[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
