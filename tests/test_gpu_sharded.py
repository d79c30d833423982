"""Sharded tick on the GPU: two contexts act as rank 0 / rank 1 of a block-sharded ready set (SURVEY.md
§8(e)).  The merged result must be IDENTICAL, placement by placement, to the single-context tick.  Runs on
one GPU (the count exchange is emulated with torch sums); with >= 2 GPUs the same check also runs with one
process per GPU over NCCL."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch

import parity as P

pytestmark = pytest.mark.gpu


def _single(wl):
    s = P.gpu_scheduler(wl)
    m = s.run_scheduling()
    s.close()
    return m


def _shard_workload(wl, lo, hi):
    import copy
    w2 = copy.copy(wl)
    w2.task_class = wl.task_class[lo:hi]
    w2.task_user_priority = wl.task_user_priority[lo:hi]
    return w2


@pytest.mark.parametrize("n,w,q,scale", [(30000, 16, 8, 1), (30000, 16, 8, 1024), (100001, 64, 16, 1)])
def test_two_shards_on_one_gpu_equal_single_context(n, w, q, scale):
    from hyperqueue_b200 import _lib as L
    from hyperqueue_b200.sharded import block_range
    wl = P.make_independent(n, w, q, seed=4, free_scale=scale)
    ref = _single(wl)
    dev = torch.device("cuda", 0)
    parts, counts = [], []
    for r in range(2):
        lo, hi = block_range(n, r, 2)
        s = P.gpu_scheduler(_shard_workload(wl, lo, hi), add_tasks=False)
        # both shards must number the priority levels identically
        from hyperqueue_b200 import priority_from_user
        lv = np.ascontiguousarray(np.unique(priority_from_user(wl.task_user_priority)))
        s._sync_classes()
        s._check(s._lib.hqs_levels_add(s._ctx, lv.size, L.ptr(lv)))
        s.add_ready_tasks(np.arange(hi - lo, dtype=np.uint32), wl.task_class[lo:hi], priority_from_user(wl.task_user_priority[lo:hi]))
        parts.append((s, lo, hi))
    # level tables must agree across ranks: seed both contexts with the full set of priorities
    workers = parts[0][0]._worker_structs(0.0)
    free = np.ascontiguousarray(wl.worker_free); total = np.ascontiguousarray(wl.worker_total)
    for s, lo, hi in parts:
        c = torch.zeros(L.HQS_MAX_GROUPS, dtype=torch.int32, device=dev)
        ng = C.c_uint32(0)
        s._check(s._lib.hqs_shard_count(s._ctx, w, L.ptr(workers), L.ptr(free), L.ptr(total), None,
                                        C.c_void_p(c.data_ptr()), c.numel(), C.byref(ng)))
        counts.append(c)
    assert parts[0][0].stats()["n_levels"] == parts[1][0].stats()["n_levels"]
    all_c = (counts[0].to(torch.int64) + counts[1].to(torch.int64)).to(torch.int32)
    befores = [torch.zeros_like(all_c), counts[0].clone()]
    torch.cuda.synchronize()
    merged = []
    for (s, lo, hi), bef in zip(parts, befores):
        s._check(s._lib.hqs_shard_solve_emit(s._ctx, C.c_void_p(all_c.data_ptr()), C.c_void_p(bef.data_ptr()), hi - lo))
        out = np.zeros(hi - lo, dtype=L.assignment_dtype)
        nn = C.c_uint32(0)
        fa = np.zeros_like(free)
        s._check(s._lib.hqs_tick_fetch(s._ctx, hi - lo, L.ptr(out), C.byref(nn), L.ptr(fa)))
        a = out[: nn.value].copy()
        a["task"] += np.uint32(lo)
        merged.append(a)
        assert np.array_equal(fa, ref.free_after)          # the solve is replicated: same free vectors
        s.close()
    got = np.concatenate(merged)
    got = got[np.argsort(got["task"], kind="stable")]
    exp = ref.assignments[np.argsort(ref.assignments["task"], kind="stable")]
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("n,w,q,scale", [(30000, 16, 8, 1), (100001, 64, 16, 1024), (200000, 1024, 16, 1024), (60000, 1024, 16, 1)])
def test_two_shards_peer_exchange_on_one_gpu(n, w, q, scale):
    """The fused sharded tick (count -> peer stores + release flag -> solver acquires and sums -> emit) with two
    contexts of ONE process attached to each other's exchange buffers; three ticks in a row exercise the
    double-buffered sequence numbers.  Must equal the single-context tick placement by placement."""
    from hyperqueue_b200 import _lib as L, priority_from_user
    from hyperqueue_b200.sharded import block_range
    wl = P.make_independent(n, w, q, seed=4, free_scale=scale)
    parts = []
    xb = (C.c_void_p * 2)()
    for r in range(2):
        lo, hi = block_range(n, r, 2)
        # the two ticks wait for each other on the device: each kernel takes half of the SMs
        s = P.gpu_scheduler(_shard_workload(wl, lo, hi), add_tasks=False, flags=L.HQS_CREATE_SHARE_DEVICE)
        lv = np.ascontiguousarray(np.unique(priority_from_user(wl.task_user_priority)))
        s._sync_classes()
        s._check(s._lib.hqs_levels_add(s._ctx, lv.size, L.ptr(lv)))
        s.add_ready_tasks(np.arange(hi - lo, dtype=np.uint32), wl.task_class[lo:hi], priority_from_user(wl.task_user_priority[lo:hi]))
        p = C.c_void_p()
        s._check(s._lib.hqs_shard_xbuf(s._ctx, C.byref(p), None))
        xb[r] = p
        parts.append((s, lo, hi))
    for r, (s, lo, hi) in enumerate(parts):
        s._check(s._lib.hqs_shard_attach(s._ctx, 2, r, xb))
        # two contexts of one process wait for each other on the device: no cudaMalloc may happen in between
        s._check(s._lib.hqs_tick_reserve(s._ctx, w, hi - lo, 0))
    single = P.gpu_scheduler(wl)
    workers = parts[0][0]._worker_structs(0.0)
    free = np.ascontiguousarray(wl.worker_free); total = np.ascontiguousarray(wl.worker_total)
    for tick in range(3):
        ref = single.run_scheduling()
        single.free = wl.worker_free.copy()                 # zero-duration: everything is free again next tick
        # both ranks launch (asynchronously, each on its own stream), then both are fetched
        for s, lo, hi in parts:
            s._check(s._lib.hqs_shard_tick_launch(s._ctx, w, L.ptr(workers), L.ptr(free), L.ptr(total), None, hi - lo))
        merged = []
        for s, lo, hi in parts:
            out = np.zeros(hi - lo, dtype=L.assignment_dtype)
            nn = C.c_uint32(0)
            fa = np.zeros_like(free)
            s._check(s._lib.hqs_tick_fetch(s._ctx, hi - lo, L.ptr(out), C.byref(nn), L.ptr(fa)))
            a = out[: nn.value].copy()
            a["task"] += np.uint32(lo)
            merged.append(a)
            assert np.array_equal(fa, ref.free_after)
        got = np.concatenate(merged)
        got = got[np.argsort(got["task"], kind="stable")]
        exp = ref.assignments[np.argsort(ref.assignments["task"], kind="stable")]
        assert np.array_equal(got, exp), tick
        if exp.size == 0:
            break
    for s, lo, hi in parts:
        s.close()
    single.close()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _nccl_worker(rank, world, port, n, w, q, ret, p2p=False):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from hyperqueue_b200 import priority_from_user
    from hyperqueue_b200.sharded import ShardedScheduler
    wl = P.make_independent(n, w, q, seed=4)
    base = P.gpu_scheduler(wl, add_tasks=False, device=rank)
    sh = ShardedScheduler(base, rank, world, n, torch.device("cuda", rank), p2p=p2p)
    sh.add_ready_tasks(np.arange(n), wl.task_class, priority_from_user(wl.task_user_priority))
    a, fa = sh.run_scheduling()
    ret[rank] = (a.tobytes(), fa.tobytes())
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_over_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from hyperqueue_b200 import _lib as L
    n, w, q = 50000, 32, 8
    wl = P.make_independent(n, w, q, seed=4)
    ref = _single(wl)
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_nccl_worker, args=(2, _free_port(), n, w, q, ret), nprocs=2, join=True)
    got = np.concatenate([np.frombuffer(ret[r][0], dtype=L.assignment_dtype) for r in range(2)])
    got = got[np.argsort(got["task"], kind="stable")]
    exp = ref.assignments[np.argsort(ref.assignments["task"], kind="stable")]
    assert np.array_equal(got, exp)


def test_two_ranks_peer_exchange_over_nvlink():
    """One process per GPU, exchange buffers mapped through CUDA IPC, no collective on the data path."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from hyperqueue_b200 import _lib as L
    n, w, q = 50000, 32, 8
    wl = P.make_independent(n, w, q, seed=4)
    ref = _single(wl)
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_nccl_worker, args=(2, _free_port(), n, w, q, ret, True), nprocs=2, join=True)
    got = np.concatenate([np.frombuffer(ret[r][0], dtype=L.assignment_dtype) for r in range(2)])
    got = got[np.argsort(got["task"], kind="stable")]
    exp = ref.assignments[np.argsort(ref.assignments["task"], kind="stable")]
    assert np.array_equal(got, exp)
