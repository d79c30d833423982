"""Multi-GPU sharding of the ready set: one process per GPU, the task table is
block-sharded by handle, worker state is replicated.

Per tick:
  1. every rank histograms ITS ready tasks per (priority level, class) group     hqs_shard_count
  2. all-gather of the count vectors (NCCL over NVLink; 4 B x groups per rank — the only data-path
     collective), from which every rank derives   counts_all = sum over ranks
                                                   ranks_before = sum over lower ranks   (shard_exchange)
  3. every rank runs the SAME deterministic solve on counts_all (replicated worker state => identical
     count segments everywhere) and emits only its own tasks: a task's global rank inside its group is
     ranks_before[g] + its local rank                                           hqs_shard_solve_emit
  4. the assignment lists are gathered where they are needed (host, or all-gather of (task, worker) pairs).
No task data moves between GPUs.  The exchange logic is device-agnostic torch code so that it is covered
by world_size-2 gloo tests on CPU (tests/test_sharded_cpu.py).

Fused form (`ShardedScheduler(..., p2p=True)`, the default on GPUs): steps 1-3 are ONE stream of kernels with no
host collective in between — the counting step's vector goes to every peer by NVLink peer stores (CUDA IPC
mapped exchange buffers, hqs_shard_xbuf / hqs_ipc_open / hqs_shard_attach) followed by a release flag, and the
solver kernel acquires all flags and sums the vectors itself (hqs_shard_tick_launch).  torch.distributed is then
used once, at set-up, to pass the 64-byte IPC handles around.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L


def shard_exchange(counts_local: torch.Tensor, rank: int, world: int,
                   group: Optional[dist.ProcessGroup] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """counts_local: int32 [G] on any device.  Returns (counts_all, ranks_before), int32 [G]."""
    if world == 1:
        return counts_local.clone(), torch.zeros_like(counts_local)
    gathered = torch.empty(world * counts_local.numel(), dtype=counts_local.dtype, device=counts_local.device)
    dist.all_gather_into_tensor(gathered, counts_local.contiguous(), group=group)
    g2 = gathered.view(world, -1).to(torch.int64)
    counts_all = g2.sum(0).to(torch.int32)
    before = g2[:rank].sum(0).to(torch.int32) if rank else torch.zeros_like(counts_local)
    return counts_all, before


def gather_peer_handles(sched, rank: int, world: int, group: Optional[dist.ProcessGroup] = None):
    """Collective half of the set-up: allocates this rank's exchange buffer and all-gathers the 64-byte CUDA IPC
    handles (on the host).  Returns (own device pointer, [handle bytes of rank r])."""
    lib = sched._lib
    own = C.c_void_p()
    handle = (C.c_uint8 * L.HQS_IPC_HANDLE_BYTES)()
    sched._check(lib.hqs_shard_xbuf(sched._ctx, C.byref(own), handle))
    if world == 1:
        return own, [bytes(handle)]
    mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    gathered = [torch.zeros(L.HQS_IPC_HANDLE_BYTES, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(gathered, mine.to(dev), group=group)
    return own, [bytes(g.cpu().tolist()) for g in gathered]


def open_and_attach(sched, rank: int, world: int, own, handles) -> None:
    """Local half: maps the other ranks' buffers (cudaIpcOpenMemHandle) and attaches the context."""
    lib = sched._lib
    ptrs = (C.c_void_p * world)()
    for r in range(world):
        if r == rank:
            ptrs[r] = own
            continue
        hb = (C.c_uint8 * L.HQS_IPC_HANDLE_BYTES)(*handles[r])
        p = C.c_void_p()
        sched._check(lib.hqs_ipc_open(sched._ctx, hb, C.byref(p)))
        ptrs[r] = p
    sched._check(lib.hqs_shard_attach(sched._ctx, world, rank, ptrs))


def attach_peers(sched, rank: int, world: int, group: Optional[dist.ProcessGroup] = None) -> None:
    """One-time set-up of the peer-to-peer count exchange: every rank allocates its exchange buffer, the CUDA IPC
    handles are all-gathered (64 bytes per rank, on the host), every rank maps the others' buffers."""
    own, handles = gather_peer_handles(sched, rank, world, group)
    open_and_attach(sched, rank, world, own, handles)


def block_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous handle range [lo, hi) of a rank; ranks are ordered by handle so that lower ranks hold the
    lower (earlier TaskId) handles — the global rank of a task inside its group needs exactly that."""
    per = (n_total + world - 1) // world
    lo = min(rank * per, n_total)
    return lo, min(lo + per, n_total)


class ShardedScheduler:
    """Wraps one GpuScheduler per rank.  Handles given to / returned from this class are GLOBAL."""

    def __init__(self, sched, rank: int, world: int, n_total: int, device: torch.device,
                 group: Optional[dist.ProcessGroup] = None, p2p: bool = False) -> None:
        self.s = sched
        self.rank, self.world, self.group = rank, world, group
        self.lo, self.hi = block_range(n_total, rank, world)
        self.device = device
        self._counts = torch.zeros(L.HQS_MAX_GROUPS, dtype=torch.int32, device=device)
        self.p2p = bool(p2p)
        if self.p2p:
            attach_peers(sched, rank, world, group)

    def add_ready_tasks(self, handles, rq_ids, priorities) -> None:
        h = np.asarray(handles, dtype=np.int64)
        # every rank sees the whole call: number the priority levels identically everywhere
        lv = np.ascontiguousarray(np.unique(np.asarray(priorities, dtype=np.uint64)))
        self.s._sync_classes()
        self.s._check(self.s._lib.hqs_levels_add(self.s._ctx, lv.size, L.ptr(lv)))
        m = (h >= self.lo) & (h < self.hi)
        if m.any():
            self.s.add_ready_tasks((h[m] - self.lo).astype(np.uint32), np.asarray(rq_ids)[m], np.asarray(priorities)[m])

    def run_scheduling(self, now: float = 0.0, out_cap: Optional[int] = None):
        s = self.s
        s._sync_classes()
        w = s._worker_structs(now)
        free = np.ascontiguousarray(s.free)
        total = np.ascontiguousarray(s.total)
        blocked = s._blocked_bytes()
        if self.p2p:
            cap = out_cap or max(self.hi - self.lo, 1)
            s._check(s._lib.hqs_shard_tick_launch(s._ctx, w.shape[0], L.ptr(w), L.ptr(free), L.ptr(total),
                                                  L.ptr(blocked) if blocked is not None else None, cap))
            out = np.zeros(cap, dtype=L.assignment_dtype)
            free_after = np.zeros_like(free)
            n = C.c_uint32(0)
            s._check(s._lib.hqs_tick_fetch(s._ctx, cap, L.ptr(out), C.byref(n), L.ptr(free_after)))
            a = out[: n.value].copy()
            self._record(a)
            a["task"] += np.uint32(self.lo)
            s.free = free_after
            return a, free_after
        ng = C.c_uint32(0)
        s._check(s._lib.hqs_shard_count(s._ctx, w.shape[0], L.ptr(w), L.ptr(free), L.ptr(total),
                                        L.ptr(blocked) if blocked is not None else None,
                                        C.c_void_p(self._counts.data_ptr()), self._counts.numel(), C.byref(ng)))
        counts_all, before = shard_exchange(self._counts, self.rank, self.world, self.group)
        torch.cuda.synchronize(self.device)
        cap = out_cap or max(self.hi - self.lo, 1)
        s._check(s._lib.hqs_shard_solve_emit(s._ctx, C.c_void_p(counts_all.data_ptr()), C.c_void_p(before.data_ptr()), cap))
        out = np.zeros(cap, dtype=L.assignment_dtype)
        free_after = np.zeros_like(free)
        n = C.c_uint32(0)
        s._check(s._lib.hqs_tick_fetch(s._ctx, cap, L.ptr(out), C.byref(n), L.ptr(free_after)))
        a = out[: n.value].copy()
        self._record(a)
        a["task"] += np.uint32(self.lo)
        s.free = free_after
        return a, free_after

    def _record(self, a_local) -> None:
        """TaskRuntimeState::Assigned{worker_id, rv_id} of this rank's tasks (local handles), as GpuScheduler.run_scheduling
        keeps it: needed to return the resources when the tasks finish."""
        if a_local.size:
            self.s._task_worker[a_local["task"]] = a_local["worker"]
            self.s._task_variant[a_local["task"]] = a_local["variant"]

    def tasks_finished(self, handles) -> None:
        """task_finished for GLOBAL handles, called with the same list on every rank: every rank holds the replicated free
        vectors, but only the owner of a task knows where it ran, so the per-worker amounts to give back are summed over the
        ranks (one all-reduce of a [W][R] matrix, or nothing with world == 1)."""
        s = self.s
        h = np.asarray(handles, dtype=np.int64)
        mine = h[(h >= self.lo) & (h < self.hi)] - self.lo
        add = np.zeros_like(s.free)
        reset = np.zeros(s.free.shape, dtype=bool)
        if mine.size:
            wi, cl, va = s._task_worker[mine], s._task_class[mine], s._task_variant[mine]
            assert (wi >= 0).all(), "a finished task of this rank was never assigned"
            np.add.at(add, wi, s._amount_tab[cl, va])
            allm = s._all_tab[cl, va]
            if allm.any():
                ws, rs = np.nonzero(allm)
                reset[wi[ws], rs] = True
            s._task_worker[mine] = -1
        if self.world > 1:
            t = torch.from_numpy(np.stack([add.astype(np.int64), reset.astype(np.int64)]))
            dev = self.device if dist.get_backend(self.group) == "nccl" else torch.device("cpu")
            t = t.to(dev)
            dist.all_reduce(t, group=self.group)
            t = t.cpu().numpy()
            add, reset = t[0].astype(np.uint64), t[1] > 0
        s.free = s.free + add.astype(np.uint64)
        s.free[reset] = s.total[reset]
