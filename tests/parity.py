"""Shared parity-test plumbing: synthetic workloads (BASELINE.md §4 shapes, scaled), drivers that run the
SAME workload through the oracle (CPU restatement of the reference) and through the CUDA path (C ABI),
the zero-duration drain simulator (SURVEY.md §8(d) mode M2) and the checks.

Test infrastructure only — never imported by hyperqueue_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from oracle import judge as J
from oracle import model as M
from oracle.core import Core, Task

FR = 10_000
MAXV = 8          # HQS_MAX_VARIANTS


@dataclass
class Workload:
    R: int
    classes: List[List[dict]]                  # class -> variants: {"amounts": {r: fractions}, "all": (), "weight", "min_time_s"}
    worker_total: np.ndarray                   # [W][R] u64
    worker_free: np.ndarray                    # [W][R] u64
    task_class: np.ndarray                     # [N] u32
    task_user_priority: np.ndarray             # [N] i32
    blocked: Optional[np.ndarray] = None       # [W][Q][MAXV] bool
    worker_remaining_s: Optional[np.ndarray] = None   # [W] float seconds, inf = none
    deps: Optional[List[List[int]]] = None     # DAG: deps[t] = list of earlier task ids
    name: str = ""

    @property
    def n_tasks(self) -> int:
        return int(self.task_class.shape[0])

    @property
    def n_workers(self) -> int:
        return int(self.worker_total.shape[0])

    # dense class tables for the judge
    def class_tables(self):
        Q = len(self.classes)
        amounts = np.zeros((Q, MAXV, self.R), dtype=np.uint64)
        allm = np.zeros((Q, MAXV, self.R), dtype=bool)
        nvar = np.zeros(Q, dtype=np.int64)
        mint = np.zeros((Q, MAXV), dtype=np.uint64)
        for c, vs in enumerate(self.classes):
            nvar[c] = len(vs)
            for v, d in enumerate(vs):
                for r, a in d["amounts"].items():
                    amounts[c, v, r] = a
                for r in d.get("all", ()):
                    allm[c, v, r] = True
                mint[c, v] = int(round(d.get("min_time_s", 0.0) * 1000))
        return amounts, allm, nvar, mint

    def remaining_ms(self) -> np.ndarray:
        W = self.n_workers
        if self.worker_remaining_s is None:
            return np.full(W, J.TIME_INF, dtype=np.uint64)
        return np.where(np.isinf(self.worker_remaining_s), np.uint64(J.TIME_INF),
                        (np.nan_to_num(self.worker_remaining_s, posinf=0) * 1000).astype(np.uint64))


# -------------------------------------------------------------------------------------------------
# synthetic inputs
# -------------------------------------------------------------------------------------------------
def _class_pool(rng, q: int, variants3: bool) -> List[List[dict]]:
    out = []
    for _ in range(q):
        cpus = int(rng.integers(1, 17))
        gpus = [0, 2500, 5000, 10000, 20000][int(rng.integers(0, 5))]     # 0, .25, .5, 1, 2 (fractional)
        mem = int(rng.integers(1, 65))
        disk = int(rng.integers(0, 33))
        base = {0: cpus * FR, 2: mem * FR}
        if gpus:
            base[1] = gpus
        if disk:
            base[3] = disk * FR
        if not variants3:
            out.append([{"amounts": base}])
        else:
            heavy = dict(base); heavy[0] = 4 * cpus * FR; heavy.pop(1, None)          # cpu-heavy, no gpu
            gpu = dict(base); gpu[0] = 1 * FR; gpu[1] = base.get(1, 0) + 1 * FR       # gpu-heavy
            out.append([{"amounts": heavy}, {"amounts": base}, {"amounts": gpu}])
    # interning: identical request lists collapse to one class in the reference (map.rs:99-109)
    uniq, seen = [], set()
    for c in out:
        key = repr(c)
        if key not in seen:
            seen.add(key); uniq.append(c)
    return uniq


def _zipf_classes(rng, n: int, q: int, s: float = 1.1) -> np.ndarray:
    w = 1.0 / np.arange(1, q + 1) ** s
    return rng.choice(q, size=n, p=w / w.sum()).astype(np.uint32)


def make_independent(n: int, w: int, q: int, seed: int = 0, free_scale: int = 1, variants3: bool = False,
                     blocked_density: float = 0.0, n_priorities: int = 8) -> Workload:
    """cfg2 / cfg3 shape: workers {cpus 128, gpus 8, mem 512, disk 2048} x free_scale, Zipf(1.1) class mix,
    user_priority U{0..n_priorities-1}."""
    rng = np.random.default_rng(seed)
    classes = _class_pool(rng, q, variants3)
    q = len(classes)
    total = np.tile(np.array([128, 8, 512, 2048], dtype=np.uint64) * np.uint64(FR) * np.uint64(free_scale), (w, 1))
    blocked = None
    if blocked_density > 0:
        blocked = np.zeros((w, q, MAXV), dtype=bool)
        nv = len(classes[0])
        blocked[:, :, :nv] = rng.random((w, q, nv)) < blocked_density
    return Workload(4, classes, total, total.copy(), _zipf_classes(rng, n, q),
                    rng.integers(0, n_priorities, size=n).astype(np.int32), blocked,
                    name=f"indep n={n} w={w} q={q} v={'3' if variants3 else '1'}")


def make_dag(n: int, w: int, q: int, seed: int = 0, window: int = 4096, max_deg: int = 8) -> Workload:
    """cfg4 shape: topological ids, in-degree U{0..8} from the previous `window` ids, out-degree <= 8 by
    rejection, unit b-level as user priority (the reference has no b-level: SURVEY.md §0)."""
    rng = np.random.default_rng(seed)
    classes = _class_pool(rng, q, False)
    q = len(classes)
    out_deg = np.zeros(n, dtype=np.int32)
    deps: List[List[int]] = []
    for t in range(n):
        k = int(rng.integers(0, max_deg + 1)) if t > 0 else 0
        lo = max(0, t - window)
        cand = np.unique(rng.integers(lo, t, size=k)) if k else np.zeros(0, dtype=np.int64)
        ds = [int(d) for d in cand if out_deg[d] < max_deg]
        for d in ds:
            out_deg[d] += 1
        deps.append(ds)
    blevel = np.ones(n, dtype=np.int32)
    consumers: List[List[int]] = [[] for _ in range(n)]
    for t, ds in enumerate(deps):
        for d in ds:
            consumers[d].append(t)
    for t in range(n - 1, -1, -1):
        if consumers[t]:
            blevel[t] = 1 + max(blevel[c] for c in consumers[t])
    total = np.tile(np.array([128, 8, 512, 2048], dtype=np.uint64) * np.uint64(FR), (w, 1))
    return Workload(4, classes, total, total.copy(), _zipf_classes(rng, n, q), blevel, deps=deps,
                    name=f"dag n={n} w={w} q={q}")


def dag_csr(deps: List[List[int]]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    n = len(deps)
    n_deps = np.array([len(d) for d in deps], dtype=np.uint32)
    cnt = np.zeros(n + 1, dtype=np.int64)
    for ds in deps:
        for d in ds:
            cnt[d + 1] += 1
    off = np.cumsum(cnt)
    cons = np.zeros(int(off[-1]), dtype=np.uint32)
    fill = off[:-1].copy()
    for t, ds in enumerate(deps):
        for d in ds:
            cons[fill[d]] = t
            fill[d] += 1
    return n_deps, off.astype(np.uint32), cons


# -------------------------------------------------------------------------------------------------
# oracle side
# -------------------------------------------------------------------------------------------------
def oracle_rqv(variants: List[dict]) -> M.ResourceRequestVariants:
    rqs = []
    for d in variants:
        es = [M.AllocRequest(r, M.COMPACT, int(a)) for r, a in d["amounts"].items()]
        es += [M.AllocRequest(r, M.ALL) for r in d.get("all", ())]
        rqs.append(M.ResourceRequest.new(es, 0, d.get("min_time_s", 0.0), d.get("weight", 1.0)))
    return M.ResourceRequestVariants(tuple(rqs))


def oracle_core(wl: Workload, add_tasks: bool = True, worker_id0: int = 0) -> Core:
    core = Core()
    for r in range(1, wl.R):
        core.get_or_create_resource_id(f"res{r}")
    for c, vs in enumerate(wl.classes):
        rid = core.get_or_create_resource_rq_id(oracle_rqv(vs))
        assert rid == c, "workload classes must be distinct"
    for w in range(wl.n_workers):
        term = None
        if wl.worker_remaining_s is not None and not np.isinf(wl.worker_remaining_s[w]):
            term = float(wl.worker_remaining_s[w])
        wk = M.Worker(worker_id0 + w, M.WorkerResources([int(x) for x in wl.worker_total[w]]),
                      free=M.WorkerResources([int(x) for x in wl.worker_free[w]]), termination_time=term)
        if wl.blocked is not None:
            ws, vs = np.nonzero(wl.blocked[w])
            wk.blocked_requests = set(zip(ws.tolist(), vs.tolist()))
        core.new_worker(wk)
    if add_tasks:
        if wl.deps is None:
            # bulk insert (identical result to on_new_tasks one by one when no prefill is outstanding)
            up = wl.task_user_priority
            for t in range(wl.n_tasks):
                core.tasks[t] = Task(t, int(wl.task_class[t]), int(up[t]))
            order = np.lexsort((np.arange(wl.n_tasks), up, wl.task_class))
            cls_s, up_s = wl.task_class[order], up[order]
            bounds = np.nonzero((np.diff(cls_s) != 0) | (np.diff(up_s) != 0))[0] + 1
            for seg in np.split(order, bounds):
                if seg.size:
                    core.task_queues.add_ready_tasks_bulk(seg.tolist(), int(wl.task_class[seg[0]]),
                                                          M.priority_from_user(int(up[seg[0]])))
        else:
            core.on_new_tasks([Task(t, int(wl.task_class[t]), int(wl.task_user_priority[t]), deps=tuple(wl.deps[t]))
                               for t in range(wl.n_tasks)])
    return core


# Practical solver settings for multi-tick oracle runs (see oracle/lp.py): 1 % optimality gap, 2 s cap,
# incumbent accepted.  The golden-vector tests run with the reference's exact defaults instead.
ORACLE_FAST = dict(time_limit=2.0, mip_rel_gap=0.01, accept_incumbent=True)


def oracle_tick(core: Core, now: float = 0.0, **solver_opts):
    """One reference tick; returns (task[], worker[], variant[]) arrays (worker = worker id)."""
    mapping = core.schedule_mapping(now, **(solver_opts or ORACLE_FAST))
    ts, ws, vs = [], [], []
    for wid, up in mapping.workers.items():
        assert not up.prefills or True
        for t, v in up.assigned:
            ts.append(t); ws.append(wid); vs.append(v)
    return (np.array(ts, dtype=np.int64), np.array(ws, dtype=np.int64), np.array(vs, dtype=np.int64), mapping)


def oracle_drain(wl: Workload, max_ticks: int = 100000, disable_prefill: bool = True):
    """Zero-duration drain (cfg(zero_worker) semantics): tick, every assigned task finishes at once,
    resources return, newly ready consumers enter the queues; repeat.  Returns (ticks, per-tick counts)."""
    core = oracle_core(wl)
    if disable_prefill:
        core.scheduler_state.config.proactive_filling_max = 0     # prefill is a latency hider, not capacity
    remaining = wl.n_tasks
    per_tick = []
    while remaining > 0 and len(per_tick) < max_ticks:
        ts, ws, vs, _ = oracle_tick(core)
        if ts.size == 0:
            raise RuntimeError(f"oracle drain stalled with {remaining} tasks left")
        for t, w in zip(ts.tolist(), ws.tolist()):
            core.task_finished(w, t)
        remaining -= ts.size
        per_tick.append(int(ts.size))
    return len(per_tick), per_tick


# -------------------------------------------------------------------------------------------------
# CUDA side (through the C ABI via the host mirror)
# -------------------------------------------------------------------------------------------------
def gpu_scheduler(wl: Workload, add_tasks: bool = True, device: int = 0):
    from hyperqueue_b200 import GpuScheduler, RequestVariant, priority_from_user
    s = GpuScheduler(wl.R, device)
    for c, vs in enumerate(wl.classes):
        rid = s.get_or_create_resource_rq_id([RequestVariant.of(d["amounts"], d.get("all", ()), d.get("weight", 1.0),
                                                                d.get("min_time_s", 0.0)) for d in vs])
        assert rid == c
    s.new_workers_bulk(np.arange(wl.n_workers, dtype=np.uint32), wl.worker_total, wl.worker_free)
    if wl.worker_remaining_s is not None:
        s.termination = wl.worker_remaining_s.astype(np.float64)
    if wl.blocked is not None:
        s.set_blocked_mask(wl.blocked)
    if add_tasks:
        prio = priority_from_user(wl.task_user_priority)
        if wl.deps is None:
            s.add_ready_tasks(np.arange(wl.n_tasks, dtype=np.uint32), wl.task_class, prio)
        else:
            n_deps, off, cons = dag_csr(wl.deps)
            s.load_dag(wl.task_class, prio, n_deps, off, cons)
    return s


def judge_tick(wl: Workload, free_before: np.ndarray, a: np.ndarray, ready_mask: Optional[np.ndarray] = None):
    amounts, allm, nvar, mint = wl.class_tables()
    blocked = None if wl.blocked is None else wl.blocked
    return J.judge_assignments(amounts, allm, nvar, mint, free_before, wl.worker_total, wl.remaining_ms(), blocked,
                               wl.task_class, a["task"], a["worker"], a["variant"], ready_mask)


def gpu_drain(wl: Workload, max_ticks: int = 100000, judge: bool = True):
    """Zero-duration drain through the C ABI.  Every tick's output goes through the feasibility judge and
    the returned free vectors are compared with an exact replay."""
    s = gpu_scheduler(wl)
    amounts, allm, _, _ = wl.class_tables()
    remaining = wl.n_tasks
    per_tick = []
    ready = np.ones(wl.n_tasks, dtype=bool)
    if wl.deps is not None:
        ready = np.array([len(d) == 0 for d in wl.deps])
    done = np.zeros(wl.n_tasks, dtype=bool)
    unfinished = None if wl.deps is None else np.array([len(d) for d in wl.deps], dtype=np.int64)
    consumers = None
    if wl.deps is not None:
        consumers = [[] for _ in range(wl.n_tasks)]
        for t, ds in enumerate(wl.deps):
            for d in ds:
                consumers[d].append(t)
    while remaining > 0 and len(per_tick) < max_ticks:
        free_before = s.free.copy()
        m = s.run_scheduling()
        a = m.assignments
        if a.size == 0:
            raise RuntimeError(f"gpu drain stalled with {remaining} tasks left")
        if judge:
            res = judge_tick(wl, free_before, a, ready)
            assert res.ok, res
            exp = J.replay_free_after(amounts, allm, free_before, wl.worker_total, wl.task_class,
                                      a["task"], a["worker"], a["variant"])
            assert np.array_equal(exp, m.free_after), "free_after differs from the exact replay"
        ready[a["task"]] = False
        done[a["task"]] = True
        n_new = s.tasks_finished(a["task"], propagate=wl.deps is not None)
        if wl.deps is not None:
            newly = 0
            for t in a["task"].tolist():
                for c in consumers[t]:
                    unfinished[c] -= 1
                    if unfinished[c] == 0:
                        ready[c] = True
                        newly += 1
            assert newly == n_new, (newly, n_new)
        remaining -= a.size
        per_tick.append(int(a.size))
    assert np.array_equal(s.free, wl.worker_free), "resources did not return to the initial free vectors"
    s.close()
    return len(per_tick), per_tick


def smoke_check() -> None:
    """__graft_entry__.smoke(): one small tick on cuda:0 checked against the oracle."""
    wl = make_independent(4000, 8, 6, seed=3)
    s = gpu_scheduler(wl)
    free_before = s.free.copy()
    m = s.run_scheduling()
    res = judge_tick(wl, free_before, m.assignments)
    assert res.ok and m.n_assigned() > 0, res
    # the oracle's own tick on the same input fills at least 98 % of what we fill, and vice versa
    core = oracle_core(wl)
    core.scheduler_state.config.proactive_filling_max = 0
    ts, _, _, _ = oracle_tick(core)
    assert abs(ts.size - m.n_assigned()) <= max(2, 0.05 * ts.size), (ts.size, m.n_assigned())
    s.close()
