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

from workloads import (FR, MAXV, Workload, dag_csr, gpu_scheduler, make_cfg1, make_dag, make_independent,  # noqa: F401
                       _class_pool, _zipf_classes)


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


def oracle_drain(wl: Workload, max_ticks: int = 100000, disable_prefill: bool = True, solver_opts: Optional[dict] = None):
    """Zero-duration drain (cfg(zero_worker) semantics): tick, every assigned task finishes at once,
    resources return, newly ready consumers enter the queues; repeat.  Returns (ticks, per-tick counts)."""
    core = oracle_core(wl)
    if disable_prefill:
        core.scheduler_state.config.proactive_filling_max = 0     # prefill is a latency hider, not capacity
    remaining = wl.n_tasks
    per_tick = []
    while remaining > 0 and len(per_tick) < max_ticks:
        opts = dict(solver_opts or {})
        ts, ws, vs, _ = oracle_tick(core, **opts)
        while ts.size == 0 and opts.get("time_limit") and opts["time_limit"] < 300:
            # HiGHS found no incumbent inside the cap (large pools): the same tick again with twice the time
            opts["time_limit"] *= 2
            ts, ws, vs, _ = oracle_tick(core, **opts)
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
    """__graft_entry__.smoke(): one small tick on cuda:0, checked (a) by the oracle's feasibility judge,
    (b) against an exact replay of the free vectors, (c) bit for bit against the sequential specification of
    the device algorithm, and (d) against the oracle's own tick on the same input: the resources the two
    ticks put to use must be comparable (the MILP may choose a different task mix)."""
    import greedy_model as G
    wl = make_independent(4000, 8, 6, seed=3)
    s = gpu_scheduler(wl)
    free_before = s.free.copy()
    m = s.run_scheduling()
    res = judge_tick(wl, free_before, m.assignments)
    assert res.ok and m.n_assigned() > 0, res
    amounts, allm, _, _ = wl.class_tables()
    exp_free = J.replay_free_after(amounts, allm, free_before, wl.worker_total, wl.task_class,
                                   m.assignments["task"], m.assignments["worker"], m.assignments["variant"])
    assert np.array_equal(exp_free, m.free_after)
    spec, spec_free = G.model_tick(wl, np.ones(wl.n_tasks, dtype=bool), free_before)
    assert np.array_equal(spec, m.assignments) and np.array_equal(spec_free, m.free_after)
    core = oracle_core(wl)
    core.scheduler_state.config.proactive_filling_max = 0
    ts, ws, vs, _ = oracle_tick(core)
    assert J.judge_assignments(amounts, allm, *wl.class_tables()[2:], free_before, wl.worker_total, wl.remaining_ms(),
                               None, wl.task_class, ts, ws, vs).ok
    used_gpu = amounts[wl.task_class[m.assignments["task"]], m.assignments["variant"]].astype(np.float64).sum(0)
    used_ref = amounts[wl.task_class[ts], vs].astype(np.float64).sum(0)
    cap = free_before.sum(0).astype(np.float64)
    u_gpu, u_ref = (used_gpu / cap).max(), (used_ref / cap).max()
    assert u_gpu >= 0.9 * u_ref, (u_gpu, u_ref, m.n_assigned(), ts.size)
    # ... and the resource-weighted work of the tick (sum over resources of used / capacity: what the reference's
    # objective adds up, solver.rs:520-549) within 15 % of the oracle's, so that the per-tick task MIX is bounded too
    w_gpu, w_ref = float((used_gpu / cap).sum()), float((used_ref / cap).sum())
    assert w_gpu >= 0.85 * w_ref, (w_gpu, w_ref)
    print(f"smoke: gpu {m.n_assigned()} tasks (bottleneck utilisation {u_gpu:.3f}, resource-weighted work {w_gpu:.3f}), "
          f"oracle {ts.size} tasks ({u_ref:.3f}, {w_ref:.3f})")
    s.close()
