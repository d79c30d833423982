"""Synthetic workloads of the BASELINE.md §4 shapes and the helper that loads one into a GpuScheduler.
Shared by the tests and bench.py; imports neither the oracle nor the sequential model."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

TIME_INF = (1 << 64) - 1

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
            return np.full(W, TIME_INF, dtype=np.uint64)
        return np.where(np.isinf(self.worker_remaining_s), np.uint64(TIME_INF),
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


def make_cfg1(n: int = 10_000, w: int = 4) -> Workload:
    """BASELINE.json configs[0]: n independent 1-cpu tasks, w workers x {cpus 128}, one class, priority 0 — the shape of
    the reference's experiment-per-task-overhead.py."""
    total = np.tile(np.array([128 * FR], dtype=np.uint64), (w, 1))
    return Workload(1, [[{"amounts": {0: 1 * FR}}]], total, total.copy(), np.zeros(n, dtype=np.uint32),
                    np.zeros(n, dtype=np.int32), name=f"cfg1 n={n} w={w}")


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



def gpu_scheduler(wl: Workload, add_tasks: bool = True, device: int = 0, flags: int = 0):
    from hyperqueue_b200 import GpuScheduler, RequestVariant, priority_from_user
    s = GpuScheduler(wl.R, device, flags)
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


