"""Multi-tick scenarios of the reference's proactive-filling tests (test_scheduler_sn.rs:849-871, 1168-1306) in a neutral
form: a list of steps, each adding workers (cpus) and / or 1-class tasks and then ticking.  Three drivers run them:
  run_oracle   the restated reference (oracle/, TestEnv)            -> per tick {worker: (assigned, prefilled, retracted)}
  run_spec     the specification of the device algorithm            -> the same + the raw records
  run_gpu      the CUDA path through GpuScheduler (tests/test_gpu_prefill.py)
Tasks never finish inside a scenario (as in the reference tests), so assigned tasks keep their resources."""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

import greedy_model as G
from workloads import FR, Workload

# name -> (reserve, max, task cpus, steps [(new worker cpus [...], new tasks)])
SCENARIOS = {
    "prefill_basic": (4, 32, 4, [([8, 8], 300)]),                       # :1168-1200  34 tasks per worker, 32 of them prefills
    "no_deps_distribute": (10, 20, 1, [([10, 10, 10], 150)]),           # :849-871    30 tasks per worker
    "prefill_choose_waiting": (3, 6, 1, [([1], 15), ([1], 0), ([1], 0)]),   # :1202-1223  prefills 6 / 4 / 0
    "prefill_steal": (3, 6, 1, [([1], 9), ([5], 0)]),                   # :1225-1306  2 retracts, 3 fresh tasks to the new worker
    "prefill_many_ticks": (2, 5, 2, [([4, 4], 40), ([6], 0), ([], 10), ([8, 2], 0), ([], 0)]),
}


def run_oracle(name: str):
    from oracle_env import TaskBuilder, TestEnv, WorkerBuilder
    reserve, pmax, cpus, steps = SCENARIOS[name]
    rt = TestEnv(); rt.set_scheduler_config(reserve, pmax)
    wids, out = [], []
    for new_w, new_t in steps:
        for c in new_w:
            wids.append(rt.new_worker(WorkerBuilder(c)))
        if new_t:
            rt.new_tasks(new_t, TaskBuilder().cpus(cpus))
        m = rt.schedule()
        tick = {}
        for i, w in enumerate(wids):
            up = m.workers.get(w)
            tick[i] = (len(up.assigned), len(up.prefills), len(up.retracts)) if up is not None else (0, 0, 0)
        out.append(tick)
    return out


class SpecDriver:
    """Keeps what the host keeps between ticks: ready mask, free vectors, prefill owner per task."""

    def __init__(self, name: str) -> None:
        self.reserve, self.pmax, self.cpus, self.steps = SCENARIOS[name]
        self.total = np.zeros((0, 1), dtype=np.uint64)
        self.free = np.zeros((0, 1), dtype=np.uint64)
        self.n_tasks = 0
        self.ready = np.zeros(0, dtype=bool)
        self.pf_worker = np.zeros(0, dtype=np.int64)

    def add(self, new_w, new_t) -> None:
        for c in new_w:
            row = np.array([[c * FR]], dtype=np.uint64)
            self.total = np.concatenate([self.total, row]); self.free = np.concatenate([self.free, row])
        if new_t:
            self.ready = np.concatenate([self.ready, np.ones(new_t, dtype=bool)])
            self.pf_worker = np.concatenate([self.pf_worker, np.full(new_t, -1, dtype=np.int64)])
            self.n_tasks += new_t

    def workload(self) -> Workload:
        return Workload(1, [[{"amounts": {0: self.cpus * FR}}]], self.total.copy(), self.free.copy(),
                        np.zeros(self.n_tasks, dtype=np.uint32), np.zeros(self.n_tasks, dtype=np.int32))

    def tick(self):
        wl = self.workload()
        pf_before = self.pf_worker.copy()
        a, fa = G.model_tick(wl, self.ready, self.free, prefill=(self.reserve, self.pmax), pf_worker=self.pf_worker)
        self.free = fa.copy()
        self.ready[a["task"][a["kind"] != 1]] = False
        return a, pf_before


def summarize(a: np.ndarray, pf_before: np.ndarray, n_workers: int) -> Dict[int, Tuple[int, int, int]]:
    out = {}
    for w in range(n_workers):
        assigned = int(np.count_nonzero((a["worker"] == w) & (a["kind"] == 0)))
        prefilled = int(np.count_nonzero((a["worker"] == w) & (a["kind"] == 1)))
        retracted = int(np.count_nonzero(pf_before[a["task"][a["kind"] == 2]] == w))
        out[w] = (assigned, prefilled, retracted)
    return out


def run_spec(name: str):
    d = SpecDriver(name)
    out, recs = [], []
    for new_w, new_t in d.steps:
        d.add(new_w, new_t)
        a, pf_before = d.tick()
        out.append(summarize(a, pf_before, d.total.shape[0]))
        recs.append(a)
    return out, recs
