"""The autoalloc what-if query (scheduler/query.rs) answered by the DEVICE algorithm: the reference's tests/test_query.rs
cases, with the ready set produced by the oracle's own tick.  The reference asks "which of these fake workers would
receive at least one task".  Two backends run the same cases:
  spec  tests/greedy_model.py, the sequential specification (CPU; this module's own tests)
  gpu   hqs_query through GpuScheduler.new_worker_query (tests/test_gpu_query.py re-runs every test of this module on the
        device and additionally requires the per-worker counts to equal the specification's)
min_utilization (query.rs:35-46 -> solver.rs:479-518) is part of the tick in both."""
import numpy as np
import pytest

import greedy_model as G
from oracle.model import AMOUNT_MAX, units
from oracle_env import TaskBuilder, TestEnv, WorkerBuilder
from workloads import Workload

_BACKEND = "spec"


def query_workload(rt, queries):
    """queries: [(resources [(name, units)], partial, time_limit_s, max_sn_workers, min_utilization)] ->
    (Workload over the ready set with the fake workers as its pool, min_utilization per worker, owner query per worker)"""
    core = rt.core
    for res, *_ in queries:
        for name, _ in res:
            core.get_or_create_resource_id(name)
    R = max(core.n_resources(), 1 + max((e.resource_id for c in range(len(core.rq_map)) for rq in core.rq_map.get(c).variants
                                          for e in rq.entries), default=0))
    ready = [t for t in core.tasks.values() if (t.state == "waiting" and t.is_ready()) or t.state == "prefilled"]     # prefilled tasks stay in the queues (taskqueue.rs:273-302)
    if not ready:
        return None, None, None
    classes = []
    for c in range(len(core.rq_map)):
        vs = []
        for rq in core.rq_map.get(c).variants:
            d = {"amounts": {e.resource_id: e.amount for e in rq.entries if not e.is_all()},
                 "all": tuple(e.resource_id for e in rq.entries if e.is_all()),
                 "weight": rq.weight / 10000.0, "min_time_s": rq.min_time}
            vs.append(d)
        classes.append(vs)
    tot, rem, mus, owner = [], [], [], []
    for qi, (res, partial, tl, n, mu) in enumerate(queries):
        for _ in range(n):
            vec = [AMOUNT_MAX if (partial and r < core.n_resources()) else 0 for r in range(R)]
            for name, amount in res:
                vec[core.resource_names.index(name)] = units(amount)
            tot.append(vec); rem.append(np.inf if tl is None else float(tl)); mus.append(mu); owner.append(qi)
    total = np.array(tot, dtype=np.uint64)
    wl = Workload(R, classes, total, total.copy(), np.array([t.rq_id for t in ready], dtype=np.uint32),
                  np.array([t.user_priority for t in ready], dtype=np.int32), worker_remaining_s=np.array(rem))
    return wl, np.array(mus, dtype=np.float32), owner


def spec_query(rt, queries):
    wl, mus, owner = query_workload(rt, queries)
    if wl is None:
        return [0 for _ in queries]
    mu_arg = mus if (mus > 0.001).any() else None
    a, free_after = G.model_tick(wl, np.ones(wl.n_tasks, dtype=bool), wl.worker_free, min_utilization=mu_arg)
    counts = np.bincount(a["worker"], minlength=wl.n_workers)
    if _BACKEND == "gpu":
        import workloads as WL
        s = WL.gpu_scheduler(wl)
        needed, got, n_total = s.new_worker_query(wl.worker_total, remaining_s=wl.worker_remaining_s, min_utilization=mus)
        # a dry run: the ready set is untouched, a real tick over the same workers still assigns the same tasks
        s.min_utilization = mus.copy()
        m = s.run_scheduling()
        s.close()
        assert np.array_equal(got, counts), (got, counts)            # device == specification, worker by worker
        assert n_total == int(counts.sum()) == m.n_assigned()
        counts = got
    loaded = counts > 0
    return [int(sum(1 for w in range(wl.n_workers) if owner[w] == qi and loaded[w])) for qi in range(len(queries))]


def q(res, partial=False, tl=None, n=1, mu=0.0):
    return (res, partial, tl, n, mu)


def cpus(c, **kw):
    return q([("cpus", c)], **kw)


def test_enough_and_not_enough_workers():                              # test_query.rs:30-155
    rt = TestEnv(); rt.new_workers_cpus([2, 3])
    for c in (3, 1, 1):
        rt.new_task_cpus(c)
    rt.schedule()
    assert spec_query(rt, [cpus(4, n=2)]) == [0]
    rt = TestEnv(); rt.new_workers_cpus([2, 3])
    for c in (3, 3, 1):
        rt.new_task_cpus(c)
    rt.schedule()
    assert spec_query(rt, [cpus(2, n=2), cpus(3, n=2)]) == [0, 1]
    rt = TestEnv(); w1 = rt.new_worker(WorkerBuilder(2))
    rt.new_task_running(TaskBuilder(), w1)
    t = rt.new_task(TaskBuilder()); rt.core.assign_task(t, w1)
    rt.new_task(TaskBuilder())
    rt.schedule()
    assert spec_query(rt, [cpus(2, n=2), cpus(3, n=2)]) == [1, 0]


def test_many_workers_needed():                                        # :157-201
    rt = TestEnv(); rt.new_workers_cpus([4, 4, 4])
    rt.new_tasks(100, TaskBuilder())
    rt.schedule()
    assert spec_query(rt, [cpus(2, n=5), cpus(1, n=1), cpus(3, n=200)]) == [5, 1, 26]


@pytest.mark.parametrize("mu,alloc,c", [(0.5, 0, 12), (0.3, 1, 12), (0.8, 0, 12), (1.0, 1, 5), (0.5, 2, 3), (0.7, 1, 3)])
def test_min_utilization1(mu, alloc, c):                               # :272-303
    rt = TestEnv()
    for x in (3, 1, 1):
        rt.new_task_cpus(x)
    rt.schedule()
    assert spec_query(rt, [cpus(c, n=2, mu=mu)]) == [alloc]


@pytest.mark.parametrize("c,secs,alloc", [(2, 75, 0), (1, 101, 1), (4, 50, 1)])
def test_min_time2(c, secs, alloc):                                    # :443-476
    rt = TestEnv()
    rt.new_task(TaskBuilder().cpus(1).time_request(100).next_variant().cpus(4).time_request(50))
    rt.schedule()
    assert spec_query(rt, [cpus(c, tl=secs, n=2)]) == [alloc]


def test_min_time1():                                                  # :478-541
    rt = TestEnv()
    rt.new_task(TaskBuilder().cpus(1).time_request(100))
    rt.new_task(TaskBuilder().cpus(10).time_request(100))
    rt.schedule()
    assert spec_query(rt, [cpus(10, tl=99, n=2)]) == [0]
    assert spec_query(rt, [cpus(10, tl=101, n=2)]) == [2]
    assert spec_query(rt, [cpus(1, tl=101, n=2)]) == [1]


@pytest.mark.parametrize("n,m", [(1, 0), (4, 0), (8, 0), (9, 1), (12, 1)])
def test_sn_leftovers1(n, m):                                          # :543-576
    rt = TestEnv(); rt.new_workers_cpus([4])
    rt.new_tasks(n, TaskBuilder().cpus(1).time_request(5000))
    rt.schedule()
    assert spec_query(rt, [cpus(2, n=2), q([], partial=True, n=2)])[1] == m


@pytest.mark.parametrize("c,out", [(1, 0), (2, 3)])
def test_sn_leftovers2(c, out):                                        # :578-597
    rt = TestEnv()
    rt.new_tasks(100, TaskBuilder().cpus(2))
    rt.schedule()
    assert spec_query(rt, [cpus(c, partial=True, n=3)]) == [out]


def test_sn_leftovers_and_partial_cpus():                              # :599-678
    rt = TestEnv()
    rt.new_task(TaskBuilder().cpus(4).time_request(750))
    rt.new_task(TaskBuilder().cpus(8).time_request(1750))
    rt.schedule()
    assert spec_query(rt, [q([], True, 1000, 3), q([], True, 50, 3), q([], True, None, 3)]) == [1, 0, 1]
    rt = TestEnv()
    rt.new_task_cpus(4)
    rt.new_tasks(4, TaskBuilder().cpus(8))
    rt.schedule()
    assert spec_query(rt, [cpus(4, partial=True, n=2), cpus(16, partial=True, tl=50, n=5), q([], True, None, 3)]) == [1, 2, 0]


@pytest.mark.parametrize("gpus,has_extra,out", [(4, False, 3), (4, True, 3), (None, False, 2), (None, True, 2),
                                                 (0, False, 0), (0, True, 0), (100, False, 2), (100, True, 2)])
def test_partial_query_gpus1(gpus, has_extra, out):                    # :680-728
    rt = TestEnv()
    rt.new_named_resource("gpus"); rt.new_named_resource("foo")
    b = TaskBuilder().cpus(1).add_resource(1, 2)
    if has_extra:
        b = b.add_resource(2, 1)
    rt.new_tasks(10, b)
    rt.schedule()
    items = [("cpus", 8)] + ([("gpus", gpus)] if gpus is not None else [])
    assert spec_query(rt, [q(items, partial=True, n=3)]) == [out]


def test_unknown_do_not_add_extra():                                   # :730-750
    rt = TestEnv()
    rt.new_task_default()
    rt.new_task(TaskBuilder().cpus(1).add_resource(1, 1))
    rt.new_task_default()
    rt.new_task(TaskBuilder().cpus(1).add_resource(1, 1))
    assert spec_query(rt, [cpus(1, partial=True, n=5)]) == [2]
