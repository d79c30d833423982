"""Pins the oracle's restatement of the autoalloc what-if query (oracle/query.py, scheduler/query.rs:12-131) on the
single-node cases of the reference's tests/test_query.rs (transcribed by hand: file:line in each test).  Multi-node
cases (:203-270) are not transcribed: multi-node requests are outside this path."""
import pytest

from oracle.query import WorkerTypeQuery as Q, compute_new_worker_query
from oracle_env import TaskBuilder, TestEnv, WorkerBuilder


def query(rt, *qs):
    return compute_new_worker_query(rt.core, list(qs), rt.now).single_node_workers_per_query


def test_query_no_tasks():                                             # :12-28
    rt = TestEnv()
    assert query(rt, Q.simple_cpus(4, max_sn_workers=2)) == [0]


def test_query_enough_workers():                                       # :30-51
    rt = TestEnv()
    rt.new_workers_cpus([2, 3])
    for c in (3, 1, 1):
        rt.new_task_cpus(c)
    rt.schedule()
    assert query(rt, Q.simple_cpus(4, max_sn_workers=2)) == [0]


def test_query_no_enough_workers1():                                   # :53-84
    rt = TestEnv()
    rt.new_workers_cpus([2, 3])
    for c in (3, 3, 1):
        rt.new_task_cpus(c)
    rt.schedule()
    assert query(rt, Q.simple_cpus(2, max_sn_workers=2), Q.simple_cpus(3, max_sn_workers=2)) == [0, 1]


def test_query_enough_workers2():                                      # :86-118
    rt = TestEnv()
    w1 = rt.new_worker(WorkerBuilder(2))
    rt.new_task_running(TaskBuilder(), w1)
    t = rt.new_task(TaskBuilder())
    rt.core.assign_task(t, w1)
    rt.schedule()
    assert query(rt, Q.simple_cpus(2, max_sn_workers=2), Q.simple_cpus(3, max_sn_workers=2)) == [0, 0]


def test_query_not_enough_workers3():                                  # :120-155
    rt = TestEnv()
    w1 = rt.new_worker(WorkerBuilder(2))
    rt.new_task_running(TaskBuilder(), w1)
    t = rt.new_task(TaskBuilder())
    rt.core.assign_task(t, w1)
    rt.new_task(TaskBuilder())
    rt.schedule()
    assert query(rt, Q.simple_cpus(2, max_sn_workers=2), Q.simple_cpus(3, max_sn_workers=2)) == [1, 0]


def test_query_many_workers_needed():                                  # :157-201
    rt = TestEnv()
    rt.new_workers_cpus([4, 4, 4])
    rt.new_tasks(100, TaskBuilder())
    rt.schedule()
    assert query(rt, Q.simple_cpus(2, max_sn_workers=5), Q.simple_cpus(1, max_sn_workers=1),
                 Q.simple_cpus(3, max_sn_workers=200)) == [5, 1, 26]


@pytest.mark.parametrize("mu,alloc,cpus", [(0.5, 0, 12), (0.3, 1, 12), (0.8, 0, 12), (1.0, 1, 5), (0.5, 2, 3), (0.7, 1, 3)])
def test_query_min_utilization1(mu, alloc, cpus):                      # :272-303
    rt = TestEnv()
    for c in (3, 1, 1):
        rt.new_task_cpus(c)
    rt.schedule()
    assert query(rt, Q.simple_cpus(cpus, max_sn_workers=2, min_utilization=mu)) == [alloc]


@pytest.mark.parametrize("mu,alloc,cpus,gpus", [(0.49, 1, 29, 40), (0.49, 0, 29, 30), (0.67, 0, 41, 30),
                                                 (0.50, 0, 41, 200), (0.45, 1, 39, 200)])
def test_query_min_utilization2(mu, alloc, cpus, gpus):                # :305-346
    rt = TestEnv()
    rt.new_named_resource("gpus")
    rt.new_tasks(2, TaskBuilder().cpus(10).add_resource(1, 20))
    rt.schedule()
    assert query(rt, Q([("cpus", cpus), ("gpus", gpus)], max_sn_workers=2, min_utilization=mu)) == [alloc]


def test_query_min_utilization3():                                     # :348-373
    rt = TestEnv()
    rt.new_tasks(2, TaskBuilder().cpus(2))
    assert query(rt, Q([("cpus", 4)], max_sn_workers=2, min_utilization=1.0)) == [1]


@pytest.mark.parametrize("cpu_tasks,gpu_tasks,alloc", [(1, 0, 0), (2, 0, 1), (3, 0, 1), (4, 1, 2), (1, 1, 1), (2, 1, 1),
                                                        (3, 1, 2), (4, 1, 2), (0, 1, 0), (0, 2, 1), (0, 3, 1), (0, 4, 2),
                                                        (0, 0, 0)])
def test_query_min_utilization_vs_partial(cpu_tasks, gpu_tasks, alloc):     # :375-417
    rt = TestEnv()
    rt.new_named_resource("gpus")
    rt.new_tasks(cpu_tasks, TaskBuilder().cpus(2))
    rt.new_tasks(gpu_tasks, TaskBuilder().cpus(2).add_resource(1, 1))
    assert query(rt, Q([("cpus", 4)], partial=True, max_sn_workers=2, min_utilization=1.0)) == [alloc]


@pytest.mark.parametrize("cpu_tasks,alloc", [(1, 1), (2, 1), (3, 1), (4, 1), (0, 0)])
def test_query_min_utilization_vs_partial2(cpu_tasks, alloc):          # :419-441
    rt = TestEnv()
    rt.new_tasks(cpu_tasks, TaskBuilder().cpus(2))
    assert query(rt, Q([], partial=True, max_sn_workers=2, min_utilization=1.0)) == [alloc]


@pytest.mark.parametrize("cpus,secs,alloc", [(2, 75, 0), (1, 101, 1), (4, 50, 1)])
def test_query_min_time2(cpus, secs, alloc):                           # :443-476
    rt = TestEnv()
    rt.new_task(TaskBuilder().cpus(1).time_request(100).next_variant().cpus(4).time_request(50))
    rt.schedule()
    assert query(rt, Q([("cpus", cpus)], time_limit=secs, max_sn_workers=2)) == [alloc]


def test_query_min_time1():                                            # :478-541
    rt = TestEnv()
    rt.new_task(TaskBuilder().cpus(1).time_request(100))
    rt.new_task(TaskBuilder().cpus(10).time_request(100))
    rt.schedule()
    assert query(rt, Q([("cpus", 10)], time_limit=99, max_sn_workers=2)) == [0]
    assert query(rt, Q([("cpus", 10)], time_limit=101, max_sn_workers=2)) == [2]
    assert query(rt, Q([("cpus", 1)], time_limit=101, max_sn_workers=2)) == [1]


@pytest.mark.parametrize("n,m", [(1, 0), (4, 0), (8, 0), (9, 1), (12, 1)])
def test_query_sn_leftovers1(n, m):                                    # :543-576
    rt = TestEnv()
    rt.new_workers_cpus([4])
    rt.new_tasks(n, TaskBuilder().cpus(1).time_request(5000))
    rt.schedule()
    assert query(rt, Q.simple_cpus(2, max_sn_workers=2), Q([], partial=True, max_sn_workers=2))[1] == m


@pytest.mark.parametrize("cpus,out", [(1, 0), (2, 3)])
def test_query_sn_leftovers2(cpus, out):                               # :578-597
    rt = TestEnv()
    rt.new_tasks(100, TaskBuilder().cpus(2))
    rt.schedule()
    assert query(rt, Q.simple_cpus(cpus, partial=True, max_sn_workers=3)) == [out]


def test_query_sn_leftovers():                                         # :599-638
    rt = TestEnv()
    rt.new_task(TaskBuilder().cpus(4).time_request(750))
    rt.new_task(TaskBuilder().cpus(8).time_request(1750))
    rt.schedule()
    assert query(rt, Q([], partial=True, time_limit=1000, max_sn_workers=3, max_workers_per_allocation=3),
                 Q([], partial=True, time_limit=50, max_sn_workers=3, max_workers_per_allocation=3),
                 Q([], partial=True, max_sn_workers=3, max_workers_per_allocation=3)) == [1, 0, 1]


def test_query_partial_query_cpus():                                   # :640-678
    rt = TestEnv()
    rt.new_task_cpus(4)
    rt.new_tasks(4, TaskBuilder().cpus(8))
    rt.schedule()
    assert query(rt, Q.simple_cpus(4, partial=True, max_sn_workers=2, max_workers_per_allocation=3),
                 Q.simple_cpus(16, partial=True, time_limit=50, max_sn_workers=5, max_workers_per_allocation=3),
                 Q([], partial=True, max_sn_workers=3, max_workers_per_allocation=3)) == [1, 2, 0]


@pytest.mark.parametrize("gpus,has_extra,out", [(4, False, 3), (4, True, 3), (None, False, 2), (None, True, 2),
                                                 (0, False, 0), (0, True, 0), (100, False, 2), (100, True, 2)])
def test_query_partial_query_gpus1(gpus, has_extra, out):              # :680-728
    rt = TestEnv()
    rt.new_named_resource("gpus")
    rt.new_named_resource("foo")
    b = TaskBuilder().cpus(1).add_resource(1, 2)
    if has_extra:
        b = b.add_resource(2, 1)
    rt.new_tasks(10, b)
    rt.schedule()
    items = [("cpus", 8)] + ([("gpus", gpus)] if gpus is not None else [])
    assert query(rt, Q(items, partial=True, max_sn_workers=3, max_workers_per_allocation=3)) == [out]


def test_query_unknown_do_not_add_extra():                             # :730-750
    rt = TestEnv()
    rt.new_task_default()
    rt.new_task(TaskBuilder().cpus(1).add_resource(1, 1))
    rt.new_task_default()
    rt.new_task(TaskBuilder().cpus(1).add_resource(1, 1))
    assert query(rt, Q.simple_cpus(1, partial=True, max_sn_workers=5, max_workers_per_allocation=3)) == [2]


def test_query_after_task_cancel():                                    # :752-771
    rt = TestEnv()
    t1 = rt.new_task_cpus(10)
    rt.new_worker(WorkerBuilder(1))
    rt.schedule()
    rt.core.remove_from_ready_queue(t1)                                # on_cancel_tasks, reactor.rs
    del rt.core.tasks[t1]
    assert query(rt, Q([], partial=True, max_sn_workers=5, max_workers_per_allocation=3)) == [0]
